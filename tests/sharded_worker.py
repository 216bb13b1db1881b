"""One rank of the multi-GPU parity test (launched by torchrun, one process per GPU): uploads this
rank's row range of the fact table plus the full build-side tables, executes the plan fixtures with
RQ_PLAN_SHARDED (partials merged inside the library over NCCL) and compares every rank's result
with the reference engine's output on the unsharded data (tests/golden/sf001)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

CASES = [("q1", "lineitem"), ("q6", "lineitem"), ("q3", "lineitem"), ("agg_nogroup_minmax", "lineitem"),
         ("agg_many_groups", "lineitem"), ("agg_empty", "lineitem"), ("agg_linenumber", "lineitem"),
         ("agg_wrap", "lineitem"), ("case_sum", "lineitem"), ("sel_or", "lineitem"),
         ("join_orders_lineitem", "lineitem"), ("sort_large", "lineitem"), ("join_dups_agg", "orders"),
         ("join_dups_rows", "orders")]


# integer-only join / aggregation shapes (string values cannot cross an exchange yet)
PART_CASES = ["micro_join_avg", "micro_join_few_groups", "q3", "q1", "q6", "agg_many_groups", "agg_empty", "sort_large",
              "join_dups_rows", "agg_linenumber"]


def main():
    import torch.distributed as dist
    from common import load_plan_dict, load_golden, plan_tables, serialize_columns, assert_same_relation
    from resql_b200 import Engine, Plan, tpch
    from resql_b200 import native as N
    from resql_b200.shard import shard_columns
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    dist.init_process_group("gloo")            # control plane only: carries the NCCL id
    eng = Engine(local_rank)
    uid = [eng.dist_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    eng.dist_init(rank, world, uid[0])
    if os.environ.get("RQ_TEST_TRACE"):
        eng.set_option("trace", 1)
    progress = bool(os.environ.get("RQ_TEST_PROGRESS"))

    def note(*a):
        if progress:
            print(f"[rank {rank}]", *a, file=sys.stderr, flush=True)
    data = tpch.generate(0.01, seed=42)
    failures = []
    for name, fact in CASES:
        d = load_plan_dict(name)
        tabs = plan_tables(d, data)
        if fact not in tabs:
            continue
        tabs[fact] = shard_columns(tabs[fact], rank, world)
        handles = {n: eng.upload(n, c) for n, c in tabs.items()}
        try:
            # three executions: careful, careful (warm memos), replayed without host waits
            for rep in range(6):
                note("sharded", name, "run", rep)
                res, tm = eng.execute(Plan(d), handles, N.RQ_PLAN_SHARDED)
                got = serialize_columns(res.columns, res.sql_types, res.sql_widths)
                _, want = load_golden(name)
                assert_same_relation(got, want, d, f"{name} on {world} GPUs (rank {rank}, run {rep})")
            if rank == 0:
                print(f"sharded {name}: {res.n_rows} rows identical on {world} GPUs, nccl_ms={tm.nccl_ms:.3f} "
                      f"host_syncs={tm.host_syncs}", flush=True)
        except Exception as e:  # noqa: BLE001 - collect, report after all ranks are through the collectives
            failures.append(f"{name}: {e}")
        finally:
            for h in handles.values():
                h.free()
    # replicated tables uploaded once (rank 0) and broadcast over NVLink: the other ranks never see the data
    for name, fact in [("q3", "lineitem"), ("q10", "lineitem"), ("join_dups_agg", "orders")]:
        d = load_plan_dict(name)
        tabs = plan_tables(d, data)
        if fact not in tabs:
            continue
        handles = {}
        for n, c in tabs.items():
            if n == fact:
                handles[n] = eng.upload(n, shard_columns(c, rank, world))
            else:
                import numpy as np
                blind = c if rank == 0 else {k: np.zeros_like(v) for k, v in c.items()}
                handles[n] = eng.upload_replicated(n, blind, 0, rank)
        try:
            res, tm = eng.execute(Plan(d), handles, N.RQ_PLAN_SHARDED)
            got = serialize_columns(res.columns, res.sql_types, res.sql_widths)
            _, want = load_golden(name)
            assert_same_relation(got, want, d, f"{name} with broadcast build tables on {world} GPUs (rank {rank})")
            if rank == 0:
                print(f"broadcast {name}: {res.n_rows} rows identical on {world} GPUs", flush=True)
        except Exception as e:  # noqa: BLE001
            failures.append(f"broadcast {name}: {e}")
        finally:
            for h in handles.values():
                h.free()
    # shared builds (replicated pure-scan build sides split over the ranks, bitmaps all-reduced), forced
    # on these small tables; q3's customer build qualifies
    eng.set_option("share_min_rows", 0)
    for name, fact in [("q3", "lineitem"), ("join_orders_lineitem", "lineitem"), ("join_dups_agg", "orders"), ("join_cust_orders", "orders")]:
        d = load_plan_dict(name)
        tabs = plan_tables(d, data)
        if fact not in tabs:
            continue
        tabs[fact] = shard_columns(tabs[fact], rank, world)
        handles = {n: eng.upload(n, c) for n, c in tabs.items()}
        try:
            for rep in range(5):
                note("shared-build", name, "run", rep)
                res, tm = eng.execute(Plan(d), handles, N.RQ_PLAN_SHARDED)
                got = serialize_columns(res.columns, res.sql_types, res.sql_widths)
                _, want = load_golden(name)
                assert_same_relation(got, want, d, f"{name} with shared builds on {world} GPUs (rank {rank}, run {rep})")
            if rank == 0:
                print(f"shared-build {name}: {res.n_rows} rows identical on {world} GPUs, nccl_ms={tm.nccl_ms:.3f} "
                      f"host_syncs={tm.host_syncs}", flush=True)
        except Exception as e:  # noqa: BLE001
            failures.append(f"shared-build {name}: {e}")
        finally:
            for h in handles.values():
                h.free()
    eng.set_option("share_min_rows", 1 << 20)
    # RQ_PLAN_PARTITIONED: EVERY table is a row range; build and probe rows meet on the rank that owns
    # hash(join key) (all-to-all), groups are merged on the rank that owns hash(group key)
    for name in PART_CASES:
        d = load_plan_dict(name)
        tabs = {n: shard_columns(c, rank, world) for n, c in plan_tables(d, data).items()}
        handles = {n: eng.upload(n, c) for n, c in tabs.items()}
        try:
            for rep in range(6):
                note("partitioned", name, "run", rep)
                res, tm = eng.execute(Plan(d), handles, N.RQ_PLAN_PARTITIONED)
                got = serialize_columns(res.columns, res.sql_types, res.sql_widths)
                _, want = load_golden(name)
                assert_same_relation(got, want, d, f"{name} partitioned over {world} GPUs (rank {rank}, run {rep})")
            if rank == 0:
                print(f"partitioned {name}: {res.n_rows} rows identical on {world} GPUs, nccl_ms={tm.nccl_ms:.3f} "
                      f"host_syncs={tm.host_syncs}", flush=True)
        except Exception as e:  # noqa: BLE001
            failures.append(f"partitioned {name}: {e}")
        finally:
            for h in handles.values():
                h.free()
    # random plans (tests/plan_fuzz.py): the probe-side table of the main pipeline is the sharded fact table
    from oracle.plan_oracle import run_plan
    from plan_fuzz import random_plan
    n_fuzz = 0
    for seed in range(60):
        d = random_plan(seed)
        tabs = plan_tables(d, data)
        try:
            want = serialize_columns(*run_plan(d, tabs))
        except ZeroDivisionError:
            continue
        fact = d["tables"][-1]["name"]
        tabs[fact] = shard_columns(tabs[fact], rank, world)
        handles = {n: eng.upload(n, c) for n, c in tabs.items()}
        try:
            note("random plan", seed)
            res, tm = eng.execute(Plan(d), handles, N.RQ_PLAN_SHARDED)
            assert_same_relation(serialize_columns(res.columns, res.sql_types, res.sql_widths), want, d,
                                 f"random plan {seed} on {world} GPUs (rank {rank})")
            n_fuzz += 1
        except N.EngineError as e:
            if e.code != 3:
                failures.append(f"random plan {seed}: {e}")
        except Exception as e:  # noqa: BLE001
            failures.append(f"random plan {seed}: {e}")
        finally:
            for h in handles.values():
                h.free()
    if rank == 0:
        print(f"sharded random plans: {n_fuzz} identical on {world} GPUs", flush=True)
    # a data-dependent failure on ONE shard (division by zero on the last rank only) must end the plan
    # with an error on EVERY rank instead of leaving the others in the merge collective
    import numpy as np
    dz = {"tables": [{"name": "t", "columns": ["a", "b"]}],
          "pipelines": [{"source_kind": 1, "source_id": 0, "sink_kind": 1, "size_hint": 0,
                         "nodes": [[1, 0, 0, 0, 0], [1, 1, 0, 0, 0], [7, 0, 1, 0, 0]], "args": [],
                         "keys": [], "vals": [[2, 1, 4, 0]]}],
          "order": [], "limit": -1, "strpool": ""}
    b = np.ones(4096, dtype=np.int64)
    if rank == world - 1:
        b[100] = 0
    t = eng.upload("t", {"a": np.arange(4096, dtype=np.int64), "b": b})
    try:
        note("error agreement")
        eng.execute(Plan(dz), {"t": t}, N.RQ_PLAN_SHARDED)
        failures.append("division by zero on one shard was not reported on rank %d" % rank)
    except N.EngineError as e:
        if rank == 0:
            print(f"sharded error agreement: every rank raised ({e})", flush=True)
    finally:
        t.free()
    flags = [None] * world
    dist.all_gather_object(flags, failures)
    eng.shutdown()
    dist.destroy_process_group()
    bad = [f for fl in flags for f in fl]
    if bad:
        print("\n".join(bad), file=sys.stderr)
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
