"""N > 1 path on CPU: two gloo ranks each run the plan on their row range of the fact table, the
merge-point relation is all-gathered and re-aggregated exactly like the library does it
(tests/shard_model.py mirrors engine_exec.inl merge_sharded), and every rank must end with the
result of the unsharded plan - i.e. the reference engine's output (tests/golden/sf001)."""
import os
import socket
import sys

import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = [("q1", "lineitem"), ("q6", "lineitem"), ("q3", "lineitem"), ("agg_neg_avg", "customer"),
         ("agg_nogroup_minmax", "lineitem"), ("agg_many_groups", "lineitem"), ("agg_empty", "lineitem"),
         ("sel_or", "lineitem"), ("join_orders_lineitem", "lineitem")]


def _worker(rank, world, port, failures):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from common import load_plan_dict, load_golden, plan_tables, serialize_columns, assert_same_relation
    from resql_b200 import tpch
    from shard_model import run_plan_sharded
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        data = tpch.generate(0.01, seed=42)

        def all_gather(obj):
            out = [None] * world
            dist.all_gather_object(out, obj)
            return out

        for name, fact in CASES:
            d = load_plan_dict(name)
            tabs = plan_tables(d, data)
            if fact not in tabs:
                fact = next(iter(tabs))
            got = serialize_columns(*run_plan_sharded(d, tabs, fact, rank, world, all_gather))
            _, want = load_golden(name)
            try:
                assert_same_relation(got, want, d, f"{name} sharded over {world} ranks (rank {rank})")
            except AssertionError as e:
                failures.put(str(e))
        # random plans (tests/plan_fuzz.py): the probe-side table of the main pipeline is sharded
        from oracle.plan_oracle import run_plan
        from plan_fuzz import random_plan
        for seed in range(40):
            d = random_plan(seed)
            tabs = plan_tables(d, data)
            try:
                want = serialize_columns(*run_plan(d, tabs))
            except ZeroDivisionError:
                continue
            got = serialize_columns(*run_plan_sharded(d, tabs, d["tables"][-1]["name"], rank, world, all_gather))
            try:
                assert_same_relation(got, want, d, f"random plan {seed} sharded over {world} ranks (rank {rank})")
            except AssertionError as e:
                failures.put(str(e))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_merge_equals_unsharded(world):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    failures = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, failures)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0, f"rank exited with {p.exitcode}"
    msgs = []
    while not failures.empty():
        msgs.append(failures.get())
    assert not msgs, "\n".join(msgs)


def test_row_ranges_tile_the_table():
    from resql_b200.shard import row_range
    for n in (0, 1, 7, 1000, 59_986_052):
        for world in (1, 2, 3, 4, 8):
            edges = [row_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
