"""Model of the library's sharded execution (engine_exec.inl merge_sharded) on top of the plan
oracle: every rank runs the plan on its row range of the fact table, the output of the merge-point
pipeline is all-gathered in rank order and re-aggregated (SUM and COUNT partials add, MIN/MAX take
min/max), the rest of the plan runs on the merged relation. Test infrastructure only."""
import numpy as np

from oracle import plan_oracle as PO
from resql_b200.shard import merge_point, shard_columns


def run_plan_sharded(plan, tables, fact, rank, world, all_gather):
    """all_gather(obj) -> list of every rank's obj, in rank order"""
    local = dict(tables)
    local[fact] = shard_columns(tables[fact], rank, world)
    pool = plan.get("strpool", "").encode("latin1")
    mp = merge_point(plan)
    outs = []
    old = np.seterr(over="ignore")
    try:
        for i, p in enumerate(plan["pipelines"]):
            out = PO._run_pipeline(plan, p, local, outs, pool)
            if i == mp:
                cols, st, sw, _ = out
                parts = all_gather([np.asarray(c) for c in cols])
                cat = [np.concatenate([pt[c] for pt in parts]) if parts else cols[c] for c in range(len(cols))]
                if p["sink_kind"] == 1:
                    nk, nv = len(p["keys"]), len(p["vals"])
                    mpipe = {"source_kind": 2, "source_id": 0, "sink_kind": 1, "size_hint": 0, "args": [],
                             "nodes": [[1, c, 0, 0, 0] for c in range(nk + nv)],
                             "keys": [[c, 0, k[2], k[3]] for c, k in enumerate(p["keys"])],
                             "vals": [[nk + c, 1 if v[1] == 2 else v[1], v[2], v[3]] for c, v in enumerate(p["vals"])]}
                    out = PO._run_pipeline(plan, mpipe, {}, [(cat, st, sw, None)], pool)
                else:
                    out = (cat, st, sw, None)
            outs.append(out)
    finally:
        np.seterr(**old)
    cols, st, sw = outs[-1][:3]
    return PO.finish(plan, cols, st, sw)
