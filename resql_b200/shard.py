"""Row-range sharding of a fact table over the ranks of one box (one process per GPU).

The merge of the per-rank partial results happens inside the library (engine_exec.inl,
merge_sharded); this module only says which rows a rank owns and which pipeline the library merges
after, so harnesses and tests agree with it."""

SINK_AGG = 1


def row_range(n_rows, rank, world):
    """contiguous [lo, hi) of rank `rank`; the ranges of all ranks tile [0, n_rows)"""
    return n_rows * rank // world, n_rows * (rank + 1) // world


def shard_columns(cols, rank, world):
    n = len(next(iter(cols.values())))
    lo, hi = row_range(n, rank, world)
    return {c: v[lo:hi] for c, v in cols.items()}


def merge_point(plan_dict):
    """index of the pipeline whose output the library exchanges under RQ_PLAN_SHARDED: the last
    aggregation, or the final relation when the plan has no aggregation"""
    last = -1
    for i, p in enumerate(plan_dict["pipelines"]):
        if p["sink_kind"] == SINK_AGG:
            last = i
    return last if last >= 0 else len(plan_dict["pipelines"]) - 1
