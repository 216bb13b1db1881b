"""The drop-in against the reference at a larger scale, both as complete programs: the reference
engine (oracle/_ref/resql-oracle: its parser, planner and x86 JIT) and resql-b200 (the same parser and
planner, our shim and the B200 engine) load the same packed tuples and run all eight statements of
the reference's tpch/queries directory plus the README microbenchmark; outputs must be identical
(same tuples; same ORDER BY key sequence)."""
import os
import subprocess

import pytest

from common import ROOT, load_plan_dict, assert_same_relation
from resql_b200 import tpch

pytestmark = pytest.mark.gpu

QUERIES = ["q1", "q6", "q3", "q5", "q10", "q12", "q14", "q19", "micro_join_avg", "agg_many_groups", "join_dups_agg"]


def test_all_tpch_queries_match_the_reference_engine_at_sf05(tmp_path):
    _program_vs_program(tmp_path, 0.5, 1)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_gpus_n_drop_in_matches_the_reference_engine(tmp_path, world):
    """`gpus=N` (the counterpart of the reference's `threads=N`, execute.h:454-474): resql-b200 starts one
    process per GPU, shards the fact table of every select by row range and merges over NCCL; rank 0's
    output files must equal the reference engine's."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    _program_vs_program(tmp_path, 0.2, world)


def test_gpus_2_with_shared_builds_forced(tmp_path):
    """the same with every eligible build split over the ranks and all-reduced (tables far below the
    size at which the engine does that by itself)"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    # (SF0.2, not smaller: at SF0.1 the REFERENCE loses matches of Q5 - its ht_get continues a probe behind
    # the last slot of the table without wrapping, qlib/hash.h:438-441, so a chain of hash-equal entries
    # that crosses the end of the table is cut short; an independent pandas evaluation agrees with the
    # GPU result there. DESIGN.md section 4 "reference defects")
    _program_vs_program(tmp_path, 0.2, 2, options="share_min_rows=0")


def _program_vs_program(tmp_path, sf, world, options=None):
    ref = os.path.join(ROOT, "oracle/_ref/resql-oracle")
    gpu = os.path.join(ROOT, "resql_b200/host/resql-b200")
    if not (os.path.exists(ref) and os.path.exists(gpu)):
        pytest.skip("needs the prebuilt reference and drop-in binaries (built where /root/reference exists)")
    from golden.queries import QUERIES as SQL
    data = tpch.generate(sf, seed=20260101)
    stm = []
    for name, schema in tpch.SCHEMAS.items():
        if name not in data:
            continue
        fields = []
        for c, k, a in schema:
            ty = {"int": "int", "date": "date", "bigint": "bigint"}.get(k) or (f"decimal(12,{a})" if k == "dec" else f"{k}({a})")
            fields.append(f"{c} {ty}")
        stm.append(f"create table {name} ( " + ", ".join(fields) + " )")
    create = tmp_path / "create.sql"
    create.write_text(";\n".join(stm) + ";\n")
    loads = [f"exec {create}"]
    for name, cols in data.items():
        p = tmp_path / f"{name}.bin"
        tpch.to_rows(name, cols).tofile(p)
        loads.append(f"binload {name} {p}")
    del data
    outs = {}
    for tag, exe in (("ref", ref), ("gpu", gpu)):
        args = [exe, "--quiet"] + ([f"gpus={world}"] if tag == "gpu" and world > 1 else []) + loads
        if tag == "ref":
            args.append(f"threads={os.cpu_count() or 1}")
        for q in QUERIES:
            args += [f"out {tmp_path / (tag + '_' + q + '.out')}", " ".join(SQL[q].split())]
        env = dict(os.environ)
        if options and tag == "gpu":
            env["RESQL_B200_OPTIONS"] = options
        r = subprocess.run(args, capture_output=True, text=True, timeout=900, env=env)
        assert r.stdout.count("#select") == len(QUERIES), f"{tag}: {r.stdout[-1500:]}{r.stderr[-1500:]}"
        outs[tag] = {q: [l for l in (tmp_path / f"{tag}_{q}.out").read_text().split("\n")[1:] if l] for q in QUERIES}
    for q in QUERIES:
        d = load_plan_dict(q)
        if d.get("limit", -1) >= 0:
            # LIMIT after an unstable sort: ties at the cut may differ (qlib/sort.h:21) - compare the key sequence
            order = d["order"]
            keys = lambda lines: [[l.split("|")[c] for c, _ in order[:1]] for l in lines]   # noqa: E731
            assert keys(outs["gpu"][q]) == keys(outs["ref"][q]), q
            assert len(outs["gpu"][q]) == len(outs["ref"][q]), q
        else:
            assert_same_relation(outs["gpu"][q], outs["ref"][q], d, f"{q} at SF{sf}, resql-b200 (gpus={world}) vs the reference engine")
