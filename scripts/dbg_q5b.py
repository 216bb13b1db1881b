"""debug: q5 at SF0.1 through the C ABI with engine features switched off one at a time, vs the plan oracle"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from resql_b200 import Engine, Plan, tpch
from oracle.plan_oracle import run_plan
from common import load_plan_dict, plan_tables, serialize_columns
sf = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 20260101
data = tpch.generate(sf, seed=seed)
eng = Engine(0)
for q in ("q5",):
    d = load_plan_dict(q)
    tabs = plan_tables(d, data)
    want = sorted(serialize_columns(*run_plan(d, tabs)))
    for opts in ({}, {"direct_joins": 0}, {"prune_builds": 0}, {"zone_skip": 0}, {"replay": 0}, {"split_min_rows": 1e18}, {"split_min_rows": 0, "split_frac": 1e18},
                 {"direct_joins": 0, "prune_builds": 0, "split_min_rows": 1e18}):
        for k, v in {"direct_joins": 1, "prune_builds": 1, "zone_skip": 1, "replay": 1, "split_min_rows": -1, "split_frac": -1, **opts}.items():
            eng.set_option(k, v)
        handles = {n: eng.upload(n, c) for n, c in tabs.items()}
        res, tm = eng.execute(Plan(d), handles)
        got = sorted(serialize_columns(res.columns, res.sql_types, res.sql_widths))
        print(q, opts, "OK" if got == want else f"DIFF {[g for g in got if g not in want]} want {[w for w in want if w not in got]}", flush=True)
        for h in handles.values(): h.free()
