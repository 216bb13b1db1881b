"""Runs random plans (tests/plan_fuzz.py) on the GPU against the oracle and prints what differs.
    python scripts/fuzz_debug.py 3 14 40"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from plan_fuzz import random_plan, COLS
from oracle.plan_oracle import run_plan
from common import plan_tables, serialize_columns
from resql_b200 import Engine, Plan, tpch, EngineError
data = tpch.generate(0.01, seed=42)
eng = Engine(0)
up = {t: eng.upload(t, {c: data[t][c] for c, _ in cols}) for t, cols in COLS.items()}
for seed in [int(x) for x in sys.argv[1:]]:
    d = random_plan(seed)
    want = serialize_columns(*run_plan(d, plan_tables(d, data)))
    try:
        res, tm = eng.execute(Plan(d), {t["name"]: up[t["name"]] for t in d["tables"]})
    except EngineError as e:
        print(f"seed {seed}: ERROR {e}")
        print(json.dumps(d)[:3000])
        continue
    got = serialize_columns(res.columns, res.sql_types, res.sql_widths)
    sg, sw = sorted(got), sorted(want)
    if sg == sw:
        print(f"seed {seed}: same multiset ({len(got)} rows); order keys {'same' if got == want or not d['order'] else 'DIFFER?'}")
        continue
    print(f"seed {seed}: MISMATCH got {len(got)} rows, want {len(want)}; launches {tm.kernel_launches}")
    only_g = [x for x in sg if x not in set(sw)][:5]; only_w = [x for x in sw if x not in set(sg)][:5]
    print("  only in got :", only_g); print("  only in want:", only_w)
    for p in d["pipelines"]:
        print("  pipeline src", p["source_kind"], p["source_id"], "sink", p["sink_kind"], "keys", p["keys"], "vals", p["vals"], "args", p["args"])
        for i, n in enumerate(p["nodes"]): print("     ", i, n)
