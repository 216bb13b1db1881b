"""Runs one TPC-H plan fixture a few times on synthetic data generated in HBM (for ncu captures
and quick timings):  python scripts/prof_one.py q1 10 3"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from resql_b200 import Engine, Plan
from resql_b200 import tpch_device as TD
from common import load_plan_dict

q = sys.argv[1]
sf = float(sys.argv[2]) if len(sys.argv) > 2 else 10
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
mode = sys.argv[4] if len(sys.argv) > 4 else "owned"     # owned: engine storage (tile-major); borrow: plain torch columns
dev = torch.device("cuda:0")
eng = Engine(0)
for k in ("stages", "warps", "narrow"):
    if os.environ.get("RQ_OPT_" + k.upper()):
        eng.set_option(k, float(os.environ["RQ_OPT_" + k.upper()]))
need_orders = q not in ("q1", "q6")
orders, li, cust = TD.gen_orders_lineitem(sf, 42, dev, want_orders=need_orders)
src = {"lineitem": li, "orders": orders, "customer": cust}
d = load_plan_dict(q)
tabs = {}
for t in d["tables"]:
    n = src[t["name"]][t["columns"][0]].shape[0]
    cols = t["columns"] if mode == "borrow" else list(src[t["name"]].keys())     # owned: the whole table, schema order
    tabs[t["name"]] = eng.upload_device(t["name"], TD.as_device_columns(src[t["name"]], cols), n, borrow=(mode == "borrow"))
n = li["l_quantity"].numel()
plan = Plan(d)
bpt = {"q1": 38, "q6": 28, "q3": 24}.get(q, 0)
for i in range(reps):
    if i == reps - 1:
        torch.cuda.cudart().cudaProfilerStart()     # ncu --profile-from-start off: only the last (warm) run
    if os.environ.get("RQ_PROF_TRACE") and i == reps - 1:
        eng.set_option("graphs", 0); eng.set_option("replay", 0); eng.set_option("trace", 1)   # per-launch times on stderr
    torch.cuda.synchronize(); _t0 = time.perf_counter()
    res, tm = eng.execute(plan, tabs)
    _wall = 1e3 * (time.perf_counter() - _t0)
    print(q, "wall_ms", round(_wall, 3), "syncs", tm.host_syncs, "rows", n, "scan_ms", round(tm.scan_kernel_ms, 3), "Gtuples/s", round(n / tm.scan_kernel_ms / 1e6, 2),
          "GB/s", round(n * bpt / tm.scan_kernel_ms / 1e6, 1), "launches", tm.kernel_launches, "kernel_ms", round(tm.kernel_ms, 3))
print([c.tolist()[:4] for c in res.columns][:4])
