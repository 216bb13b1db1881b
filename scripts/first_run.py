"""First-execution latency (cold capacity memo) vs steady state of one plan fixture on HBM-resident
synthetic data:  python scripts/first_run.py q3 100"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from resql_b200 import Engine, Plan
from resql_b200 import tpch_device as TD
from common import load_plan_dict
q = sys.argv[1]; sf = float(sys.argv[2])
dev = torch.device("cuda:0")
eng = Engine(0)
orders, li, cust = TD.gen_orders_lineitem(sf, 42, dev, want_orders=q not in ("q1", "q6"))
src = {"lineitem": li, "orders": orders, "customer": cust}
d = load_plan_dict(q)
tabs = {t["name"]: eng.upload_device(t["name"], TD.as_device_columns(src[t["name"]], t["columns"]),
                                     src[t["name"]][t["columns"][0]].shape[0], borrow=True) for t in d["tables"]}
torch.cuda.synchronize()
for i in range(4):
    t0 = time.perf_counter()
    res, tm = eng.execute(Plan(d), tabs)
    print(f"{q} run {i}: wall {1e3*(time.perf_counter()-t0):8.2f} ms, kernels {tm.kernel_ms:8.2f} ms, launches {tm.kernel_launches}")
