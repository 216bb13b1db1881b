"""debug helper: the microbenchmark plan on one GPU, N rows / G groups, with or without replay"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from resql_b200 import Engine, Plan
from common import load_plan_dict
n, g, replay = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
dev = torch.device("cuda:0")
eng = Engine(0)
eng.set_option("replay", replay)
d = load_plan_dict("micro_join_avg")
idx = torch.arange(0, n, device=dev, dtype=torch.int64)
a = (idx * 2654435761) % n + 1
c = ((idx * 2654435769) & 0xFFFFFFFF) % g
dd = (idx * 1000003 + 12345) % n + 1
cols = {"foo": {"a": a, "c": c}, "bar": {"d": dd}}
tabs = {t["name"]: eng.upload_device(t["name"], {k: (cols[t["name"]][k].data_ptr(), 3, 8) for k in t["columns"]}, n, borrow=True) for t in d["tables"]}
for i in range(4):
    res, tm = eng.execute(Plan(d), tabs)
    print("run", i, "rows", res.n_rows, "syncs", tm.host_syncs, "launches", tm.kernel_launches, flush=True)
