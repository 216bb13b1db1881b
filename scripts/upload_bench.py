"""Times rq_table_upload from pinned host columns (the e2e load path): SF100 lineitem columns of
Q1+Q6+Q3, with and without PCIe narrowing.  python scripts/upload_bench.py [sf]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from resql_b200 import Engine
from resql_b200 import tpch_device as TD

sf = float(sys.argv[1]) if len(sys.argv) > 1 else 100
dev = torch.device("cuda:0")
eng = Engine(0)
orders, li, cust = TD.gen_orders_lineitem(sf, 42, dev, want_orders=False)
cols = ["l_orderkey", "l_quantity", "l_extendedprice", "l_discount", "l_tax", "l_shipdate", "l_returnflag", "l_linestatus"]
host = {}
for c in cols:
    hp = torch.empty(li[c].shape, dtype=li[c].dtype, pin_memory=True)
    hp.copy_(li[c]); host[c] = hp.numpy()
del li; torch.cuda.empty_cache(); torch.cuda.synchronize()
nbytes = sum(a.nbytes for a in host.values())
eng.set_option("trace", 1)
t = eng.upload("lineitem", host); t.free()       # first use: staging buffers, threads
for opts in ({}, {"up_chunk_krows": 1024}, {"up_chunk_krows": 256}, {"up_threads": 12}, {"up_threads": 24}, {"narrow": 0}):
    for k, v in {"narrow": 1, "up_chunk_krows": 512, "up_threads": 0, **opts}.items():
        eng.set_option(k, v)
    for _ in range(2):
        t0 = time.perf_counter()
        t = eng.upload("lineitem", host)
        ms = 1e3 * (time.perf_counter() - t0)
        print(f"{opts} upload {nbytes/1e9:.2f} GB in {ms:.1f} ms = {nbytes/1e6/ms:.1f} GB/s of table bytes", flush=True)
        t.free()
eng.shutdown()
