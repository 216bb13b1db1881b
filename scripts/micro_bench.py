"""BASELINE config 5: the reference's README microbenchmark
    select c, avg(d * a) from foo, bar where a = d group by c order by c
on one GPU, tables generated in HBM (a, d permutations of 1..N, c uniform in [0, G)); every result
is checked against an independent torch int64 evaluation (wrap-around sums, truncating AVG).
    python scripts/micro_bench.py [N G]...        default: a sweep up to N = 1e8"""
import json
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from resql_b200 import Engine, Plan
from common import load_plan_dict

dev = torch.device("cuda:0")
eng = Engine(0)
d = load_plan_dict("micro_join_avg")
args = [int(float(x)) for x in sys.argv[1:]]
cases = list(zip(args[0::2], args[1::2])) or [(10**6, 1000), (10**7, 4), (10**7, 10**6), (10**8, 4), (10**8, 1000), (10**8, 10**6), (10**8, 10**7)]
for n, g in cases:
    gen = torch.Generator(device=dev); gen.manual_seed(n ^ g)
    a = torch.randperm(n, device=dev, generator=gen) + 1
    dd = torch.randperm(n, device=dev, generator=gen) + 1
    c = torch.randint(0, g, (n,), device=dev, generator=gen)
    cols = {"foo": {"a": a, "c": c}, "bar": {"d": dd}}
    tabs = {t["name"]: eng.upload_device(t["name"], {k: (cols[t["name"]][k].data_ptr(), 3, 8) for k in t["columns"]}, n, borrow=True)
            for t in d["tables"]}
    best, first = None, None
    for i in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        res, tm = eng.execute(Plan(d), tabs)
        wall = 1e3 * (time.perf_counter() - t0)
        if i == 0: first = wall
        best = wall if best is None else min(best, wall)
    # torch check: every a matches exactly one d, so d * a = a * a
    s = torch.zeros(g, dtype=torch.int64, device=dev).index_add_(0, c, a * a)
    k = torch.zeros(g, dtype=torch.int64, device=dev).index_add_(0, c, torch.ones_like(a))
    keep = k > 0
    want_c = torch.nonzero(keep).flatten()
    num = s[keep] * 100
    want_avg = torch.div(num.abs(), k[keep], rounding_mode="floor") * torch.sign(num)      # truncation toward zero
    got_c = torch.from_numpy(res.columns[0]).to(dev); got_avg = torch.from_numpy(res.columns[1]).to(dev)
    ok = res.n_rows == want_c.numel() and bool((got_c == want_c).all()) and bool((got_avg == want_avg).all())
    print(json.dumps({"rows": n, "groups": g, "result_rows": res.n_rows, "identical_to_torch_int64": ok,
                      "best_wall_ms": round(best, 3), "first_wall_ms": round(first, 2), "kernel_ms": round(tm.kernel_ms, 3),
                      "probe_tuples_per_s": round(n / (best / 1e3)), "launches": tm.kernel_launches}), flush=True)
    for t in tabs.values(): t.free()
    del a, dd, c, s, k
    torch.cuda.empty_cache()
