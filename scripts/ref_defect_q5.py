"""Evidence for a REFERENCE defect: Q5 at SF0.1 through resql-b200 (1 and 2 GPUs, option variants)
against the reference engine, whose ht_get continues a probe behind the last slot without wrapping (qlib/hash.h:438-441)
and so loses matches of hash-equal chains that cross the end of its table; pandas agrees with the GPU result."""
import os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from resql_b200 import tpch
from golden.queries import QUERIES as SQL
sf = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
qs = sys.argv[2].split(",") if len(sys.argv) > 2 else ["q5"]
tmp = tempfile.mkdtemp()
data = tpch.generate(sf, seed=20260101)
stm = []
for name, schema in tpch.SCHEMAS.items():
    if name not in data: continue
    fields = []
    for c, k, a in schema:
        ty = {"int": "int", "date": "date", "bigint": "bigint"}.get(k) or (f"decimal(12,{a})" if k == "dec" else f"{k}({a})")
        fields.append(f"{c} {ty}")
    stm.append(f"create table {name} ( " + ", ".join(fields) + " )")
create = os.path.join(tmp, "create.sql"); open(create, "w").write(";\n".join(stm) + ";\n")
loads = [f"exec {create}"]
for name, cols in data.items():
    p = os.path.join(tmp, f"{name}.bin"); tpch.to_rows(name, cols).tofile(p); loads.append(f"binload {name} {p}")
def run(tag, exe, pre, env_opts):
    args = [exe, "--quiet"] + pre + loads
    for q in qs: args += [f"out {tmp}/{tag}_{q}.out", " ".join(SQL[q].split())]
    env = dict(os.environ)
    if env_opts: env["RESQL_B200_OPTIONS"] = env_opts
    r = subprocess.run(args, capture_output=True, text=True, timeout=600, env=env)
    return {q: open(f"{tmp}/{tag}_{q}.out").read().split("\n")[1:] for q in qs}, r.stderr[-1500:]
ref, _ = run("ref", os.path.join(ROOT, "oracle/_ref/resql-oracle"), [], None)
for tag, pre, opts in [("g1", [], None), ("g2", ["gpus=2"], None), ("g2s", ["gpus=2"], "share_min_rows=0"), ("g2sz", ["gpus=2"], "share_min_rows=0,zone_skip=0"),
                       ("g2z", ["gpus=2"], "zone_skip=0"), ("g2st", ["gpus=2"], "share_min_rows=0,trace=1,replay=0")]:
    got, err = run(tag, os.path.join(ROOT, "resql_b200/host/resql-b200"), pre, opts)
    for q in qs:
        print(tag, q, "OK" if sorted(got[q]) == sorted(ref[q]) else f"DIFF got {got[q][:6]} want {ref[q][:6]}", flush=True)
    if "t" in tag[3:]: print(err)
