#!/usr/bin/env bash
TAG=${1:-it}
G=${2:-2}
O=gpurun_out/$TAG
mkdir -p $O
echo "== pytest"; timeout 1800 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; tail -n 25 $O/pytest_gpu.log
cp gpurun_out/sharded_worker_n*.log $O/ 2>/dev/null
echo "== q3 trace"; RQ_PROF_TRACE=1 timeout 300 python scripts/prof_one.py q3 100 4 owned 2>&1 | grep -v "t=\s" | tail -n 16
echo "== bench N=1"; timeout 900 python bench.py --steps 10 --warmup 4 --no-cpu --no-e2e > $O/bench_n1.json 2> $O/bench_n1.err; python - <<PY
import json
d=json.load(open("$O/bench_n1.json"))
print("value %.1f G/s" % (d["value"]/1e9), "ms/step", round(d["ms_per_step"],3), d["checks"], "frac", round(d["roofline"]["frac"],3))
for q,v in d["queries"].items(): print("  ",q,"kernel",round(v["kernel_ms"],3),"scan",round(v["lineitem_scan_kernel_ms"],3),"syncs",v["host_syncs_per_execution"])
PY
tail -n 3 $O/bench_n1.err
echo "== bench N=$G"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $G --steps 10 --warmup 4 --no-cpu --no-e2e > $O/bench_n$G.json 2> $O/bench_n$G.err; python - <<PY
import json
d=json.load(open("$O/bench_n$G.json"))
print("value %.1f G/s" % (d["value"]/1e9), "ms/step", round(d["ms_per_step"],3), d["checks"])
for q,v in d["queries"].items(): print("  ",q,"kernel",round(v["kernel_ms"],3),"scan",round(v["lineitem_scan_kernel_ms"],3),"nccl",round(v["nccl_ms"],3),"syncs",v["host_syncs_per_execution"])
PY
grep -v "^\s*$\|OMP_NUM\|\*\*\*" $O/bench_n$G.err | tail -n 5
echo "== micro N=$G"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --workload micro --gpus $G --micro-cases 1e7:4,1e8:1e6 --steps 3 --warmup 3 > $O/micro_n$G.jsonl 2> $O/micro_n$G.err; python - <<PY
import json
for l in open("$O/micro_n$G.jsonl"):
    x = json.loads(l); print(x["config"]["workload"][22:100], "ms", round(x["ms_per_step"],3), x["checks"], "kernel", round(x["kernel_ms"],3), "nccl", round(x["nccl_ms"],3), "syncs", x["host_syncs"])
PY
grep -v "^\s*$\|OMP_NUM\|\*\*\*" $O/micro_n$G.err | tail -n 5
ls $O
