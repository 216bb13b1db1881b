#!/usr/bin/env bash
# One-GPU visit: parity tests, per-query timings, bench with e2e (no CPU baseline).
TAG=${1:-one}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > $O/gpu.txt 2>&1; nproc >> $O/gpu.txt; free -g >> $O/gpu.txt
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; tail -n 25 $O/pytest_gpu.log
for q in q1 q6 q3; do echo "== prof_one $q"; timeout 300 python scripts/prof_one.py $q 100 4 owned 2>&1 | tail -n 3; done
echo "== q3 trace"; RQ_PROF_TRACE=1 timeout 300 python scripts/prof_one.py q3 100 4 owned 2>&1 | grep "launch\|plan:" | tail -n 12
echo "== bench N=1"; timeout 900 python bench.py --steps 10 --warmup 4 --no-cpu > $O/bench_n1.json 2> $O/bench_n1.err; python - <<PY
import json
d=json.load(open("$O/bench_n1.json"))
print("value %.1f G/s" % (d["value"]/1e9), "ms/step", round(d["ms_per_step"],3), d["checks"], "frac", round(d["roofline"]["frac"],3))
for q,v in d["queries"].items(): print("  ",q,"kernel",round(v["kernel_ms"],3),"scan",round(v["lineitem_scan_kernel_ms"],3),"syncs",v["host_syncs_per_execution"], "cold", round(v["cold"]["first_execution_wall_ms"],1))
print("e2e", d.get("e2e"))
PY
tail -n 3 $O/bench_n1.err
ls $O
