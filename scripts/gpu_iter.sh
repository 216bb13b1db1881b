#!/usr/bin/env bash
# One short GPU-box visit while iterating on a kernel: parity tests, quick timings, one
# source-level ncu capture. usage: bash scripts/gpu_iter.sh <tag> [query-to-profile|none] [micro cases]
TAG=${1:-it}
Q=${2:-q1}
MC=${3:-"1e7:4,1e7:1e6,1e8:1e6"}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > $O/gpu.txt 2>&1; nproc >> $O/gpu.txt; free -g >> $O/gpu.txt
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; tail -n 15 $O/pytest_gpu.log
for q in q1 q6 q3; do
  echo "== prof_one $q owned"; timeout 300 python scripts/prof_one.py $q 100 5 owned 2>&1 | tail -n 4 | tee -a $O/prof_one.log
done
echo "== micro N=1"; timeout 900 python bench.py --workload micro --micro-cases $MC --steps 3 --warmup 2 > $O/micro_n1.jsonl 2> $O/micro_n1.err; cut -c1-120 $O/micro_n1.jsonl; python - <<PY
import json
for l in open("$O/micro_n1.jsonl"):
    x = json.loads(l); print(x["config"]["workload"][22:100], "ms", round(x["ms_per_step"],3), x["checks"], "kernel", round(x["kernel_ms"],3), "syncs", x["host_syncs"])
PY
tail -n 3 $O/micro_n1.err
if [ "$Q" != "none" ]; then
echo "== ncu source $Q"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:rq_scan \
    -o $O/full_$Q -f python scripts/prof_one.py $Q 10 3 > $O/full_$Q.log 2>&1
tail -n 2 $O/full_$Q.log
fi
ls -la $O
