#!/usr/bin/env bash
# One short GPU-box visit while iterating on a kernel: parity tests, quick timings, one
# source-level ncu capture. usage: bash scripts/gpu_iter.sh <tag> [query-to-profile] [sf-list]
TAG=${1:-it}
Q=${2:-q1}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > $O/gpu.txt 2>&1; nproc >> $O/gpu.txt; free -g >> $O/gpu.txt
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; tail -n 15 $O/pytest_gpu.log
for q in q1 q6 q3; do
  echo "== prof_one $q owned"; timeout 300 python scripts/prof_one.py $q 100 4 owned 2>&1 | tail -n 3 | tee -a $O/prof_one.log
done
if [ "$Q" != "none" ]; then
echo "== ncu source $Q"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:rq_scan \
    -o $O/full_$Q -f python scripts/prof_one.py $Q 10 3 > $O/full_$Q.log 2>&1
tail -n 2 $O/full_$Q.log
fi
ls -la $O
