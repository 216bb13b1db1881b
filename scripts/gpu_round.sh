#!/usr/bin/env bash
# One GPU-box visit: parity tests, bench (both arms), ncu launch list of the bench command and
# ncu --set full captures of the three lineitem scan kernels. Outputs under gpurun_out/.
# usage (from the repo root, through gpurun):  bash scripts/gpu_round.sh [tag]
TAG=${1:-r01}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem --format=csv > $O/gpu.txt 2>&1
free -g >> $O/gpu.txt; nproc >> $O/gpu.txt
cp MEASURED_PEAKS.json $O/ 2>/dev/null
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -n 3 $O/pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -n 2 $O/smoke.log
echo "== bench ours"; timeout 1200 python bench.py --steps 10 --warmup 3 > $O/bench_ours.json 2> $O/bench_ours.err; tail -c 1500 $O/bench_ours.json; tail -n 5 $O/bench_ours.err
echo "== bench reference"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; tail -c 600 $O/bench_reference.json
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:rq_ -c 400 --csv \
    --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > $O/launches_bench.log 2>&1
wc -l $O/launches.csv
echo "== ncu full"
for q in q1 q6 q3; do
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:rq_scan \
      -o $O/full_$q -f python scripts/prof_one.py $q 10 3 > $O/full_$q.log 2>&1
  tail -n 2 $O/full_$q.log
done
ls -la $O
