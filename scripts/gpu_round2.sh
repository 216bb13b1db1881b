#!/usr/bin/env bash
# Final visit of a round: smoke, bench (ours, with the SF10 CPU baseline and e2e), ncu launch list of the bench
# command and ncu --set full captures of the three lineitem scan kernels. Outputs under gpurun_out/<tag>.
TAG=${1:-r20}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem --format=csv > $O/gpu.txt 2>&1
free -g >> $O/gpu.txt; nproc >> $O/gpu.txt
cp MEASURED_PEAKS.json $O/ 2>/dev/null
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -n 2 $O/smoke.log
echo "== bench ours"; timeout 1500 python bench.py --steps 20 --warmup 4 > $O/bench_ours.json 2> $O/bench_ours.err; tail -c 1200 $O/bench_ours.json; tail -n 5 $O/bench_ours.err
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:rq_ -c 600 --csv \
    --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > $O/launches_bench.log 2>&1
wc -l $O/launches.csv
echo "== ncu full"
for q in q1 q6 q3; do
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:rq_scan \
      -o $O/full_$q -f python scripts/prof_one.py $q 10 3 > $O/full_$q.log 2>&1
  tail -n 2 $O/full_$q.log
done
ls -la $O
