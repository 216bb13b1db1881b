#!/usr/bin/env bash
# last visit of the round: bench line (with CPU baseline and e2e) and ncu launch list of the final build
O=gpurun_out/${1:-r21}
mkdir -p $O
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem --format=csv > $O/gpu.txt 2>&1; free -g >> $O/gpu.txt; nproc >> $O/gpu.txt
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -n 2 $O/smoke.log
echo "== bench ours"; timeout 1500 python bench.py --steps 20 --warmup 4 > $O/bench_ours.json 2> $O/bench_ours.err; tail -c 600 $O/bench_ours.json; tail -n 5 $O/bench_ours.err
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:rq_ -c 600 --csv \
    --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > $O/launches_bench.log 2>&1
wc -l $O/launches.csv
