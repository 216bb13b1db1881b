#!/usr/bin/env bash
# Multi-GPU iteration visit: parity tests (incl. the sharded worker at the box's GPU count), short
# bench at N=1 and N=<gpus>. usage: bash scripts/gpu_iter2.sh <tag> <gpus> [sf]
TAG=${1:-it}
G=${2:-2}
SF=${3:-100}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > $O/gpu.txt 2>&1; nproc >> $O/gpu.txt; free -g >> $O/gpu.txt
echo "== pytest"; timeout 1800 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; tail -n 15 $O/pytest_gpu.log
echo "== bench N=1"; timeout 900 python bench.py --steps 5 --warmup 3 --sf $SF --no-cpu > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 2500 $O/bench_n1.json; tail -n 5 $O/bench_n1.err
for n in $G; do
echo "== bench N=$n"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $n --steps 5 --warmup 3 --sf $SF --no-cpu > $O/bench_n$n.json 2> $O/bench_n$n.err; tail -c 2500 $O/bench_n$n.json; tail -n 5 $O/bench_n$n.err
done
ls -la $O
