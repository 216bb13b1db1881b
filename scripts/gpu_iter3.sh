#!/usr/bin/env bash
# Multi-GPU visit for the partitioned path: parity tests, micro bench at N=1 and N=<gpus>, TPC-H bench.
TAG=${1:-it}
G=${2:-2}
CASES1=${3:-"1e7:4,1e7:1e6,1e8:1e6"}
CASESN=${4:-"1e7:4,1e8:1e6,1e8:1e8"}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > $O/gpu.txt 2>&1; nproc >> $O/gpu.txt; free -g >> $O/gpu.txt
echo "== pytest"; timeout 1800 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; tail -n 15 $O/pytest_gpu.log
cp gpurun_out/sharded_worker_n*.log $O/ 2>/dev/null
echo "== micro N=1"; timeout 900 python bench.py --workload micro --micro-cases $CASES1 --steps 3 --warmup 2 > $O/micro_n1.jsonl 2> $O/micro_n1.err; cut -c1-700 $O/micro_n1.jsonl; tail -n 5 $O/micro_n1.err
echo "== micro N=$G"; timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --workload micro --gpus $G --micro-cases $CASESN --steps 3 --warmup 2 > $O/micro_n$G.jsonl 2> $O/micro_n$G.err; cut -c1-700 $O/micro_n$G.jsonl; tail -n 8 $O/micro_n$G.err
echo "== bench N=$G"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $G --steps 5 --warmup 3 --no-cpu --no-e2e > $O/bench_n$G.json 2> $O/bench_n$G.err; tail -c 1500 $O/bench_n$G.json; tail -n 5 $O/bench_n$G.err
ls -la $O
