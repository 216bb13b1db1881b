#!/usr/bin/env bash
# 8-GPU visit (charged 8x): sharded/partitioned parity worker, micro bench at 1e9, TPC-H bench at N=8 (and 4).
TAG=${1:-n8}
CASES=${2:-"1e9:1e6"}
EXTRA=${3:-""}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > $O/gpu.txt 2>&1; nproc >> $O/gpu.txt; free -g >> $O/gpu.txt
echo "== sharded worker N=8"; RQ_TEST_TRACE=${RQ_TEST_TRACE:-} timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 \
    tests/sharded_worker.py > $O/sharded_worker_n8.log 2> $O/sharded_worker_n8.err; tail -n 30 $O/sharded_worker_n8.log; grep -v "^\s*$\|OMP_NUM\|\*\*\*" $O/sharded_worker_n8.err | tail -n 15
echo "== bench N=8"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu $EXTRA > $O/bench_n8.json 2> $O/bench_n8.err; tail -c 300 $O/bench_n8.json; tail -n 3 $O/bench_n8.err
echo "== micro N=8"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --workload micro --gpus 8 --micro-cases $CASES --steps 3 --warmup 2 > $O/micro_n8.jsonl 2> $O/micro_n8.err; cut -c1-400 $O/micro_n8.jsonl; grep -v "^\s*$\|OMP_NUM\|\*\*\*" $O/micro_n8.err | tail -n 8
ls -la $O
