#!/usr/bin/env bash
# where does the multi-GPU parity worker stop? (progress lines of every rank on stderr, bounded run)
G=${2:-2}
O=gpurun_out/${1:-shd}
mkdir -p $O
RQ_TEST_PROGRESS=1 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29531 \
    tests/sharded_worker.py > $O/worker.out 2> $O/worker.err
echo "rc=$?"
tail -n 6 $O/worker.out
grep "^\[rank" $O/worker.err | tail -n 12
grep -v "^\[rank" $O/worker.err | grep -v "^$" | tail -n 15
