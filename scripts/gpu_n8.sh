#!/usr/bin/env bash
# 8-GPU visit (expensive: 8x box time): multi-GPU parity worker + bench with e2e
O=gpurun_out/${1:-n8}
G=${2:-8}
mkdir -p $O
nvidia-smi --query-gpu=name --format=csv,noheader | sort | uniq -c > $O/gpu.txt; nproc >> $O/gpu.txt
echo "== sharded parity N=$G"; timeout 400 python -m pytest tests/test_gpu_sharded.py -m gpu -q -x -k "$G" 2>&1 | tail -n 4
cp gpurun_out/sharded_worker_n$G.log $O/ 2>/dev/null; grep -c "identical" $O/sharded_worker_n$G.log
echo "== bench N=$G (with e2e)"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $G --steps 20 --warmup 4 --no-cpu > $O/bench_n$G.json 2> $O/bench_n$G.err; python - <<PY
import json
d=json.load(open("$O/bench_n$G.json"))
print("value %.1f G/s" % (d["value"]/1e9), "ms/step", round(d["ms_per_step"],3), d["checks"])
for q,v in d["queries"].items(): print("  ",q,"kernel",round(v["kernel_ms"],3),"scan",round(v["lineitem_scan_kernel_ms"],3),"nccl",round(v["nccl_ms"],3),"syncs",v["host_syncs_per_execution"],"cold",round(v["cold"]["first_execution_wall_ms"],1))
e=d.get("e2e",{}); print("e2e", e.get("value"), e.get("ms_per_step"), e.get("phases_ms_last_step_rank0"))
PY
grep -v "^\s*$\|OMP_NUM\|\*\*\*" $O/bench_n$G.err | tail -n 5
