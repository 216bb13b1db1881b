#!/usr/bin/env bash
# 2-GPU visit: forced shared-build program test, bench N=2 (Q3 build sharing / zone skipping), N=1 for reference
O=gpurun_out/${1:-n2}
G=${2:-2}
mkdir -p $O
echo "== pytest (gpus_2 shared builds)"; timeout 600 python -m pytest tests/test_gpu_binaries_sf.py -m gpu -q -x -k "shared_builds" 2>&1 | tail -n 4
echo "== bench N=$G"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $G --steps 10 --warmup 4 --no-cpu --no-e2e > $O/bench_n$G.json 2> $O/bench_n$G.err; python - <<PY
import json
d=json.load(open("$O/bench_n$G.json"))
print("value %.1f G/s" % (d["value"]/1e9), "ms/step", round(d["ms_per_step"],3), d["checks"])
for q,v in d["queries"].items(): print("  ",q,"kernel",round(v["kernel_ms"],3),"scan",round(v["lineitem_scan_kernel_ms"],3),"nccl",round(v["nccl_ms"],3),"syncs",v["host_syncs_per_execution"],"cold",round(v["cold"]["first_execution_wall_ms"],1))
PY
grep -v "^\s*$\|OMP_NUM\|\*\*\*" $O/bench_n$G.err | tail -n 5
echo "== q3 trace N=$G"; RQ_TEST_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29513 scripts/trace_q3_dist.py 2>&1 | grep "launch\|plan:\|q3" | tail -n 14
