#!/usr/bin/env bash
O=gpurun_out/${1:-n2b}
G=${2:-2}
mkdir -p $O
echo "== pytest multi-GPU"; timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_binaries_sf.py -m gpu -q -x -k "sharded or gpus_" 2>&1 | tail -n 6
cp gpurun_out/sharded_worker_n$G.log $O/ 2>/dev/null; grep "broadcast\|shared-build" $O/sharded_worker_n$G.log
echo "== bench N=$G (with e2e)"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $G --steps 10 --warmup 4 --no-cpu > $O/bench_n$G.json 2> $O/bench_n$G.err; python - <<PY
import json
d=json.load(open("$O/bench_n$G.json"))
print("value %.1f G/s" % (d["value"]/1e9), "ms/step", round(d["ms_per_step"],3), d["checks"])
for q,v in d["queries"].items(): print("  ",q,"kernel",round(v["kernel_ms"],3),"scan",round(v["lineitem_scan_kernel_ms"],3),"nccl",round(v["nccl_ms"],3),"syncs",v["host_syncs_per_execution"],"cold",round(v["cold"]["first_execution_wall_ms"],1))
print("e2e", {k: v for k, v in d.get("e2e", {}).items() if k != "path"})
PY
grep -v "^\s*$\|OMP_NUM\|\*\*\*" $O/bench_n$G.err | tail -n 5
