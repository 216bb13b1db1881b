#!/usr/bin/env bash
# upload timing + targeted tests + bench
O=gpurun_out/${1:-two}
mkdir -p $O
echo "== upload bench"; timeout 600 python scripts/upload_bench.py 100 2>&1 | grep -v "^$" | tail -n 12 | tee $O/upload_bench.txt
echo "== pytest (narrowing)"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "transfer_narrowing or sf1" 2>&1 | tail -n 5
echo "== bench N=1"; timeout 900 python bench.py --steps 10 --warmup 4 --no-cpu > $O/bench_n1.json 2> $O/bench_n1.err; python - <<PY
import json
d=json.load(open("$O/bench_n1.json"))
print("value %.1f G/s" % (d["value"]/1e9), "ms/step", round(d["ms_per_step"],3), d["checks"], "frac", round(d["roofline"]["frac"],3))
for q,v in d["queries"].items(): print("  ",q,"kernel",round(v["kernel_ms"],3),"scan",round(v["lineitem_scan_kernel_ms"],3),"syncs",v["host_syncs_per_execution"], "cold", round(v["cold"]["first_execution_wall_ms"],1))
print("e2e", d.get("e2e"))
PY
tail -n 3 $O/bench_n1.err
