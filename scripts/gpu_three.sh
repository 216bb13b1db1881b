#!/usr/bin/env bash
O=gpurun_out/${1:-three}
mkdir -p $O
echo "== pytest (new)"; timeout 900 python -m pytest tests/test_gpu_loader.py tests/test_dbgen.py tests/test_gpu_binaries_sf.py -m gpu -q -x 2>&1 | tail -n 8
for q in q1 q6; do echo "== prof_one $q"; timeout 300 python scripts/prof_one.py $q 100 4 owned 2>&1 | grep wall_ms | tail -n 2; done
echo "== bench N=1"; timeout 900 python bench.py --steps 10 --warmup 4 --no-cpu > $O/bench_n1.json 2> $O/bench_n1.err; python - <<PY
import json
d=json.load(open("$O/bench_n1.json"))
print("value %.1f G/s" % (d["value"]/1e9), "ms/step", round(d["ms_per_step"],3), d["checks"], "frac", round(d["roofline"]["frac"],3))
for q,v in d["queries"].items(): print("  ",q,"kernel",round(v["kernel_ms"],3),"scan",round(v["lineitem_scan_kernel_ms"],3),"syncs",v["host_syncs_per_execution"], "cold", round(v["cold"]["first_execution_wall_ms"],1))
e=d.get("e2e",{}); print("e2e", e.get("ms_per_step"), e.get("phases_ms_last_step_rank0"))
PY
tail -n 3 $O/bench_n1.err
