#!/usr/bin/env bash
# stage-depth / physical-width sweep of the scan kernel on Q1, Q6, Q3 at SF100
O=gpurun_out/${1:-sweep}
mkdir -p $O
for q in q6 q1 q3; do
  for nar in 1 0; do
    for st in 0 1 2 4 8; do
      echo "== $q narrow=$nar stages=$st"
      RQ_OPT_NARROW=$nar RQ_OPT_STAGES=$st timeout 200 python scripts/prof_one.py $q 100 3 owned 2>&1 | grep wall_ms | tail -n 1
    done
  done
done 2>&1 | tee $O/sweep.txt
