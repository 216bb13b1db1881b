#!/usr/bin/env bash
# stage/warp sweep of the scan kernel (tuning aid):  bash scripts/sweep.sh q1 10
q=$1; sf=${2:-10}
for cfg in "2 99" "1 99" "1 12" "1 10" "2 8" "2 6" "3 99"; do
  set -- $cfg
  echo "== $q stages=$1 warps<=$2"
  RQ_OPT_STAGES=$1 RQ_OPT_WARPS=$2 timeout 200 python scripts/prof_one.py $q $sf 3 | tail -n 2 | head -n 1
done
