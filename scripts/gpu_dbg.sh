#!/usr/bin/env bash
O=gpurun_out/dbg
mkdir -p $O
echo "== plain"; timeout 600 python bench.py --workload micro --micro-cases 1e7:1e6 --steps 3 --warmup 2 > $O/micro.jsonl 2> $O/micro.err; cut -c1-300 $O/micro.jsonl; tail -n 3 $O/micro.err
echo "== no replay"; RQ_NO_REPLAY=1 timeout 600 python scripts/micro_dbg.py 10000000 1000000 0 2>&1 | tail -n 8
echo "== replay"; timeout 600 python scripts/micro_dbg.py 10000000 1000000 1 2>&1 | tail -n 8
echo "== sanitizer"; timeout 1500 compute-sanitizer --tool memcheck --print-limit 5 python scripts/micro_dbg.py 3000000 1000000 1 > $O/memcheck.log 2>&1; grep -v "^=========     at\|^=========     by" $O/memcheck.log | head -60
echo "== reftests-gpu"; timeout 600 resql_b200/host/resql-reftests-gpu > $O/reftests_gpu.log 2>&1; tail -n 5 $O/reftests_gpu.log
