#!/usr/bin/env bash
O=gpurun_out/${1:-four}
mkdir -p $O
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; tail -n 6 $O/pytest_gpu.log
for v in 1 0; do echo "== q3 one_pass_semi=$v"; RQ_OPT_ONE_PASS_SEMI=$v RQ_PROF_TRACE=1 timeout 300 python scripts/prof_one.py q3 100 4 owned 2>&1 | grep "launch\|plan:\|wall_ms" | tail -n 10; done
echo "== q1"; timeout 300 python scripts/prof_one.py q1 100 4 owned 2>&1 | grep wall_ms | tail -n 2
