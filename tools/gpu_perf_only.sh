#!/bin/bash
mkdir -p gpurun_out
for cfg in ${CFG_LIST:-"3 448"}; do
set -- $cfg
echo "== GROUPS=$1 THREADS=$2"
B200DVB_GROUPS=$1 B200DVB_THREADS=$2 timeout 600 python tools/quick_perf.py 262144 2>&1 | grep -v "^demap\|mc_generate"
done > gpurun_out/quick_perf.txt
cat gpurun_out/quick_perf.txt
[ -n "$SKIP_TESTS" ] || timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
