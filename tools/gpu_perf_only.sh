#!/bin/bash
# usage: CFGS="3:448 3:192" bash tools/gpu_perf_only.sh   (groups:threads)
mkdir -p gpurun_out
for cfg in ${CFGS:-3:448}; do
g=${cfg%%:*}; t=${cfg##*:}
echo "== GROUPS=$g THREADS=$t"
B200DVB_GROUPS=$g B200DVB_THREADS=$t timeout 600 python tools/quick_perf.py 262144 2>&1 | grep -v "^demap\|mc_generate"
done > gpurun_out/quick_perf.txt
cat gpurun_out/quick_perf.txt
[ -n "$SKIP_TESTS" ] || timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
