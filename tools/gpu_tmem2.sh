#!/bin/bash
mkdir -p gpurun_out
for y in 1 0; do
echo "== YTMEM=$y"
B200DVB_YTMEM=$y timeout 120 python tools/quick_perf.py 262144 2>&1 | grep -v "^demap\|mc_generate\|N=48" | head -2
done
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
