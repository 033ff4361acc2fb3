#!/bin/bash
mkdir -p gpurun_out
for t in ${THREADS_LIST:-448 320 192}; do
echo "== B200DVB_THREADS=$t"
B200DVB_THREADS=$t timeout 600 python tools/quick_perf.py ${1:-262144} 2>&1 | grep -v "^demap\|mc_generate"
done > gpurun_out/quick_perf.txt
cat gpurun_out/quick_perf.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
