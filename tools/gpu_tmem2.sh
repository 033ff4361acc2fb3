#!/bin/bash
mkdir -p gpurun_out
for tm in 1 0; do
echo "== TMEM=$tm"
B200DVB_TMEM=$tm timeout 120 python tools/quick_perf.py 262144 2>&1 | grep -v "^demap\|mc_generate"
done
B200DVB_TMEM=1 timeout 600 python -m pytest tests/test_gpu_codec.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
