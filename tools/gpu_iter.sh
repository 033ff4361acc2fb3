#!/bin/bash
# development loop: TPF parity tests, then timing with the phase breakdown
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then
  timeout 900 python -m pytest tests/test_gpu_tpf.py tests/test_gpu_codec.py -x -q 2>&1 | tail -8 > gpurun_out/iter_tests.txt
  cat gpurun_out/iter_tests.txt
fi
timeout 300 python tools/tpf_perf.py ${FRAMES:-262144} ${SMALL:+1} 2>&1 > gpurun_out/iter_perf.txt
cat gpurun_out/iter_perf.txt
