#!/bin/bash
# TPF decoder: parity tests, then timing (quad kernel for comparison with B200DVB_KERNEL=quad)
mkdir -p gpurun_out
[ -n "$SKIP_TESTS" ] || timeout 600 python -m pytest tests/test_gpu_codec.py -x -q 2>&1 | tail -15 > gpurun_out/tpf_tests.txt
cat gpurun_out/tpf_tests.txt
timeout 300 python tools/tpf_perf.py ${FRAMES:-262144} 2>&1 > gpurun_out/tpf_perf.txt
cat gpurun_out/tpf_perf.txt
