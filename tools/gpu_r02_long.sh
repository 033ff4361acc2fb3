#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_codec.py tests/test_gpu_tpf.py tests/test_gpu_hardening.py -x -q -k "424 or 752 or 848 or 220 or large_batch or properties or randomised" > gpurun_out/r02_long_tests.txt 2>&1; echo "rc=$?" >> gpurun_out/r02_long_tests.txt
tail -5 gpurun_out/r02_long_tests.txt
timeout 200 python tools/r02_measure.py long > gpurun_out/r02_long.txt 2>&1; cat gpurun_out/r02_long.txt
