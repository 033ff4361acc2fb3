#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_nii.py -x -q > gpurun_out/r02_nii16_tests.txt 2>&1; echo "rc=$?" >> gpurun_out/r02_nii16_tests.txt
tail -25 gpurun_out/r02_nii16_tests.txt
timeout 300 python tools/nii_perf.py 262144 > gpurun_out/r02_nii16_perf.txt 2>&1; cat gpurun_out/r02_nii16_perf.txt
