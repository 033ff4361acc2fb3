#!/bin/bash
# round 2: first light of the nii kernel + the hardened host side
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_nii.py -x -q > gpurun_out/r02_nii_tests.txt 2>&1; echo "rc=$?" >> gpurun_out/r02_nii_tests.txt
tail -15 gpurun_out/r02_nii_tests.txt
timeout 300 python tools/nii_perf.py 262144 > gpurun_out/r02_nii_perf.txt 2>&1; echo "rc=$?" >> gpurun_out/r02_nii_perf.txt
cat gpurun_out/r02_nii_perf.txt
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_nii.py > gpurun_out/r02_gpu_tests.txt 2>&1; echo "rc=$?" >> gpurun_out/r02_gpu_tests.txt
tail -15 gpurun_out/r02_gpu_tests.txt
