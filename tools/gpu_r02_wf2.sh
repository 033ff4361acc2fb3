#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_waveform.py -x -q > gpurun_out/r02_wf_tests.txt 2>&1; echo "rc=$?" >> gpurun_out/r02_wf_tests.txt
tail -4 gpurun_out/r02_wf_tests.txt
timeout 200 python tools/wf_perf.py 0 1 2 > gpurun_out/r02_wf_perf.txt 2>&1; cat gpurun_out/r02_wf_perf.txt
