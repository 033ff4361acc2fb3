#!/bin/bash
# development run on the B200 box: parity tests, then quick timings
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.txt
cat gpurun_out/pytest_gpu.txt
timeout 600 python tools/quick_perf.py > gpurun_out/quick_perf.txt 2>&1
cat gpurun_out/quick_perf.txt
