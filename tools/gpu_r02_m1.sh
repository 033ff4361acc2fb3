#!/bin/bash
# round-2 measurement pack 1 (one GPU)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02_m1_smi.txt 2>&1
timeout 900 python tools/r02_measure.py > gpurun_out/r02_m1.txt 2>&1; echo "measure rc=$?" >> gpurun_out/r02_m1.txt
timeout 300 python tools/h2d_ceiling.py > gpurun_out/r02_h2d_n1.txt 2>&1; echo "rc=$?" >> gpurun_out/r02_h2d_n1.txt
lscpu | head -30 > gpurun_out/r02_lscpu.txt; nvidia-smi topo -m >> gpurun_out/r02_lscpu.txt 2>&1
tail -50 gpurun_out/r02_m1.txt; cat gpurun_out/r02_h2d_n1.txt
