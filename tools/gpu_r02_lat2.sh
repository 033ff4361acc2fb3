#!/bin/bash
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_hardening.py tests/test_gpu_tpf.py -x -q -k "latency or tile_boundaries or lat or warmup" > gpurun_out/r02_lat_tests.txt 2>&1; echo "rc=$?" >> gpurun_out/r02_lat_tests.txt
tail -5 gpurun_out/r02_lat_tests.txt
timeout 100 python tools/r02_measure.py latvar 2>&1 | grep -v "warm-up  *[1-9]" 2>&1 | tee gpurun_out/r02_latvar.txt
