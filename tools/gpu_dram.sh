#!/bin/bash
# parity + timing + DRAM bytes of the decode kernel (light ncu pass)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_codec.py tests/test_gpu_tpf.py -x -q 2>&1 | tail -4
timeout 300 python tools/tpf_perf.py 262144 2>&1 | tail -3
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:tpf_kernel -s 1 -c 1 python tools/tpf_perf.py 65536 2>&1 | grep -E "dram__|lts__|gpu__time" 
