#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/wf_perf.py > gpurun_out/wf_plain.txt 2>&1 && cat gpurun_out/wf_plain.txt &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:matched_filter_kernel -s 2 -c 1 -o gpurun_out/prof_mf python tools/wf_perf.py > gpurun_out/ncu_mf.log 2>&1
ncu -i gpurun_out/prof_mf.ncu-rep --page raw --csv > gpurun_out/mf_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_mf.ncu-rep --page source --csv > gpurun_out/mf_src.csv 2>/dev/null
tail -2 gpurun_out/ncu_mf.log
