#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/wf_perf.py once > gpurun_out/wf_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:matched_filter -s 1 -c 1 -f -o gpurun_out/prof_mf2 python tools/wf_perf.py once > gpurun_out/ncu_mf2.log 2>&1
ncu -i gpurun_out/prof_mf2.ncu-rep --page source --csv > gpurun_out/mf2_src.csv 2>/dev/null
ncu -i gpurun_out/prof_mf2.ncu-rep --page raw --csv > gpurun_out/mf2_raw.csv 2>/dev/null
python tools/ncu_src_summary.py gpurun_out/mf2_src.csv | head -30
