#!/bin/bash
# ncu source-level capture of the decoder kernel on a small batch; summary CSVs come back
mkdir -p gpurun_out
CMD="python tools/quick_perf.py 65536"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section LaunchStats --section Occupancy --section MemoryWorkloadAnalysis --clock-control none --import-source on -k regex:quad_kernel -s 1 -c 1 -o gpurun_out/prof_quad2 $CMD > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/prof_quad2.ncu-rep --page source --csv > gpurun_out/quad_src2.csv 2>/dev/null
ncu -i gpurun_out/prof_quad2.ncu-rep --page raw --csv > gpurun_out/quad_raw2.csv 2>/dev/null
tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out | tail -5
