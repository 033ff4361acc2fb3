#!/bin/bash
# ncu source-level capture of the TPF decoder kernel on a small batch; summary CSVs come back
mkdir -p gpurun_out
CMD="python tools/tpf_perf.py ${FRAMES:-37888}"
timeout 300 $CMD > gpurun_out/tpf_plain.log 2>&1 &&
timeout 900 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section LaunchStats --section Occupancy --section MemoryWorkloadAnalysis --section SpeedOfLight --clock-control none --import-source on -k regex:tpf_kernel -s 1 -c 1 -o gpurun_out/prof_tpf $CMD > gpurun_out/ncu_tpf.log 2>&1
ncu -i gpurun_out/prof_tpf.ncu-rep --page source --csv > gpurun_out/tpf_src.csv 2>/dev/null
ncu -i gpurun_out/prof_tpf.ncu-rep --page raw --csv > gpurun_out/tpf_raw.csv 2>/dev/null
cat gpurun_out/tpf_plain.log; tail -3 gpurun_out/ncu_tpf.log; python tools/ncu_src_summary.py gpurun_out/tpf_src.csv | head -70
