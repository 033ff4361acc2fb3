#!/bin/bash
# ncu full capture of the nii decoder kernel on a 4-wave batch; the report and its CSV pages come back
mkdir -p gpurun_out
CMD="python tools/nii_prof_cmd.py ${FRAMES:-37888}"
timeout 300 $CMD > gpurun_out/nii_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nii_kernel -s 1 -c 1 -o gpurun_out/prof_nii $CMD > gpurun_out/ncu_nii.log 2>&1
ncu -i gpurun_out/prof_nii.ncu-rep --page source --csv > gpurun_out/nii_src.csv 2>/dev/null
ncu -i gpurun_out/prof_nii.ncu-rep --page raw --csv > gpurun_out/nii_raw.csv 2>/dev/null
cat gpurun_out/nii_plain.log; tail -3 gpurun_out/ncu_nii.log; python tools/ncu_src_summary.py gpurun_out/nii_src.csv | head -70
