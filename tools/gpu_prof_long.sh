#!/bin/bash
# ncu full capture of the quad kernel (global-record geometry) on one wave of N=752 frames
mkdir -p gpurun_out
CMD="python tools/nii_prof_cmd.py ${FRAMES:-4736} double-pass 752 1/2"
timeout 300 $CMD > gpurun_out/long_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:quad_kernel -s 1 -c 1 -o gpurun_out/prof_long $CMD > gpurun_out/ncu_long.log 2>&1
ncu -i gpurun_out/prof_long.ncu-rep --page source --csv > gpurun_out/long_src.csv 2>/dev/null
cat gpurun_out/long_plain.log; tail -3 gpurun_out/ncu_long.log
python tools/ncu_key_metrics.py gpurun_out/prof_long.ncu-rep | tee gpurun_out/long_key.txt
ncu -i gpurun_out/prof_long.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; r=rows[2]
for k,v in zip(h,r):
    if any(t in k for t in ('l1tex__t_sector_hit_rate','l1tex__t_sectors_pipe_lsu_mem_global_op_ld','l1tex__t_requests_pipe_lsu_mem_global_op_ld','lsu_mem_global_op_ld_lookup','l1tex__m_xbar2l1tex','lts__t_sectors_op_read.sum','lts__t_sectors_op_write.sum','smsp__average_warp')): print(k, v)
" | tee -a gpurun_out/long_key.txt
python tools/ncu_src_summary.py gpurun_out/long_src.csv | head -60 | tee gpurun_out/long_src_top.txt
