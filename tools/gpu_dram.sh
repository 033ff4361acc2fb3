#!/bin/bash
# `ncu --set full` captures behind profiles/r02_traffic.json: one launch of each decode kernel, of the six demapper
# launches and of the waveform kernels, on the library in the tree (its sha256 is recorded beside the reports).
mkdir -p gpurun_out
sha256sum modulations_b200/libb200dvb.so > gpurun_out/dram_lib_sha256.txt
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 python tools/nii_prof_cmd.py 37888 double-pass > gpurun_out/dram_plain_tpf.log 2>&1 &&
timeout 900 $NCU -k regex:tpf_kernel -s 1 -c 1 -f -o gpurun_out/prof_dram_tpf python tools/nii_prof_cmd.py 37888 double-pass > gpurun_out/ncu_dram_tpf.log 2>&1
timeout 300 python tools/nii_prof_cmd.py 37888 nii > gpurun_out/dram_plain_nii.log 2>&1 &&
timeout 900 $NCU -k regex:nii_kernel -s 1 -c 1 -f -o gpurun_out/prof_dram_nii python tools/nii_prof_cmd.py 37888 nii > gpurun_out/ncu_dram_nii.log 2>&1
timeout 300 python tools/nii_prof_cmd.py 4736 double-pass 752 1/2 > gpurun_out/dram_plain_quad.log 2>&1 &&
timeout 900 $NCU -k regex:quad_kernel -s 1 -c 1 -f -o gpurun_out/prof_dram_quad752 python tools/nii_prof_cmd.py 4736 double-pass 752 1/2 > gpurun_out/ncu_dram_quad.log 2>&1
timeout 300 python tools/nii_prof_cmd.py 16 double-pass 212 1/3 lat > gpurun_out/dram_plain_lat.log 2>&1 &&
timeout 900 $NCU -k regex:lat_kernel -s 1 -c 1 -f -o gpurun_out/prof_dram_lat python tools/nii_prof_cmd.py 16 double-pass 212 1/3 lat > gpurun_out/ncu_dram_lat.log 2>&1
timeout 300 python tools/wf_perf.py demap > gpurun_out/dram_plain_demap.log 2>&1 &&
timeout 900 $NCU -k regex:demap -s 6 -c 6 -f -o gpurun_out/prof_dram_demap python tools/wf_perf.py demap > gpurun_out/ncu_dram_demap.log 2>&1
timeout 300 python tools/wf_perf.py once > gpurun_out/dram_plain_mf.log 2>&1 &&
timeout 900 $NCU -k regex:"matched_filter|pulse_shape" -s 2 -c 2 -f -o gpurun_out/prof_dram_mf python tools/wf_perf.py once > gpurun_out/ncu_dram_mf.log 2>&1
cat gpurun_out/dram_plain_*.log | tail -30; tail -2 gpurun_out/ncu_dram_*.log
