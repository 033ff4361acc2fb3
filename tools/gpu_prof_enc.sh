#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/enc_prof_cmd.py > gpurun_out/enc_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:encode_kernel -s 1 -c 1 -f -o gpurun_out/prof_enc python tools/enc_prof_cmd.py > gpurun_out/ncu_enc.log 2>&1
ncu -i gpurun_out/prof_enc.ncu-rep --page source --csv > gpurun_out/enc_src.csv 2>/dev/null
ncu -i gpurun_out/prof_enc.ncu-rep --page raw --csv > gpurun_out/enc_raw.csv 2>/dev/null
cat gpurun_out/enc_plain.log; python tools/ncu_src_summary.py gpurun_out/enc_src.csv | head -40
