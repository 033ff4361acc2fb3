#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r02_nii_ablation.txt
timeout 120 python tools/abl_run.py modulations_b200/libb200dvb.so shipped >> gpurun_out/r02_nii_ablation.txt 2>&1
for f in NOCHAN NOGATHER NOY NOCK NOREC NORAW NOPREP ALL; do
  timeout 120 python tools/abl_run.py tools/scratch/abl/lib_$f.so $f >> gpurun_out/r02_nii_ablation.txt 2>&1
done
cat gpurun_out/r02_nii_ablation.txt
timeout 600 python tools/ber_compare.py 524288 > gpurun_out/r02_ber_three_modes.txt 2>&1; cat gpurun_out/r02_ber_three_modes.txt
timeout 900 python -m pytest tests/test_gpu_nii.py -x -q -k "confidence" 2>&1 | tail -3
