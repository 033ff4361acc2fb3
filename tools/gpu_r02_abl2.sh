#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r02_nii_bulk.txt
timeout 120 python tools/abl_run.py modulations_b200/libb200dvb.so shipped >> gpurun_out/r02_nii_bulk.txt 2>&1
for f in tools/scratch/abl/lib_*.so; do
  n=$(basename $f .so); timeout 60 python tools/abl_run.py $f ${n#lib_} >> gpurun_out/r02_nii_bulk.txt 2>&1
done
cat gpurun_out/r02_nii_bulk.txt
