#!/bin/bash
# timing-only ablation builds of the TPF kernel (tools/scratch/abl/*.so), phase breakdown each
mkdir -p gpurun_out
: > gpurun_out/abl.txt
for so in "" tools/scratch/abl/*.so; do
  echo "== ${so:-baseline}" >> gpurun_out/abl.txt
  B200DVB_LIB=${so:+$PWD/$so} timeout 120 python tools/tpf_perf.py 131072 2>&1 | grep -v "^$" >> gpurun_out/abl.txt
done
cat gpurun_out/abl.txt
