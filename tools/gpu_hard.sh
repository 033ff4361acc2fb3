#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/hard.txt
for v in "" "NOREF=1" "OUT=bits" "OUT=packed NOREF=1"; do
  echo "== $v" >> gpurun_out/hard.txt
  env $v timeout 120 python tools/tpf_perf.py 131072 2>&1 | grep -v "^$" >> gpurun_out/hard.txt
done
cat gpurun_out/hard.txt
