#!/bin/bash
# experiment: 2 CTAs per SM (15 frames each) vs 1 CTA of 32 frames; pipe-mix microbench
mkdir -p gpurun_out
./tools/scratch/mb2 > gpurun_out/mb2.txt 2>&1
cat gpurun_out/mb2.txt
for cfg in 4:448:32 2:256:15 2:224:15 2:192:15 2:128:15; do
IFS=: read g t f <<< "$cfg"
echo "== GROUPS=$g THREADS=$t FRAMES=$f"
B200DVB_GROUPS=$g B200DVB_THREADS=$t B200DVB_FRAMES=$f timeout 300 python tools/quick_perf.py 262144 2>&1 | grep -v "^demap\|mc_generate" | head -3
done > gpurun_out/exp1.txt 2>&1
cat gpurun_out/exp1.txt
