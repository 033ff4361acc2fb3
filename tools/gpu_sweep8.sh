#!/bin/bash
# BASELINE configs[2]/[4] shape on N GPUs: coded 16QAM and 8PSK BER/FER sweeps, frames sharded by rank, counters all-reduced
N=${1:-8}
mkdir -p gpurun_out
for MOD in 16QAM 8PSK; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 \
    tools/ber_sweep.py --mod $MOD --ebn0 2 4 6 8 10 --frames 8388608 --batch 262144 --min-frame-errors 1000000000 --modes double-pass nii nii16 \
    > gpurun_out/sweep_${MOD}_n$N.json 2> gpurun_out/sweep_${MOD}_n$N.err
  grep "^#" gpurun_out/sweep_${MOD}_n$N.err
done
