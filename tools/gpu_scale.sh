#!/bin/bash
# weak-scaling bench line at $1 GPUs (torchrun, NCCL counters all-reduce), plus the reference arm
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
head -c 600 gpurun_out/bench_n$N.json; echo; tail -2 gpurun_out/bench_n$N.err
