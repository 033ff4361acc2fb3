#!/bin/bash
# 8-GPU box: pinned-copy ceiling at 1/2/4/8 ranks, then bench.py at 8 and 4 ranks
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo8.txt 2>&1; lscpu | grep -E "^CPU\(s\)|NUMA|Model name" >> gpurun_out/r02_topo8.txt
: > gpurun_out/r02_h2d_ceiling.jsonl
timeout 200 python tools/h2d_ceiling.py >> gpurun_out/r02_h2d_ceiling.jsonl 2>> gpurun_out/r02_h2d.err
for n in 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2950$n tools/h2d_ceiling.py >> gpurun_out/r02_h2d_ceiling.jsonl 2>> gpurun_out/r02_h2d.err
done
cat gpurun_out/r02_h2d_ceiling.jsonl
for n in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/r02_bench_n$n.json 2> gpurun_out/r02_bench_n$n.err
  python -c "
import json
d=json.load(open('gpurun_out/r02_bench_n$n.json'))
print('N=$n value', d['value'], 'ms', d['ms_per_step'])
print(' e2e', {k:v for k,v in d['e2e'].items() if k not in ('api','ceiling_what','int32_layout')})
print(' e2e int32', d['e2e']['int32_layout'])
print(' e2e_mc', d['e2e_mc']['value'], d['e2e_mc']['frac_of_resident'])
print(' nonparity', {k:v.get('info_gbit_per_s') for k,v in d['nonparity_modes'].items() if k[0]!='_'})
" || tail -5 gpurun_out/r02_bench_n$n.err
done
