#!/bin/bash
# 4-GPU box: bench.py at 4 and 2 ranks (the driver's own launch line) with the final library of the round
mkdir -p gpurun_out
for n in 4 2; do
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/r02_final_bench_n$n.json 2> gpurun_out/r02_final_bench_n$n.err
  python -c "
import json
d=json.load(open('gpurun_out/r02_final_bench_n$n.json'))
print('N=$n value', round(d['value'],3), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],3), 'ceiling', round(d['e2e']['copy_ceiling_gbit_per_s'],3), 'frac', round(d['e2e']['frac_of_ceiling'],3), 'e2e_mc', round(d['e2e_mc']['value'],3), round(d['e2e_mc']['frac_of_resident'],3), 'nonparity', {k:round(v['info_gbit_per_s'],2) for k,v in d['nonparity_modes'].items() if k[0]!='_'})
" || tail -5 gpurun_out/r02_final_bench_n$n.err
done
