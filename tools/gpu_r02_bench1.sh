#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err ) 2> gpurun_out/r02_bench_a.time; echo "rc=$?" >> gpurun_out/r02_bench_a.time
tail -5 gpurun_out/r02_bench_a.err; cat gpurun_out/r02_bench_a.time
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err ) 2> gpurun_out/r02_bench_ref.time
cat gpurun_out/r02_bench_ref.json; tail -3 gpurun_out/r02_bench_ref.err; cat gpurun_out/r02_bench_ref.time
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_a.json'))
for k in ('value','ms_per_step'): print(k, d[k])
print('roofline', {k:v for k,v in d['roofline'].items() if k not in ('note','traffic_source')})
print('e2e', {k:v for k,v in d['e2e'].items() if k not in ('api','ceiling_what')})
print('e2e_mc', d['e2e_mc']['value'], d['e2e_mc']['frac_of_resident'])
print('nonparity', {k:(v.get('info_gbit_per_s'), v.get('speedup_over_parity_mode')) for k,v in d['nonparity_modes'].items() if k[0]!='_'})
print('cpu', d['cpu_baseline'])
for k in ('latency','n752_r12','config0_n48_qpsk'): print(k, d.get(k))
print('demap', {k:(v['gsym_per_s'], v['roofline']['frac']) for k,v in d['demap'].items() if k!='_what'})
print('mapper', {k:(v['gsym_per_s'], v['roofline']['frac']) for k,v in d['mapper'].items()})
print('waveform', {k:v['roofline']['frac'] for k,v in d['waveform'].items()})
print('clocks', d['clocks'])
"
