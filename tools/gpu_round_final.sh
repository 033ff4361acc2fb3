#!/bin/bash
# round-end evidence: GPU test suite, bench line, ncu launch list, ncu --set full of the decode kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.txt
cat gpurun_out/pytest_gpu.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
head -c 1500 gpurun_out/bench.json; echo; tail -2 gpurun_out/bench.err
timeout 300 python __graft_entry__.py > gpurun_out/smoke.txt 2>&1; tail -3 gpurun_out/smoke.txt
SMALL="python bench.py --steps 2 --warmup 1 --frames 65536 --e2e-frames 32768"
timeout 600 $SMALL > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $SMALL > gpurun_out/ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tpf_kernel -s 1 -c 1 -o gpurun_out/prof_tpf_full python tools/tpf_perf.py 65536 > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/prof_tpf_full.ncu-rep --page raw --csv > gpurun_out/tpf_full_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_tpf_full.ncu-rep --page source --csv > gpurun_out/tpf_full_src.csv 2>/dev/null
timeout 300 python tools/tpf_perf.py 262144 48 > gpurun_out/tpf_perf_final.txt 2>&1; cat gpurun_out/tpf_perf_final.txt
ls -la gpurun_out | tail -12
