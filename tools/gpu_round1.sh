#!/bin/bash
# round-1 evidence run: tests, smoke, bench, ncu launch list + full capture of the decoder kernel
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.txt; cat gpurun_out/pytest_gpu.txt
timeout 300 python __graft_entry__.py > gpurun_out/smoke.txt 2>&1; tail -3 gpurun_out/smoke.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
SMALL="python bench.py --steps 2 --warmup 1 --frames 65536 --e2e-frames 32768"
timeout 600 $SMALL > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $SMALL > gpurun_out/ncu_list.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:quad_kernel -s 1 -c 1 -o gpurun_out/prof_quad $SMALL > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:demap_pwl -c 2 -o gpurun_out/prof_demap $SMALL > gpurun_out/ncu_demap.log 2>&1
ls -la gpurun_out
