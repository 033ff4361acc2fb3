#!/bin/bash
# round-2 final evidence, part 1 (one GPU): full GPU test suite + smoke, hash-stamped ncu captures, launch list
mkdir -p gpurun_out
( timeout 2400 python -m pytest tests -m gpu -q; echo "pytest rc=$?" ) > gpurun_out/r02_final_gpu_tests.txt 2>&1
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()"; echo "smoke rc=$?" ) >> gpurun_out/r02_final_gpu_tests.txt 2>&1
tail -6 gpurun_out/r02_final_gpu_tests.txt
timeout 200 python tools/mc_perf.py > gpurun_out/r02_mc_source_perf.txt 2>&1; cat gpurun_out/r02_mc_source_perf.txt
timeout 900 python tools/r02_measure.py lat > gpurun_out/r02_latency.txt 2>&1; grep "^lat\|^auto" gpurun_out/r02_latency.txt
timeout 900 python tools/r02_measure.py long > gpurun_out/r02_long.txt 2>&1; cat gpurun_out/r02_long.txt
bash tools/gpu_dram.sh > gpurun_out/r02_dram_capture.log 2>&1; tail -12 gpurun_out/r02_dram_capture.log
timeout 600 python bench.py --steps 2 --warmup 3 --frames 262144 > gpurun_out/r02_launch_plain.json 2> gpurun_out/r02_launch_plain.err &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --frames 262144 > gpurun_out/r02_launch_ncu.log 2>&1
ls -la gpurun_out/r02_launches.csv gpurun_out/prof_dram_*.ncu-rep
