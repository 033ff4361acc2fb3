#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_hardening.py tests/test_gpu_tpf.py -x -q -k "latency or tile_boundaries" > gpurun_out/r02_lat_tests.txt 2>&1; echo "rc=$?" >> gpurun_out/r02_lat_tests.txt
tail -5 gpurun_out/r02_lat_tests.txt
timeout 600 python tools/r02_measure.py lat > gpurun_out/r02_latency.txt 2>&1; cat gpurun_out/r02_latency.txt
timeout 300 python - <<'PY' >> gpurun_out/r02_latency.txt 2>&1
import sys, time; sys.path.insert(0, '.')
import numpy as np, torch
from modulations_b200 import dvb_rcs2_turbo as turbo
for N, rate in ((48, '1/3'), (212, '1/3'), (752, '1/2')):
    c = turbo.DVBRCS2_Turbo(N, rate, 8)
    x = (np.random.RandomState(1).randn(c.n_llr) * 3).astype(np.float32)
    for _ in range(5): c.decode(x)
    ts = []
    for _ in range(30):
        t0 = time.perf_counter(); c.decode(x); ts.append(time.perf_counter() - t0)
    print(f"decode() one frame N={N} R={rate}: median {np.median(ts)*1e6:.1f} us wall (min {min(ts)*1e6:.1f})")
PY
tail -4 gpurun_out/r02_latency.txt
