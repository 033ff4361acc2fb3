#!/bin/bash
# first-light run on the B200 box: device info, microbench, parity tests
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
python - <<'PY' > gpurun_out/microbench.txt 2>&1
import ctypes, numpy as np, torch
from modulations_b200 import _lib
lib = _lib.load()
print("devices", lib.b200dvb_device_count())
r = np.zeros(8)
rc = lib.b200dvb_microbench(_lib.host_ptr(r))
print("rc", rc, lib.b200dvb_last_cuda_error())
for n, v in zip(["FADD","FMNMX","ACS(2FADD+FMNMX)","SHFL","DADD","F2F","FADD2x2","clkMHz"], r):
    print(f"{n:18s} {v:10.2f}")
PY
cat gpurun_out/microbench.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.txt
cat gpurun_out/pytest_gpu.txt
