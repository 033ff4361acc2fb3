#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_tpf.py tests/test_gpu_nii.py tests/test_gpu_codec.py tests/test_gpu_hardening.py -x -q > gpurun_out/r02_tma_tests.txt 2>&1; echo "rc=$?" >> gpurun_out/r02_tma_tests.txt
tail -6 gpurun_out/r02_tma_tests.txt
timeout 300 python tools/nii_perf.py 262144 > gpurun_out/r02_tma_perf.txt 2>&1; cat gpurun_out/r02_tma_perf.txt
timeout 200 python - <<'PY' > gpurun_out/r02_stream_bw.txt 2>&1
# unidirectional streaming calibration: what does a read-only / write-only kernel reach on this part, next to the
# copy figure MEASURED_PEAKS.json is based on?
import torch
n = 1 << 30
a = torch.empty(n, dtype=torch.float32, device="cuda").normal_()
b = torch.empty_like(a)
def t(fn, reps=10):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); best = min(best, s.elapsed_time(e))
    return best
print(f"copy  (read+write 8 GB): {8 * n / t(lambda: b.copy_(a)) / 1e6:.0f} GB/s")
print(f"read  (sum of 4 GB)    : {4 * n / t(lambda: a.sum()) / 1e6:.0f} GB/s")
print(f"read  (amax of 4 GB)   : {4 * n / t(lambda: a.amax()) / 1e6:.0f} GB/s")
print(f"write (fill 4 GB)      : {4 * n / t(lambda: b.fill_(1.0)) / 1e6:.0f} GB/s")
print(f"write (zero 4 GB)      : {4 * n / t(lambda: b.zero_()) / 1e6:.0f} GB/s")
PY
cat gpurun_out/r02_stream_bw.txt
