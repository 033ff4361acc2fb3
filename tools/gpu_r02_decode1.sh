#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests -m gpu -x -q -k "low_latency_kernel or facade or error_behaviour or iterations_attribute" 2>&1 | tail -3
timeout 60 python - <<'PY'
import sys, time; sys.path.insert(0, '.')
import numpy as np, torch
from modulations_b200 import dvb_rcs2_turbo as turbo
for N, rate in ((48, '1/3'), (212, '1/3'), (752, '1/2')):
    c = turbo.DVBRCS2_Turbo(N, rate, 8)
    info = np.random.RandomState(2).randint(0, 2, 2 * N)
    cw = c.encode(info)
    nv = 1.0 / (2 * (c.k_info / c.n_llr) * 10 ** 0.2)
    x = (2 * ((1.0 - 2.0 * cw) + np.sqrt(nv) * np.random.RandomState(1).randn(c.n_llr)) / nv).astype(np.float32)
    ref = c.decode_batch(x[None, :])[0]
    assert np.array_equal(c.decode(x), ref)
    for _ in range(10): c.decode(x)
    ts = []
    for _ in range(50):
        t0 = time.perf_counter(); c.decode(x); ts.append(time.perf_counter() - t0)
    tb = []
    for _ in range(50):
        t0 = time.perf_counter(); c.decode_batch(x[None, :]); tb.append(time.perf_counter() - t0)
    print(f"decode() one frame N={N} R={rate}: median {np.median(ts)*1e6:.1f} us wall (min {min(ts)*1e6:.1f}); through decode_batch: median {np.median(tb)*1e6:.1f} us")
PY
