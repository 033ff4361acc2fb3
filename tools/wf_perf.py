#!/usr/bin/env python3
"""Device-side timing of the waveform kernels through the C ABI (development aid; ncu target)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from modulations_b200 import _lib
from modulations_b200.modulators import Modulator

lib = _lib.load()
if len(sys.argv) > 1 and sys.argv[1] == "demap":      # ncu target: two rounds of the six demapper launches, 2^27 symbols
    from modulations_b200.sdr_modem import gray_modem
    n = 1 << 27
    iq = (torch.randn(n, 2, device="cuda") * 0.7).view(torch.complex64).reshape(-1)
    for rnd in range(2):
        for name in ('BPSK', 'QPSK', '8PSK', '16QAM', '64QAM', '256QAM'):
            m = gray_modem(name)
            out = torch.empty(n * m.bps, dtype=torch.float32, device="cuda")
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); _lib.check(lib.b200dvb_demap(m.h, n, _lib.ptr(iq), 0.05, 1.0, _lib.ptr(out), _lib.stream_ptr()), name); b.record()
            torch.cuda.synchronize()
            print(f"demap {name}: {a.elapsed_time(b):.3f} ms")
            del out
    sys.exit(0)
mo = Modulator()
n, nt, sps = 1 << 24, len(mo.rrc_filter), mo.sps
taps = np.ascontiguousarray(mo.rrc_filter, np.float64)
sy = (torch.randn(n, 2, device="cuda") * 0.7).view(torch.complex64).reshape(-1)
shaped = torch.empty((n - 1) * sps + nt, dtype=torch.complex64, device="cuda")
start = 2 * mo.filter_delay
n_mf = (shaped.numel() + nt - 1 - start + sps - 1) // sps
mf = torch.empty(n_mf, dtype=torch.complex64, device="cuda")
calls = (("pulse_shape", n * (8 + 8 * sps), lambda: lib.b200dvb_pulse_shape(n, _lib.ptr(sy), _lib.host_ptr(taps), nt, sps, _lib.ptr(shaped), _lib.stream_ptr())),
                       ("matched_filter", n_mf * (8 * sps + 8), lambda: lib.b200dvb_matched_filter(shaped.numel(), _lib.ptr(shaped), _lib.host_ptr(taps), nt, sps, start, n_mf, _lib.ptr(mf), _lib.stream_ptr())))
if len(sys.argv) > 1 and sys.argv[1] == "once":       # ncu target: (pulse_shape, matched_filter) twice
    for rnd in range(2):
        for name, by, call in calls:
            _lib.check(call(), name)
        torch.cuda.synchronize()
    sys.exit(0)
variants = [int(a) for a in sys.argv[1:] if a.isdigit()] or [0]
for var in variants:
  lib.b200dvb_debug_set_option(1, var)
  print(f"# matched-filter variant {var}")
  for name, by, call in calls:
    ts = []
    for i in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); _lib.check(call(), name); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    t = min(ts[2:])
    print(f"{name}: {t:.3f} ms  {n / t / 1e6:.1f} Gsym/s  {by / t / 1e6:.0f} GB/s = {by / t / 1e6 / 6545.3 * 100:.1f}% of the measured copy bandwidth")
