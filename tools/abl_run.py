#!/usr/bin/env python3
"""Timing of one (ablation) build of the library: argv[1] = path of the .so, argv[2] = label (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from modulations_b200 import _lib
_lib.LIB_PATH = os.path.abspath(sys.argv[1])         # an explicit, in-process choice of the binary under test
from modulations_b200 import dvb_rcs2_turbo as turbo
lib = _lib.load()
N, rate, B = 212, '1/3', 131072
for mode in ("nii", "nii16"):
    c = turbo.DVBRCS2_Turbo(N, rate, 8, boundary=mode)
    h = c.handle
    info = torch.empty((B, 2 * N), dtype=torch.uint8, device="cuda")
    coded = torch.empty((B, h.n_llr), dtype=torch.uint8, device="cuda")
    llr = torch.empty((B, h.n_llr), dtype=torch.float32, device="cuda")
    h.mc_generate_bpsk(B, 0.95, 1234, 0, info, coded, llr)
    cnt = torch.zeros(4, dtype=torch.int64, device="cuda")
    ts = []
    for i in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); c.decode_batch(llr, ref_bits=info, counters=cnt, out="none"); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ph = np.zeros(8)
    lib.b200dvb_debug_nii_cycles(_lib.host_ptr(ph), 1)
    h.set_option(_lib.OPT_PHASE_TIMERS, 1)
    c.decode_batch(llr, ref_bits=info, counters=cnt, out="none"); torch.cuda.synchronize()
    h.set_option(_lib.OPT_PHASE_TIMERS, 0)
    lib.b200dvb_debug_nii_cycles(_lib.host_ptr(ph), 1)
    tiles = B / (32 if mode == "nii16" else 16)
    cc = cnt.cpu().numpy()
    print(f"{sys.argv[2]:9s} {mode:6s}: BER {cc[0]/cc[3]:.6f} {min(ts[1:]):7.2f} ms   cycles per tile-SISO: in-pass {ph[1]/tiles/16:7.0f}  out {(ph[4]+ph[5])/tiles/16:7.0f}  other {ph[2]/tiles/16:6.0f}")
