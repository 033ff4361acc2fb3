#!/usr/bin/env python3
"""Throughput of the Monte-Carlo source and of the encoder alone (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from modulations_b200 import dvb_rcs2_turbo as turbo
for N, rate, B in ((212, '1/3', 1 << 20), (48, '1/3', 1 << 21)):
    c = turbo.DVBRCS2_Turbo(N, rate, 8); h = c.handle
    info = torch.empty((B, 2 * N), dtype=torch.uint8, device="cuda")
    coded = torch.empty((B, h.n_llr), dtype=torch.uint8, device="cuda")
    llr = torch.empty((B, h.n_llr), dtype=torch.float32, device="cuda")
    for what, fn in (("mc_generate", lambda: h.mc_generate_bpsk(B, 0.9, 1, 0, info, coded, llr)), ("encode", lambda: h.encode(info))):
        fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        print(f"N={N} {what}: {min(ts):.2f} ms per {B} frames = {B/min(ts)/1e3:.1f} Mframes/s")
