#!/usr/bin/env python3
"""Three launches of one decode kernel on resident frames (the command the ncu captures wrap).
argv: frames [mode: nii|double-pass] [N] [rate] [kernel: tpf|quad|lat]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from modulations_b200 import dvb_rcs2_turbo as turbo
B = int(sys.argv[1]) if len(sys.argv) > 1 else 37888
mode = sys.argv[2] if len(sys.argv) > 2 else "nii"
N = int(sys.argv[3]) if len(sys.argv) > 3 else 212
rate = sys.argv[4] if len(sys.argv) > 4 else '1/3'
kern = sys.argv[5] if len(sys.argv) > 5 else ("tpf" if N <= 212 else None)
c = turbo.DVBRCS2_Turbo(N, rate, 8, kernel=kern, boundary=mode)
h = c.handle
info = torch.empty((B, 2 * N), dtype=torch.uint8, device="cuda")
coded = torch.empty((B, h.n_llr), dtype=torch.uint8, device="cuda")
llr = torch.empty((B, h.n_llr), dtype=torch.float32, device="cuda")
h.mc_generate_bpsk(max(B, 16), 1.0 / (2 * (c.k_info / h.n_llr) * 10 ** 0.2), 1234, 0, info, coded, llr)
cnt = torch.zeros(4, dtype=torch.int64, device="cuda")
for _ in range(3):
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(); c.decode_batch(llr, ref_bits=info, counters=cnt, out="none"); b.record(); torch.cuda.synchronize()
    print(f"{mode} N={N} B={B}: {a.elapsed_time(b):.3f} ms")
