#!/usr/bin/env python3
"""BER / FER of the three decoder modes on the SAME device-generated frames, bijective interleaver (the committed
table floors FER at 1), N=212 R=1/3 BPSK/AWGN, 8 iterations.  99 % binomial half-widths beside each figure."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from modulations_b200 import dvb_rcs2_turbo as turbo

N, rate = 212, '1/3'
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 19
perm = turbo.bijective_interleaver(N)
codecs = {m: turbo.DVBRCS2_Turbo(N, rate, 8, perm=perm, boundary=m) for m in ("double-pass", "nii", "nii16")}
h = codecs["double-pass"].handle
info = torch.empty((B, 2 * N), dtype=torch.uint8, device="cuda")
coded = torch.empty((B, h.n_llr), dtype=torch.uint8, device="cuda")
llr = torch.empty((B, h.n_llr), dtype=torch.float32, device="cuda")
print(f"# {B} frames per point, N={N} R={rate}, bijective interleaver; value +- 99% binomial half-width")
print(f"{'Eb/N0':>6} | " + " | ".join(f"{m + ' BER':>24} {m + ' FER':>22}" for m in codecs))
for ebn0 in (0.0, 2.0, 4.0, 6.0, 8.0, 10.0, 12.0):
    h.mc_generate_bpsk(B, 1.0 / (2.0 * (1 / 3) * 10 ** (ebn0 / 10)), 1000 + int(ebn0), 0, info, coded, llr)
    cells = []
    for m, c in codecs.items():
        cnt = torch.zeros(4, dtype=torch.int64, device="cuda")
        c.decode_batch(llr, ref_bits=info, counters=cnt, out="none")
        be, fe, fr, bits = [float(v) for v in cnt.cpu().numpy()]
        ber, fer = be / bits, fe / fr
        cells.append(f"{ber:12.6e} +-{2.576 * np.sqrt(max(ber * (1 - ber), 1e-15) / bits):9.2e} {fer:10.5f} +-{2.576 * np.sqrt(max(fer * (1 - fer), 1e-15) / fr):8.5f}")
    print(f"{ebn0:6.1f} | " + " | ".join(cells))
