#!/usr/bin/env python3
"""Three launches of the encoder on resident frames (ncu target)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from modulations_b200 import dvb_rcs2_turbo as turbo
B = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
c = turbo.DVBRCS2_Turbo(212, '1/3', 8); h = c.handle
info = torch.randint(0, 2, (B, 424), dtype=torch.uint8, device="cuda")
for _ in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); h.encode(info); b.record(); torch.cuda.synchronize()
    print(f"encode B={B}: {a.elapsed_time(b):.3f} ms")
