#!/usr/bin/env python3
"""End-to-end (pinned host in -> pinned host out) decode throughput vs pipeline chunk size (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from modulations_b200 import dvb_rcs2_turbo as turbo, _lib
c = turbo.DVBRCS2_Turbo(212, '1/3', 8); h = c.handle; lib = _lib.load()
B = 262144
wave = int(lib.b200dvb_codec_frames_per_wave(h.h))
hin = torch.randn((B, h.n_llr), dtype=torch.float32).pin_memory()
hout = torch.empty((B, 424), dtype=torch.int32, pin_memory=True)
for mult in (1, 2, 3, 4, 8):
    chunk = mult * wave
    for _ in range(2): c.decode_batch_host(hin, hout, chunk=chunk)
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3): c.decode_batch_host(hin, hout, chunk=chunk)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    print(f"chunk = {mult} waves = {chunk} frames: {ms:.2f} ms per {B} frames = {B*424/ms/1e6:.3f} Gbit/s")
