#!/usr/bin/env python3
"""Device-side timing + per-phase cycle breakdown of the thread-per-frame decoder (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from modulations_b200 import dvb_rcs2_turbo as turbo, _lib

def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), sum(ts) / len(ts)

cfgs = [(212, '1/3', int(sys.argv[1]) if len(sys.argv) > 1 else 262144)]
if len(sys.argv) > 2: cfgs.append((48, '1/3', 1 << 20))
for (N, rate, B) in cfgs:
    c = turbo.DVBRCS2_Turbo(N, rate, 8, kernel='tpf')
    h = c.handle
    lib = _lib.load()
    info = torch.empty((B, 2 * N), dtype=torch.uint8, device="cuda")
    coded = torch.empty((B, h.n_llr), dtype=torch.uint8, device="cuda")
    llr = torch.empty((B, h.n_llr), dtype=torch.float32, device="cuda")
    nv = 1.0 / (2 * (1 / 3) * 10 ** 0.2)
    _lib.check(lib.b200dvb_mc_generate_bpsk(h.h, B, nv, 1234, 0, _lib.ptr(info), _lib.ptr(coded), _lib.ptr(llr), _lib.stream_ptr()), "mc")
    torch.cuda.synchronize()
    counters = torch.zeros(4, dtype=torch.int64, device="cuda")
    ref = None if os.environ.get("NOREF") else info
    outm = os.environ.get("OUT", "none")
    best, avg = timeit(lambda: c.decode_batch(llr, ref_bits=ref, counters=counters, out=outm))
    cnt = counters.cpu().numpy()
    fps = B / (best * 1e-3)
    acs = 320 * N * 2 * 8
    print(f"N={N} B={B}: {best:.2f} ms (avg {avg:.2f})  {fps/1e6:.3f} Mframes/s  {fps*2*N/1e9:.3f} Gbit/s info  "
          f"{fps*acs/1e12:.2f} TACS/s = {fps*acs/(64*148*1.965e9)*100:.1f}% of nominal ALU roofline; "
          f"BER={cnt[0]/max(cnt[3],1):.4f} FER={cnt[1]/max(cnt[2],1):.4f}")
    ph = np.zeros(8); lib.b200dvb_debug_tpf_cycles(_lib.host_ptr(ph), 1)
    h.set_option(_lib.OPT_PHASE_TIMERS, 1)        # the instance of the kernel with per-phase clock64() accounting
    c.decode_batch(llr, ref_bits=ref, counters=counters, out=outm); torch.cuda.synchronize()
    h.set_option(_lib.OPT_PHASE_TIMERS, 0)
    lib.b200dvb_debug_tpf_cycles(_lib.host_ptr(ph), 1)
    tot = ph[7]
    if tot > 0:
        names = ["transpose", "pass1a+prep", "pass1b", "pass2", "out_smem", "out_tmem", "hard"]
        tiles = B / 16
        print("   phases: " + "  ".join(f"{n}={v/tot*100:.1f}%" for n, v in zip(names, ph[:7])))
        print("   cycles per tile-SISO: " + "  ".join(f"{n}={v/tiles/16:.0f}" for n, v in zip(names[1:6], ph[1:6]))
              + f"   per tile: transpose={ph[0]/tiles:.0f} hard={ph[6]/tiles:.0f} total={tot/tiles:.0f}")
    del info, coded, llr
