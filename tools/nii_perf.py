#!/usr/bin/env python3
"""Device-side timing + per-phase cycle breakdown of the decoders: parity (thread-per-frame kernel) against the
non-parity "nii" mode, same frames (development aid)."""
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

lib = _lib.load()
B0 = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
for (N, rate, B) in ((212, '1/3', B0), (48, '1/3', 4 * B0)):
    res = {}
    for mode in ("double-pass", "nii", "nii16"):
        c = turbo.DVBRCS2_Turbo(N, rate, 8, kernel="tpf", boundary=mode)
        h = c.handle
        info = torch.empty((B, 2 * N), dtype=torch.uint8, device="cuda")
        coded = torch.empty((B, h.n_llr), dtype=torch.uint8, device="cuda")
        llr = torch.empty((B, h.n_llr), dtype=torch.float32, device="cuda")
        h.mc_generate_bpsk(B, 1.0 / (2 * (1 / 3) * 10 ** 0.2), 1234, 0, info, coded, llr)
        torch.cuda.synchronize()
        counters = torch.zeros(4, dtype=torch.int64, device="cuda")
        best, avg = timeit(lambda: c.decode_batch(llr, ref_bits=info, counters=counters, out="none"))
        cnt = counters.cpu().numpy()
        fps = B / (best * 1e-3)
        acs = 320 * N * 2 * 8
        res[mode] = fps
        print(f"{mode:11s} N={N} B={B}: {best:.2f} ms (avg {avg:.2f})  {fps/1e6:.3f} Mframes/s  {fps*2*N/1e9:.3f} Gbit/s info  "
              f"(reference-algorithm ACS/s equivalent {fps*acs/(64*148*1.965e9)*100:.1f}% of the ALU roofline); "
              f"BER={cnt[0]/max(cnt[3],1):.4f} FER={cnt[1]/max(cnt[2],1):.4f}")
        ph = np.zeros(8)
        rd = lib.b200dvb_debug_nii_cycles if mode != "double-pass" else lib.b200dvb_debug_tpf_cycles
        rd(_lib.host_ptr(ph), 1)
        h.set_option(_lib.OPT_PHASE_TIMERS, 1)
        c.decode_batch(llr, ref_bits=info, counters=counters, out="none"); torch.cuda.synchronize()
        h.set_option(_lib.OPT_PHASE_TIMERS, 0)
        rd(_lib.host_ptr(ph), 1)
        tot = ph[7]
        if tot > 0:
            names = (["transpose", "in-pass+prep", "boundary+crossing", "-", "out_smem", "out_tmem", "hard"] if mode != "double-pass"
                     else ["transpose", "pass1a+prep", "pass1b", "pass2", "out_smem", "out_tmem", "hard"])
            tiles = B / (32 if mode == "nii16" else 16)
            print("   phases: " + "  ".join(f"{n}={v/tot*100:.1f}%" for n, v in zip(names, ph[:7])))
            print("   cycles per tile-SISO: " + "  ".join(f"{n}={v/tiles/16:.0f}" for n, v in zip(names[1:6], ph[1:6]))
                  + f"   per tile: transpose={ph[0]/tiles:.0f} hard={ph[6]/tiles:.0f} total={tot/tiles:.0f}")
        del info, coded, llr
    print(f"   nii / parity throughput: {res['nii']/res['double-pass']:.3f}x    nii16 / parity: {res['nii16']/res['double-pass']:.3f}x")
