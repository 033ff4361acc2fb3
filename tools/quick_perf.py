#!/usr/bin/env python3
"""Quick device-side timing of the decoder and demapper (development aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from modulations_b200 import dvb_rcs2_turbo as turbo, _lib
from modulations_b200.sdr_modem import gray_modem

def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), sum(ts) / len(ts)

for (N, rate, B) in ((212, '1/3', int(sys.argv[1]) if len(sys.argv) > 1 else 262144), (48, '1/3', 1 << 20)):
    c = turbo.DVBRCS2_Turbo(N, rate, 8)
    h = c.handle
    lib = _lib.load()
    info = torch.empty((B, 2 * N), dtype=torch.uint8, device="cuda")
    coded = torch.empty((B, h.n_llr), dtype=torch.uint8, device="cuda")
    llr = torch.empty((B, h.n_llr), dtype=torch.float32, device="cuda")
    nv = 1.0 / (2 * (1 / 3) * 10 ** 0.2)
    rc = lib.b200dvb_mc_generate_bpsk(h.h, B, nv, 1234, 0, _lib.ptr(info), _lib.ptr(coded), _lib.ptr(llr), _lib.stream_ptr())
    _lib.check(rc, "mc")
    torch.cuda.synchronize()
    counters = torch.zeros(4, dtype=torch.int64, device="cuda")
    best, avg = timeit(lambda: c.decode_batch(llr, ref_bits=info, counters=counters, out="none"))
    cnt = counters.cpu().numpy()
    fps = B / (best * 1e-3)
    acs = 320 * N * 2 * 8
    print(f"N={N} B={B}: {best:.2f} ms (avg {avg:.2f})  {fps/1e6:.3f} Mframes/s  {fps*2*N/1e9:.3f} Gbit/s info  "
          f"{fps*acs/1e12:.2f} TACS/s = {fps*acs/(64*148*1.965e9)*100:.1f}% of nominal ALU roofline; "
          f"BER={cnt[0]/cnt[3]:.4f} FER={cnt[1]/cnt[2]:.4f}")
    ph = np.zeros(8); lib.b200dvb_debug_phase_cycles(_lib.host_ptr(ph), 1)
    c.decode_batch(llr, ref_bits=info, counters=counters, out="none"); torch.cuda.synchronize()
    lib.b200dvb_debug_phase_cycles(_lib.host_ptr(ph), 1)
    tot = ph[5]
    print("   phases: " + "  ".join(f"{n}={v/tot*100:.1f}%" for n, v in zip(["prep","rec_in","rec_out","epi","hard"], ph[:5])) + f"  (CTA-cycles total {tot:.3g})")
    t_gen, _ = timeit(lambda: lib.b200dvb_mc_generate_bpsk(h.h, B, nv, 1234, 0, _lib.ptr(info), _lib.ptr(coded), _lib.ptr(llr), _lib.stream_ptr()))
    print(f"   mc_generate: {t_gen:.2f} ms ({B/t_gen/1e3:.2f} Mframes/s)")
    del info, coded, llr

n = 1 << 26
iq = torch.randn(n, 2, device="cuda").view(torch.complex64).reshape(-1) * 0.7
for name in ('BPSK', 'QPSK', '8PSK', '16QAM', '64QAM', '256QAM'):
    m = gray_modem(name)
    out = torch.empty(n * m.bps, dtype=torch.float32, device="cuda")
    lib = _lib.load()
    best, avg = timeit(lambda: lib.b200dvb_demap(m.h, n, _lib.ptr(iq), 0.05, 1.0, _lib.ptr(out), _lib.stream_ptr()), 5)
    by = n * (8 + 4 * m.bps)
    print(f"demap {name:7s}: {best:.3f} ms  {n/best/1e6:.1f} Gsym/s  {by/best/1e6:.0f} GB/s = {by/best/1e6/6545.3*100:.1f}% of measured HBM")
    del out
