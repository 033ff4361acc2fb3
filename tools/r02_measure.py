#!/usr/bin/env python3
"""Round-2 measurement pack (development aid, one GPU): packed-16/DPX issue rates, decode latency against
batch size for both decode kernels, the long frames of the table, every demapper modulation and the mapper."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from modulations_b200 import dvb_rcs2_turbo as turbo, _lib
from modulations_b200.sdr_modem import gray_modem

lib = _lib.load()

def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), float(np.median(ts))

which = sys.argv[1:] or ["mb", "lat", "long", "demap"]

if "mb" in which:
    r = np.zeros(16); _lib.check(lib.b200dvb_microbench2(_lib.host_ptr(r)), "mb2")
    names = ["add.f16x2", "max.f16x2", "mix f16x2 (2add+max+sub)", "mix f32 (2add+max+sub)", "add.s32", "max.s32",
             "viaddmax.s32", "add.s16x2", "max.s16x2", "viaddmax.s16x2", "mix s16x2 (add+viaddmax+add)", "mix s32 (add+viaddmax+sub)"]
    print("# microbench2: lane-instructions per clock per SM")
    for n, v in zip(names, r): print(f"{n:32s} {v:8.2f}")

def gen(c, B, seed=1234):
    h = c.handle
    info = torch.empty((B, 2 * c.N), dtype=torch.uint8, device="cuda")
    coded = torch.empty((B, h.n_llr), dtype=torch.uint8, device="cuda")
    llr = torch.empty((B, h.n_llr), dtype=torch.float32, device="cuda")
    nv = 1.0 / (2 * turbo_rate(c) * 10 ** 0.2)
    B16 = B
    _lib.check(lib.b200dvb_mc_generate_bpsk(h.h, B16, nv, seed, 0, _lib.ptr(info), _lib.ptr(coded), _lib.ptr(llr), _lib.stream_ptr()), "mc")
    torch.cuda.synchronize()
    return info, llr

def turbo_rate(c):
    return c.k_info / c.handle.n_llr

if "lat" in which:
    print("# decode latency vs batch, N=212 R=1/3 8 it: resident (device tensors, CUDA events) and decode()/decode_batch on numpy (wall)")
    for kern in ("tpf", "quad", "lat", "auto"):
        c = turbo.DVBRCS2_Turbo(212, '1/3', 8, kernel=kern)
        for B in (1, 16, 148, 296, 1024, 4096, 65536):
            if kern == "lat" and B > 4096: continue
            info, llr = gen(c, max(B, 16))
            llr = llr[:B].contiguous()
            best, med = timeit(lambda: c.decode_batch(llr, out="packed"))
            xh = llr.cpu().numpy()
            for _ in range(2): c.decode_batch(xh)
            t = []
            for _ in range(5):
                t0 = time.perf_counter(); c.decode_batch(xh); t.append(time.perf_counter() - t0)
            print(f"{kern:5s} B={B:6d}: resident {best*1e3:9.1f} us (median {med*1e3:9.1f})  = {B/best/1e3:8.3f} Mframes/s | numpy in/out wall {min(t)*1e6:9.1f} us")

if "latvar" in which:
    print("# low-latency kernel, one frame resident, 8 it")
    for N, rate in ((48, '1/3'), (212, '1/3'), (752, '1/2')):
        ref = turbo.DVBRCS2_Turbo(N, rate, 8, kernel="quad")
        c = turbo.DVBRCS2_Turbo(N, rate, 8, kernel="lat")
        info, llr = gen(c, 16)
        want = ref.decode_batch(llr)
        for var in (32, 48, 64, 80, 96, 128, 0):
            _lib.check(lib.b200dvb_debug_set_option(3, var), "dbg")
            ok = bool(torch.equal(c.decode_batch(llr), want))
            one = llr[:1].contiguous()
            best, med = timeit(lambda: c.decode_batch(one, out="packed"))
            print(f"N={N:4d} R={rate} warm-up {var:3d}: {best*1e3:8.1f} us (median {med*1e3:8.1f})  bit-exact vs quad: {ok}")
            c.handle.set_option(_lib.OPT_PHASE_TIMERS, 1)
            ph = np.zeros(8); lib.b200dvb_debug_lat_cycles(_lib.host_ptr(ph), 1)
            c.decode_batch(one, out="packed"); torch.cuda.synchronize()
            lib.b200dvb_debug_lat_cycles(_lib.host_ptr(ph), 1)
            c.handle.set_option(_lib.OPT_PHASE_TIMERS, 0)
            print("        cycles: setup %d  P0 %d  P1 %d  P2 %d  hard %d  total %d  (per SISO: P0 %.0f P1 %.0f P2 %.0f)" % (*ph[:6], ph[1]/16, ph[2]/16, ph[3]/16))

if "long" in which:
    print("# long frames (table N), 8 it, resident, B sized to ~4 waves; % of the 64 ACS/clk/SM roofline at 1965 MHz")
    for N, rate in ((48, '1/3'), (64, '1/3'), (212, '1/3'), (212, '1/2'), (220, '1/3'), (424, '1/3'), (752, '1/2'), (752, '1/3'), (848, '1/3')):
        c = turbo.DVBRCS2_Turbo(N, rate, 8)
        wave = int(lib.b200dvb_codec_frames_per_wave(c.handle.h))
        B = max(16, (4 * wave // 16) * 16)
        info, llr = gen(c, B)
        cnt = torch.zeros(4, dtype=torch.int64, device="cuda")
        best, med = timeit(lambda: c.decode_batch(llr, ref_bits=info, counters=cnt, out="none"), n=3, warm=1)
        fps = B / (best * 1e-3); acs = 320 * N * 2 * 8
        print(f"N={N:4d} R={rate}: wave={wave:5d} B={B:6d} {best:8.2f} ms  {fps/1e6:7.3f} Mframes/s  {fps*2*N/1e9:6.3f} Gbit/s  {fps*acs/(64*148*1.965e9)*100:5.1f}% of ALU roofline")
        if N > 212:
            c.handle.set_option(_lib.OPT_PHASE_TIMERS, 1)
            ph = np.zeros(8); lib.b200dvb_debug_phase_cycles(_lib.host_ptr(ph), 1)
            c.decode_batch(llr, out="none"); torch.cuda.synchronize()
            lib.b200dvb_debug_phase_cycles(_lib.host_ptr(ph), 1)
            c.handle.set_option(_lib.OPT_PHASE_TIMERS, 0)
            tot = ph[5] if ph[5] else 1.0
            print("        phases (share of CTA cycles): prep %.1f%%  recursion in %.1f%%  out %.1f%%  epilogue %.1f%%  hard decision %.1f%%" % tuple(100 * ph[i] / tot for i in range(5)))
        del info, llr

if "demap" in which:
    print("# demapper / mapper, 2^27 symbols, resident; HBM fraction against MEASURED_PEAKS hbm_gbs 6545.3")
    n = 1 << 27
    iq = torch.randn(n, 2, device="cuda").view(torch.complex64).reshape(-1) * 0.7
    for name in ('BPSK', 'QPSK', '8PSK', '16QAM', '64QAM', '256QAM'):
        m = gray_modem(name)
        out = torch.empty(n * m.bps, dtype=torch.float32, device="cuda")
        best, med = timeit(lambda: lib.b200dvb_demap(m.h, n, _lib.ptr(iq), 0.05, 1.0, _lib.ptr(out), _lib.stream_ptr()))
        by = n * (8 + 4 * m.bps)
        print(f"demap {name:7s}: {med:.3f} ms  {n/med/1e6:.1f} Gsym/s  {by/med/1e6:.0f} GB/s = {by/med/1e6/6545.3*100:.1f}%")
        del out
        bits = torch.randint(0, 2, (n * m.bps,), dtype=torch.uint8, device="cuda")
        sy = torch.empty(n, dtype=torch.complex64, device="cuda")
        by = n * (m.bps + 8)
        for var, what in ((1, "one symbol per lane"), (2, "four per lane, 256-bit store"), (0, "shipped choice")):
            if var == 1 and m.bps in (3, 6): continue
            _lib.check(lib.b200dvb_debug_set_option(2, var), "dbg")
            best, med = timeit(lambda: lib.b200dvb_map(m.h, n, _lib.ptr(bits), _lib.ptr(sy), 0, _lib.stream_ptr()))
            print(f"map   {name:7s} [{what:28s}]: {med:.3f} ms  {n/med/1e6:.1f} Gsym/s  {by/med/1e6:.0f} GB/s (uint8 bit per byte in) = {by/med/1e6/6545.3*100:.1f}%")
        del bits, sy
