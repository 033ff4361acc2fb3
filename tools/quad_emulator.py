#!/usr/bin/env python3
"""Lane-level numpy emulation of the "quad" SISO kernel (decode_quad.cu).

Design aid, not product and not the oracle: it replays the CUDA kernel's exact
data flow (4 lanes x 4 states, radix-4 local steps, smem transposes, G-format
branch-metric records, checkpoint/recompute windows, role-mapped extrinsic) in
float32 numpy so the index algebra can be checked against oracle/ on the CPU
before any GPU time is spent.  Run:  python tools/quad_emulator.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
f32 = np.float32


def cidx(s):
    """(wb, yb) class of state s: wb = s0^s1^s2, yb = s1  ->  c = 2*wb + yb."""
    s0, s1, s2 = s & 1, (s >> 1) & 1, (s >> 2) & 1
    return 2 * (s0 ^ s1 ^ s2) + s1


def prep_record(k, sysA, sysB, parW, parY, LaA, LaB):
    """-> (record[8] float32 in smem layout, YA, YB).  record[2c..2c+1] =
    (GP[c], GM[c]), exchanged when k is odd and iw^iy = 1 (flavour 1)."""
    YA = np.float64(sysA) + np.float64(LaA)
    YB = np.float64(sysB) + np.float64(LaB)
    a, b = YA * 0.5, YB * 0.5
    w, y = np.float64(parW) * 0.5, np.float64(parY) * 0.5
    s_, d_ = a + b, a - b
    P = np.zeros(4, f32); M = np.zeros(4, f32)
    for c in range(4):
        sw = -1.0 if (c >> 1) else 1.0
        sy = -1.0 if (c & 1) else 1.0
        P[c] = f32((s_ + sw * w) + sy * y)
        M[c] = f32((d_ + sw * w) + sy * y)
    rec = np.zeros(8, f32)
    for c in range(4):
        GP = max(P[c], -P[3 - c])
        GM = max(M[3 - c], -M[c])
        swap = (k & 1) and (((c >> 1) ^ c) & 1)
        rec[2 * c], rec[2 * c + 1] = (GM, GP) if swap else (GP, GM)
    return rec, YA, YB


def bf(x0, x1, g0, g1):
    return max(f32(x0 + g0), f32(x1 + g1)), max(f32(x0 + g1), f32(x1 + g0))


def lane_classes(p, k_parity):
    """state class c used by butterfly A and B of lane p for a step that uses
    gamma[k] with k of the given parity (same for alpha and beta)."""
    if k_parity == 0:      # butterflies q = p (A), p+4 (B)
        return cidx(p), cidx(p + 4)
    return cidx(2 * p), cidx(2 * p + 1)


def step(R, rec, kpar, direction):
    """One trellis step on a quad.  R[p] = [r0..r3].  Returns un-normalised R'."""
    out = np.zeros((4, 4), f32)
    for p in range(4):
        cA, cB = lane_classes(p, kpar)
        gA = (rec[2 * cA], rec[2 * cA + 1])
        gB = (rec[2 * cB + 1], rec[2 * cB])          # static exchange for butterfly B
        r = R[p]
        if direction == 'a':   # A on (r0,r2) -> (r0',r1');  B on (r1,r3) -> (r2',r3')
            o0, o1 = bf(r[0], r[2], *gA)
            o2, o3 = bf(r[1], r[3], *gB)
        else:                  # beta: same wiring
            o0, o1 = bf(r[0], r[2], *gA)
            o2, o3 = bf(r[1], r[3], *gB)
        out[p] = [o0, o1, o2, o3]
    return out


def normalise(R):
    n = R[0][0]
    return (R - n).astype(f32)


def transpose_alpha(R):
    # (lane j, reg i) -> (lane i, reg j)
    return R.T.copy()


def transpose_beta(R):
    # plain transpose applied to the register tuple (r0, r2, r1, r3)
    perm = [0, 2, 1, 3]
    T = R[:, perm].T.copy()
    out = np.zeros_like(R)
    out[:, perm] = T
    return out


def alpha_layout(k_even):
    """state held by (lane p, reg i)."""
    L = np.zeros((4, 4), int)
    for p in range(4):
        L[p] = [p, p + 4, p + 8, p + 12] if k_even else [2 * p, 2 * p + 1, 2 * p + 8, 2 * p + 9]
    return L


def beta_layout(k_even):
    L = np.zeros((4, 4), int)
    for p in range(4):
        L[p] = [4 * p, 4 * p + 2, 4 * p + 1, 4 * p + 3] if k_even else [2 * p, 2 * p + 8, 2 * p + 1, 2 * p + 9]
    return L


def alpha_step(R, k, recs):
    """alpha[k] -> alpha[k+1] (normalised, laid out for index k+1)."""
    R2 = normalise(step(R, recs[k], k & 1, 'a'))
    if (k + 1) % 2 == 0:
        R2 = transpose_alpha(R2)
    return R2


def beta_step(R, k, recs):
    """beta[k+1] -> beta[k]."""
    R2 = normalise(step(R, recs[k], k & 1, 'b'))
    if k % 2 == 0:
        R2 = transpose_beta(R2)
    return R2


def extrinsic(Ra, Rb, rec, k):
    """alpha[k] (Ra), beta[k+1] (Rb) -> (U0, U3, V1, V2) after the cross-lane max."""
    kpar = k & 1
    part = np.zeros((4, 4), f32)          # per lane T0..T3
    for p in range(4):
        cA, cB = lane_classes(p, kpar)
        T = [None] * 4
        for (c, xi, exch) in ((cA, (0, 2), False), (cB, (1, 3), True)):
            g = (rec[2 * c], rec[2 * c + 1]); h = (rec[2 * (3 - c)], rec[2 * (3 - c) + 1])
            if exch:
                g = g[::-1]; h = h[::-1]
            x0, x1 = Ra[p][xi[0]], Ra[p][xi[1]]
            y0, y1 = Rb[p][xi[0]], Rb[p][xi[1]]
            E0 = max(f32(f32(x0 + g[0]) + y0), f32(f32(x1 + g[0]) + y1))
            E1 = max(f32(f32(x0 + g[1]) + y1), f32(f32(x1 + g[1]) + y0))
            F0 = max(f32(f32(x0 - h[0]) + y0), f32(f32(x1 - h[0]) + y1))
            F1 = max(f32(f32(x0 - h[1]) + y1), f32(f32(x1 - h[1]) + y0))
            if kpar == 0 and exch:       # even k: butterfly B has t=1 -> roles (V1,U0,V2,U3)
                vals = [E1, E0, F1, F0]
            else:
                vals = [E0, E1, F0, F1]
            T = vals if T[0] is None else [max(a, b) for a, b in zip(T, vals)]
        part[p] = T
    # roles per lane: even k -> (U0,V1,U3,V2); odd k: lanes with p1=1 hold (V1,U0,V2,U3)
    res = []
    for r in range(4):
        v = None
        for j in range(4):
            col = r ^ 1 if (kpar == 1 and (j >> 1)) else r
            v = part[j][col] if v is None else max(v, part[j][col])
        res.append(v)
    U0, V1, U3, V2 = res
    return U0, U3, V1, V2


def siso_quad(sysA, sysB, parW, parY, LaA, LaB, sf, N):
    recs = np.zeros((N, 8), f32); YA = np.zeros(N); YB = np.zeros(N)
    for k in range(N):
        recs[k], YA[k], YB[k] = prep_record(k, sysA[k], sysB[k], parW[k], parY[k], LaA[k], LaB[k])
    M = ((N // 2) // 8) * 8
    if M == 0:
        M = N // 2 if (N // 2) % 2 == 0 else N // 2 - 1
    # phase 1
    Ra = np.zeros((4, 4), f32); Rb = np.zeros((4, 4), f32)
    for k in range(N):
        Ra = alpha_step(Ra, k, recs)
    for k in range(N - 1, -1, -1):
        Rb = beta_step(Rb, k, recs)
    # phase 2a: checkpoints
    W = 8
    ckA = {}; ckB = {}
    for k in range(M):
        if k % W == 0:
            ckA[k] = Ra.copy()
        Ra = alpha_step(Ra, k, recs)
    ends = sorted(set(min(j + W, N) for j in range(M, N, W)))
    for j in range(N, M, -1):
        if j in ends:
            ckB[j] = Rb.copy()
        Rb = beta_step(Rb, j - 1, recs)
    out = np.zeros((N, 4), f32)
    # phase 2b, alpha warp: windows ascending over [M, N)
    for j0 in range(M, N, W):
        j1 = min(j0 + W, N)
        ln = j1 - j0
        wb = [None] * ln
        wb[ln - 1] = ckB[j1]
        for i in range(ln - 2, -1, -1):
            wb[i] = beta_step(wb[i + 1], j0 + 1 + i, recs)
        for i in range(ln):
            k = j0 + i
            out[k] = extrinsic(Ra, wb[i], recs[k], k)
            Ra = alpha_step(Ra, k, recs)
    # phase 2b, beta warp: windows descending over [0, M)
    for j1 in range(M, 0, -W):
        j0 = j1 - W
        wa = [None] * W
        wa[0] = ckA[j0]
        for i in range(1, W):
            wa[i] = alpha_step(wa[i - 1], j0 + i - 1, recs)
        for i in range(W - 1, -1, -1):
            k = j0 + i
            out[k] = extrinsic(wa[i], Rb, recs[k], k)
            Rb = beta_step(Rb, k, recs)
    # epilogue
    LeA = np.zeros(N); LeB = np.zeros(N)
    for k in range(N):
        U0, U3, V1, V2 = out[k]
        a, b = YA[k] * 0.5, YB[k] * 0.5
        sP, sM = (a + b) < 0, (a - b) < 0
        app0, app3 = (U3, U0) if sP else (U0, U3)
        app1, app2 = (V2, V1) if sM else (V1, V2)
        LA = f32(max(app0, app1) - max(app2, app3))
        LB = f32(max(app0, app2) - max(app1, app3))
        ea = (np.float64(LA) - YA[k]) * sf
        eb = (np.float64(LB) - YB[k]) * sf
        LeA[k] = min(max(ea, -300.0), 300.0); LeB[k] = min(max(eb, -300.0), 300.0)
    return LeA, LeB


def main():
    from oracle import oracle
    from tests import vectors
    # layout self-check against the trellis
    for N, rate in ((48, '1/3'), (212, '1/3'), (48, '1/2'), (220, '1/3')):
        c = oracle.OracleTurbo(N, rate, 8)
        info, llrs = vectors.codec_inputs(N, rate, 2, [1.0], c.encode, c.n_coded)
        Lc = vectors.depuncture(llrs[0][1], N, c.punct)
        for (LaA, LaB, sf, W, Y) in ((np.zeros(N), np.zeros(N), 0.7, Lc[2], Lc[3]),
                                     (*vectors.siso_apriori(N), 1.0, Lc[4], Lc[5])):
            ref = c.siso(Lc[0], Lc[1], W, Y, LaA, LaB, sf)
            got = siso_quad(Lc[0], Lc[1], W, Y, LaA, LaB, sf, N)
            ok = np.array_equal(ref[0], got[0]) and np.array_equal(ref[1], got[1])
            print(N, rate, sf, "bit-exact" if ok else
                  f"MISMATCH nA={np.sum(ref[0] != got[0])} nB={np.sum(ref[1] != got[1])}")
            assert ok


if __name__ == "__main__":
    main()
