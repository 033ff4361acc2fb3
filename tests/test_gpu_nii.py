"""GPU tests of the NON-PARITY decoder mode boundary="nii" (SURVEY 8(f) N2): the kernel (csrc/decode_nii.cu) must
equal ITS OWN model (oracle/nii_model.c) bit for bit, and its BER/FER must sit inside the binomial confidence
interval of the parity mode's.  Nothing here claims parity with the reference."""
import os

import numpy as np
import pytest

from oracle import oracle
from tests import vectors

pytestmark = pytest.mark.gpu


def _llrs(o, N, rate, nfr, ebn0, seed):
    rs = np.random.RandomState(seed)
    info = rs.randint(0, 2, (nfr, 2 * N))
    llr = np.stack([vectors.awgn_llr(rs, np.asarray(o.encode(info[i])), rate, ebn0) for i in range(nfr)])
    return info, llr


MODELS = {"nii": oracle.NiiModel, "nii16": oracle.Nii16Model}


@pytest.mark.parametrize("mode", ["nii", "nii16"])
@pytest.mark.parametrize("N,rate,iters", [(212, '1/3', 8), (212, '1/2', 3), (48, '1/3', 8), (48, '3/4', 2), (64, '1/3', 4),
                                          (64, '2/3', 1), (212, '3/4', 2)])
@pytest.mark.parametrize("nfr", [1, 16, 17, 32, 37, 70])
def test_nii_kernel_equals_its_model(mode, N, rate, iters, nfr):
    from modulations_b200 import dvb_rcs2_turbo as turbo
    g = turbo.DVBRCS2_Turbo(N, rate, iters, boundary=mode)
    m = MODELS[mode](N, rate, iters, perm=g.perm, inv_perm=g.inv_perm)
    info, llr = _llrs(m, N, rate, nfr, 2.0, 31 + N + nfr)
    if llr.shape[1] < g.n_llr:                       # rate 2/3: n_coded is short of what the depuncturer consumes (reference bug kept)
        llr = np.pad(llr, ((0, 0), (0, g.n_llr - llr.shape[1])))
    ref = m.decode_batch(llr)
    dec = g.decode_batch(llr)
    assert dec.dtype == np.int32 and dec.shape == (nfr, 2 * N)
    assert np.array_equal(dec, ref), f"N={N} R={rate} B={nfr}: {np.sum(dec != ref)} bits differ"
    assert np.array_equal(g.decode(llr[0]), ref[0])


@pytest.mark.parametrize("mode", ["nii", "nii16"])
def test_nii_strided_packed_counters_and_host_pipeline(mode):
    import torch
    from modulations_b200 import dvb_rcs2_turbo as turbo
    N, rate, iters, nfr = 212, '1/3', 4, 50
    g = turbo.DVBRCS2_Turbo(N, rate, iters, boundary=mode)
    m = MODELS[mode](N, rate, iters, perm=g.perm, inv_perm=g.inv_perm)
    info, llr = _llrs(m, N, rate, nfr, 1.0, 4242)
    ref = m.decode_batch(llr)
    n = llr.shape[1]
    for pitch in (n, n + 4, n + 3):
        buf = torch.zeros((nfr, pitch), dtype=torch.float32, device="cuda")
        buf[:, :n] = torch.from_numpy(llr).cuda()
        counters = torch.zeros(4, dtype=torch.int64, device="cuda")
        dec = g.decode_batch(buf[:, :n], ref_bits=torch.from_numpy(info.astype(np.uint8)).cuda(), counters=counters)
        assert np.array_equal(dec.cpu().numpy(), ref), f"pitch {pitch}"
        cnt = counters.cpu().numpy()
        assert cnt[0] == np.sum(ref != info) and cnt[1] == np.sum(np.any(ref != info, axis=1))
        assert cnt[2] == nfr and cnt[3] == nfr * 2 * N
    got = turbo.unpack_bits(g.decode_batch(llr, out="packed"), 2 * N)
    assert np.array_equal(got, ref)
    big = torch.from_numpy(llr).repeat(300, 1).contiguous().pin_memory()
    out = g.decode_batch_host(big, out="packed")
    assert np.array_equal(turbo.unpack_bits(out.numpy(), 2 * N), np.tile(ref, (300, 1)))


@pytest.mark.parametrize("mode", ["nii", "nii16"])
@pytest.mark.parametrize("N,rate,B", [(212, '1/3', 60_000), (48, '1/2', 60_000)])
def test_nii_randomised_large_batch_vs_model(mode, N, rate, B):
    """Distinct frames in every tile and wave (device Philox source) against the model on every host thread."""
    import torch
    from modulations_b200 import dvb_rcs2_turbo as turbo
    g = turbo.DVBRCS2_Turbo(N, rate, 8, boundary=mode)
    m = MODELS[mode](N, rate, 8, perm=g.perm, inv_perm=g.inv_perm)
    h = g.handle
    info = torch.empty((B, g.k_info), dtype=torch.uint8, device="cuda")
    coded = torch.empty((B, h.n_llr), dtype=torch.uint8, device="cuda")
    llr = torch.empty((B, h.n_llr), dtype=torch.float32, device="cuda")
    h.mc_generate_bpsk(B, 1.0 / (2.0 * (g.k_info / h.n_llr) * 10 ** 0.2), 99 + N, 0, info, coded, llr)
    got = turbo.unpack_bits(g.decode_batch(llr, out="packed").cpu(), g.k_info)
    ref = m.decode_batch(llr.cpu().numpy(), threads=os.cpu_count() or 1)
    bad = np.flatnonzero(np.any(got != ref, axis=1))
    assert bad.size == 0, f"{bad.size} of {B} frames differ (first: {bad[:8]})"
    again = turbo.unpack_bits(g.decode_batch(llr, out="packed").cpu(), g.k_info)
    assert np.array_equal(again, got)


@pytest.mark.parametrize("mode", ["nii", "nii16"])
def test_nii_ber_inside_parity_confidence_interval(mode):
    """BER / FER of a non-parity mode against the parity mode on the SAME frames, with a bijective interleaver (the
    committed table floors FER at 1): each mode's counts must lie inside the 99 % binomial interval
    (normal approximation, 2.576 sigma of the pooled estimate) of the other's.  (The frames are the same, so the
    two estimates are positively correlated and the test is conservative in the right direction: a real difference
    larger than the interval of INDEPENDENT samples fails.)  The fixed-point mode additionally gets a stated
    quantisation allowance: 1 % of the BER and 4 % of the FER, RELATIVE.  With 55 M bits per point the binomial
    interval is 0.17 % of the BER at 2 dB and resolves the format's real cost: +0.3 ... +0.7 % of the BER and +1 ... +3 % of
    the FER in the waterfall (a few hundredths of a dB; profiles/r02_ber_three_modes.txt).  The clamps play no part in
    it (the C model gives identical decisions with wider ones); it is the 1/4-LLR resolution: a-posteriori values that
    come out EXACTLY zero are decided arbitrarily, which a float decoder never sees."""
    import torch
    from modulations_b200 import dvb_rcs2_turbo as turbo
    N, rate, B = 212, '1/3', 1 << 17
    perm = turbo.bijective_interleaver(N)
    par = turbo.DVBRCS2_Turbo(N, rate, 8, perm=perm)
    nii = turbo.DVBRCS2_Turbo(N, rate, 8, perm=perm, boundary=mode)
    h = par.handle
    info = torch.empty((B, par.k_info), dtype=torch.uint8, device="cuda")
    coded = torch.empty((B, h.n_llr), dtype=torch.uint8, device="cuda")
    llr = torch.empty((B, h.n_llr), dtype=torch.float32, device="cuda")
    for ebn0 in (2.0, 5.0, 8.0):
        h.mc_generate_bpsk(B, 1.0 / (2.0 * (1 / 3) * 10 ** (ebn0 / 10)), 7 + int(ebn0), 0, info, coded, llr)
        c0 = torch.zeros(4, dtype=torch.int64, device="cuda")
        c1 = torch.zeros(4, dtype=torch.int64, device="cuda")
        par.decode_batch(llr, ref_bits=info, counters=c0, out="none")
        nii.decode_batch(llr, ref_bits=info, counters=c1, out="none")
        c0, c1 = c0.cpu().numpy().astype(float), c1.cpu().numpy().astype(float)
        assert c0[2] == B and c1[2] == B
        for what, e0, e1, n in (("BER", c0[0], c1[0], c0[3]), ("FER", c0[1], c1[1], c0[2])):
            p = (e0 + e1) / (2 * n)
            half = 2.576 * np.sqrt(max(p * (1 - p), 1e-12) * 2 / n)
            if mode == "nii16":
                half = max(half, (0.01 if what == "BER" else 0.04) * e0 / n)
            assert abs(e0 / n - e1 / n) <= half + 1e-12, (ebn0, what, e0 / n, e1 / n, half)


def test_nii_is_refused_where_it_has_no_kernel():
    from modulations_b200 import dvb_rcs2_turbo as turbo
    with pytest.raises(ValueError):
        turbo.DVBRCS2_Turbo(752, '1/2', 8, boundary="nii").handle
    with pytest.raises(ValueError):
        turbo.DVBRCS2_Turbo(212, '1/3', 8, boundary="sliding")
