"""GPU parity tests for the turbo codec: CUDA path (through the C-ABI) vs oracle/
and vs the committed golden outputs of the reference.  Bit-exact everywhere."""
import numpy as np
import pytest

from oracle import oracle
from tests import vectors

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "these tests need the B200"
    return torch


def _codecs(N, rate, iters, golden):
    from modulations_b200 import dvb_rcs2_turbo as turbo
    t = golden["tables"]
    g = turbo.DVBRCS2_Turbo(N, rate, iters)
    o = oracle.OracleTurbo(N, rate, iters)
    # the fixture was generated with the authoring host's argsort tie order; the GPU
    # box's numpy may order ties differently, so both sides use ITS table
    # when they disagree and the golden `dec` is then checked through the oracle only.
    same_host = np.array_equal(g.inv_perm, t[f"inv_perm_{N}"])
    return g, o, same_host


@pytest.mark.parametrize("case", vectors.CODEC_CASES, ids=lambda c: vectors.case_tag(*c[:3]))
def test_decode_matches_reference(torch_cuda, golden, case):
    N, rate, iters, nfr, ebn0s = case
    g, o, same_host = _codecs(N, rate, iters, golden)
    k = golden["codec_kat"]
    tag = vectors.case_tag(N, rate, iters)
    info, llrs = vectors.codec_inputs(N, rate, nfr, ebn0s, o.encode, o.n_coded)
    coded = g.encode_batch(info)
    assert coded.dtype == np.uint8
    assert np.array_equal(np.packbits(coded, axis=1), k[f"{tag}/coded"])          # encoder: bit-exact
    assert np.array_equal(g.encode(info[0]), o.encode(info[0]))
    from modulations_b200 import dvb_rcs2_turbo as turbo
    forced = turbo.DVBRCS2_Turbo(N, rate, iters, kernel="tpf") if N <= 212 else None   # small batches go to the quad kernel
    for e, llr in zip(ebn0s, llrs):
        dec = g.decode_batch(llr)
        assert dec.dtype == np.int32 and dec.shape == (nfr, 2 * N)
        ref = o.decode_batch(llr)
        assert np.array_equal(dec, ref), f"{tag} Eb/N0={e}: {np.sum(dec != ref)} bits differ"
        if forced is not None:
            assert np.array_equal(forced.decode_batch(llr), ref), f"{tag} Eb/N0={e}: thread-per-frame kernel differs"
        if same_host:
            assert np.array_equal(np.packbits(dec.astype(np.uint8), axis=1), k[f"{tag}/ebn0_{e}/dec"])
    assert np.array_equal(g.decode(llrs[0][0]), ref[0] if len(ebn0s) == 1 else o.decode(llrs[0][0]))


@pytest.mark.parametrize("case", vectors.CODEC_CASES, ids=lambda c: vectors.case_tag(*c[:3]))
def test_siso_bit_exact(torch_cuda, golden, case):
    from modulations_b200 import dvb_rcs2_turbo as turbo
    N, rate, iters, nfr, ebn0s = case
    g, o, _ = _codecs(N, rate, iters, golden)
    k = golden["codec_kat"]
    tag = vectors.case_tag(N, rate, iters)
    info, llrs = vectors.codec_inputs(N, rate, nfr, ebn0s, o.encode, o.n_coded)
    Lc = vectors.depuncture(llrs[0][0], N, g.punct)
    z = np.zeros(N)
    args = (g.next_state, g.out_W, g.out_Y, g.prev_state, g.prev_input, N)
    LeA, LeB = turbo.bcjr_max_log_map(Lc[0], Lc[1], Lc[2], Lc[3], z, z, *args, 0.7)
    assert LeA.dtype == np.float64 and LeA.shape == (N,)
    assert np.array_equal(LeA, k[f"{tag}/siso0_LeA"]) and np.array_equal(LeB, k[f"{tag}/siso0_LeB"])
    LaA, LaB = vectors.siso_apriori(N)
    LeA, LeB = turbo.bcjr_max_log_map(Lc[0], Lc[1], Lc[4], Lc[5], LaA, LaB, *args, 1.0)
    assert np.array_equal(LeA, k[f"{tag}/siso1_LeA"]) and np.array_equal(LeB, k[f"{tag}/siso1_LeB"])
    # historic aliases run the same committed arithmetic
    a7 = turbo.bcjr_decode_circular(Lc[0], Lc[1], Lc[4], Lc[5], LaA, LaB, 1.0)
    assert np.array_equal(a7[0], LeA)
    a11 = turbo.max_log_map_decode(Lc[0], Lc[1], Lc[4], Lc[5], LaA, LaB, g.next_state, g.prev_state,
                                   g.out_W, g.out_Y, 1.0)
    assert np.array_equal(a11[1], LeB)


def test_siso_batch_ragged(torch_cuda, golden):
    """B not a multiple of the 8 frames a CTA holds; every frame checked vs oracle."""
    from modulations_b200 import dvb_rcs2_turbo as turbo
    N = 48
    o = oracle.OracleTurbo(N, '1/3', 1)
    rs = np.random.RandomState(5)
    B = 19
    Lc = [(rs.randn(B, N) * 4).astype(np.float32) for _ in range(4)]
    La = [rs.randn(B, N) * 2 for _ in range(2)]
    LeA, LeB = turbo.bcjr_max_log_map(*Lc, *La, o.next_state, o.out_W, o.out_Y, o.prev_state,
                                      o.prev_input, N, 0.7)
    for b in range(B):
        ra, rb = o.siso(*[x[b] for x in Lc], *[x[b] for x in La], 0.7)
        assert np.array_equal(LeA[b], ra) and np.array_equal(LeB[b], rb), b


def test_decode_large_batch_and_counters(torch_cuda, golden):
    """More groups than resident CTAs (persistent loop), in-kernel error counters,
    packed output and torch-tensor in/out."""
    torch = torch_cuda
    from modulations_b200 import dvb_rcs2_turbo as turbo
    N, rate = 48, '1/3'
    g = turbo.DVBRCS2_Turbo(N, rate, 8)
    o = oracle.OracleTurbo(N, rate, 8, perm=g.perm, inv_perm=g.inv_perm)
    B = 8 * 1500 + 3
    rs = np.random.RandomState(11)
    info = rs.randint(0, 2, (64, 2 * N))
    coded = o.encode_batch(info)
    llr64 = np.stack([vectors.awgn_llr(rs, coded[i], rate, 1.0) for i in range(64)])
    reps = (B + 63) // 64
    llr = np.tile(llr64, (reps, 1))[:B]
    ref_bits = np.tile(info, (reps, 1))[:B].astype(np.uint8)
    want = o.decode_batch(llr64, threads=4)
    counters = torch.zeros(4, dtype=torch.int64, device="cuda")
    x = torch.from_numpy(llr).cuda()
    dec = g.decode_batch(x, ref_bits=torch.from_numpy(ref_bits).cuda(), counters=counters)
    assert dec.is_cuda and dec.dtype == torch.int32
    dec = dec.cpu().numpy()
    assert np.array_equal(dec, np.tile(want, (reps, 1))[:B])
    c = counters.cpu().numpy()
    errs = (dec != ref_bits)
    assert c[0] == errs.sum() and c[1] == errs.any(axis=1).sum() and c[2] == B and c[3] == B * 2 * N
    packed = g.decode_batch(x, out="packed").cpu().numpy().view(np.uint32)
    bits = ((packed[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(B, -1)[:, :2 * N]
    assert np.array_equal(bits, dec)


def test_decode_properties_full_size(torch_cuda):
    """Headline size (N=212, R=1/3, 8 it): noiseless frames decode to what the
    oracle decodes them to, duplicates decode identically, a strided input view
    gives the same bits as a contiguous one."""
    torch = torch_cuda
    from modulations_b200 import dvb_rcs2_turbo as turbo
    N, rate = 212, '1/3'
    g = turbo.DVBRCS2_Turbo(N, rate, 8)
    o = oracle.OracleTurbo(N, rate, 8, perm=g.perm, inv_perm=g.inv_perm)
    rs = np.random.RandomState(3)
    info = rs.randint(0, 2, (32, 2 * N))
    coded = g.encode_batch(info)
    assert np.array_equal(coded, o.encode_batch(info).astype(np.uint8))
    llr32 = np.stack([vectors.awgn_llr(rs, coded[i].astype(np.int32), rate, 2.0) for i in range(32)])
    want = o.decode_batch(llr32, threads=4)
    B = 20000
    idx = rs.randint(0, 32, B)
    dec = g.decode_batch(llr32[idx])
    assert np.array_equal(dec, want[idx])
    wide = torch.zeros((64, g.n_llr + 40), dtype=torch.float32, device="cuda")
    wide[:, :g.n_llr] = torch.from_numpy(llr32[idx[:64]]).cuda()
    assert np.array_equal(g.decode_batch(wide).cpu().numpy(), want[idx[:64]])


def test_decode_error_behaviour(torch_cuda):
    from modulations_b200 import dvb_rcs2_turbo as turbo
    g = turbo.DVBRCS2_Turbo(212, '2/3', 1)
    assert g.n_coded == 630 and g.n_llr == 636            # reference n_coded bug
    with pytest.raises(IndexError):
        g.decode(np.zeros(g.n_coded, np.float32))         # reference: IndexError at :484
    o = oracle.OracleTurbo(212, '2/3', 1, perm=g.perm, inv_perm=g.inv_perm)
    rs = np.random.RandomState(1)
    llr = (rs.randn(g.n_llr) * 3).astype(np.float32)
    assert np.array_equal(g.decode(llr), o.decode(llr))


def test_facade_and_circular_state(torch_cuda, golden):
    from modulations_b200 import dvb_rcs2_turbo as turbo
    c = turbo.DVB_RCS2_TurboCodec(block_length=212, code_rate='1/2', n_iterations=2)
    o = oracle.OracleTurbo(212, '1/2', 2, perm=c.interleaver.perm, inv_perm=c.interleaver.inv_perm)
    assert (c.k_info, c.n_coded, c.N, c.code_rate) == (424, 848, 212, 0.5)
    rs = np.random.RandomState(42)
    info = rs.randint(0, 2, c.k_info)
    coded, circ = o.encode(info, return_circ=True)
    assert np.array_equal(c.encode(info), coded)
    A, B = info[0::2], info[1::2]
    assert turbo.determine_circular_state(A, B) == circ[0]
    Ai, Bi = c.interleaver.interleave(A, B)
    assert turbo.determine_circular_state(Ai, Bi) == circ[1]
    W, Y = c.encoder1.encode(A, B)
    full = oracle.OracleTurbo(212, '1/3', 1, perm=o.perm, inv_perm=o.inv_perm).encode(info).reshape(212, 6)
    assert np.array_equal(W, full[:, 2]) and np.array_equal(Y, full[:, 3])
    W2, Y2 = c.encoder2.encode(Ai, Bi)
    assert np.array_equal(W2, full[:, 4]) and np.array_equal(Y2, full[:, 5])
    llr = vectors.awgn_llr(rs, coded, '1/2', 3.0)
    Lc = vectors.depuncture(llr, 212, o.punct)
    z = np.zeros(212)
    LeA, LeB = c.decoder1.decode(Lc[0], Lc[1], Lc[2], Lc[3], z, z)
    ra, rb = o.siso(Lc[0], Lc[1], Lc[2], Lc[3], z, z, 0.7)
    assert np.array_equal(LeA, ra) and np.array_equal(LeB, rb)
    assert np.array_equal(c.decode(llr), o.decode(llr))
    lut = golden["tables"]["circ_lut_212"]
    import ctypes
    from modulations_b200 import _lib
    got = np.zeros(16, np.int32)
    _lib.check(_lib.load().b200dvb_codec_circular_lut(c._codec.handle.h, _lib.host_ptr(got)))
    assert np.array_equal(got, lut)


@pytest.mark.parametrize("N", [48, 212])
def test_siso_edge_values(torch_cuda, N):
    """Ties, zeros, signed zeros, saturating and clipping inputs: still bit-exact."""
    from modulations_b200 import dvb_rcs2_turbo as turbo
    o = oracle.OracleTurbo(N, '1/3', 1)
    rs = np.random.RandomState(N)
    cases = []
    z32 = np.zeros(N, np.float32); z64 = np.zeros(N)
    cases.append(([z32] * 4, [z64] * 2, 0.7))                                   # all ties
    cases.append(([np.full(N, -0.0, np.float32)] * 4, [np.full(N, -0.0)] * 2, 1.0))
    big = [(rs.choice([-50.0, 50.0], N)).astype(np.float32) for _ in range(4)]
    cases.append((big, [rs.choice([-300.0, 300.0], N) for _ in range(2)], 0.7))  # saturated, clips at +-300
    huge = [(rs.randn(N) * 1e4).astype(np.float32) for _ in range(4)]
    cases.append((huge, [rs.randn(N) * 1e4 for _ in range(2)], 1.0))
    tiny = [(rs.randn(N) * 1e-30).astype(np.float32) for _ in range(4)]
    cases.append((tiny, [rs.randn(N) * 1e-200 for _ in range(2)], 0.7))
    quant = [(rs.randint(-4, 5, N) * 0.5).astype(np.float32) for _ in range(4)]   # many exact ties
    cases.append((quant, [rs.randint(-4, 5, N) * 0.25 for _ in range(2)], 0.7))
    punct = [quant[0], quant[1], z32, quant[3]]                                  # a punctured parity stream
    cases.append((punct, [z64, z64], 0.7))
    args = (o.next_state, o.out_W, o.out_Y, o.prev_state, o.prev_input, N)
    for i, (Lc, La, sf) in enumerate(cases):
        ga, gb = turbo.bcjr_max_log_map(*Lc, *La, *args, sf)
        ra, rb = o.siso(*Lc, *La, sf)
        assert np.array_equal(ga, ra) and np.array_equal(gb, rb), f"case {i}"


def test_decode_all_zero_and_saturated_frames(torch_cuda):
    from modulations_b200 import dvb_rcs2_turbo as turbo
    for N, rate in ((48, '1/2'), (212, '1/3')):
        g = turbo.DVBRCS2_Turbo(N, rate, 8)
        o = oracle.OracleTurbo(N, rate, 8, perm=g.perm, inv_perm=g.inv_perm)
        rs = np.random.RandomState(9)
        llr = np.stack([np.zeros(g.n_llr, np.float32),
                        rs.choice([-50.0, 50.0], g.n_llr).astype(np.float32),
                        (rs.randint(-3, 4, g.n_llr) * 1.0).astype(np.float32),
                        np.full(g.n_llr, 50.0, np.float32)])
        assert np.array_equal(g.decode_batch(llr), o.decode_batch(llr))
