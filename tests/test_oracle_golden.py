"""Pins oracle/ (the CPU restatement) against outputs of the REFERENCE ITSELF,
frozen in tests/golden/ by oracle/make_golden.py.  CPU only."""
import hashlib

import numpy as np
import pytest

from oracle import oracle
from tests import vectors


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def test_trellis_tables(golden):
    c = oracle.OracleTurbo(48, '1/3', 1)
    for k in ("next_state", "out_W", "out_Y", "prev_state", "prev_input", "G_matrix"):
        assert np.array_equal(getattr(c, k), golden["tables"][k]), k


@pytest.mark.parametrize("N", sorted(oracle.INTERLEAVER_PARAMS))
def test_interleaver_and_circular_lut(golden, N):
    t = golden["tables"]
    c = oracle.OracleTurbo(N, '1/3', 1)
    assert np.array_equal(c.perm, t[f"perm_{N}"])
    # inv_perm = host argsort with unstable tie order: same numpy here
    assert np.array_equal(c.inv_perm, t[f"inv_perm_{N}"])
    GN = oracle.mat_pow_gf2(c.G_matrix, N)
    assert np.array_equal(GN, t[f"Gpow_{N}"])
    lut = [oracle.solve_circular_state_gf2(GN, z) for z in range(16)]
    assert np.array_equal(lut, t[f"circ_lut_{N}"])
    for rate in oracle.PUNCTURE_PATTERNS:
        cc = oracle.OracleTurbo(N, rate, 1)
        assert cc.n_coded == int(t[f"n_coded_{N}_{rate.replace('/', '_')}"])


@pytest.mark.parametrize("case", vectors.CODEC_CASES, ids=lambda c: vectors.case_tag(*c[:3]))
def test_codec_kat(golden, case):
    N, rate, iters, nfr, ebn0s = case
    k = golden["codec_kat"]
    tag = vectors.case_tag(N, rate, iters)
    t = golden["tables"]
    c = oracle.OracleTurbo(N, rate, iters, perm=t[f"perm_{N}"], inv_perm=t[f"inv_perm_{N}"])
    info, llrs = vectors.codec_inputs(N, rate, nfr, ebn0s, c.encode, c.n_coded)
    assert sha(info.astype(np.uint8)) == str(k[f"{tag}/info_sha"])
    coded = c.encode_batch(info)
    assert np.array_equal(np.packbits(coded.astype(np.uint8), axis=1), k[f"{tag}/coded"])
    for e, llr in zip(ebn0s, llrs):
        assert sha(llr) == str(k[f"{tag}/ebn0_{e}/llr_sha"])
        dec = c.decode_batch(llr, threads=2)
        assert np.array_equal(np.packbits(dec.astype(np.uint8), axis=1), k[f"{tag}/ebn0_{e}/dec"])
    Lc = vectors.depuncture(llrs[0][0], N, c.punct)
    z = np.zeros(N)
    LeA, LeB = c.siso(Lc[0], Lc[1], Lc[2], Lc[3], z, z, 0.7)
    assert np.array_equal(LeA, k[f"{tag}/siso0_LeA"]) and np.array_equal(LeB, k[f"{tag}/siso0_LeB"])
    LaA, LaB = vectors.siso_apriori(N)
    LeA, LeB = c.siso(Lc[0], Lc[1], Lc[4], Lc[5], LaA, LaB, 1.0)
    assert np.array_equal(LeA, k[f"{tag}/siso1_LeA"]) and np.array_equal(LeB, k[f"{tag}/siso1_LeB"])


def test_decode_trace_consistent():
    c = oracle.OracleTurbo(48, '1/3', 3)
    info, llrs = vectors.codec_inputs(48, '1/3', 2, [2], c.encode, c.n_coded)
    dec, tr, lf = c.decode(llrs[0][0], trace=True)
    assert np.array_equal(dec, c.decode(llrs[0][0]))
    assert np.array_equal(dec[0::2], (lf[0] < 0).astype(np.int32))
    assert tr.shape == (3, 4, 48)


def test_short_llr_raises():
    c = oracle.OracleTurbo(212, '2/3', 1)          # n_coded bug: 630 < 636 consumed
    assert c.n_coded == 630 and c.n_emit == 636
    with pytest.raises(IndexError):
        c.decode(np.zeros(c.n_coded, np.float32))
    c.decode(np.zeros(c.n_emit, np.float32))


@pytest.mark.parametrize("name", list(vectors.BPS))
def test_mapper_and_slicer(golden, name):
    m = golden["modem_kat"]
    const = oracle.sdr_constellation(name)
    assert str(const.dtype) == str(m[f"{name}/const_dtype"])
    assert np.array_equal(const, m[f"{name}/const"])
    bits = vectors.mapper_bits(name)
    syms = oracle.modulate(bits, name)
    assert syms.dtype == m[f"{name}/syms"].dtype and np.array_equal(syms, m[f"{name}/syms"])
    rx = vectors.noisy_symbols(syms, name)
    assert sha(rx) == str(m[f"{name}/rx_sha"])
    hard = oracle.demodulate(rx, name)
    assert np.array_equal(hard, m[f"{name}/hard"])
    # hard round trip on clean symbols
    n = len(syms) * vectors.BPS[name]
    padded = np.append(bits, [0] * (n - len(bits)))
    assert np.array_equal(oracle.demodulate(syms, name), padded if name != 'BPSK' else bits)


@pytest.mark.parametrize("name", ['BPSK', 'QPSK', '8PSK', '16QAM'])
def test_compute_llr(golden, name):
    m = golden["modem_kat"]
    rx = vectors.noisy_symbols(oracle.modulate(vectors.mapper_bits(name), name), name)[:vectors.DEMAP_N]
    for nv in vectors.DEMAP_NOISE_VARS:
        for key, x in ((f"{name}/llr_nv{nv}", rx), (f"{name}/llr32_nv{nv}", rx.astype(np.complex64))):
            lit = oracle.compute_llr_literal(x, name, nv)
            vec = oracle.compute_llr(x, name, nv)
            assert np.array_equal(lit, m[key]), key
            assert np.array_equal(vec, m[key]), key
    # sign convention: positive LLR <=> bit 1
    bits = vectors.mapper_bits(name)
    clean = oracle.modulate(bits, name)
    llr = oracle.compute_llr(clean, name, 0.1)
    n = len(clean) * vectors.BPS[name]
    padded = np.append(bits, [0] * (n - len(bits)))
    assert np.array_equal((llr > 0).astype(int), padded)


@pytest.mark.parametrize("name", ['64QAM', '256QAM'])
def test_compute_llr_highorder_unpinned(name):
    """No reference soft demapper exists for 64/256QAM: literal == vectorised only."""
    rx = vectors.noisy_symbols(oracle.modulate(vectors.mapper_bits(name), name), name)[:64]
    assert np.array_equal(oracle.compute_llr_literal(rx, name, 0.01), oracle.compute_llr(rx, name, 0.01))


@pytest.mark.parametrize("name", ['BPSK', 'QPSK', '8PSK', '16QAM', '64QAM'])
def test_modulator_alt_tables(golden, name):
    m = golden["modem_kat"]
    bits = vectors.mapper_bits(name)
    syms = oracle.modulator_mod(bits, name)
    assert np.array_equal(syms, m[f"alt_{name}/syms"])
    rx = vectors.noisy_symbols(syms, name)
    assert np.array_equal(oracle.modulator_demod(rx.copy(), name), m[f"alt_{name}/hard"])
