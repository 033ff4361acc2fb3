"""GPU parity tests for mapper / slicer / soft demapper vs oracle/ and golden."""
import numpy as np
import pytest

from oracle import oracle
from tests import vectors

pytestmark = pytest.mark.gpu

# float32 kernel vs the reference's float64.  Tolerance = 2 x the largest error OBSERVED on a B200 per modulation, over the
# fixture symbols and 200 000 random ones, noise variances 1e-4 ... 0.5 (profiles/r02_demap_observed_error.txt, made by
# tools/demap_err.py); every entry is below the 1e-4 that BASELINE.md section 4 states for +-30-clipped LLRs.
LLR_ATOL = {'BPSK': 9e-5, 'QPSK': 1e-5, '8PSK': 3.2e-5, '16QAM': 2.7e-5, '64QAM': 2.2e-5, '256QAM': 2.4e-5}


def llr_atol(name):
    return LLR_ATOL[name]


@pytest.mark.parametrize("name", list(vectors.BPS))
def test_mapper_and_slicer(golden, name):
    from modulations_b200.sdr_modem import SDRModem
    m = golden["modem_kat"]
    sm = SDRModem()
    bits = vectors.mapper_bits(name)
    syms = sm.modulate(bits, name)
    assert syms.dtype == m[f"{name}/syms"].dtype
    assert np.array_equal(syms, m[f"{name}/syms"])                       # bit-exact mapping
    rx = vectors.noisy_symbols(syms, name)
    hard = sm.demodulate(rx, name)
    assert np.array_equal(hard, m[f"{name}/hard"])
    assert np.array_equal(sm.demodulate(rx.astype(np.complex64), name),
                          oracle.demodulate(rx.astype(np.complex64), name))


@pytest.mark.parametrize("name", ['BPSK', 'QPSK', '8PSK', '16QAM', '64QAM'])
def test_modulator_alt_tables(golden, name):
    from modulations_b200.modulators import Modulator
    m = golden["modem_kat"]
    mo = Modulator()
    fn = getattr(mo, "mod_" + name.lower())
    dfn = getattr(mo, "demod_" + name.lower())
    bits = vectors.mapper_bits(name)
    syms = fn(bits)
    assert np.array_equal(syms, m[f"alt_{name}/syms"])
    rx = vectors.noisy_symbols(syms, name)
    assert np.array_equal(dfn(rx), m[f"alt_{name}/hard"])


@pytest.mark.parametrize("name", list(vectors.BPS))
def test_compute_llr(golden, name):
    from modulations_b200.soft_demod import compute_llr, decoder_llr
    m = golden["modem_kat"]
    rx = vectors.noisy_symbols(oracle.modulate(vectors.mapper_bits(name), name), name)
    for nv in vectors.DEMAP_NOISE_VARS + [0.0001]:
        got = compute_llr(rx, name, nv)
        assert got.dtype == np.float64 and got.shape == (len(rx) * vectors.BPS[name],)
        want = oracle.compute_llr(rx, name, nv)
        key = f"{name}/llr_nv{nv}"
        if key in m.files:                                               # reference's own output
            assert np.array_equal(want[:vectors.DEMAP_N * vectors.BPS[name]], m[key])
        err = np.abs(got - want)
        assert err.max() <= llr_atol(name), (name, nv, "max |dLLR| = %.3e" % err.max())
        # identical hard decisions except on near-zero |LLR| ties
        bad = (got > 0) != (want > 0)
        assert np.all(np.abs(want[bad]) < llr_atol(name))
        assert np.array_equal(decoder_llr(rx, name, nv), -got.astype(np.float32))
    assert np.all(np.abs(compute_llr(rx, name, 0.001)) <= 30.0)


@pytest.mark.parametrize("name", ['QPSK', '16QAM', '64QAM', '256QAM'])
def test_generic_and_separable_kernels_agree(name):
    """The piecewise-linear per-axis kernel and the brute-force table kernel are two
    implementations of the same max-log rule."""
    from modulations_b200.sdr_modem import ModemHandle, gray_constellation
    rx = vectors.noisy_symbols(oracle.modulate(vectors.mapper_bits(name), name), name)
    table = gray_constellation(name)
    fast = ModemHandle(name, table)
    # rotating the table by a hair breaks separability -> generic kernel
    rot = np.exp(1j * 1e-9)
    slow = ModemHandle(name, table.astype(np.complex128) * rot)
    a, b = fast.llr(rx, 0.05), slow.llr(rx, 0.05)
    assert np.abs(a - b).max() < 2e-4


def test_modulator_llr_nongray_table():
    """compute_llr over a caller-supplied (non-Gray) table."""
    from modulations_b200.modulators import natural_constellation
    from modulations_b200.soft_demod import compute_llr
    c = natural_constellation('16QAM')
    rs = np.random.RandomState(2)
    bits = rs.randint(0, 2, 4 * 500)
    rx = oracle.modulator_mod(bits, '16QAM') + 0.1 * (rs.randn(500) + 1j * rs.randn(500))
    got = compute_llr(rx, '16QAM', 0.02, constellation=c)
    want = oracle.compute_llr(rx, '16QAM', 0.02, constellation=c)
    assert np.abs(got - want).max() <= llr_atol('16QAM')


def test_coded_16qam_pipeline(golden):
    """Config 3 shape: encode -> 16QAM map -> AWGN -> demap (sign flipped, F4) ->
    pad/trim to n_coded (test_sdr_with_coding.py:474-478) -> decode; the GPU chain
    must give the oracle chain's bits when fed the oracle's float32 LLRs, and its
    own LLRs must agree within tolerance."""
    from modulations_b200 import dvb_rcs2_turbo as turbo
    from modulations_b200.sdr_modem import SDRModem
    from modulations_b200.soft_demod import decoder_llr
    N, rate = 212, '1/2'
    g = turbo.DVBRCS2_Turbo(N, rate, 8)
    o = oracle.OracleTurbo(N, rate, 8, perm=g.perm, inv_perm=g.inv_perm)
    rs = np.random.RandomState(8)
    info = rs.randint(0, 2, 2 * N)
    coded = g.encode(info)
    syms = SDRModem().modulate(coded, '16QAM')
    assert np.array_equal(syms, oracle.modulate(coded, '16QAM'))
    nv = 1.0 / (4 * 0.5 * 10 ** 0.6)
    rx = syms + np.sqrt(nv / 2) * (rs.randn(len(syms)) + 1j * rs.randn(len(syms)))
    llr_o = (-oracle.compute_llr(rx, '16QAM', nv))[:g.n_coded].astype(np.float32)
    llr_g = decoder_llr(rx, '16QAM', nv)[:g.n_coded]
    assert np.abs(llr_g - llr_o).max() <= llr_atol('16QAM')
    assert np.array_equal(g.decode(llr_o), o.decode(llr_o))


@pytest.mark.parametrize("name", list(vectors.BPS))
def test_bf16_symbol_input(name):
    """bf16x2 I/Q input (4 bytes per symbol): every component is widened exactly to float32 on the device, so the
    result equals the float32-input demapper on the widened symbols bit for bit, and the oracle on them within the
    modulation's tolerance."""
    import torch
    from modulations_b200.sdr_modem import gray_modem
    rs = np.random.RandomState(77)
    n = 100_003
    tx = oracle.modulate(rs.randint(0, 2, vectors.BPS[name] * n), name)
    rx = tx + 0.1 * (rs.randn(n) + 1j * rs.randn(n))
    xb = torch.from_numpy(np.stack([rx.real, rx.imag], axis=1).astype(np.float32)).to(torch.bfloat16).cuda()
    wide = xb.float().contiguous()                                  # exact widening
    m = gray_modem(name)
    for nv, scale in ((0.05, 1.0), (0.5, -1.0)):
        got = m.llr_bf16(xb, nv, scale)
        same = m.llr(torch.view_as_complex(wide), nv, scale)
        assert got.dtype == torch.float32 and got.shape == same.shape
        assert torch.equal(got, same)
        wc = wide.cpu().numpy()
        want = scale * oracle.compute_llr(wc[:, 0].astype(np.float64) + 1j * wc[:, 1].astype(np.float64), name, nv)
        assert np.abs(got.cpu().numpy() - want).max() <= llr_atol(name)
    with pytest.raises(ValueError):
        m.llr_bf16(wide, 0.05)


@pytest.mark.parametrize("name", list(vectors.BPS))
@pytest.mark.parametrize("nsym", [1, 2, 3, 5, 31, 1000, 4097, 100_003])
def test_mapper_word_load_kernel(name, nsym):
    """complex64 mapping goes through the word-load kernel (modem.cu map_words_kernel: a symbol's bit-bytes fetched as
    aligned words, label gathered by one multiply, 4 symbols per thread in flight): every symbol count down to 1,
    including the 3- and 6-bit orders whose last symbols must not read past the array; bit-exact against the
    reference mapping, and identical to the generic kernel (complex128 output, rounded)."""
    import torch
    from modulations_b200.sdr_modem import gray_modem
    bps = vectors.BPS[name]
    rng = np.random.RandomState(nsym * 7 + bps)
    bits = rng.randint(0, 2, nsym * bps).astype(np.uint8)
    g = gray_modem(name)
    got = g.map(bits, out_complex128=False)
    assert got.dtype == np.complex64 and got.shape == (nsym,)
    want = np.asarray(oracle.modulate(bits, name)).astype(np.complex64)
    assert np.array_equal(got, want)
    assert np.array_equal(got, g.map(bits, out_complex128=True).astype(np.complex64))
    # bit values other than 0 / 1 count by their LSB in both kernels
    odd = (bits + 2 * rng.randint(0, 100, bits.size)).astype(np.uint8)
    assert np.array_equal(g.map(odd, out_complex128=False), want)
    # an unaligned device view takes the generic kernel: same symbols
    if nsym >= 31:
        t = torch.zeros(bits.size + 1, dtype=torch.uint8, device="cuda")
        t[1:] = torch.from_numpy(bits).cuda()
        assert np.array_equal(g.map(t[1:], out_complex128=False).cpu().numpy(), want)
