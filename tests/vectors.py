"""Seeded synthetic inputs shared by oracle/make_golden.py, the tests and bench.py.

Everything is drawn from ``np.random.RandomState`` (legacy MT19937: the stream
is stable across numpy versions), following the reference's own Monte-Carlo
shape: BPSK 0->+1 over AWGN, ``sigma^2 = 1/(2 R Eb/N0)``, ``llr = 2y/sigma^2``
clipped to +-50 (turbo_test_suite.py:132-161).
"""
import numpy as np

RATE = {'1/3': 1 / 3, '1/2': 1 / 2, '2/3': 2 / 3, '3/4': 3 / 4}

# (N couples, rate, iterations, frames, Eb/N0 list in dB)
CODEC_CASES = [
    (48, '1/3', 8, 24, [0, 3, 6]),
    (48, '1/2', 8, 16, [2]),
    (48, '2/3', 8, 8, [4]),
    (48, '3/4', 8, 8, [4]),
    (64, '1/3', 8, 8, [2]),
    (212, '1/3', 8, 24, [0, 2, 4]),
    (212, '1/3', 1, 4, [2]),
    (212, '1/2', 8, 12, [3]),
    (220, '1/3', 4, 4, [2]),
    (424, '1/3', 8, 4, [2]),
    (752, '1/2', 8, 4, [1.5]),
    (848, '1/3', 2, 2, [2]),
]

DEMAP_NOISE_VARS = [0.001, 0.05, 0.5]
DEMAP_N = 384
BPS = {'BPSK': 1, 'QPSK': 2, '8PSK': 3, '16QAM': 4, '64QAM': 6, '256QAM': 8}


def case_tag(N, rate, iters):
    return f"N{N}_R{rate.replace('/', '_')}_it{iters}"


def awgn_llr(rs, coded, rate, ebn0_db):
    """turbo_test_suite.py:132-161 for one frame."""
    noise_var = 1.0 / (2.0 * RATE[rate] * 10 ** (ebn0_db / 10))
    tx = 1.0 - 2.0 * coded.astype(float)
    rx = tx + np.sqrt(noise_var) * rs.randn(len(coded))
    return np.clip(2.0 * rx / noise_var, -50, 50).astype(np.float32)


def codec_inputs(N, rate, nfr, ebn0s, encode, n_coded):
    """info bits [nfr, 2N] and, per Eb/N0, float32 LLRs [nfr, n_coded]."""
    rs = np.random.RandomState(1234 + N)
    info = rs.randint(0, 2, (nfr, 2 * N))
    out = []
    for e in ebn0s:
        llr = np.zeros((nfr, n_coded), np.float32)
        for i in range(nfr):
            coded = np.asarray(encode(info[i]))
            assert len(coded) == n_coded, (len(coded), n_coded)
            llr[i] = awgn_llr(rs, coded, rate, e)
        out.append(llr)
    return info, out


def depuncture(llr, N, punct):
    """dvb_rcs2_turbo.py:466-487 -> (Lc_A, Lc_B, Lc_W1, Lc_Y1, Lc_W2, Lc_Y2)."""
    llr = np.array(llr, dtype=np.float32)
    L = [np.zeros(N, np.float32) for _ in range(6)]
    idx = 0
    period = punct['period']
    for i in range(N):
        p = i % period
        L[0][i] = llr[idx]; idx += 1
        L[1][i] = llr[idx]; idx += 1
        for j, key in enumerate(('W1', 'Y1', 'W2', 'Y2')):
            if punct[key][p]:
                L[2 + j][i] = llr[idx]; idx += 1
    return L


def siso_apriori(N):
    rs = np.random.RandomState(77 + N)
    return rs.randn(N) * 3.0, rs.randn(N) * 3.0


def mapper_bits(name):
    rs = np.random.RandomState(4242 + BPS[name])
    n = 2048 * BPS[name] - (1 if BPS[name] > 1 else 0)      # forces the zero-pad branch
    return rs.randint(0, 2, n)


def noisy_symbols(syms, name):
    """complex128 symbols + complex AWGN, sigma picked so a few hard errors occur."""
    rs = np.random.RandomState(999 + BPS[name])
    sigma = {'BPSK': 0.5, 'QPSK': 0.4, '8PSK': 0.2, '16QAM': 0.15, '64QAM': 0.07, '256QAM': 0.03}[name]
    return np.asarray(syms).astype(np.complex128) + sigma * (rs.randn(len(syms)) + 1j * rs.randn(len(syms)))


# ---- waveform stage (modulators.py:19-117): RRC pulse shaping and matched filter -------------------
WAVEFORM_CASES = [(8, 0.35, 6), (4, 0.25, 8), (5, 0.5, 3)]       # (samples per symbol, roll-off, span in symbols)
WAVEFORM_NSYM = 1000


def waveform_symbols(sps):
    """Seeded unit-power QPSK-like complex64 symbols."""
    rs = np.random.RandomState(7000 + sps)
    s = (rs.randint(0, 2, WAVEFORM_NSYM) * 2 - 1) + 1j * (rs.randint(0, 2, WAVEFORM_NSYM) * 2 - 1)
    return (s / np.sqrt(2)).astype(np.complex64)


def waveform_noise(shaped, sps):
    """The received samples fed to the matched filter: the shaped waveform + seeded complex noise, complex64."""
    rs = np.random.RandomState(7100 + sps)
    n = 0.05 * (rs.randn(len(shaped)) + 1j * rs.randn(len(shaped)))
    return (np.asarray(shaped) + n).astype(np.complex64)
