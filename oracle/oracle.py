"""oracle/oracle.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the reference hot path, used only as the checker by
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` leg.  Nothing under ``modulations_b200/`` imports it.

* Turbo codec (reference ``dvb_rcs2_turbo.py``): thin ctypes wrapper over
  ``turbo_oracle.c`` (plain C, exact numba arithmetic, see its header).
* Mapper / hard slicer (reference ``sdr_modem.py:101-266``), alternate
  ``Modulator`` tables (``modulators.py:119-200``) and the max-log soft demapper
  ``compute_llr`` (``test_sdr_with_coding.py:200-225``): numpy restatements.

Parity status: PINNED against outputs of the reference itself, generated in the
authoring container by ``oracle/make_golden.py`` (committed under
``tests/golden/``).  The one exception is the 64QAM / 256QAM *soft* demapper:
the reference has no implementation of it (``MODULATIONS`` in
``test_sdr_with_coding.py:101-106`` stops at 16QAM), so for those two the oracle
is ``compute_llr``'s algorithm applied to ``SDRModem._qam64_mod/_qam256_mod``'s
constellations — "parity unpinned" by the reference, pinned only by that
construction.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

# dvb_rcs2_turbo.py:12-17
INTERLEAVER_PARAMS = {
    48: (31, 4, 2, 0, 3), 64: (41, 2, 6, 4, 1),
    212: (137, 0, 6, 4, 9), 220: (143, 4, 2, 8, 5),
    424: (277, 2, 4, 0, 7), 752: (491, 0, 8, 2, 5),
    848: (553, 4, 6, 0, 3),
}
# dvb_rcs2_turbo.py:21-26
PUNCTURE_PATTERNS = {
    '1/3': {'period': 1, 'W1': [1], 'Y1': [1], 'W2': [1], 'Y2': [1]},
    '1/2': {'period': 2, 'W1': [1, 0], 'Y1': [0, 1], 'W2': [1, 0], 'Y2': [0, 1]},
    '2/3': {'period': 3, 'W1': [1, 0, 0], 'Y1': [0, 1, 0], 'W2': [0, 0, 1], 'Y2': [0, 0, 0]},
    '3/4': {'period': 4, 'W1': [1, 0, 0, 0], 'Y1': [0, 1, 0, 0], 'W2': [0, 0, 1, 0], 'Y2': [0, 0, 0, 0]},
}


def build_lib(force: bool = False) -> str:
    """Compile turbo_oracle.c -> oracle/liboracle.so (gcc, a second or two)."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("turbo_oracle.c", "nii_model.c", "nii16_model.c", "Makefile")]
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"],
                              stdout=subprocess.DEVNULL)
    return so


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build_lib())
    return _LIB


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


_I32, _U8, _F32, _F64 = ctypes.c_int32, ctypes.c_uint8, ctypes.c_float, ctypes.c_double


class OracleTurbo:
    """Restatement of ``DVBRCS2_Turbo`` (dvb_rcs2_turbo.py:287-537)."""

    def __init__(self, N_couples, code_rate, iterations=8, perm=None, inv_perm=None):
        self.N = int(N_couples)
        self.k_info = self.N * 2
        self.iterations = int(iterations)
        self.punct = PUNCTURE_PATTERNS[code_rate]            # KeyError like :292
        if self.N not in INTERLEAVER_PARAMS:                 # :295-296
            raise ValueError(f"Block size {self.N} not in standard tables.")
        L = lib()
        self.next_state = np.zeros((16, 4), np.int32)
        self.out_W = np.zeros((16, 4), np.int32)
        self.out_Y = np.zeros((16, 4), np.int32)
        self.prev_state = np.zeros((16, 4), np.int32)
        self.prev_input = np.zeros((16, 4), np.int32)
        self.G_matrix = np.zeros((4, 4), np.int32)
        L.orc_build_trellis(_p(self.next_state, _I32), _p(self.out_W, _I32), _p(self.out_Y, _I32),
                            _p(self.prev_state, _I32), _p(self.prev_input, _I32),
                            _p(self.G_matrix, _I32))
        if perm is None:
            self.perm = np.zeros(self.N, np.int32)
            L.orc_interleaver(self.N, *INTERLEAVER_PARAMS[self.N], _p(self.perm, _I32))
            # :325 — host argsort; NOT an inverse (perm is not a bijection)
            self.inv_perm = np.argsort(self.perm).astype(np.int32)
        else:
            self.perm = np.ascontiguousarray(perm, np.int32)
            self.inv_perm = np.ascontiguousarray(inv_perm, np.int32)
        period = self.punct['period']
        self.period = period
        self.punct_u8 = np.array([self.punct[k] for k in ('W1', 'Y1', 'W2', 'Y2')], np.uint8)
        bpp = 2 * period + int(self.punct_u8.sum())
        self.n_coded = (self.N // period) * bpp              # :398-402 (keeps the N//period bug)
        self.n_emit = sum(2 + int(self.punct_u8[:, i % period].sum()) for i in range(self.N))

    # -- encoder (:431-462) -------------------------------------------------
    def encode(self, bits, return_circ=False):
        bits = np.ascontiguousarray(np.array(bits, dtype=np.int32))
        coded = np.zeros(self.n_emit, np.int32)
        circ = np.zeros(2, np.int32)
        n = lib().orc_encode(self.N, _p(self.next_state, _I32), _p(self.out_W, _I32),
                             _p(self.out_Y, _I32), _p(self.G_matrix, _I32), _p(self.perm, _I32),
                             _p(self.punct_u8, _U8), self.period, _p(bits, _I32),
                             _p(coded, _I32), _p(circ, _I32))
        assert n == self.n_emit
        return (coded, circ) if return_circ else coded

    def encode_batch(self, bits):
        bits = np.ascontiguousarray(bits, np.int32)
        B = bits.shape[0]
        coded = np.zeros((B, self.n_emit), np.int32)
        rc = lib().orc_encode_batch(B, self.N, _p(self.next_state, _I32), _p(self.out_W, _I32),
                                    _p(self.out_Y, _I32), _p(self.G_matrix, _I32),
                                    _p(self.perm, _I32), _p(self.punct_u8, _U8), self.period,
                                    _p(bits, _I32), _p(coded, _I32), self.n_emit)
        assert rc == 0
        return coded

    # -- SISO (:116-281) ------------------------------------------------------
    def siso(self, Lc_A, Lc_B, Lc_W, Lc_Y, La_A, La_B, scaling_factor):
        N = self.N
        f = [np.ascontiguousarray(x, np.float32) for x in (Lc_A, Lc_B, Lc_W, Lc_Y)]
        d = [np.ascontiguousarray(x, np.float64) for x in (La_A, La_B)]
        Le_A = np.zeros(N); Le_B = np.zeros(N)
        scratch = np.zeros(N * 64 + 2 * (N + 1) * 16, np.float32)
        fn = lib().orc_bcjr_max_log_map
        fn.argtypes = [ctypes.c_void_p] * 11 + [ctypes.c_int, ctypes.c_double] + [ctypes.c_void_p] * 3
        fn.restype = None
        fn(*(x.ctypes.data for x in f), *(x.ctypes.data for x in d),
           self.next_state.ctypes.data, self.out_W.ctypes.data, self.out_Y.ctypes.data,
           self.prev_state.ctypes.data, self.prev_input.ctypes.data, N, float(scaling_factor),
           Le_A.ctypes.data, Le_B.ctypes.data, scratch.ctypes.data)
        return Le_A, Le_B

    # -- decoder (:464-537) ---------------------------------------------------
    def decode(self, llr, trace=False):
        llr = np.ascontiguousarray(np.array(llr, dtype=np.float32))
        dec = np.zeros(self.k_info, np.int32)
        tr = np.zeros((self.iterations, 4, self.N)) if trace else None
        lf = np.zeros((2, self.N)) if trace else None
        rc = lib().orc_decode(self.N, self.iterations, _p(self.next_state, _I32),
                              _p(self.out_W, _I32), _p(self.out_Y, _I32),
                              _p(self.prev_state, _I32), _p(self.prev_input, _I32),
                              _p(self.perm, _I32), _p(self.inv_perm, _I32),
                              _p(self.punct_u8, _U8), self.period, _p(llr, _F32), int(llr.size),
                              _p(dec, _I32), _p(tr, _F64) if trace else None,
                              _p(lf, _F64) if trace else None)
        if rc != 0:
            raise IndexError("llr shorter than the depuncturer consumes (reference: IndexError)")
        return (dec, tr, lf) if trace else dec

    def decode_batch(self, llr, threads=1):
        llr = np.ascontiguousarray(llr, np.float32)
        B, n = llr.shape
        dec = np.zeros((B, self.k_info), np.int32)
        fn = lib().orc_decode_batch

        def run(lo, hi):
            if hi <= lo:
                return 0
            return fn(int(hi - lo), self.N, self.iterations, _p(self.next_state, _I32),
                      _p(self.out_W, _I32), _p(self.out_Y, _I32), _p(self.prev_state, _I32),
                      _p(self.prev_input, _I32), _p(self.perm, _I32), _p(self.inv_perm, _I32),
                      _p(self.punct_u8, _U8), self.period, _p(llr[lo:hi], _F32), int(n),
                      _p(dec[lo:hi], _I32))
        if threads <= 1:
            rc = run(0, B)
        else:
            # ctypes drops the GIL during the foreign call, so these run in parallel
            cuts = np.linspace(0, B, threads + 1).astype(int)
            with ThreadPoolExecutor(threads) as ex:
                rc = min(ex.map(lambda i: run(int(cuts[i]), int(cuts[i + 1])), range(threads)))
        if rc != 0:
            raise IndexError("llr shorter than the depuncturer consumes")
        return dec


class NiiModel(OracleTurbo):
    """Model of the NON-PARITY decoder mode "nii" (oracle/nii_model.c): same tables and encoder as the reference
    codec, single-pass float32 SISOs with next-iteration initialisation.  Not a restatement of the reference."""

    def decode_batch(self, llr, threads=1, sf_inner=0.7, sf_last=1.0):
        llr = np.ascontiguousarray(llr, np.float32)
        B, n = llr.shape
        dec = np.zeros((B, self.k_info), np.int32)
        fn = lib().nii_decode_batch
        fn.restype = ctypes.c_int
        fn.argtypes = [ctypes.c_int] * 3 + [ctypes.c_void_p] * 8 + [ctypes.c_int, ctypes.c_float, ctypes.c_float,
                                                                   ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]

        def run(lo, hi):
            if hi <= lo:
                return 0
            return fn(int(hi - lo), self.N, self.iterations, self.next_state.ctypes.data, self.out_W.ctypes.data,
                      self.out_Y.ctypes.data, self.prev_state.ctypes.data, self.prev_input.ctypes.data,
                      self.perm.ctypes.data, self.inv_perm.ctypes.data, self.punct_u8.ctypes.data, self.period,
                      float(sf_inner), float(sf_last), llr[lo:hi].ctypes.data, int(n), dec[lo:hi].ctypes.data)
        if threads <= 1:
            rc = run(0, B)
        else:
            cuts = np.linspace(0, B, threads + 1).astype(int)
            with ThreadPoolExecutor(threads) as ex:
                rc = min(ex.map(lambda i: run(int(cuts[i]), int(cuts[i + 1])), range(threads)))
        if rc != 0:
            raise IndexError("llr shorter than the depuncturer consumes")
        return dec

    def decode(self, llr):
        return self.decode_batch(np.asarray(llr, np.float32)[None, :])[0]


class Nii16Model(OracleTurbo):
    """Model of the NON-PARITY fixed-point decoder mode "nii16" (oracle/nii16_model.c).  Not the reference."""

    def decode_batch(self, llr, threads=1, sf_inner_q=45, sf_last_q=64):
        llr = np.ascontiguousarray(llr, np.float32)
        B, n = llr.shape
        dec = np.zeros((B, self.k_info), np.int32)
        fn = lib().nii16_decode_batch
        fn.restype = ctypes.c_int
        fn.argtypes = [ctypes.c_int] * 3 + [ctypes.c_void_p] * 8 + [ctypes.c_int] * 3 + [ctypes.c_void_p, ctypes.c_int,
                                                                                       ctypes.c_void_p]

        def run(lo, hi):
            if hi <= lo:
                return 0
            return fn(int(hi - lo), self.N, self.iterations, self.next_state.ctypes.data, self.out_W.ctypes.data,
                      self.out_Y.ctypes.data, self.prev_state.ctypes.data, self.prev_input.ctypes.data,
                      self.perm.ctypes.data, self.inv_perm.ctypes.data, self.punct_u8.ctypes.data, self.period,
                      int(sf_inner_q), int(sf_last_q), llr[lo:hi].ctypes.data, int(n), dec[lo:hi].ctypes.data)
        if threads <= 1:
            rc = run(0, B)
        else:
            # the model's overflow flag is a file-scope variable: worker threads only ever SET it, which is benign
            cuts = np.linspace(0, B, threads + 1).astype(int)
            with ThreadPoolExecutor(threads) as ex:
                rc = min(ex.map(lambda i: run(int(cuts[i]), int(cuts[i + 1])), range(threads)))
        if rc == -1:
            raise IndexError("llr shorter than the depuncturer consumes")
        if rc == -2:
            raise OverflowError("a 16-bit quantity of the nii16 format overflowed")
        return dec

    def decode(self, llr):
        return self.decode_batch(np.asarray(llr, np.float32)[None, :])[0]


def solve_circular_state_gf2(G_pow_N, Z_N):
    """dvb_rcs2_turbo.py:63-114."""
    G = np.ascontiguousarray(G_pow_N, np.int32)
    return int(lib().orc_solve_circular_state_gf2(_p(G, _I32), int(Z_N)))


def mat_pow_gf2(A, power):
    """dvb_rcs2_turbo.py:50-61."""
    A = np.ascontiguousarray(A, np.int32)
    R = np.zeros((4, 4), np.int32)
    lib().orc_mat_pow_gf2(_p(A, _I32), ctypes.c_longlong(int(power)), _p(R, _I32))
    return R


# =============================================================================
# Mapper / slicer — sdr_modem.py:66-73, 101-266 (SDRModem)
# =============================================================================
GRAY2 = [0, 1, 3, 2]
GRAY3 = [0, 1, 3, 2, 6, 7, 5, 4]
GRAY4 = [0, 1, 3, 2, 6, 7, 5, 4, 12, 13, 15, 14, 10, 11, 9, 8]
BPS = {'BPSK': 1, 'QPSK': 2, '8PSK': 3, '16QAM': 4, '64QAM': 6, '256QAM': 8}


def _label_bits(bps):
    """Rows = labels 0..2^bps-1, MSB first (test_sdr_with_coding.py:207)."""
    return np.array([list(map(int, format(i, f'0{bps}b'))) for i in range(1 << bps)])


def sdr_constellation(mod):
    """Constellation indexed by MSB-first label, built with the same scalar
    expressions and dtypes as SDRModem._*_mod (sdr_modem.py:101-207): complex64
    everywhere except QPSK, which the reference returns as complex128."""
    if mod == 'BPSK':                                   # :101-102
        return (2.0 * np.array([0, 1], dtype=np.complex64) - 1.0)
    if mod == 'QPSK':                                   # :107-112
        b = _label_bits(2)
        I = 1 - 2 * b[:, 0]
        Q = 1 - 2 * b[:, 1]
        return (I + 1j * Q).astype(np.complex64) / np.sqrt(2)
    if mod == '8PSK':                                   # :120-131
        return np.array([np.exp(1j * (GRAY3[i] * np.pi / 4)) for i in range(8)], dtype=np.complex64)
    if mod in ('16QAM', '64QAM', '256QAM'):            # :142-154, :168-181, :195-207
        gray, half, norm = {'16QAM': (GRAY2, 2, 10), '64QAM': (GRAY3, 3, 42),
                            '256QAM': (GRAY4, 4, 170)}[mod]
        L = 1 << half
        pts = []
        for lab in range(L * L):
            i_idx, q_idx = lab >> half, lab & (L - 1)
            I = (2 * gray[i_idx] - (L - 1)) / np.sqrt(norm)
            Q = (2 * gray[q_idx] - (L - 1)) / np.sqrt(norm)
            pts.append(I + 1j * Q)
        return np.array(pts, dtype=np.complex64)
    raise ValueError(f"Unknown modulation: {mod}")


def modulate(bits, modulation='QPSK'):
    """SDRModem.modulate (sdr_modem.py:222-243): zero-pad to a multiple of bps,
    MSB-first label -> constellation point."""
    if modulation not in BPS:
        raise ValueError(f"Unknown modulation: {modulation}")
    bps = BPS[modulation]
    bits = np.array(bits)
    pad = (bps - len(bits) % bps) % bps
    if pad:
        bits = np.append(bits, [0] * pad)
    lab = bits.reshape(-1, bps).astype(np.int64) @ (1 << np.arange(bps)[::-1])
    return sdr_constellation(modulation)[lab]


def demodulate(symbols, modulation='QPSK'):
    """SDRModem.demodulate hard slicers (sdr_modem.py:104-266)."""
    s = np.asarray(symbols)
    if modulation == 'BPSK':                            # :104-105
        return (np.real(s) > 0).astype(int)
    if modulation == 'QPSK':                            # :114-118
        bits = np.zeros(len(s) * 2, dtype=int)
        bits[0::2] = (np.real(s) < 0).astype(int)
        bits[1::2] = (np.imag(s) < 0).astype(int)
        return bits
    if modulation == '8PSK':                            # :133-140
        inv = [GRAY3.index(i) for i in range(8)]
        out = []
        for v in s:
            ph = np.angle(v)
            if ph < 0:
                ph += 2 * np.pi
            o = inv[int(np.round(ph / (np.pi / 4))) % 8]
            out.extend([(o >> 2) & 1, (o >> 1) & 1, o & 1])
        return np.array(out)
    if modulation in ('16QAM', '64QAM', '256QAM'):     # :156-166, :183-193, :209-220
        gray, half, norm = {'16QAM': (GRAY2, 2, 10), '64QAM': (GRAY3, 3, 42),
                            '256QAM': (GRAY4, 4, 170)}[modulation]
        L = 1 << half
        inv = [gray.index(i) for i in range(L)]
        out = []
        for v in s:
            I = np.real(v) * np.sqrt(norm)
            Q = np.imag(v) * np.sqrt(norm)
            ii = int(np.clip(np.round((I + (L - 1)) / 2), 0, L - 1))
            qi = int(np.clip(np.round((Q + (L - 1)) / 2), 0, L - 1))
            for o in (inv[ii], inv[qi]):
                out.extend([(o >> j) & 1 for j in range(half - 1, -1, -1)])
        return np.array(out)
    raise ValueError(f"Unknown modulation: {modulation}")


# =============================================================================
# Alternate (non-Gray) tables — modulators.py:119-200 (Modulator)
# =============================================================================
def modulator_constellation(mod):
    """Label (MSB-first) -> point, as Modulator.mod_* (modulators.py:119-200)."""
    if mod == 'BPSK':                                   # :119-120
        return (2 * np.array([0, 1]) - 1).astype(np.complex64)
    if mod == 'QPSK':                                   # :125-131
        b = _label_bits(2)
        return ((1 - 2 * b[:, 0]) + 1j * (1 - 2 * b[:, 1])) / np.sqrt(2)
    if mod == '8PSK':                                   # :139-145
        return np.exp(1j * 2 * np.pi * np.arange(8) / 8)
    if mod in ('16QAM', '64QAM'):                       # :157-163
        M = int(mod[:-3])
        m = int(np.sqrt(M))
        axis = np.arange(-m + 1, m, 2)
        xv, yv = np.meshgrid(axis, axis)
        c = xv.flatten() + 1j * yv.flatten()
        c /= np.sqrt(np.mean(np.abs(c) ** 2))
        return c
    raise ValueError(f"Unknown modulation: {mod}")


def modulator_mod(bits, mod):
    bps = BPS[mod]
    bits = np.asarray(bits, int)
    pad = (bps - len(bits) % bps) % bps
    if pad:
        bits = np.append(bits, [0] * pad)
    dec = bits.reshape(-1, bps).dot(1 << np.arange(bps)[::-1])
    return modulator_constellation(mod)[dec]


def modulator_demod(symbols, mod):
    """modulators.py:121-171 hard demod."""
    s = np.asarray(symbols)
    if mod == 'BPSK':
        return (np.real(s) > 0).astype(int)
    if mod == 'QPSK':
        return np.column_stack([(np.real(s) < 0).astype(int), (np.imag(s) < 0).astype(int)]).flatten()
    if mod == '8PSK':
        phi = np.angle(s)
        phi[phi < 0] += 2 * np.pi
        dec = np.round(phi / (np.pi / 4)).astype(int) % 8
        return np.array([[(d >> 2) & 1, (d >> 1) & 1, d & 1] for d in dec]).flatten()
    c = modulator_constellation(mod)
    idxs = np.argmin(np.abs(s[:, None] - c[None, :]), axis=1)     # :165-171
    k = BPS[mod]
    return np.array([[(i >> j) & 1 for j in range(k - 1, -1, -1)] for i in idxs]).flatten()


# =============================================================================
# Soft demapper — test_sdr_with_coding.py:200-225 (compute_llr)
# =============================================================================
def compute_llr_literal(syms, mod_type, noise_var, constellation=None):
    """Line-for-line restatement (per-symbol Python loop; small inputs only)."""
    noise_var = max(noise_var, 0.005)                   # :202
    bps = BPS[mod_type]
    order = 1 << bps
    all_bits = _label_bits(bps)                         # :207
    if constellation is None:
        constellation = modulate(all_bits.flatten(), mod_type).reshape(-1)   # :208
    n_syms = len(syms)
    llr = np.zeros(n_syms * bps)
    for i, s in enumerate(syms):                        # :213-223
        distances = np.abs(s - constellation) ** 2
        for b in range(bps):
            idx0 = np.where(all_bits[:, b] == 0)[0]
            idx1 = np.where(all_bits[:, b] == 1)[0]
            llr[i * bps + b] = (np.min(distances[idx0]) - np.min(distances[idx1])) / noise_var
    assert order == len(constellation)
    return np.clip(llr, -30, 30)                        # :225


def compute_llr(syms, mod_type, noise_var, constellation=None, chunk=1 << 16):
    """Same arithmetic as compute_llr_literal, vectorised over symbols (checked
    equal to the literal loop in tests/test_oracle_golden.py)."""
    noise_var = max(noise_var, 0.005)
    bps = BPS[mod_type]
    all_bits = _label_bits(bps)
    if constellation is None:
        constellation = modulate(all_bits.flatten(), mod_type).reshape(-1)
    syms = np.asarray(syms)
    out = np.zeros((len(syms), bps))
    sets = [(np.where(all_bits[:, b] == 0)[0], np.where(all_bits[:, b] == 1)[0]) for b in range(bps)]
    for lo in range(0, len(syms), chunk):
        s = syms[lo:lo + chunk]
        d = np.abs(s[:, None] - constellation[None, :]) ** 2
        for b, (i0, i1) in enumerate(sets):
            out[lo:lo + chunk, b] = (d[:, i0].min(axis=1) - d[:, i1].min(axis=1)) / noise_var
    return np.clip(out.reshape(-1), -30, 30)


# =============================================================================
# Waveform stage — modulators.py:19-117 (rrcosfilter, apply_pulse_shaping, matched_filter)
# =============================================================================
def rrcosfilter(N, alpha, Ts, Fs):
    """Root-raised-cosine taps, unit energy (modulators.py:19-48): odd tap count int(N*Fs)|1, the two
    singular points handled as the reference does, denominator clamped at 1e-10."""
    num_taps = int(N * Fs) | 1
    t = (np.arange(num_taps) - (num_taps - 1) / 2) * (1.0 / float(Fs))
    h = np.zeros(num_taps, dtype=float)
    for x in range(num_taps):
        tt = t[x]
        if tt == 0.0:
            h[x] = 1.0 - alpha + (4 * alpha / np.pi)
        elif alpha != 0 and abs(tt) == Ts / (4 * alpha):
            h[x] = (alpha / np.sqrt(2)) * (((1 + 2 / np.pi) * (np.sin(np.pi / (4 * alpha))))
                                           + ((1 - 2 / np.pi) * (np.cos(np.pi / (4 * alpha)))))
        else:
            denom = (1 - (4 * alpha * tt / Ts) ** 2)
            if abs(denom) < 1e-10:
                denom = 1e-10
            num = (np.sin(np.pi * tt / Ts * (1 - alpha)) + 4 * alpha * tt / Ts * np.cos(np.pi * tt / Ts * (1 + alpha)))
            h[x] = num / (np.pi * tt / Ts * denom)
    return h / np.sqrt(np.sum(h ** 2))


def pulse_shape(symbols, h, sps):
    """apply_pulse_shaping (modulators.py:85-100) = scipy.signal.upfirdn(h, syms, up=sps, down=1):
    zero-stuff by sps, convolve 'full', keep (n-1)*sps + len(h) samples.  complex64 in, complex128 out."""
    syms = np.array(symbols, dtype=np.complex64)
    if len(syms) == 0:
        return np.zeros(0, np.complex128)
    up = np.zeros((len(syms) - 1) * sps + 1, dtype=np.complex128)
    up[::sps] = syms
    return np.convolve(up, np.asarray(h, float), mode='full')


def matched_filter(samples, h, sps):
    """matched_filter (modulators.py:102-117): convolve(samples, h, 'full')[2*delay::sps] with
    delay = (len(h)-1)//2; an empty complex64 array when the start index is past the end."""
    samples = np.asarray(samples)
    filtered = np.convolve(samples, np.asarray(h, float), mode='full') if len(samples) else np.zeros(0, np.complex128)
    start = 2 * ((len(h) - 1) // 2)
    if start >= len(filtered):
        return np.array([], dtype=np.complex64)
    return filtered[start::sps]
