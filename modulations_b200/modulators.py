"""Drop-in for the reference ``modulators.Modulator``: the natural-binary (non-Gray)
BPSK/QPSK/8PSK/16/64QAM mapping (``modulators.py:119-200`` — same kernels as ``SDRModem``,
different constellation tables) and the waveform stage next to it (``modulators.py:19-117``:
``rrcosfilter``, ``apply_pulse_shaping``, ``matched_filter`` — SURVEY 8(f) N4), as streaming
FIR kernels behind ``b200dvb_pulse_shape`` / ``b200dvb_matched_filter``.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .sdr_modem import ModemHandle


def rrcosfilter(N, alpha, Ts, Fs):
    """Root-raised-cosine taps with unit energy (``modulators.py:19-48``): ``int(N*Fs) | 1`` taps,
    the singular points t = 0 and abs(t) = Ts/(4 alpha) in closed form, denominator clamped at 1e-10.
    Host-side table generation (49 numbers); the FIRs that use the table run on the GPU."""
    num_taps = int(N * Fs) | 1
    t = (np.arange(num_taps) - (num_taps - 1) / 2) * (1.0 / float(Fs))
    h = np.zeros(num_taps, dtype=float)
    for i, tt in enumerate(t):
        if tt == 0.0:
            h[i] = 1.0 - alpha + (4 * alpha / np.pi)
        elif alpha != 0 and abs(tt) == Ts / (4 * alpha):
            h[i] = (alpha / np.sqrt(2)) * (((1 + 2 / np.pi) * (np.sin(np.pi / (4 * alpha))))
                                           + ((1 - 2 / np.pi) * (np.cos(np.pi / (4 * alpha)))))
        else:
            denom = (1 - (4 * alpha * tt / Ts) ** 2)
            if abs(denom) < 1e-10:
                denom = 1e-10
            num = np.sin(np.pi * tt / Ts * (1 - alpha)) + 4 * alpha * tt / Ts * np.cos(np.pi * tt / Ts * (1 + alpha))
            h[i] = num / (np.pi * tt / Ts * denom)
    return h / np.sqrt(np.sum(h ** 2))


def natural_constellation(modulation):
    """Label (MSB-first) -> point, as Modulator.mod_* builds it."""
    if modulation == 'BPSK':                                 # :119-120
        return np.array([-1, 1]).astype(np.complex64)
    if modulation == 'QPSK':                                 # :125-131
        return np.array([1 + 1j, 1 - 1j, -1 + 1j, -1 - 1j]) / np.sqrt(2)
    if modulation == '8PSK':                                 # :139-145
        return np.exp(1j * 2 * np.pi * np.arange(8) / 8)
    if modulation in ('16QAM', '64QAM'):                     # _qam_const :157-163
        m = int(np.sqrt(int(modulation[:-3])))
        axis = np.arange(-m + 1, m, 2)
        xv, yv = np.meshgrid(axis, axis)
        c = xv.flatten() + 1j * yv.flatten()
        c /= np.sqrt(np.mean(np.abs(c) ** 2))
        return c
    raise ValueError(f"Unknown modulation: {modulation}")


class Modulator:
    def __init__(self, samples_per_symbol=8, bt=0.3, rrc_alpha=0.35, rrc_span=6):
        """Same arguments and attributes as the reference constructor (``modulators.py:52-63``)."""
        self.sps = int(samples_per_symbol)
        self.bt = float(bt)
        self.rrc_alpha = rrc_alpha
        self.rrc_span = rrc_span
        self.rrc_filter = rrcosfilter(self.rrc_span, self.rrc_alpha, 1, self.sps)
        self.filter_delay = (len(self.rrc_filter) - 1) // 2
        self._h = {}

    # ============ PULSE SHAPING (modulators.py:85-117) ============
    def _taps(self):
        """the reference's float64 tap array, as the C ABI takes it (a host pointer)"""
        self._taps_h = np.ascontiguousarray(self.rrc_filter, np.float64)
        return _lib.host_ptr(self._taps_h)

    def apply_pulse_shaping(self, symbols):
        """Upsample by ``sps`` and apply the TX RRC filter: ``upfirdn(rrc_filter, complex64(symbols), up=sps)``
        (``:85-100``), length ``(n-1)*sps + len(rrc_filter)``.  numpy in -> complex128 numpy out (the
        reference's dtype; the values carry float32 accuracy), CUDA tensor in -> complex64 CUDA tensor out."""
        torch = _lib.require_cuda()
        is_torch = isinstance(symbols, torch.Tensor)
        s = _lib.to_device(symbols if is_torch else np.asarray(symbols).astype(np.complex64), torch.complex64).reshape(-1)
        n, nt = s.numel(), len(self.rrc_filter)
        out = torch.empty((n - 1) * self.sps + nt if n else 0, dtype=torch.complex64, device=s.device)
        if n:
            _lib.check(_lib.load().b200dvb_pulse_shape(n, _lib.ptr(s), self._taps(), nt, self.sps,
                                                       _lib.ptr(out), _lib.stream_ptr()), "pulse_shape")
        return out if is_torch else out.cpu().numpy().astype(np.complex128)

    def matched_filter(self, samples):
        """RX RRC filter and symbol-rate decimation: ``convolve(samples, rrc_filter, 'full')[2*filter_delay::sps]``
        (``:102-117``); an empty complex64 array when the start index is past the end, as the reference returns."""
        torch = _lib.require_cuda()
        is_torch = isinstance(samples, torch.Tensor)
        x = _lib.to_device(samples if is_torch else np.asarray(samples).astype(np.complex64), torch.complex64).reshape(-1)
        n, nt = x.numel(), len(self.rrc_filter)
        start = 2 * self.filter_delay
        full = n + nt - 1 if n else 0
        if start >= full:
            return torch.empty(0, dtype=torch.complex64, device=x.device) if is_torch else np.array([], dtype=np.complex64)
        n_out = (full - start + self.sps - 1) // self.sps
        out = torch.empty(n_out, dtype=torch.complex64, device=x.device)
        _lib.check(_lib.load().b200dvb_matched_filter(n, _lib.ptr(x), self._taps(), nt, self.sps, start,
                                                      n_out, _lib.ptr(out), _lib.stream_ptr()), "matched_filter")
        return out if is_torch else out.cpu().numpy().astype(np.complex128)

    def _modem(self, name):
        if name not in self._h:
            self._h[name] = ModemHandle(name, natural_constellation(name))
        return self._h[name]

    def _mod(self, bits, name, k, c128=True):
        bits = np.asarray(bits, int)
        pad = (k - len(bits) % k) % k
        if pad:
            bits = np.append(bits, [0] * pad)
        return self._modem(name).map(bits, out_complex128=c128)

    def _demod(self, symbols, name):
        return self._modem(name).hard(np.asarray(symbols)).astype(int)

    def mod_bpsk(self, bits): return self._mod(bits, 'BPSK', 1, c128=False)
    def demod_bpsk(self, symbols): return self._demod(symbols, 'BPSK')
    def mod_qpsk(self, bits): return self._mod(bits, 'QPSK', 2)
    def demod_qpsk(self, symbols): return self._demod(symbols, 'QPSK')
    def mod_8psk(self, bits): return self._mod(bits, '8PSK', 3)
    def demod_8psk(self, symbols): return self._demod(symbols, '8PSK')
    def mod_16qam(self, bits): return self._mod(bits, '16QAM', 4)
    def demod_16qam(self, symbols): return self._demod(symbols, '16QAM')
    def mod_64qam(self, bits): return self._mod(bits, '64QAM', 6)
    def demod_64qam(self, symbols): return self._demod(symbols, '64QAM')
