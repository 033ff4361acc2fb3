"""Drop-in for the mapping half of the reference ``modulators.Modulator``
(``modulators.py:119-200``): natural-binary (non-Gray) BPSK/QPSK/8PSK/16/64QAM.
Same kernels as ``SDRModem``, different constellation tables.  Pulse shaping
(``modulators.py:19-115``) is out of scope.
"""
from __future__ import annotations

import numpy as np

from .sdr_modem import ModemHandle


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
    def __init__(self, sps=4, alpha=0.35, span=6):
        self.sps, self.alpha, self.span = sps, alpha, span
        self._h = {}

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
