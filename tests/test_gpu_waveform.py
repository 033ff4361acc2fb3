"""GPU: RRC pulse shaping and matched filter (waveform.cu through b200dvb_pulse_shape /
b200dvb_matched_filter and the Modulator drop-in) against the oracle and the frozen reference outputs.
float32 FMAs vs the reference's float64 convolution: |delta| <= 2e-6 * sum|taps| * max|input| (stated)."""
import os

import numpy as np
import pytest

from oracle import oracle
from tests import vectors

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "waveform_kat.npz")


def _tol(h, x):
    return 2e-6 * np.sum(np.abs(h)) * max(1.0, float(np.max(np.abs(x))))


@pytest.mark.parametrize("sps,alpha,span", vectors.WAVEFORM_CASES)
def test_waveform_matches_reference(sps, alpha, span):
    from modulations_b200.modulators import Modulator
    gold = np.load(GOLD)
    tag = f"sps{sps}_a{alpha}_n{span}"
    mo = Modulator(samples_per_symbol=sps, rrc_alpha=alpha, rrc_span=span)
    assert np.array_equal(mo.rrc_filter, gold[f"{tag}/taps"]) and mo.filter_delay == int(gold[f"{tag}/delay"])
    syms = vectors.waveform_symbols(sps)
    shaped = mo.apply_pulse_shaping(syms)
    ref = gold[f"{tag}/shaped"]
    assert shaped.shape == ref.shape and shaped.dtype == ref.dtype
    assert np.max(np.abs(shaped - ref)) <= _tol(mo.rrc_filter, syms)
    rx = vectors.waveform_noise(ref, sps)
    mf = mo.matched_filter(rx)
    refm = gold[f"{tag}/mf"]
    assert mf.shape == refm.shape and mf.dtype == refm.dtype
    assert np.max(np.abs(mf - refm)) <= _tol(mo.rrc_filter, rx)
    # perfect-sync loopback (modulators.py:102-117's use case): the reference's own output on its own waveform
    clean = mo.matched_filter(ref)
    assert clean.shape == gold[f"{tag}/mf_clean"].shape
    assert np.max(np.abs(clean - gold[f"{tag}/mf_clean"])) <= _tol(mo.rrc_filter, ref)
    assert np.max(np.abs(clean[:len(syms)] - syms)) < 0.1           # RRC * RRC is (nearly, for a short span) Nyquist


@pytest.mark.parametrize("n", [0, 1, 3, 255, 256, 257, 5000, 100003])
def test_waveform_edges_against_oracle(n):
    """empty, shorter than the filter, tile boundaries of the 256-output blocks, a long odd length."""
    import torch
    from modulations_b200.modulators import Modulator
    mo = Modulator()                                            # reference defaults: sps 8, alpha 0.35, span 6
    rs = np.random.RandomState(n + 11)
    x = (rs.randn(n) + 1j * rs.randn(n)).astype(np.complex64)
    want = oracle.pulse_shape(x, mo.rrc_filter, mo.sps)
    got = mo.apply_pulse_shaping(x)
    assert got.shape == want.shape
    if n:
        assert np.max(np.abs(got - want)) <= _tol(mo.rrc_filter, x)
    wantm = oracle.matched_filter(x, mo.rrc_filter, mo.sps)
    gotm = mo.matched_filter(x)
    assert gotm.shape == wantm.shape
    if wantm.size:
        assert np.max(np.abs(gotm - wantm)) <= _tol(mo.rrc_filter, x)
        t = mo.matched_filter(torch.from_numpy(x).cuda())       # device-resident path
        assert t.is_cuda and t.dtype == torch.complex64
        assert np.array_equal(t.cpu().numpy().astype(np.complex128), gotm)
