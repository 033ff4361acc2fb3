"""CPU: the oracle's restatement of the reference waveform stage (modulators.py:19-117: rrcosfilter,
apply_pulse_shaping, matched_filter) against outputs of the unmodified reference frozen by
oracle/make_golden_waveform.py; the package's host-side tap generator against both."""
import os

import numpy as np
import pytest

from modulations_b200.modulators import rrcosfilter
from oracle import oracle
from tests import vectors

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "waveform_kat.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.mark.parametrize("sps,alpha,span", vectors.WAVEFORM_CASES)
def test_waveform_oracle_matches_reference(gold, sps, alpha, span):
    tag = f"sps{sps}_a{alpha}_n{span}"
    h = oracle.rrcosfilter(span, alpha, 1, sps)
    assert np.array_equal(h, gold[f"{tag}/taps"])                       # bit-exact taps
    assert np.array_equal(rrcosfilter(span, alpha, 1, sps), h)          # product-side generator == oracle
    assert (len(h) - 1) // 2 == int(gold[f"{tag}/delay"])
    syms = vectors.waveform_symbols(sps)
    shaped = oracle.pulse_shape(syms, h, sps)
    ref = gold[f"{tag}/shaped"]
    assert shaped.shape == ref.shape and shaped.dtype == ref.dtype
    assert np.max(np.abs(shaped - ref)) <= 4e-16                        # upfirdn sums in polyphase order
    rx = vectors.waveform_noise(ref, sps)
    assert np.array_equal(oracle.matched_filter(rx, h, sps), gold[f"{tag}/mf"])
    assert np.array_equal(oracle.matched_filter(ref, h, sps), gold[f"{tag}/mf_clean"])


def test_waveform_short_input(gold):
    h = oracle.rrcosfilter(6, 0.35, 1, 8)
    got = oracle.matched_filter(np.zeros(3, np.complex64), h, 8)
    assert got.shape == gold["short/mf_empty"].shape
    assert oracle.matched_filter(np.zeros(0, np.complex64), h, 8).size == 0
