#!/usr/bin/env python3
"""Observed float32-vs-float64 error of the soft demapper per modulation and noise variance (sets the test tolerance)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle
from tests import vectors
from modulations_b200.soft_demod import compute_llr
for name in vectors.BPS:
    rx = vectors.noisy_symbols(oracle.modulate(vectors.mapper_bits(name), name), name)
    rs = np.random.RandomState(5)
    big = oracle.modulate(rs.randint(0, 2, vectors.BPS[name] * 200000), name)
    big = big + 0.15 * (rs.randn(len(big)) + 1j * rs.randn(len(big)))
    for nv in vectors.DEMAP_NOISE_VARS + [0.0001]:
        e1 = np.abs(compute_llr(rx, name, nv) - oracle.compute_llr(rx, name, nv)).max()
        e2 = np.abs(compute_llr(big, name, nv) - oracle.compute_llr(big, name, nv)).max()
        print(f"{name:7s} nv={nv:<7g} max|dLLR| fixture {e1:.3e}  200k random {e2:.3e}   x max(nv,0.005) = {max(e1, e2) * max(nv, 0.005):.3e}")
