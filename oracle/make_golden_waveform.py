#!/usr/bin/env python3
"""oracle/make_golden_waveform.py — freezes outputs of the UNMODIFIED reference's waveform stage
(/root/reference/modulators.py: rrcosfilter, Modulator.apply_pulse_shaping, Modulator.matched_filter)
on seeded inputs into tests/golden/waveform_kat.npz.  Test infrastructure; needs /root/reference + scipy.

    python oracle/make_golden_waveform.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import vectors  # noqa: E402

REF = os.environ.get("REFERENCE_DIR", "/root/reference")


def main():
    sys.path.insert(0, REF)
    import modulators                                             # the reference module, imported as it is
    out = {}
    for (sps, alpha, span) in vectors.WAVEFORM_CASES:
        tag = f"sps{sps}_a{alpha}_n{span}"
        mo = modulators.Modulator(samples_per_symbol=sps, rrc_alpha=alpha, rrc_span=span)
        out[f"{tag}/taps"] = np.asarray(mo.rrc_filter, float)
        out[f"{tag}/delay"] = np.int64(mo.filter_delay)
        syms = vectors.waveform_symbols(sps)
        shaped = mo.apply_pulse_shaping(syms)
        out[f"{tag}/shaped"] = np.asarray(shaped)
        rx = vectors.waveform_noise(shaped, sps)
        out[f"{tag}/mf"] = np.asarray(mo.matched_filter(rx))
        out[f"{tag}/mf_clean"] = np.asarray(mo.matched_filter(shaped))
    out["short/mf_empty"] = np.asarray(modulators.Modulator().matched_filter(np.zeros(3, np.complex64)))
    path = os.path.join(ROOT, "tests", "golden", "waveform_kat.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), sorted(out)[:6])


if __name__ == "__main__":
    main()
