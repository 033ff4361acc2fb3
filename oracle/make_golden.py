#!/usr/bin/env python3
"""oracle/make_golden.py — generates tests/golden/*.npz by RUNNING THE UNMODIFIED
REFERENCE (/root/reference) in the authoring container.  Test infrastructure.

The reference cannot travel to the GPU box, so its outputs on seeded inputs are
frozen here; tests/test_oracle_golden.py pins oracle/ against them and the
`-m gpu` tests pin the CUDA path against oracle/ and against these files.

    python oracle/make_golden.py            # needs /root/reference + numba

Inputs are produced by tests/vectors.py (np.random.RandomState, stable across
numpy versions); only reference OUTPUTS (and input hashes) are stored.
"""
import ast
import hashlib
import os
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import vectors  # noqa: E402

REF = os.environ.get("REFERENCE_DIR", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def load_reference():
    """Import dvb_rcs2_turbo / sdr_modem / modulators from a writable copy (numba
    cache=True wants to write __pycache__), and exec the mapper + compute_llr
    definitions of test_sdr_with_coding.py (its module top imports matplotlib and
    a class that does not exist, so it cannot be imported whole)."""
    tmp = tempfile.mkdtemp(prefix="refcopy_")
    for f in os.listdir(REF):
        if f.endswith(".py"):
            shutil.copy(os.path.join(REF, f), tmp)
    sys.path.insert(0, tmp)
    import dvb_rcs2_turbo as turbo
    import sdr_modem
    import modulators
    src = open(os.path.join(REF, "test_sdr_with_coding.py")).read()
    tree = ast.parse(src)
    keep = []
    want_fn = {"bpsk_mod", "bpsk_demod", "qpsk_mod", "qpsk_demod", "psk8_mod", "psk8_demod",
               "qam16_mod", "qam16_demod", "compute_llr"}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in want_fn:
            keep.append(node)
        if isinstance(node, ast.Assign) and getattr(node.targets[0], "id", "") == "MODULATIONS":
            keep.append(node)
    ns = {"np": np}
    exec(compile(ast.Module(body=keep, type_ignores=[]), "test_sdr_with_coding.py", "exec"), ns)
    return turbo, sdr_modem, modulators, ns


def main():
    os.makedirs(OUT, exist_ok=True)
    turbo, sdr_modem, modulators, tswc = load_reference()

    # ---- 1. tables --------------------------------------------------------
    tab = {}
    for N in sorted(turbo.INTERLEAVER_PARAMS):
        c = turbo.DVBRCS2_Turbo(N, '1/3', 1)
        tab[f"perm_{N}"] = c.perm
        tab[f"inv_perm_{N}"] = c.inv_perm
        GN = turbo.mat_pow_gf2(c.G_matrix, N)
        tab[f"Gpow_{N}"] = GN
        tab[f"circ_lut_{N}"] = np.array([turbo.solve_circular_state_gf2(GN, z) for z in range(16)], np.int32)
        for rate in turbo.PUNCTURE_PATTERNS:
            cc = turbo.DVBRCS2_Turbo.__new__(turbo.DVBRCS2_Turbo)
            cc.N = N
            cc.punct = turbo.PUNCTURE_PATTERNS[rate]
            cc._calc_coded_size()
            tab[f"n_coded_{N}_{rate.replace('/', '_')}"] = np.int64(cc.n_coded)
    for k in ("next_state", "out_W", "out_Y", "prev_state", "prev_input", "G_matrix"):
        tab[k] = getattr(c, k)
    np.savez_compressed(os.path.join(OUT, "tables.npz"), **tab)

    # ---- 2. codec known-answer vectors --------------------------------------
    kat = {}
    for (N, rate, iters, nfr, ebn0s) in vectors.CODEC_CASES:
        codec = turbo.DVBRCS2_Turbo(N, rate, iters)
        tag = vectors.case_tag(N, rate, iters)
        info, llr_by_snr = vectors.codec_inputs(N, rate, nfr, ebn0s, encode=codec.encode,
                                                n_coded=codec.n_coded)
        coded = np.stack([codec.encode(b) for b in info]).astype(np.uint8)
        kat[f"{tag}/info_sha"] = sha(info.astype(np.uint8))
        kat[f"{tag}/coded"] = np.packbits(coded, axis=1)
        for e, llr in zip(ebn0s, llr_by_snr):
            dec = np.stack([codec.decode(x) for x in llr]).astype(np.uint8)
            kat[f"{tag}/ebn0_{e}/llr_sha"] = sha(llr)
            kat[f"{tag}/ebn0_{e}/dec"] = np.packbits(dec, axis=1)
        # SISO KATs on frame 0 of the first SNR: (a) first half-iteration (La = 0,
        # sf = 0.7); (b) a-priori = seeded float64 noise, sf = 1.0.
        Lc = vectors.depuncture(llr_by_snr[0][0], N, codec.punct)
        z = np.zeros(N)
        LeA, LeB = turbo.bcjr_max_log_map(Lc[0], Lc[1], Lc[2], Lc[3], z, z, codec.next_state,
                                          codec.out_W, codec.out_Y, codec.prev_state,
                                          codec.prev_input, N, 0.7)
        kat[f"{tag}/siso0_LeA"], kat[f"{tag}/siso0_LeB"] = LeA, LeB
        LaA, LaB = vectors.siso_apriori(N)
        LeA, LeB = turbo.bcjr_max_log_map(Lc[0], Lc[1], Lc[4], Lc[5], LaA, LaB, codec.next_state,
                                          codec.out_W, codec.out_Y, codec.prev_state,
                                          codec.prev_input, N, 1.0)
        kat[f"{tag}/siso1_LeA"], kat[f"{tag}/siso1_LeB"] = LeA, LeB
        print(tag, "done")
    np.savez_compressed(os.path.join(OUT, "codec_kat.npz"), **kat)

    # ---- 3. mapper / slicer / soft demapper ---------------------------------
    m = sdr_modem.SDRModem()
    mod = {}
    for name in m.MODULATIONS:
        bps = m.MODULATIONS[name]['bps']
        labels = np.array([list(map(int, format(i, f'0{bps}b'))) for i in range(1 << bps)]).flatten()
        const = m.modulate(labels, name)
        mod[f"{name}/const"] = const
        mod[f"{name}/const_dtype"] = str(const.dtype)
        bits = vectors.mapper_bits(name)
        syms = m.modulate(bits, name)
        mod[f"{name}/syms"] = syms
        rx = vectors.noisy_symbols(syms, name)
        mod[f"{name}/hard"] = np.asarray(m.demodulate(rx, name)).astype(np.uint8)
        mod[f"{name}/rx_sha"] = sha(rx)
        if name in tswc["MODULATIONS"]:
            # duplicated mapper in the script must equal SDRModem's
            assert np.array_equal(tswc["MODULATIONS"][name]['mod'](bits), syms)
            for nv in vectors.DEMAP_NOISE_VARS:
                mod[f"{name}/llr_nv{nv}"] = tswc["compute_llr"](rx[:vectors.DEMAP_N], name, nv)
                mod[f"{name}/llr32_nv{nv}"] = tswc["compute_llr"](
                    rx[:vectors.DEMAP_N].astype(np.complex64), name, nv)
    mo = modulators.Modulator()
    for name, fn, dfn in (("BPSK", mo.mod_bpsk, mo.demod_bpsk), ("QPSK", mo.mod_qpsk, mo.demod_qpsk),
                          ("8PSK", mo.mod_8psk, mo.demod_8psk), ("16QAM", mo.mod_16qam, mo.demod_16qam),
                          ("64QAM", mo.mod_64qam, mo.demod_64qam)):
        bits = vectors.mapper_bits(name)
        syms = fn(bits)
        mod[f"alt_{name}/syms"] = syms
        rx = vectors.noisy_symbols(syms, name)
        mod[f"alt_{name}/hard"] = np.asarray(dfn(rx.copy())).astype(np.uint8)
    np.savez_compressed(os.path.join(OUT, "modem_kat.npz"), **mod)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
