"""CPU-only checks of the C-ABI boundary: the library loads and exports exactly the
entry points include/b200dvb.h declares; host-side table logic matches the oracle."""
import os
import re

import numpy as np
import pytest

from modulations_b200 import _lib
from modulations_b200 import dvb_rcs2_turbo as turbo
from modulations_b200.modulators import natural_constellation
from modulations_b200.sdr_modem import gray_constellation, SDRModem
from oracle import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "b200dvb.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(b200dvb_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_lib.SIGNATURES) == names          # the ctypes table mirrors the header
    assert lib.b200dvb_version() >= 100
    assert lib.b200dvb_error_string(-2).decode().startswith("no kernel specialisation")


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        turbo.DVBRCS2_Turbo(48, '1/3').decode(np.zeros(288, np.float32))
    with pytest.raises(RuntimeError):
        SDRModem().modulate([0, 1, 1, 0], 'QPSK')


def test_constructor_errors_match_reference():
    with pytest.raises(ValueError):
        turbo.DVBRCS2_Turbo(50, '1/3')               # dvb_rcs2_turbo.py:295-296
    with pytest.raises(KeyError):
        turbo.DVBRCS2_Turbo(48, '5/6')               # :292
    with pytest.raises(ValueError):
        SDRModem().modulate([0, 1], '1024QAM')       # sdr_modem.py:241-242
    with pytest.raises(ValueError):
        SDRModem().demodulate(np.zeros(2, complex), 'nope')


@pytest.mark.parametrize("N", sorted(turbo.INTERLEAVER_PARAMS))
def test_host_tables_match_golden(golden, N):
    t = golden["tables"]
    c = turbo.DVBRCS2_Turbo(N, '1/3')
    assert np.array_equal(c.perm, t[f"perm_{N}"]) and c.perm.dtype == np.int32
    assert np.array_equal(c.inv_perm, t[f"inv_perm_{N}"])
    for k in ("next_state", "out_W", "out_Y", "prev_state", "prev_input", "G_matrix"):
        assert np.array_equal(getattr(c, k), t[k]), k
    for rate in turbo.PUNCTURE_PATTERNS:
        assert turbo.DVBRCS2_Turbo(N, rate).n_coded == int(t[f"n_coded_{N}_{rate.replace('/', '_')}"])
    GN = turbo.mat_pow_gf2(c.G_matrix, N)
    assert np.array_equal(GN, t[f"Gpow_{N}"])
    assert [turbo.solve_circular_state_gf2(GN, z) for z in range(16)] == list(t[f"circ_lut_{N}"])


def test_gf2_helpers_random():
    rs = np.random.RandomState(0)
    for _ in range(50):
        A = rs.randint(0, 2, (4, 4)); B = rs.randint(0, 2, (4, 4))
        assert np.array_equal(turbo.mat_mul_gf2(A, B), (A @ B) % 2)
        z = int(rs.randint(0, 16))
        assert turbo.solve_circular_state_gf2(A, z) == oracle.solve_circular_state_gf2(A, z)
    assert turbo.max_star(1.0, 2.0) == 2.0 and turbo.max_star(3.0, 2.0) == 3.0
    ns, ow, oy = turbo.build_trellis()
    assert ns.shape == (16, 4) and ow.max() == 1 and oy.max() == 1


def test_constellations_match_golden(golden):
    m = golden["modem_kat"]
    for name in SDRModem.MODULATIONS:
        c = gray_constellation(name)
        assert str(c.dtype) == str(m[f"{name}/const_dtype"]) and np.array_equal(c, m[f"{name}/const"])
    for name in ('BPSK', 'QPSK', '8PSK', '16QAM', '64QAM'):
        assert np.array_equal(natural_constellation(name), oracle.modulator_constellation(name))
