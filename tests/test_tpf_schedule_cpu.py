"""CPU check of the thread-per-frame decoder's schedule: tools/tpf_emulator.cu replays, for one
frame, exactly what an (alpha lane, beta lane) pair of decode_tpf.cu does — bit-reversed labels
for the backward lane, meet in the middle, checkpoints, recompute windows, fused epilogue — with
the kernel's own arithmetic core (tpf_core.cuh compiled for the host) and compares one SISO with
the oracle bit for bit.  No GPU needed; skipped if nvcc is not on this machine."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"),
                    reason="needs nvcc to compile the host side of tpf_core.cuh")
def test_tpf_emulator_matches_oracle(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    orc = tmp_path / "orc.o"
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-c", os.path.join(ROOT, "oracle", "turbo_oracle.c"), "-o", str(orc)],
                   check=True)
    emu_o = tmp_path / "emu.o"
    subprocess.run([nvcc, "-O1", "--fmad=false", "-Xcompiler", "-ffp-contract=off", "-c",
                    os.path.join(ROOT, "tools", "tpf_emulator.cu"), "-o", str(emu_o)], check=True,
                   capture_output=True)
    exe = tmp_path / "emu"
    subprocess.run([nvcc, "-o", str(exe), str(emu_o), str(orc)], check=True, capture_output=True)
    res = subprocess.run([str(exe)], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout[-2000:]
    assert "FAIL" not in res.stdout and res.stdout.count(" ok ") >= 20
