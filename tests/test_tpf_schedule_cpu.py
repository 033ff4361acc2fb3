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


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"),
                    reason="needs nvcc to compile the host side of nii_core.cuh")
def test_nii_emulator_matches_its_model(tmp_path):
    """The non-parity "nii" mode: the kernel's arithmetic header (nii_core.cuh: merged-branch float32 records,
    re-associated a-posteriori maxima per parity class, float32 epilogue) replayed in the kernel's lane schedule,
    four chained SISOs with the boundary metrics carried over, against the naive model oracle/nii_model.c."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    objs = []
    for src in ("turbo_oracle.c", "nii_model.c"):
        o = tmp_path / (src + ".o")
        subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-c", os.path.join(ROOT, "oracle", src), "-o", str(o)], check=True)
        objs.append(str(o))
    emu_o = tmp_path / "nii_emu.o"
    subprocess.run([nvcc, "-O1", "--fmad=false", "-Xcompiler", "-ffp-contract=off", "-c",
                    os.path.join(ROOT, "tools", "nii_emulator.cu"), "-o", str(emu_o)], check=True, capture_output=True)
    exe = tmp_path / "nii_emu"
    subprocess.run([nvcc, "-o", str(exe), str(emu_o)] + objs, check=True, capture_output=True)
    res = subprocess.run([str(exe)], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout[-2000:]
    assert "FAIL" not in res.stdout and res.stdout.count(" ok ") >= 20


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"),
                    reason="needs nvcc to compile the host side of nii16_core.cuh")
def test_nii16_emulator_matches_its_model(tmp_path):
    """The fixed-point non-parity mode "nii16": the kernel's packed-s16x2 arithmetic header (host definitions of the DPX
    operations) replayed in the lane schedule with two different frames per register, against the naive integer model
    oracle/nii16_model.c (which also range-checks every 16-bit quantity), incl. saturated inputs."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    objs = []
    for src in ("turbo_oracle.c", "nii16_model.c"):
        o = tmp_path / (src + ".o")
        subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-c", os.path.join(ROOT, "oracle", src), "-o", str(o)], check=True)
        objs.append(str(o))
    emu_o = tmp_path / "nii16_emu.o"
    subprocess.run([nvcc, "-O1", "-c", os.path.join(ROOT, "tools", "nii16_emulator.cu"), "-o", str(emu_o)],
                   check=True, capture_output=True)
    exe = tmp_path / "nii16_emu"
    subprocess.run([nvcc, "-o", str(exe), str(emu_o)] + objs, check=True, capture_output=True)
    res = subprocess.run([str(exe)], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout[-2000:]
    assert "FAIL" not in res.stdout and res.stdout.count(" ok ") >= 20


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"),
                    reason="needs nvcc to compile the host side of tpf_core.cuh")
def test_lat_emulator_matches_oracle(tmp_path):
    """The low-latency kernel (decode_lat.cu): one state metric per lane (per-lane operand / record selection), lap 1
    of each recursion in four speculative segments verified by a carrier, lap 2 ended where it re-joins lap 1 —
    replayed on the CPU with the kernel's arithmetic core for warm-up lengths from 1 (nearly every guess wrong) to
    beyond N (one segment), against oracle/turbo_oracle.c bit for bit."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    orc = tmp_path / "orc.o"
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-c", os.path.join(ROOT, "oracle", "turbo_oracle.c"), "-o", str(orc)],
                   check=True)
    emu_o = tmp_path / "lat_emu.o"
    subprocess.run([nvcc, "-O1", "--fmad=false", "-Xcompiler", "-ffp-contract=off", "-c",
                    os.path.join(ROOT, "tools", "lat_emulator.cu"), "-o", str(emu_o)], check=True, capture_output=True)
    exe = tmp_path / "lat_emu"
    subprocess.run([nvcc, "-o", str(exe), str(emu_o), str(orc)], check=True, capture_output=True)
    res = subprocess.run([str(exe)], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout[-2000:]
    assert "FAIL" not in res.stdout and res.stdout.count(" ok ") >= 100
