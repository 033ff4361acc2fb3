"""Builds libb200dvb.so in-tree with nvcc for sm_100a (no torch, no JIT cache).

    python -m modulations_b200.build [--force] [--verbose]

The shared library is git-ignored but travels to the GPU box with the repo
snapshot; `__graft_entry__.build()` calls :func:`build`.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libb200dvb.so")
SOURCES = ["api.cu", "decode_quad.cu", "decode_tpf.cu", "decode_nii.cu", "decode_lat.cu", "encode.cu", "modem.cu", "waveform.cu", "microbench.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "--shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--fmad=false",
    "-ccbin", "g++",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def source_sha256():
    """sha256 over everything the library is built from (csrc/*, include/b200dvb.h, the nvcc flags): identifies a
    build independently of the machine that compiled it.  profiles/r02_traffic.json is stamped with it."""
    import hashlib
    h = hashlib.sha256()
    files = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)) + [os.path.join(os.path.dirname(HERE), "include", "b200dvb.h")]
    for f in files:
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS + SOURCES).encode())
    return h.hexdigest()


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(os.path.dirname(HERE), "include", "b200dvb.h"), os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ into modulations_b200/libb200dvb.so."""
    if not force and not needs_build():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if verbose or res.returncode != 0:
        print(log)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libb200dvb.so (see modulations_b200/build.log)")
    _check_sass()
    return LIB


def _check_sass():
    """ptxas 12.9 may fold a warp-uniform base into the shared-memory operand of cp.async
    (`LDGSTS [R+UR+imm]`); that form raises "illegal instruction" on sm_100a (found on the B200,
    DESIGN.md §4.1).  The kernels route such bases through memory; make sure it stays that way."""
    import re
    try:
        sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    except OSError:
        return
    bad = [l for l in sass.splitlines() if re.search(r"LDGSTS.*\[R\d+\+UR", l)]
    if bad:
        raise RuntimeError("libb200dvb.so contains LDGSTS with a [R+UR+imm] destination (faults on sm_100a):\n"
                           + "\n".join(bad[:5]))


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv or True)
