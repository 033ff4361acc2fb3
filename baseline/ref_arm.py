"""CPU arm of bench.py: the reference's OWN decoder timed on the host cores (TEST / MEASUREMENT INFRASTRUCTURE).

`baseline/_ref/` is a verbatim copy of the reference's importable modules, staged by `__graft_entry__.build()` in the
authoring container (git-ignored, travels to the GPU box with the snapshot).  When it is there and numba imports,
`ReferencePool` runs the committed `DVBRCS2_Turbo.decode` (numba JIT, `/root/reference/dvb_rcs2_turbo.py:464-537`),
one single-threaded worker process per host core, JIT warm-up outside the timed region.  When it is
not, the caller falls back to the C port of the same algorithm (oracle/turbo_oracle.c) and says so (`kind: "port"`).
Nothing under modulations_b200/ imports this file.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available():
    """-> (ok, why)."""
    if not os.path.exists(os.path.join(REF_DIR, "dvb_rcs2_turbo.py")):
        return False, "baseline/_ref/dvb_rcs2_turbo.py not staged"
    try:
        import numba  # noqa: F401
    except Exception as e:                                    # pragma: no cover
        return False, f"numba not importable: {e!r}"
    return True, ""


def _worker(conn, N, rate, iters, path, lo, hi):
    try:
        sys.path.insert(0, REF_DIR)
        import dvb_rcs2_turbo as ref                           # the unmodified reference module
        codec = ref.DVBRCS2_Turbo(N, rate, iters)              # constructor JIT-compiles and warms up (:303-309)
        llr = np.load(path, mmap_mode="r")[lo:hi]
        llr = np.ascontiguousarray(llr)
        if len(llr):
            codec.decode(llr[0])                               # make sure the jitted SISO is hot
        conn.send(("ready", codec.n_coded))
        if conn.recv() != "go":
            return
        t0 = time.perf_counter()
        out = np.empty((len(llr), codec.k_info), np.uint8)
        for i in range(len(llr)):
            out[i] = codec.decode(llr[i])
        dt = time.perf_counter() - t0
        conn.send(("done", dt, out))
    except Exception as e:                                    # pragma: no cover
        conn.send(("error", repr(e)))


def _run_pool(N, rate, iters, path, procs):
    """In THIS process (no CUDA context here): fork one worker per core, time them, return the result dict."""
    F = np.load(path, mmap_mode="r").shape[0]
    procs = max(1, min(int(procs), F))
    ctx = mp.get_context("fork")
    cuts = np.linspace(0, F, procs + 1).astype(int)
    workers = []
    try:
        for i in range(procs):
            a, b = ctx.Pipe()
            p = ctx.Process(target=_worker, args=(b, N, rate, iters, path, int(cuts[i]), int(cuts[i + 1])), daemon=True)
            p.start()
            workers.append((p, a))
        for p, a in workers:
            msg = a.recv()
            if msg[0] != "ready":
                raise RuntimeError(f"reference worker failed: {msg}")
        t0 = time.perf_counter()
        for p, a in workers:
            a.send("go")
        outs, worker_s = [], []
        for p, a in workers:
            msg = a.recv()
            if msg[0] != "done":
                raise RuntimeError(f"reference worker failed: {msg}")
            worker_s.append(msg[1])
            outs.append(msg[2])
        wall = time.perf_counter() - t0
        for p, a in workers:
            p.join(timeout=10)
        return {"seconds": wall, "worker_seconds_max": max(worker_s), "frames": int(F), "procs": procs,
                "decoded": np.concatenate(outs, axis=0)}
    finally:
        for p, a in workers:
            if p.is_alive():
                p.terminate()


def decode_timed(N, rate, iters, llr, procs=None):
    """Decode `llr` float32 [F, n] with the reference on `procs` single-threaded worker processes.
    -> dict(seconds = wall time of the timed region (from "go" to the last worker's result; JIT warm-up and input
    loading excluded), worker_seconds_max, frames, procs, decoded = uint8 [F, 2N]).
    Runs in a child interpreter (`python baseline/ref_arm.py ...`) so that the caller's CUDA context is never forked."""
    import json
    import subprocess
    procs = int(procs or os.cpu_count() or 1)
    shm_dir = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
    tag = f"b200dvb_ref_{os.getpid()}_{time.time_ns()}"
    path_in, path_out = os.path.join(shm_dir, tag + "_in.npy"), os.path.join(shm_dir, tag + "_out.npy")
    np.save(path_in, np.ascontiguousarray(llr, np.float32))
    try:
        res = subprocess.run([sys.executable, os.path.abspath(__file__), str(N), rate, str(iters), path_in, path_out,
                              str(procs)], capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("reference arm failed: " + res.stderr[-2000:])
        out = json.loads(res.stdout.strip().splitlines()[-1])
        out["decoded"] = np.load(path_out)
        return out
    finally:
        for f in (path_in, path_out):
            try:
                os.remove(f)
            except OSError:
                pass


if __name__ == "__main__":
    import json
    N_, rate_, iters_, pin, pout, procs_ = int(sys.argv[1]), sys.argv[2], int(sys.argv[3]), sys.argv[4], sys.argv[5], int(sys.argv[6])
    r = _run_pool(N_, rate_, iters_, pin, procs_)
    np.save(pout, r.pop("decoded"))
    print(json.dumps(r))
