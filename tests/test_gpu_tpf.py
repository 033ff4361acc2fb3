"""GPU parity tests specific to the thread-per-frame decoder kernel (decode_tpf.cu):
both decode kernels must agree with the oracle and with each other, bit for bit, on tile
boundaries (16 frames per warp), strided input, packed output and the in-kernel counters."""
import os

import numpy as np
import pytest

from oracle import oracle
from tests import vectors

pytestmark = pytest.mark.gpu


def _llrs(o, N, rate, nfr, ebn0, seed):
    rs = np.random.RandomState(seed)
    info = rs.randint(0, 2, (nfr, 2 * N))
    llr = np.stack([vectors.awgn_llr(rs, np.asarray(o.encode(info[i])), rate, ebn0) for i in range(nfr)])
    return info, llr


def _codec(N, rate, iters, kernel):
    """kernel: 'tpf' (default selection) or 'quad' (B200DVB_KERNEL=quad at codec creation)."""
    from modulations_b200 import dvb_rcs2_turbo as turbo
    old = os.environ.get("B200DVB_KERNEL")
    try:
        if kernel == "quad":
            os.environ["B200DVB_KERNEL"] = "quad"
        else:
            os.environ.pop("B200DVB_KERNEL", None)
        return turbo.DVBRCS2_Turbo(N, rate, iters)
    finally:
        if old is None:
            os.environ.pop("B200DVB_KERNEL", None)
        else:
            os.environ["B200DVB_KERNEL"] = old


@pytest.mark.parametrize("N,rate,iters", [(212, '1/3', 8), (220, '1/3', 3), (48, '1/2', 8), (64, '1/3', 2), (424, '1/3', 1)])
@pytest.mark.parametrize("nfr", [1, 15, 16, 17, 37])
def test_kernels_agree_on_tile_boundaries(N, rate, iters, nfr):
    """1 / 15 / 16 / 17 / 37 frames: partial, exact and multiple tiles of 16 frames."""
    o = oracle.OracleTurbo(N, rate, iters)
    info, llr = _llrs(o, N, rate, nfr, 2.0, 99 + N + nfr)
    ref = o.decode_batch(llr)
    for kernel in ("tpf", "quad"):
        g = _codec(N, rate, iters, kernel)
        dec = g.decode_batch(llr)
        assert np.array_equal(dec, ref), f"{kernel} kernel, N={N} R={rate} B={nfr}: {np.sum(dec != ref)} bits differ"


def test_tpf_strided_packed_and_counters():
    import torch
    N, rate, iters, nfr = 212, '1/3', 8, 50
    o = oracle.OracleTurbo(N, rate, iters)
    info, llr = _llrs(o, N, rate, nfr, 1.0, 4242)
    ref = o.decode_batch(llr)
    g = _codec(N, rate, iters, "tpf")
    n = llr.shape[1]
    for pitch in (n, n + 4, n + 3):            # 16-byte row pitch (row staging), and a pitch that forbids it
        buf = torch.zeros((nfr, pitch), dtype=torch.float32, device="cuda")
        buf[:, :n] = torch.from_numpy(llr).cuda()
        view = buf[:, :n]
        counters = torch.zeros(4, dtype=torch.int64, device="cuda")
        dec = g.decode_batch(view, ref_bits=torch.from_numpy(info.astype(np.uint8)).cuda(), counters=counters)
        assert np.array_equal(dec.cpu().numpy(), ref), f"pitch {pitch}"
        cnt = counters.cpu().numpy()
        assert cnt[0] == np.sum(ref != info) and cnt[1] == np.sum(np.any(ref != info, axis=1))
        assert cnt[2] == nfr and cnt[3] == nfr * 2 * N
    packed = g.decode_batch(llr, out="packed")
    want = np.packbits(ref.astype(np.uint8), axis=1, bitorder="little")
    got = np.ascontiguousarray(packed).view(np.uint8)[:, :want.shape[1]]
    assert np.array_equal(got, want)


def test_tpf_large_batch_is_position_independent():
    """More frames than one wave of warps holds: a frame must decode identically wherever it lands."""
    import torch
    N, rate, iters = 212, '1/3', 8
    o = oracle.OracleTurbo(N, rate, iters)
    info, llr = _llrs(o, N, rate, 24, 2.0, 7)
    ref = o.decode_batch(llr)
    g = _codec(N, rate, iters, "tpf")
    reps = 900                                  # 21 600 frames = 1 350 tiles > 592 resident warps
    big = torch.from_numpy(llr).cuda().repeat(reps, 1)
    dec = g.decode_batch(big).reshape(reps, 24, 2 * N)
    want = torch.from_numpy(ref).cuda()
    assert bool((dec == want[None]).all())


@pytest.mark.parametrize("chunk", [None, 1000, 4096])
def test_host_pipeline_matches_resident_decode(chunk):
    """decode_batch_host (pinned host in -> pinned host out, 3-stream pipeline; the call bench.py's e2e
    figure times) must return exactly what the resident decode returns, for the default chunk (one kernel
    wave) and for chunks that do not divide the batch."""
    import torch
    N, rate, iters = 48, '1/3', 2
    o = oracle.OracleTurbo(N, rate, iters)
    info, llr = _llrs(o, N, rate, 16, 2.0, 11)
    g = _codec(N, rate, iters, "tpf")
    B = 21000                                    # more than two default chunks (9 472 frames on a B200)
    reps = (B + 15) // 16
    big = torch.from_numpy(llr).repeat(reps, 1)[:B].contiguous().pin_memory()
    want = g.decode_batch(big.cuda())
    out = g.decode_batch_host(big) if chunk is None else g.decode_batch_host(big, chunk=chunk)
    assert out.dtype == torch.int32 and tuple(out.shape) == (B, 2 * N)
    assert bool((out.cuda() == want).all())
    ref = torch.from_numpy(o.decode_batch(llr))
    assert bool((out[:16] == ref).all())


def test_timed_kernel_instance_is_bit_exact_too(monkeypatch):
    """B200DVB_TPF_TIMERS selects the template instance with per-phase clock64() accounting
    (tools/tpf_perf.py); it must decode exactly like the production instance and fill the counters."""
    from modulations_b200 import _lib
    N, rate, iters, nfr = 212, '1/3', 2, 33
    o = oracle.OracleTurbo(N, rate, iters)
    info, llr = _llrs(o, N, rate, nfr, 2.0, 777)
    ref = o.decode_batch(llr)
    g = _codec(N, rate, iters, "tpf")
    lib = _lib.load()
    ph = np.zeros(8)
    _lib.check(lib.b200dvb_debug_tpf_cycles(_lib.host_ptr(ph), 1), "reset")
    assert np.array_equal(g.decode_batch(llr), ref)
    _lib.check(lib.b200dvb_debug_tpf_cycles(_lib.host_ptr(ph), 1), "read")
    assert ph[7] == 0, "the production instance keeps no phase counters"
    monkeypatch.setenv("B200DVB_TPF_TIMERS", "1")
    assert np.array_equal(g.decode_batch(llr), ref)
    monkeypatch.delenv("B200DVB_TPF_TIMERS")
    _lib.check(lib.b200dvb_debug_tpf_cycles(_lib.host_ptr(ph), 1), "read")
    assert ph[7] > 0 and abs(ph[:7].sum() - ph[7]) <= 0.05 * ph[7]
