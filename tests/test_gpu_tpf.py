"""GPU parity tests specific to the thread-per-frame decoder kernel (decode_tpf.cu):
both decode kernels must agree with the oracle and with each other, bit for bit, on tile
boundaries (16 frames per warp), strided input, packed output and the in-kernel counters."""
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
    """kernel: 'tpf' / 'quad' force one decode kernel (B200DVB_OPT_KERNEL), 'auto' is the production choice."""
    from modulations_b200 import dvb_rcs2_turbo as turbo
    if kernel == "tpf" and N > 212:
        kernel = "auto"                          # no thread-per-frame geometry for the long frames
    return turbo.DVBRCS2_Turbo(N, rate, iters, kernel=kernel)


@pytest.mark.parametrize("N,rate,iters", [(212, '1/3', 8), (220, '1/3', 3), (48, '1/2', 8), (64, '1/3', 2), (424, '1/3', 1)])
@pytest.mark.parametrize("nfr", [1, 15, 16, 17, 37])
def test_kernels_agree_on_tile_boundaries(N, rate, iters, nfr):
    """1 / 15 / 16 / 17 / 37 frames: partial, exact and multiple tiles of 16 frames."""
    o = oracle.OracleTurbo(N, rate, iters)
    info, llr = _llrs(o, N, rate, nfr, 2.0, 99 + N + nfr)
    ref = o.decode_batch(llr)
    for kernel in ("tpf", "quad", "lat", "auto"):
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
    # read on the HOST straight after the call: the device->host copies must have completed (no stream-ordered
    # access in between that could hide a missing synchronisation)
    got = out.numpy().copy()
    assert np.array_equal(got, want.cpu().numpy())
    ref = o.decode_batch(llr)
    assert np.array_equal(got[:16], ref)


@pytest.mark.parametrize("mode", ["packed", "uint8"])
def test_host_pipeline_compact_outputs(mode):
    """out="packed" copies the kernel's packed words to the host (56 B instead of 1 696 B per N=212 frame),
    out="uint8" one byte per bit; both must carry exactly the bits of the int32 layout."""
    import torch
    from modulations_b200.dvb_rcs2_turbo import unpack_bits
    N, rate, iters = 212, '1/3', 2
    o = oracle.OracleTurbo(N, rate, iters)
    info, llr = _llrs(o, N, rate, 16, 2.0, 12)
    g = _codec(N, rate, iters, "auto")
    B = 20000
    big = torch.from_numpy(llr).repeat((B + 15) // 16, 1)[:B].contiguous().pin_memory()
    ref = np.tile(o.decode_batch(llr), ((B + 15) // 16, 1))[:B]
    out = g.decode_batch_host(big, out=mode)
    got = out.numpy().copy()
    if mode == "packed":
        assert out.dtype == torch.int32 and tuple(out.shape) == (B, (2 * N + 31) // 32)
        got = unpack_bits(got, 2 * N)
    else:
        assert out.dtype == torch.uint8 and tuple(out.shape) == (B, 2 * N)
    assert np.array_equal(got, ref)
    with pytest.raises(ValueError):
        g.decode_batch_host(big, out_host=torch.empty((B, 2 * N), dtype=torch.int32), out=mode)


def test_timed_kernel_instance_is_bit_exact_too():
    """B200DVB_OPT_PHASE_TIMERS selects the template instance with per-phase clock64() accounting
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
    g.handle.set_option(_lib.OPT_PHASE_TIMERS, 1)
    assert np.array_equal(g.decode_batch(llr), ref)
    g.handle.set_option(_lib.OPT_PHASE_TIMERS, 0)
    _lib.check(lib.b200dvb_debug_tpf_cycles(_lib.host_ptr(ph), 1), "read")
    assert ph[7] > 0 and abs(ph[:7].sum() - ph[7]) <= 0.05 * ph[7]
