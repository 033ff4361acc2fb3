"""Host-side logic of the multi-GPU path, on CPU with the gloo backend
(world_size 2): frame sharding tiles the batch, and the counter all-reduce
reproduces the single-rank totals."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from modulations_b200 import montecarlo as mc


def test_shard_range_tiles_and_aligns():
    for total in (0, 16, 100, 1000, 1 << 20, 1_000_000):
        for world in (1, 2, 3, 4, 8):
            spans = [mc.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b
            assert all(lo % mc.ALIGN == 0 for lo, _ in spans)


def test_noise_var_matches_reference_formula():
    # turbo_test_suite.py:132-134: 1 / (2 * R * 10^(snr/10))
    assert mc.noise_var('1/3', 3.0) == pytest.approx(1.0 / (2.0 * (1 / 3) * 10 ** 0.3))
    assert mc.noise_var('1/2', 0.0, bps=4) == pytest.approx(1.0 / (2 * 0.5 * 4))


def _fake_counters(lo, hi, n_points):
    """Deterministic per-frame 'errors' so any sharding must sum to the same totals."""
    out = np.zeros(4 * n_points, np.int64)
    f = np.arange(lo, hi, dtype=np.int64)
    for p in range(n_points):
        e = (f * 2654435761 + p) % 7
        out[4 * p:4 * p + 4] = [e.sum(), (e > 0).sum(), len(f), len(f) * 424]
    return out


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = mc.shard_range(total, rank, world)
    t = torch.from_numpy(_fake_counters(lo, hi, 3))
    mc.reduce_counters(t)
    q.put((rank, t.numpy().tolist()))
    dist.destroy_process_group()


def test_counter_allreduce_gloo_world2():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    total = 100_003
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = _fake_counters(0, total, 3).tolist()
    assert got[0] == want and got[1] == want
    cfg = mc.SweepConfig(ebn0_db=[0.0, 1.0, 2.0])
    s = mc.summarise(cfg, np.array(want))
    assert s["points"][1]["frames"] == total and 0 <= s["points"][0]["ber"] <= 7
