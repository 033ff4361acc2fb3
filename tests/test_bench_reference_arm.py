"""`bench.py --impl reference` (the CPU arm the driver times beside the GPU arm): prints ONE JSON line with the
contract's keys on rank 0, nothing on the other ranks, and needs no GPU.  It times the reference's own numba decoder
(baseline/_ref, staged by __graft_entry__.build(): kind "reference") when that is present and numba imports, else the
oracle port on the host cores (kind "port") — one of the two places outside tests/ allowed to execute oracle/."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                           "--warmup", "1"], capture_output=True, text=True, cwd=ROOT, env=env, timeout=300)


def test_reference_arm_line_on_rank0():
    res = _run({})
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "turbo_info_throughput_N212_R1/3_8it" and d["unit"] == "Gbit/s"
    assert d["steps"] == 2 and d["warmup"] == 1 and d["higher_is_better"] is True and d["gpu_launches"] == 0
    assert d["value"] > 0 and d["ms_per_step"] > 0
    cb = d["cpu_baseline"]
    have_ref = os.path.exists(os.path.join(ROOT, "baseline", "_ref", "dvb_rcs2_turbo.py"))
    assert cb["kind"] == ("reference" if have_ref else "port"), cb
    assert cb["cores"] >= 1 and cb["value"] == d["value"] and "N=212" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "Gbit/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_do_no_work():
    res = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert res.returncode == 0, res.stderr[-2000:]
    assert not [l for l in res.stdout.splitlines() if l.startswith("{")]
