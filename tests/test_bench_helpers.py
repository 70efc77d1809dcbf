"""CPU-side pieces of bench.py (no GPU): the CPU baseline legs and the clock sampler degrade gracefully."""
import json
import os
import subprocess
import sys
import time

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def bench():
    sys.path.insert(0, ROOT)
    import bench as b
    return b


def test_reference_sample_times_the_references_own_loop(bench, maps):
    from oracle.pyoracle import RefBackend
    if not RefBackend.available(fast=True):
        pytest.skip("oracle/_ref/libhitl_ref_fast.so not built")
    g = maps("small")
    r = bench.ref_sample(g, steps=1)
    assert r["kind"] == "reference" and r["unit"] == bench.UNIT and r["cores"] >= 1 and r["value"] > 0
    assert "JointOpt::FindSTFCorrespondences" in r["sample"] and r["seconds"] > 0
    # the sample is FIXED by the map size alone (same in every run, at every N) and covers the full map's targets
    a, b = bench.RefSampler(g), bench.RefSampler(g)
    assert a.stride == b.stride == 1 and np.array_equal(a.ids, b.ids) and a.queries == b.queries > 0
    big = {"poses": np.zeros((5000, 3), np.float32), "offsets": np.arange(5001, dtype=np.uint64) * 715}
    assert max(1, int(round(5000 * float(big["offsets"][-1]) / bench.REF_SAMPLE_UNIT))) == 16


def test_cpu_legs_use_every_host_thread_even_under_torchrun(bench, monkeypatch):
    monkeypatch.setenv("OMP_NUM_THREADS", "1")                 # what torch.distributed.run exports to its workers
    n = bench.use_all_host_threads()
    assert n == bench.host_threads() >= 1 and os.environ["OMP_NUM_THREADS"] == str(n)


def test_port_sample_and_reference_arm_line(bench, maps, tmp_path):
    g = maps("small")
    r = bench.cpu_sample(g, seconds=0.3)
    assert r["kind"] == "port" and r["value"] > 0 and r["search_Mq_per_s"] > 0
    # the reference arm prints exactly one JSON line with the contract's keys (small map, short steps)
    env = dict(os.environ, HITL_SYNTH_DIR=str(tmp_path))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c2", "--poses", "120", "--beams", "90", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == bench.UNIT and d["higher_is_better"] is True
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["value"] == d["value"] and d["config"]["n_poses"] == 120
    assert d["cpu_baseline"]["cores"] == bench.host_threads() and d["scaling"] == "strong" and d["config"]["parallelism"] == "single GPU"
    # torchrun's OMP_NUM_THREADS=1 must not reach the reference arm, and rank 0 of an N-rank launch reports the same sample
    env8 = dict(env, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="8", LOCAL_RANK="0")
    out8 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "8", "--workload", "c2", "--poses", "120", "--beams", "90", "--steps", "1", "--warmup", "0"],
                          capture_output=True, text=True, env=env8, timeout=600)
    assert out8.returncode == 0, out8.stderr[-2000:]
    d8 = json.loads([l for l in out8.stdout.splitlines() if l.strip()][0])
    assert d8["cpu_baseline"]["cores"] == bench.host_threads() and d8["n_gpus"] == 8
    assert d8["cpu_baseline"]["sample"].split(" in ")[0] == d["cpu_baseline"]["sample"].split(" in ")[0]      # same fixed sample, same counts


def test_reference_arm_other_ranks_exit_quietly(tmp_path):
    env = dict(os.environ, RANK="3", WORLD_SIZE="8", LOCAL_RANK="3", HITL_SYNTH_DIR=str(tmp_path))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "8"], capture_output=True, text=True, env=env, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_reference_correction_baseline(bench, maps):
    from hitl_slam_b200 import synth
    from oracle.pyoracle import RefBackend
    if not RefBackend.available(fast=True):
        pytest.skip("oracle/_ref/libhitl_ref_fast.so not built")
    g = maps("small", drift_xy=0.012, drift_th=0.004)
    r = bench.ref_correction_cpu(g, synth.pick_strokes(g, min_sep=0.045))
    assert r["ms"] > 0 and r["human_blocks"] > 0 and abs(r["ms"] - (r["ms_world"] + r["ms_em"] + r["ms_correct_backprop_blocks"])) < 1e-6


def test_clock_sampler_without_a_gpu_returns_an_empty_record(bench):
    s = bench.ClockSampler(0, period_s=0.01)
    s.start()
    time.sleep(0.1)
    s.mark_begin()
    time.sleep(0.05)
    c = s.stop()
    assert set(c) >= {"sm_mhz", "sm_max_mhz", "reasons", "samples", "source"}
    assert c["samples"] == 0 or c["sm_mhz"] > 0          # no NVML / nvidia-smi here: nothing sampled, nothing invented
