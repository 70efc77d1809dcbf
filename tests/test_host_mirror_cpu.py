"""CPU-side checks of the C++ host mirror (no GPU): the Ceres-shaped solver, SegFitEM, the float
problem-building arithmetic, the session-log format, exported symbols, and the 2-rank sharding
logic over gloo."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_library_exports_every_declared_symbol(host):
    hdr = open(os.path.join(ROOT, "hitl_slam_b200", "host", "hitl_host.h")).read()
    names = sorted(set(re.findall(r"\b(hitl_host_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(host.lib, n)]
    assert not missing, missing


def test_solver_powell_dense_cg_and_constant_block(host):
    x0 = [3.0, -1.0, 0.0, 1.0]
    for cg in (False, True):
        x, s = host.solver_selftest(x0, force_cg=cg)
        assert s["initial_cost"] == pytest.approx(107.5)
        assert s["final_cost"] < 1e-15 and np.abs(x).max() < 1e-4 and s["termination"] == 0
    x, s = host.solver_selftest(x0, hold_x1=True)
    assert x[0] == 3.0                                  # SetParameterBlockConstant honoured
    assert s["final_cost"] < s["initial_cost"]
    # cross-check the constrained minimum with scipy
    from scipy.optimize import least_squares

    def f(v):
        x1, (x2, x3, x4) = 3.0, v
        return [x1 + 10 * x2, np.sqrt(5) * (x3 - x4), (x2 - 2 * x3) ** 2, np.sqrt(10) * (x1 - x4) ** 2]
    ref = least_squares(f, x0[1:], xtol=1e-15, ftol=1e-15, gtol=1e-15)
    assert np.abs(x[1:] - ref.x).max() < 1e-6


def test_problem_evaluate_keeps_constant_blocks_as_zero_columns(host):
    """Ceres' Problem::Evaluate (ProblemEvaluateTest.ConstantParameterBlock): a constant block keeps its gradient entries and Jacobian
    columns, filled with zeros — the reference's consumers index gradients_ / ceres_jacobian_ by 3 * pose with pose 0 held constant."""
    x0 = np.array([3.0, -1.0, 0.5, 1.0])
    r = np.array([x0[0] + 10 * x0[1], np.sqrt(5) * (x0[2] - x0[3]), (x0[1] - 2 * x0[2]) ** 2, np.sqrt(10) * (x0[0] - x0[3]) ** 2])
    J = np.array([[1, 10, 0, 0], [0, 0, np.sqrt(5), -np.sqrt(5)], [0, 2 * (x0[1] - 2 * x0[2]), -4 * (x0[1] - 2 * x0[2]), 0],
                  [2 * np.sqrt(10) * (x0[0] - x0[3]), 0, 0, -2 * np.sqrt(10) * (x0[0] - x0[3])]])
    g, dims, cost = host.evaluate_selftest(x0)
    assert len(g) == 4 and dims == (4, 4, 8) and cost == pytest.approx(0.5 * (r ** 2).sum()) and np.allclose(g, J.T @ r, rtol=1e-12)
    g, dims, cost = host.evaluate_selftest(x0, hold_x1=True)
    want = J.T @ r
    want[0] = 0.0
    assert len(g) == 4 and dims[:2] == (4, 4) and dims[2] == 6 and g[0] == 0.0 and np.allclose(g, want, rtol=1e-12)


def test_problem_bookkeeping_off_the_chain_shape(host):
    """The stand-in Problem indexes parameter blocks by bisection while they arrive in ascending address order (the pose chain) and
    by hash map afterwards; cost functions shared between residual blocks are deleted once; block lists longer than two spill."""
    import ctypes as C
    lib = host.lib
    lib.hitl_host_problem_selftest.argtypes = [np.ctypeslib.ndpointer(np.float64, flags="C")]
    lib.hitl_host_problem_selftest.restype = C.c_int
    out = np.zeros(5)
    assert lib.hitl_host_problem_selftest(out) == 0
    r = np.array([2.75, 5.5, -5.5])
    assert out[0] == pytest.approx(0.5 * (r ** 2).sum(), rel=1e-15) and out[1] == 6 and out[2] == 3
    assert out[3] == 0.0        # the constant block keeps a zero gradient entry
    assert out[4] == 0          # both cost functions died with the problem, the shared one once


def test_solver_chain_direct_matches_dense_and_cg(host):
    """Beyond the dense limit the stand-in LM eliminates block-tridiagonal systems (the odometry chain + unary human factors)
    directly; same minimiser as the dense path and as PCG."""
    import ctypes as C
    lib = host.lib
    lib.hitl_host_solver_chain_selftest.argtypes = [np.ctypeslib.ndpointer(np.float64, flags="C"), C.c_int, C.c_int, np.ctypeslib.ndpointer(np.float64, flags="C")]
    for n in (2, 3, 40, 400):
        rng = np.random.default_rng(n)
        x0 = np.cumsum(rng.normal(size=(n, 2)) * 0.3, 0).reshape(-1)
        sols, costs = [], []
        for mode in (0, 1, 2):
            x, out = x0.copy(), np.zeros(4)
            assert lib.hitl_host_solver_chain_selftest(x, n, mode, out) == 0
            assert out[1] <= out[0] and out[3] in (0.0, 1.0)
            sols.append(x); costs.append(out[1])
        assert np.abs(sols[1] - sols[0]).max() <= 1e-7 and np.abs(sols[2] - sols[0]).max() <= 1e-5
        assert abs(costs[1] - costs[0]) <= 1e-10 * max(1.0, costs[0])
        assert np.array_equal(sols[1][:2], x0[:2])                # the constant block stays put


@pytest.mark.parametrize("seed", range(6))
def test_seg_fit_em_matches_oracle(host, oracle, seed):
    rng = np.random.default_rng(seed)
    ang = rng.uniform(0.02, 1.5)
    n = int(rng.integers(8, 400))
    t = rng.uniform(-0.2, 2.2, n)
    data = np.stack([1 + t * np.cos(ang), 2 + t * np.sin(ang)], 1) + rng.normal(size=(n, 2)) * 0.01
    p1 = np.array([1.0, 2.0]) + rng.normal(size=2) * 0.03
    p2 = np.array([1 + 2 * np.cos(ang), 2 + 2 * np.sin(ang)]) + rng.normal(size=2) * 0.03
    a, b = host.seg_fit_em(p1, p2, data), oracle.seg_fit(p1, p2, data)
    assert np.abs(a - b).max() <= 1e-5                  # two LM implementations, float endpoints: tolerance-level
    assert np.allclose((a[0] + a[1]) / 2, (p1 + p2) / 2, atol=1e-6)   # midpoint and length are kept
    assert np.linalg.norm(a[0] - a[1]) == pytest.approx(np.linalg.norm(p1 - p2), abs=1e-5)


def test_seg_fit_em_without_inliers_keeps_the_initial_angle(host, oracle):
    p1, p2 = np.array([0.0, 0.0]), np.array([1.0, -1.0])   # negative slope: acos(|dx|/h) drops the sign (reference quirk)
    a, b = host.seg_fit_em(p1, p2, np.zeros((0, 2))), oracle.seg_fit(p1, p2, np.zeros((0, 2)))
    assert np.array_equal(a, b)
    assert a[0, 1] > a[1, 1]                                # endpoints come back on the positive-slope diagonal


def test_odometry_constants_bit_exact(host, oracle, maps):
    g = maps("small")
    poses = g["poses"].copy()
    poses[7] = poses[6]                                   # a pair that did not move: the heading-axes branch
    poses[7, 2] += np.float32(0.3)
    poses[20, 2] = np.float32(3.1)
    poses[21, 2] = np.float32(-3.1)                       # wrap-around of the measured rotation
    assert np.array_equal(host.odometry_consts(poses), oracle.odometry_consts(poses))


def test_human_targets_bit_exact(host, oracle, maps):
    g = maps("small")
    n = len(g["poses"])
    rng = np.random.default_rng(5)
    m = 64
    ids = np.stack([rng.choice([2, 4, 5, 6], m), rng.integers(0, n, m), rng.integers(0, n, m)], 1).astype(np.int32)
    deltas = rng.normal(size=(m, 4)).astype(np.float32) * 2
    blk_i, blk_d = oracle.human_blocks(g["poses"], ids, deltas)
    tg = host.human_targets(g["poses"], ids, deltas)
    assert np.array_equal(blk_i[:, 0], ids[:, 0]) and np.array_equal(blk_i[:, 1], ids[:, 1])
    assert np.array_equal(tg, blk_d)


def test_session_log_round_trip_and_format(host, tmp_path):
    entries = [(4, 0, [[1.23456, 2.0], [3.0, 4.00004], [5.5, 6.5], [7.25, -8.125]]),
               (5, 1, [[0, 0], [1, 1], [2, 2], [3, 3]]),
               (1, 0, [[9, 9], [8, 8]]),
               (3, 0, [[i, -i] for i in range(8)]),
               (2, 0, [[0.1, 0.2], [0.3, 0.4], [0.5, 0.6], [0.7, 0.8]])]
    p = str(tmp_path / "session_logged.log")
    host.save_log(p, entries)
    text = open(p).read().split("\n")
    assert text[0] == "5 " and text[1] == "4, 0" and text[2] == "1.2346, 2.0000"   # "%d \n", "%d, %d\n", "%.4f, %.4f\n"
    back = host.load_log(p)
    assert [(t, u, len(x)) for t, u, x in back] == [(4, 0, 4), (5, 1, 4), (1, 0, 2), (3, 0, 8), (2, 0, 4)]
    assert np.allclose(back[0][2], np.round(np.array(entries[0][2]), 4), atol=1e-6)
    with pytest.raises(IOError):
        host.load_log(str(tmp_path / "missing.log"))


def test_stfs_covars_io_threads_and_fast_parser(host, tmp_path, monkeypatch):
    """The parallel reader / writer (HitLSLAM_main.cpp:192-300, vector_mapping_main.cpp:1855-1928 mirrors) give the
    same bytes and the same arrays as the single-threaded path; the Clinger fast path equals strtof (= fscanf %f);
    a malformed line stops the load where the reference's fscanf loop would."""
    rng = np.random.default_rng(7)
    n = 300
    poses = (rng.normal(size=(n, 3)) * 20).astype(np.float32)
    cov = np.abs(rng.normal(size=(n, 9)) * 1e-4).astype(np.float32)
    cnt = rng.integers(0, 900, n)
    cnt[5] = 0
    off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.uint32)
    obs = (rng.normal(size=(off[-1], 2)) * 25).astype(np.float32)
    nrm = rng.normal(size=(off[-1], 2)).astype(np.float32)
    obs[::11] = 0
    obs[3] = [123456.7, -0.00004]           # rounds to 8 digits / to zero at 4 decimals
    obs[4] = [1.6777216e7, 3.3554432e7]     # mantissa >= 2^24: strtof path
    paths = {}
    for tag, threads in (("mt", "8"), ("st", "1")):
        monkeypatch.setenv("HITL_IO_THREADS", threads)
        paths[tag] = str(tmp_path / (tag + ".stfs.covars"))
        host.save_stfs_covars(paths[tag], poses, cov, off, obs, nrm, map_name="io", timestamp=12.5)
    a, b = open(paths["mt"], "rb").read(), open(paths["st"], "rb").read()
    assert a == b and len(a) > (1 << 20)
    text = a.decode().split("\n")
    assert text[0] == "io" and float(text[1]) == 12.5
    # python's float() is correctly rounded like strtof: check the parsed world-frame values through a pose with theta = 0
    loads = {}
    for tag, threads in (("mt", "8"), ("st", "1")):
        monkeypatch.setenv("HITL_IO_THREADS", threads)
        loads[tag] = host.load_pose_graph(paths["mt"])
    for k in ("poses", "offsets", "pts", "nrm"):
        assert np.array_equal(np.asarray(loads["mt"][k]).view(np.uint32), np.asarray(loads["st"][k]).view(np.uint32)), k
    g = loads["mt"]
    want_poses = np.array([[np.float32(x) for x in line.split(",")[:3]] for line in text[2:-1]], np.float32)
    change = np.ones(len(want_poses), bool)
    change[1:] = (want_poses[1:] != want_poses[:-1]).any(1)
    assert np.array_equal(g["poses"], want_poses[change])
    assert int(g["offsets"][-1]) == len(want_poses)
    # malformed line in the middle: everything from it on is ignored, in both modes
    lines = text[:]
    cutline = 2 + len(want_poses) // 2
    lines[cutline] = "oops"
    bad = str(tmp_path / "bad.stfs.covars")
    open(bad, "w").write("\n".join(lines))
    for threads in ("8", "1"):
        monkeypatch.setenv("HITL_IO_THREADS", threads)
        gb = host.load_pose_graph(bad)
        assert int(gb["offsets"][-1]) == cutline - 2
        assert np.array_equal(gb["pts"], g["pts"][:cutline - 2])


def test_pose_graph_binary_cache(host, tmp_path):
    """a binary cache of the parsed scans beside the .stfs.covars text file.  A cache hit returns exactly what the text
    parser returns; a cache that no longer matches the text file (content, size or mtime changed, truncated, foreign bytes) is ignored
    and rewritten."""
    import os
    rng = np.random.default_rng(17)
    n = 120
    poses = (rng.normal(size=(n, 3)) * 10).astype(np.float32)
    cov = np.abs(rng.normal(size=(n, 9)) * 1e-4).astype(np.float32)
    cnt = rng.integers(1, 200, n)
    off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.uint32)
    obs = (rng.normal(size=(off[-1], 2)) * 25).astype(np.float32)
    nrm = rng.normal(size=(off[-1], 2)).astype(np.float32)
    path = str(tmp_path / "map.stfs.covars")
    host.save_stfs_covars(path, poses, cov, off, obs, nrm, map_name="cache", timestamp=3.25)
    want = host.load_pose_graph(path)

    def same(a, b):
        return all(np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)) for k in ("poses", "cov", "offsets", "pts", "nrm"))

    first = host.load_pose_graph(path, cache=True)
    assert not first["from_cache"] and same(first, want) and os.path.exists(path + ".hitlcache")
    second = host.load_pose_graph(path, cache=True)
    assert second["from_cache"] and same(second, want)
    # explicit cache location
    other = str(tmp_path / "elsewhere.bin")
    assert not host.load_pose_graph(path, cache=other)["from_cache"] and host.load_pose_graph(path, cache=other)["from_cache"]
    # truncated cache: falls back to the text and repairs the cache
    blob = open(path + ".hitlcache", "rb").read()
    open(path + ".hitlcache", "wb").write(blob[:len(blob) // 2])
    third = host.load_pose_graph(path, cache=True)
    assert not third["from_cache"] and same(third, want)
    assert host.load_pose_graph(path, cache=True)["from_cache"]
    # foreign bytes
    open(path + ".hitlcache", "wb").write(b"not a cache at all" * 10)
    assert not host.load_pose_graph(path, cache=True)["from_cache"]
    # the text changes (one more scan): the old cache must not be served
    poses2 = np.concatenate([poses, poses[-1:] + np.float32(1.0)])
    cov2 = np.concatenate([cov, cov[-1:]])
    off2 = np.concatenate([off, [off[-1] + 7]]).astype(np.uint32)
    obs2 = np.concatenate([obs, (rng.normal(size=(7, 2)) * 5).astype(np.float32)])
    nrm2 = np.concatenate([nrm, rng.normal(size=(7, 2)).astype(np.float32)])
    host.save_stfs_covars(path, poses2, cov2, off2, obs2, nrm2, map_name="cache", timestamp=3.25)
    fresh = host.load_pose_graph(path, cache=True)
    assert not fresh["from_cache"] and len(fresh["poses"]) == n + 1 and same(fresh, host.load_pose_graph(path))
    assert host.load_pose_graph(path, cache=True)["from_cache"]
    with pytest.raises(IOError):
        host.load_pose_graph(str(tmp_path / "missing.stfs.covars"), cache=True)


def _same_bits(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    return a.shape == b.shape and bool(((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))).all())


@pytest.mark.parametrize("ctype", [2, 4, 5, 6])
def test_explicit_correction_bit_exact(host, oracle, ctype):
    """AppExpCorrect (ApplyExplicitCorrection.cpp:150-181, 229-316, 360-445): corrected poses move rigidly about feature A,
    later poses follow the last corrected pose, the first correction of the first contiguous group goes to Backprop."""
    rng = np.random.default_rng(100 + ctype)
    for trial in range(25):
        n = int(rng.integers(5, 120))
        poses = np.cumsum(rng.normal(size=(n, 3)) * [0.3, 0.3, 0.05], 0).astype(np.float32)
        a0 = rng.normal(size=2) * 5
        b0 = a0 + rng.normal(size=2) * 0.4
        da, db = rng.normal(size=2), rng.normal(size=2)
        if trial % 5 == 0:
            db = np.array([-da[1], da[0]])                       # exactly perpendicular strokes
        if trial % 7 == 0:
            db = da.copy()                                       # parallel strokes
        sel = np.array([a0, a0 + da, b0, b0 + db], np.float32)
        k = int(rng.integers(1, max(2, n // 2)))
        corrected = np.sort(rng.choice(n, k, replace=False))
        if trial % 3 == 0:
            corrected = np.arange(n // 3, n // 3 + k) % n        # one contiguous run
        if trial % 11 == 0:
            corrected = corrected[::-1]                          # unsorted list
        got_p, got_c = host.app_exp_correct(ctype, sel, poses, corrected)
        want_p, want_c = oracle.app_exp_corrections(ctype, sel, poses, corrected)
        # bit-exact; exactly parallel strokes can give A.B = 1 + 1 ulp and acosf -> NaN in the reference arithmetic too (any NaN matches)
        assert _same_bits(got_p, want_p), (ctype, trial)
        assert _same_bits(got_c, want_c)
    p, c = host.app_exp_correct(ctype, sel, poses, [])            # nothing to correct: poses untouched
    assert c is None and np.array_equal(p, poses)
    p, c = host.app_exp_correct(1, sel, poses, [1, 2])            # point correction: unsupported, as in the reference
    assert c is None and np.array_equal(p, poses)


def test_mirror_refuses_to_run_without_a_context(host):
    host._bind_mirror()
    assert not host.lib.hitl_host_session_create(None)     # no ctx, no session: there is no CPU implementation of the stages


# ---- 2-rank sharding over gloo (host logic of the multi-GPU path; the oracle stands in for the GPUs) ----
_WORKER = r'''
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from hitl_slam_b200 import synth
from hitl_slam_b200.sharding import shard_ranges, gather_stf, allreduce_normal_equations
from oracle.pyoracle import Oracle

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
g = synth.generate("tiny")
poses = g["poses"].astype(np.float64)
n = len(poses)
orc = Oracle()
S = orc.scans(g["offsets"], g["pts"], g["nrm"])
lo, hi = shard_ranges(g["offsets"], world)[rank]
local = S.find_stf(poses, src_lo=lo, src_hi=hi)
full = gather_stf(local)
x = poses + np.random.default_rng(0).normal(size=poses.shape) * 0.01

def packed(corr):
    r, J = S.eval_stf(x, corr)
    H, gv = np.zeros((n, 3, 3)), np.zeros((n, 3))
    for b in range(len(corr["pair_i"])):
        for side, p in ((0, int(corr["pair_i"][b])), (1, int(corr["pair_j"][b]))):
            H[p] += J[b, side].T @ J[b, side]
            gv[p] += J[b, side].T @ r[b]
    return np.concatenate([H.reshape(-1), gv.reshape(-1), [0.5 * (r ** 2).sum()]])

t = torch.from_numpy(packed(local))
allreduce_normal_equations(t)
if rank == 0:
    ref = S.find_stf(poses)
    for k in ("pair_i", "pair_j", "pair_off", "k", "idx"):
        assert np.array_equal(np.asarray(full[k]), np.asarray(ref[k])), k
    assert full["n_queries"] == ref["n_queries"]
    want = packed(ref)
    assert np.abs(t.numpy() - want).max() <= 1e-9 * np.abs(want).max()
    assert len(ref["pair_i"]) > 0 and hi > lo
    print("SHARDING_OK", len(ref["pair_i"]), world)
dist.destroy_process_group()
'''


def test_two_rank_sharding_over_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1", HITL_SYNTH_DIR=str(tmp_path))
    port = 29500 + (os.getpid() % 2000)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(script), ROOT], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
    assert "SHARDING_OK" in r.stdout


def test_shard_ranges_by_work():
    from hitl_slam_b200.sharding import shard_ranges_by_work
    work = np.array([0, 0, 10, 10, 10, 10, 0, 40, 0, 0], np.uint64)
    r = shard_ranges_by_work(work, 2)
    assert r[0][0] == 0 and r[-1][1] == len(work) and r[0][1] == r[1][0]
    assert abs(int(work[r[0][0]:r[0][1]].sum()) - int(work[r[1][0]:r[1][1]].sum())) <= 40
    r4 = shard_ranges_by_work(work, 4)
    assert [a for a, _ in r4][0] == 0 and all(r4[i][1] == r4[i + 1][0] for i in range(3)) and r4[-1][1] == len(work)
    assert shard_ranges_by_work(np.zeros(8), 4) == [(0, 2), (2, 4), (4, 6), (6, 8)]      # nothing measured yet: equal pose counts
    assert shard_ranges_by_work(np.zeros(0), 2) == [(0, 0), (0, 0)]
