"""GPU parity against the REFERENCE'S OWN code (-m gpu): the sm_100a kernels, called through the C ABI / the C++ host mirror,
against oracle/_ref/libhitl_ref.so — JointOptimization.cpp, EMinput.cpp, ApplyExplicitCorrection.cpp, Backprop.cpp, HitLSLAM.cpp
and kdtree.cpp compiled from /root/reference where they lie (stand-in Eigen/Ceres/glog/CImg headers: oracle/shim3).  The prebuilt
library travels to the GPU box; nothing here reads /root/reference.  Correspondence sets, observation sets, world clouds, trees'
answers and the back-propagated poses: bit-exact.  Residuals / Jacobians: <= 1e-9 relative (FP64), <= 1e-5 (FP32 mode).
Solved poses: tolerance written in the test."""
import os

import numpy as np
import pytest

from conftest import assert_same_stf, random_scans
from oracle.pyoracle import RefBackend

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not RefBackend.available(), reason="oracle/_ref/libhitl_ref.so not built (needs /root/reference at build time)")]

REL64, REL32 = 1e-9, 1e-5
DRIFTY = dict(drift_xy=0.012, drift_th=0.004)


@pytest.fixture(scope="module")
def ref():
    return RefBackend()


def same_bits(a, b):
    a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
    return a.shape == b.shape and bool(((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))).all())


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300) if b.size else 0.0


def load_map(gpu, g):
    gpu.set_scans(g["offsets"], g["pts"], g["nrm"])
    gpu.build_kdtrees()


def jittered(g, seed):
    return g["poses"].astype(np.float64) + np.random.default_rng(seed).normal(size=g["poses"].shape) * [0.02, 0.02, 0.01]


@pytest.mark.parametrize("name,normals", [("tiny", "compensated"), ("tiny", "faithful"), ("small", "compensated"), ("small", "faithful")])
def test_find_stf_equals_reference_FindSTFCorrespondences(gpu, ref, maps, name, normals):
    g = maps(name, normals=normals)
    load_map(gpu, g)
    J = ref.joint_opt(g["offsets"], g["pts"], g["nrm"], g["poses"])
    for poses in (g["poses"].astype(np.float64), jittered(g, 1)):
        want = J.find_stf(poses)
        assert_same_stf(gpu.find_stf(poses), want)
        assert_same_stf(gpu.find_stf(poses, opts=gpu.stf_opts(disable_culling=1)), want)


@pytest.mark.parametrize("opts", [dict(cap=1), dict(cap=3, skip=2), dict(skip=5), dict(thr=0.05), dict(thr=0.4, cap=2)])
def test_find_stf_options_equal_reference(gpu, ref, maps, opts):
    g = maps("small")
    load_map(gpu, g)
    J = ref.joint_opt(g["offsets"], g["pts"], g["nrm"], g["poses"])
    J.set_options(**opts)
    poses = jittered(g, 2)
    for lo, hi in ((0, None), (20, 90)):
        assert_same_stf(gpu.find_stf(poses, min_pose=lo, max_pose=hi, opts=gpu.stf_opts(**opts)), J.find_stf(poses, min_pose=lo, max_pose=hi if hi is not None else len(poses) - 1))


def test_find_stf_c1_full_size_equals_reference(gpu, ref, maps):
    # BASELINE config 1 (500 poses x 360 beams, 90 M queries) through the reference's own OpenMP loop on the box's host cores
    g = maps("c1")
    load_map(gpu, g)
    poses = g["poses"].astype(np.float64)
    J = ref.joint_opt(g["offsets"], g["pts"], g["nrm"], g["poses"])
    want = J.find_stf(poses)
    assert len(want["pair_i"]) > 1000
    assert_same_stf(gpu.find_stf(poses), want)


def test_find_stf_random_scans_with_ties_equal_reference(gpu, ref):
    rng = np.random.default_rng(3)
    off, pts, nrm = random_scans(rng, 24, 1, 200)          # no empty scans: the reference reads an uninitialised tree for those
    poses32 = (rng.normal(size=(24, 3)) * 0.05).astype(np.float32)
    poses = poses32.astype(np.float64) + rng.normal(size=(24, 3)) * 1e-3
    gpu.set_scans(off, pts, nrm)
    gpu.build_kdtrees()
    J = ref.joint_opt(off, pts, nrm, poses32)
    J.set_options(thr=0.3, cap=4)
    assert_same_stf(gpu.find_stf(poses, opts=gpu.stf_opts(thr=0.3, cap=4)), J.find_stf(poses))
    # the device-built trees answer single queries like the reference's BuildKDTrees trees
    for scan in (0, 7, 23):
        q = (rng.normal(size=(2000, 2)) * 2.0).astype(np.float32)
        q[::4] = pts[off[scan]:off[scan + 1]][rng.integers(0, off[scan + 1] - off[scan], len(q[::4]))]
        for mode in (0, 1):
            for thr in (0.05, 0.5):
                d0, i0 = J.kd_query(scan, q, thr, mode)
                d1, i1 = gpu.kd_query(scan, q, thr, mode)
                assert np.array_equal(i0, i1) and same_bits(d0, d1), (scan, mode, thr)


def test_find_vo_equals_reference(gpu, ref, maps):
    g = maps("small")
    load_map(gpu, g)
    J = ref.joint_opt(g["offsets"], g["pts"], g["nrm"], g["poses"])
    for poses, lo, hi in ((g["poses"].astype(np.float64), 0, 159), (jittered(g, 4), 20, 60), (jittered(g, 4), 5, 5)):
        for a, b in zip(J.find_vo(poses, lo, hi), gpu.find_vo(poses, lo, hi)):
            assert np.array_equal(a, b)


def test_relative_pose_and_world_clouds_equal_reference(gpu, ref, maps):
    g = maps("small")
    load_map(gpu, g)
    poses = jittered(g, 5)
    poses[3, 2] = 3.14159; poses[4, 2] = -3.14159; poses[5, 2] = 100.25; poses[6] = [1e3, -2e3, -7.5]
    J = ref.joint_opt(g["offsets"], g["pts"], g["nrm"], g["poses"])
    J.set_pose_array(poses)
    rng = np.random.default_rng(6)
    src, dst = rng.integers(0, len(poses), 4000).astype(np.uint32), rng.integers(0, len(poses), 4000).astype(np.uint32)
    assert same_bits(gpu.debug_relative_pose(poses, src, dst), J.relative_pose(src, dst))
    p32 = g["poses"].copy()
    p32[:, 2] += np.float32(0.37)
    J.set_poses(p32)
    assert same_bits(gpu.world_transform(p32), J.world_clouds())


def test_observation_sets_equal_reference(gpu, ref, maps):
    from hitl_slam_b200 import synth
    g = maps("small", **DRIFTY)
    load_map(gpu, g)
    world = gpu.world_transform(g["poses"])
    strokes = synth.pick_strokes(g, min_sep=0.045)
    rng = np.random.default_rng(7)
    for trial in range(5):
        s = strokes + (rng.normal(size=strokes.shape) * 0.01 * trial).astype(np.float32)
        want, got = ref.em_observation_sets(g["offsets"], world, s), gpu.em_assign(s)
        for f in range(2):
            for a, b in zip(want[f], got[f]):
                assert np.array_equal(a, b), (trial, f)


def test_verify_input_equals_reference(gpu, ref, oracle, host, maps):
    from hitl_slam_b200 import HostSession, synth
    from test_oracle_ref_backend import _verify_cases
    g = maps("small", **DRIFTY)
    load_map(gpu, g)
    world = gpu.world_transform(g["poses"])
    sess = ref.session(g["offsets"], g["pts"], g["nrm"], g["poses"])
    strokes = synth.pick_strokes(g, min_sep=0.045)
    for sel in _verify_cases(g, world, strokes):
        got, mask = gpu.verify_input(sel)
        assert got == sess.verify(4, sel), sel                    # HitLSLAM::verifyUserInput on the reference's own code
        assert (got, mask) == oracle.verify_input(g["offsets"], world, sel)
    # odd point counts and fewer than four selected points (the reference only ever passes four)
    off, pts, nrm = random_scans(np.random.default_rng(5), 9, 1, 40, empty=(2,))
    gpu.set_scans(off, pts, nrm)
    gpu.build_kdtrees()
    w = gpu.world_transform(np.zeros((9, 3), np.float32))
    for k in (1, 2, 3, 8):
        sel = np.concatenate([w[-1:], w[:1], w[3:4] + np.float32(10.0), w[5:6] + np.float32(0.01), w[-1:] + np.float32(0.2), w[1:4]])[:k]
        assert gpu.verify_input(sel) == oracle.verify_input(off, w, sel)
    # the session drops an unverified input like HitLSLAM::replayLog does
    s = HostSession(gpu, host)
    try:
        s.set_map(g["poses"], g["offsets"], g["pts"], g["nrm"])
        s.world_transform(keep_host_copy=False)
        assert s.verify_input(strokes) == 4
        bad = strokes.copy(); bad[0] = [999, 999]
        out = s.correct(4, bad, solve=False, verify=True)
        assert not out["applied"] and out["verified"] is False
    finally:
        s.close()


def test_residual_blocks_equal_the_blocks_the_reference_builds(gpu, ref, host, maps):
    """GPU evaluation of odometry / human / STF blocks against AutoDiffCostFunction over the reference's own functors, in the blocks
    JointOpt::AddOdometryConstraints / AddHumanConstraints / AddSTFConstraints build (constants frozen by the reference code)."""
    g = maps("small")
    load_map(gpu, g)
    n = len(g["poses"])
    rng = np.random.default_rng(8)
    m = 64
    ids = np.stack([rng.choice([2, 4, 5, 6], m), rng.integers(0, n, m), rng.integers(0, n, m)], 1).astype(np.int32)
    deltas = (rng.normal(size=(m, 4)) * 2).astype(np.float32)
    J = ref.joint_opt(g["offsets"], g["pts"], g["nrm"], g["poses"])
    J.set_human_constraints([(ids, deltas)])
    poses = g["poses"].astype(np.float64)
    corr = J.find_stf(poses)
    x = jittered(g, 9)
    r_odo, J_odo, _ = J.eval_blocks(0, x, n)
    r_hum, J_hum, nr = J.eval_blocks(1, x, m)
    r_stf, J_stf, _ = J.eval_blocks(2, x, len(corr["pair_i"]))
    gpu.find_stf(poses)
    gpu.set_odometry_blocks(host.odometry_consts(g["poses"]))
    tg = host.human_targets(g["poses"], ids, deltas)
    gpu.set_human_blocks(np.stack([ids[:, 0], ids[:, 1]], 1).astype(np.int32), tg)
    gpu.set_stf_blocks_from_search()
    for precision, tol in ((0, REL64), (1, REL32)):
        out = gpu.eval(x, precision=precision)
        assert rel_err(out["r_odometry"], r_odo) <= tol and rel_err(out["J_odometry"].reshape(len(r_odo), -1), J_odo) <= tol
        assert rel_err(out["r_stf"], r_stf) <= tol and rel_err(out["J_stf"].reshape(len(r_stf), -1), J_stf) <= tol
        for b in range(m):
            k = nr[b]
            assert rel_err(out["r_human"][b, :k], r_hum[b, :k]) <= tol and np.abs(out["J_human"][b, :k].reshape(-1) - J_hum[b, :3 * k]).max() <= tol, b


def test_back_propagation_kernel_equals_reference(gpu, ref, host):
    rng = np.random.default_rng(10)
    for n, lo, hi in ((60, 5, 50), (700, 100, 650), (2500, 0, 2499), (9, 3, 5)):
        poses = np.cumsum(rng.normal(size=(n, 3)) * [0.25, 0.25, 0.03], 0).astype(np.float32)
        cov = np.zeros((n, 9), np.float32)
        cov[:, 0] = cov[:, 4] = 1e-4 * (1 + np.arange(n) / 100) * rng.uniform(0.5, 1.5, n)
        cov[:, 8] = 1e-5 * (1 + np.arange(n) / 100)
        c3 = np.array([0.31, -0.22, 0.07], np.float32)
        want_p, want_c = ref.backprop(poses, cov, lo, hi, c3)
        got_p, got_c, _ = host.backprop(gpu, poses, cov, lo, hi, c3)
        assert same_bits(got_p, want_p) and same_bits(got_c, want_c), n


@pytest.mark.parametrize("name", ["small", "c1"])
def test_whole_correction_equals_HitLSLAM_replayLog(gpu, ref, host, maps, name, monkeypatch):
    """One replayed colinear correction (BASELINE config 1 is the c1 case) through the product — C++ host mirror session over the GPU
    context — and through the reference's HitLSLAM::replayLog.  The refit stroke endpoints (SegFitEM) agree to 1e-5; both solves are
    run to convergence (the reference's through HITL_SHIM_LM_TIGHT, the mirror's through its solver options), so the final poses are
    the same minimiser up to the endpoint difference and float32 storage: <= 5e-5."""
    from hitl_slam_b200 import HostSession, synth
    monkeypatch.setenv("HITL_SHIM_LM_TIGHT", "1")
    g = maps(name, **DRIFTY)
    n = len(g["poses"])
    strokes = synth.pick_strokes(g, min_sep=0.045)
    cov0 = np.tile(np.array([1e-4, 0, 0, 0, 1e-4, 0, 0, 0, 1e-5], np.float32), (n, 1))
    sess = ref.session(g["offsets"], g["pts"], g["nrm"], g["poses"], cov0)
    assert sess.verify(4, strokes) == 4 and sess.replay(4, strokes) == 1
    p_ref, cov_ref, w_ref = sess.state()
    hc_i, hc_f = sess.constraints(0)

    s = HostSession(gpu, host)
    try:
        s.set_map(g["poses"], g["offsets"], g["pts"], g["nrm"])
        s.world_transform(keep_host_copy=False)
        s.solver_options(0, max_iterations=2000, function_tolerance=1e-16, gradient_tolerance=1e-14, parameter_tolerance=1e-14)
        cov = cov0.copy()
        out = s.correct(4, strokes, cov=cov, solve=True)
        got, _ = s.poses()
    finally:
        s.close()
    em_ref = ref.em_run(g["offsets"], gpu.world_transform(g["poses"]), strokes)
    assert out["applied"] and out["n_corrected"] == len(em_ref["corrected"]) and out["n_anchor"] == len(em_ref["anchor"]) and out["backprop"] == em_ref["backprop"]
    assert np.abs(out["segs"] - em_ref["segs"]).max() <= 1e-5
    assert out["n_constraints"] == len(hc_i)
    assert np.abs(cov - cov_ref).max() <= 1e-6 * np.abs(cov_ref).max()     # covariances: same recurrences, endpoints differ by <= 1e-5
    assert np.abs(got - p_ref).max() <= 5e-5
    assert np.abs(gpu.world_transform(got) - w_ref).max() <= 5e-4


@pytest.mark.parametrize("name", ["tiny", "small"])
def test_reference_jointopt_with_the_hot_path_bound_to_the_c_abi(gpu, ref, maps, name):
    """The reference's own JointOpt::PostHumanOptimization with BuildKDTrees / FindSTFCorrespondences / FindVisualOdometryCorrespondences
    re-bound to the C ABI (oracle/_ref/libhitl_ref_dropin.so) against the same function of the unmodified CPU library: same blocks, same
    consecutive-pose matches, bit-identical optimised poses and gradient (everything after the search is the same code on the same inputs)."""
    from oracle.pyoracle import RefDropin
    if not RefDropin.available():
        pytest.skip("oracle/_ref/libhitl_ref_dropin.so not built")
    g = maps(name)
    n = len(g["poses"])
    want = ref.joint_opt(g["offsets"], g["pts"], g["nrm"], g["poses"]).post_human_optimization()
    d = RefDropin()
    h = d.create(gpu.ctx, g["offsets"], g["pts"], g["nrm"], g["poses"])
    try:
        got = d.post_human_optimization(h, n)
    finally:
        d.destroy(h)
    assert (got["n_blocks"], got["n_matches"], got["n_vo"]) == (want["n_blocks"], want["n_matches"], want["n_vo"])
    assert got["termination"] == want["termination"]
    assert np.array_equal(got["pose_array"], want["pose_array"]) and np.array_equal(got["gradient"], want["gradient"])


def test_reference_jointopt_with_gpu_cost_blocks(gpu, ref, maps):
    """The north-star claim, literally: the reference's own JointOpt builds its STF problem with GPU-backed cost blocks (AddSTFConstraints
    re-bound, one batched hitl_eval per evaluation point behind SizedCostFunction<2,3,3>::Evaluate) and evaluates / solves it through the
    Ceres-shaped API.  Problem::Evaluate at a perturbed point: cost, residuals and gradient within 1e-9 of the reference's Jet blocks.
    PostHumanOptimization end to end: the optimised poses agree to 1e-6 (two evaluations that differ at 1e-12 through 100 LM iterations)."""
    from oracle.pyoracle import RefDropin
    if not RefDropin.available(blocks=True):
        pytest.skip("oracle/_ref/libhitl_ref_dropin_blocks.so not built")
    g = maps("tiny")
    n = len(g["poses"])
    x = jittered(g, 31)
    J = ref.joint_opt(g["offsets"], g["pts"], g["nrm"], g["poses"])
    cost0, res0, grad0 = J.evaluate_stf_problem(x, 4096)
    d = RefDropin(blocks=True)
    h = d.create(gpu.ctx, g["offsets"], g["pts"], g["nrm"], g["poses"])
    try:
        cost1, res1, grad1 = d.evaluate_stf_problem(h, x, 4096)
        assert d.lib.dropin_last_batches() == 1              # one batched hitl_eval served every block of the evaluation
        assert res1.shape == res0.shape and rel_err(res1, res0) <= REL64 and abs(cost1 - cost0) <= REL64 * cost0 and rel_err(grad1, grad0) <= REL64
        want = J.post_human_optimization(g["poses"].astype(np.float64))
        got = d.post_human_optimization(h, n, g["poses"].astype(np.float64))
    finally:
        d.destroy(h)
    assert (got["n_blocks"], got["n_matches"], got["n_vo"]) == (want["n_blocks"], want["n_matches"], want["n_vo"])
    assert np.abs(got["pose_array"] - want["pose_array"]).max() <= 1e-6
