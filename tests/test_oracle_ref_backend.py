"""Pins the oracle restatement (and the host mirror's O(#poses) stages) to the REFERENCE's own back-end code:
JointOptimization.cpp, EMinput.cpp, ApplyExplicitCorrection.cpp, Backprop.cpp, HitLSLAM.cpp and kdtree.cpp compiled
where they lie against the stand-in headers of oracle/shim3 (oracle/_ref/libhitl_ref.so, built by oracle/Makefile when
/root/reference is present; the prebuilt library travels to the GPU box).  Integer / index / float32 outputs are compared
bit for bit; double residuals and Jacobians at 1e-12 (same Jet arithmetic, accumulation order inside a block may differ);
anything that passes through a Levenberg-Marquardt solve at the tolerance written in the test (two LM implementations)."""
import numpy as np
import pytest

from conftest import assert_same_stf, random_scans
from oracle.pyoracle import RefBackend, default_min_cos

pytestmark = pytest.mark.skipif(not RefBackend.available(), reason="oracle/_ref/libhitl_ref.so not built (needs /root/reference)")


@pytest.fixture(scope="module")
def ref():
    return RefBackend()


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def same_bits(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    return a.shape == b.shape and bool(((bits(a) == bits(b)) | (np.isnan(a) & np.isnan(b))).all())


def close(a, b, tol=1e-12):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    if a.size == 0:
        return b.size == 0
    return np.abs(a - b).max() <= tol * max(np.abs(b).max(), 1.0)


DRIFTY = dict(drift_xy=0.012, drift_th=0.004)   # enough odometry drift for a visible loop-closure error on the small maps


def jittered(g, seed, s_xy=0.02, s_th=0.01):
    rng = np.random.default_rng(seed)
    return g["poses"].astype(np.float64) + rng.normal(size=g["poses"].shape) * [s_xy, s_xy, s_th]


# ---- a6 / a7 / a8: the search loops ---------------------------------------------------------------------------------
def test_cosine_gate_constant(ref):
    # cos(kMaxStfAngleError) as JointOptimization.cpp:564 forms it == the float the oracle / C ABI are given
    assert ref.min_cos() == default_min_cos()


@pytest.mark.parametrize("name,normals", [("tiny", "compensated"), ("tiny", "faithful"), ("small", "compensated")])
def test_find_stf_is_the_references_own_loop(oracle, ref, maps, name, normals):
    g = maps(name, normals=normals)
    S = oracle.scans(g["offsets"], g["pts"], g["nrm"])
    J = ref.joint_opt(g["offsets"], g["pts"], g["nrm"], g["poses"])
    for poses in (g["poses"].astype(np.float64), jittered(g, 1)):
        want = J.find_stf(poses, with_points=True)
        got = S.find_stf(poses)
        assert len(want["k"]) > 1000
        assert_same_stf(got, want)
        # the point / normal copies every reference correspondence carries are the scan entries the indices name
        off = g["offsets"].astype(np.int64)
        bi = np.repeat(want["pair_i"].astype(np.int64), np.diff(want["pair_off"].astype(np.int64)))
        bj = np.repeat(want["pair_j"].astype(np.int64), np.diff(want["pair_off"].astype(np.int64)))
        xy = want["xy8"]
        assert np.array_equal(xy[:, 0], g["pts"][off[bi] + want["k"]]) and np.array_equal(xy[:, 1], g["pts"][off[bj] + want["idx"]])
        assert np.array_equal(xy[:, 2], g["nrm"][off[bi] + want["k"]]) and np.array_equal(xy[:, 3], g["nrm"][off[bj] + want["idx"]])


@pytest.mark.parametrize("opts", [dict(cap=1), dict(cap=3, skip=2), dict(skip=5), dict(thr=0.05), dict(thr=0.4, cap=2), dict(max_angle=np.float32(0.1))])
def test_find_stf_options(oracle, ref, maps, opts):
    g = maps("tiny")
    S = oracle.scans(g["offsets"], g["pts"], g["nrm"])
    J = ref.joint_opt(g["offsets"], g["pts"], g["nrm"], g["poses"])
    J.set_options(**opts)
    poses = jittered(g, 2)
    kw = {k: v for k, v in opts.items() if k in ("cap", "skip", "thr")}
    if "max_angle" in opts:
        kw["min_cos"] = ref.min_cos(opts["max_angle"])
    assert_same_stf(S.find_stf(poses, **kw), J.find_stf(poses))


@pytest.mark.parametrize("lo,hi", [(0, 39), (5, 20), (10, 10), (12, 13), (0, 1000), (30, 39)])
def test_find_stf_pose_ranges(oracle, ref, maps, lo, hi):
    g = maps("tiny")
    S = oracle.scans(g["offsets"], g["pts"], g["nrm"])
    J = ref.joint_opt(g["offsets"], g["pts"], g["nrm"], g["poses"])
    poses = g["poses"].astype(np.float64)
    assert_same_stf(S.find_stf(poses, min_pose=lo, max_pose=hi), J.find_stf(poses, min_pose=lo, max_pose=hi))


@pytest.mark.parametrize("seed", range(4))
def test_find_stf_random_ragged_scans_with_ties(oracle, ref, seed):
    # random clouds (duplicated coordinates -> tie cases of the tree build, 1-point scans), overlapping poses.  No EMPTY scans here:
    # the reference queries a default-constructed KDTree for them (JointOptimization.cpp:533-535), which reads uninitialised members;
    # the oracle / CUDA path define that case as "no match" (tests/test_gpu_parity.py covers it against the oracle).
    rng = np.random.default_rng(seed)
    off, pts, nrm = random_scans(rng, 14, 1, 90)
    poses32 = (rng.normal(size=(14, 3)) * [0.05, 0.05, 0.05]).astype(np.float32)
    poses = poses32.astype(np.float64) + rng.normal(size=(14, 3)) * 1e-3
    S = oracle.scans(off, pts, nrm)
    J = ref.joint_opt(off, pts, nrm, poses32)
    J.set_options(thr=0.3, cap=4)
    got, want = S.find_stf(poses, thr=0.3, cap=4), J.find_stf(poses)
    assert_same_stf(got, want)
    J.set_options(thr=0.3, cap=64, max_angle=np.float32(3.0))
    assert_same_stf(S.find_stf(poses, thr=0.3, cap=64, min_cos=ref.min_cos(np.float32(3.0))), J.find_stf(poses))


def test_find_vo_is_the_references_own_loop(oracle, ref, maps):
    g = maps("small")
    S = oracle.scans(g["offsets"], g["pts"], g["nrm"])
    J = ref.joint_opt(g["offsets"], g["pts"], g["nrm"], g["poses"])
    for poses, lo, hi in ((g["poses"].astype(np.float64), 0, None), (jittered(g, 3), 0, None), (jittered(g, 4), 17, 90), (jittered(g, 4), 5, 5), (jittered(g, 4), 100, 5000)):
        a, b = S.find_vo(poses, min_pose=lo, max_pose=hi), J.find_vo(poses, min_pose=lo, max_pose=hi)
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
    assert len(b[0]) > 0


def test_relative_pose_transform_bits(ref, host, maps):
    g = maps("small")
    rng = np.random.default_rng(5)
    poses = jittered(g, 5)
    poses[3, 2] = 3.14159; poses[4, 2] = -3.14159; poses[5, 2] = 100.25; poses[6] = [1e3, -2e3, -7.5]
    J = ref.joint_opt(g["offsets"], g["pts"], g["nrm"], g["poses"])
    J.set_pose_array(poses)
    src, dst = rng.integers(0, len(poses), 1500).astype(np.uint32), rng.integers(0, len(poses), 1500).astype(np.uint32)
    src[:8], dst[:8] = [3, 4, 5, 6, 3, 5, 6, 0], [4, 3, 6, 5, 0, 3, 4, 6]
    want = J.relative_pose(src, dst)
    flat = poses.reshape(-1)
    got = np.stack([host.relative_pose(flat, int(a), int(b)) for a, b in zip(src, dst)])
    assert same_bits(got, want)


def test_trees_answer_like_the_references(oracle, ref):
    rng = np.random.default_rng(6)
    off, pts, nrm = random_scans(rng, 6, 1, 300)
    S = oracle.scans(off, pts, nrm)
    J = ref.joint_opt(off, pts, nrm, np.zeros((6, 3), np.float32))
    for scan in range(6):
        q = (rng.normal(size=(500, 2)) * 2.0).astype(np.float32)
        q[::3] = pts[off[scan]:off[scan + 1]][rng.integers(0, off[scan + 1] - off[scan], len(q[::3]))]    # exact hits (distance 0 -> early exit)
        for mode in (0, 1):
            for thr in (0.05, 0.5, 10.0):
                d0, i0 = S.query(scan, q, thr, mode)
                d1, i1 = J.kd_query(scan, q, thr, mode)
                assert np.array_equal(i0, i1) and same_bits(d0, d1), (scan, mode, thr)


# ---- f1: world-frame clouds -----------------------------------------------------------------------------------------
def test_world_clouds_bits(oracle, ref, maps):
    g = maps("small")
    S = oracle.scans(g["offsets"], g["pts"], g["nrm"])
    poses = g["poses"].copy()
    poses[:, 2] += np.float32(0.37)
    J = ref.joint_opt(g["offsets"], g["pts"], g["nrm"], poses)
    want = J.world_clouds()                                   # JointOpt::CopyTempLaserScans
    assert same_bits(S.world_transform(poses), want)
    sess = ref.session(g["offsets"], g["pts"], g["nrm"], poses)
    assert same_bits(sess.state()[2], want)                   # HitLSLAM::transformPointCloudsToWorldFrame gives the same bits


def _verify_cases(g, world, strokes):
    rng = np.random.default_rng(21)
    far = np.array([[500, 500], [501, 500], [-300, 2], [-300, 3]], np.float32)
    one_off = strokes.copy(); one_off[2] = [400, -400]
    degenerate_a = strokes.copy(); degenerate_a[1] = degenerate_a[0]
    degenerate_b = strokes.copy(); degenerate_b[3] = degenerate_b[2]
    on_points = world[rng.integers(0, len(world), 4)].copy()
    near = on_points + np.float32(0.03)                         # 0.042 m away: inside the 0.05 m selection radius
    edge = on_points + np.array([0.05, 0.0], np.float32)        # right at the radius: decided by float rounding
    return [strokes, far, one_off, degenerate_a, degenerate_b, on_points, near, edge]


def test_verify_user_input_is_the_references(oracle, ref, maps):
    from hitl_slam_b200 import synth
    g = maps("small", **DRIFTY)
    S = oracle.scans(g["offsets"], g["pts"], g["nrm"])
    world = S.world_transform(g["poses"])
    sess = ref.session(g["offsets"], g["pts"], g["nrm"], g["poses"])
    seen = set()
    for sel in _verify_cases(g, world, synth.pick_strokes(g, min_sep=0.045)):
        want = sess.verify(4, sel)                                # HitLSLAM::verifyUserInput
        got, _ = oracle.verify_input(g["offsets"], world, sel)
        assert got == want, sel
        seen.add(want)
    assert {0, 4} <= seen


# ---- a15 - a19: the residual blocks the reference's Add*Constraints build -------------------------------------------
def test_odometry_blocks_built_by_the_reference(oracle, ref, maps):
    g = maps("small")
    poses32 = g["poses"].copy()
    poses32[7] = poses32[6]; poses32[7, 2] += np.float32(0.3)            # a pair that did not move: the heading-axes branch
    poses32[20, 2] = np.float32(3.1); poses32[21, 2] = np.float32(-3.1)   # wrap-around of the measured rotation
    J = ref.joint_opt(g["offsets"], g["pts"], g["nrm"], poses32)
    x = poses32.astype(np.float64) + np.random.default_rng(7).normal(size=poses32.shape) * 0.02
    r1, J1, nr = J.eval_blocks(0, x, len(x))
    r0, J0 = oracle.eval_odometry(oracle.odometry_consts(poses32), x)
    assert len(r1) == len(x) - 1 and np.all(nr == 3)
    assert close(r0, r1) and close(J0.reshape(len(r0), -1), J1)


def test_human_blocks_built_by_the_reference(oracle, ref, maps):
    g = maps("small")
    n = len(g["poses"])
    rng = np.random.default_rng(8)
    m = 96
    ids = np.stack([rng.choice([2, 4, 5, 6], m), rng.integers(0, n, m), rng.integers(0, n, m)], 1).astype(np.int32)
    deltas = (rng.normal(size=(m, 4)) * 2).astype(np.float32)
    J = ref.joint_opt(g["offsets"], g["pts"], g["nrm"], g["poses"])
    J.set_human_constraints([(ids[:40], deltas[:40]), (ids[40:], deltas[40:])])
    x = jittered(g, 9)
    r1, J1, nr = J.eval_blocks(1, x, m)
    blk_i, blk_d = oracle.human_blocks(g["poses"], ids, deltas)
    r0, J0 = oracle.eval_human(blk_i, blk_d, x)
    assert len(r1) == m
    assert np.array_equal(nr, np.array([{2: 3, 4: 2, 5: 1, 6: 1}[t] for t in ids[:, 0]]))
    for b in range(m):
        k = nr[b]
        assert close(r0[b, :k], r1[b, :k]) and close(J0[b, :k].reshape(-1), J1[b, :3 * k]), b


def test_stf_blocks_built_by_the_reference(oracle, ref, maps):
    g = maps("tiny")
    S = oracle.scans(g["offsets"], g["pts"], g["nrm"])
    J = ref.joint_opt(g["offsets"], g["pts"], g["nrm"], g["poses"])
    poses = g["poses"].astype(np.float64)
    corr = J.find_stf(poses)
    x = jittered(g, 10)
    r1, J1, nr = J.eval_blocks(2, x, len(corr["pair_i"]))     # AddSTFConstraints over point_point_glob_correspondences_
    r0, J0 = S.eval_stf(x, corr)
    assert len(r1) == len(corr["pair_i"]) and np.all(nr == 2)
    assert close(r0, r1) and close(J0.reshape(len(r0), -1), J1)


# ---- a9 - a14: EM ---------------------------------------------------------------------------------------------------
def test_dist_to_line_seg_bits(oracle, ref):
    rng = np.random.default_rng(11)
    p1, p2, p = rng.normal(size=(3, 3000, 2)).astype(np.float32)
    p2[:50] = p1[:50]                                         # zero-length stroke: 0/0 -> NaN compares false twice -> projection branch
    p[50:100] = p1[50:100]
    for k in list(range(100)) + list(range(100, 3000, 7)):
        seg = np.array([p1[k], p2[k]], np.float32).reshape(-1)
        want = ref.dist_to_line_seg(p1[k], p2[k], p[k])
        got = oracle.lib.orc_dist_to_line_seg(seg, float(p[k, 0]), float(p[k, 1]))
        assert (np.isnan(want) and np.isnan(got)) or want == got, k


def test_observation_sets_are_the_references(oracle, ref, maps):
    from hitl_slam_b200 import synth
    g = maps("small", **DRIFTY)
    S = oracle.scans(g["offsets"], g["pts"], g["nrm"])
    w = S.world_transform(g["poses"])
    segs = synth.pick_strokes(g, min_sep=0.045)
    rng = np.random.default_rng(12)
    for trial in range(6):
        s = segs + (rng.normal(size=segs.shape) * 0.01 * trial).astype(np.float32)
        a, b = oracle.em_assign(g["offsets"], w, s), ref.em_observation_sets(g["offsets"], w, s)
        for f in range(2):
            assert all(np.array_equal(x, y) for x, y in zip(a[f], b[f])), (trial, f)
        assert len(b[0][0]) > 0 and len(b[1][0]) > 0


@pytest.mark.parametrize("name", ["small", "c1"])
def test_em_run_is_the_references(oracle, ref, maps, name):
    from hitl_slam_b200 import synth
    g = maps(name, **DRIFTY)
    S = oracle.scans(g["offsets"], g["pts"], g["nrm"])
    w = S.world_transform(g["poses"])
    segs = synth.pick_strokes(g, min_sep=0.045)
    for s in (segs, segs[[2, 3, 0, 1]]):                      # drawn late-visit first ("user was good") and the other way round (swapped by the reference)
        a, b = oracle.em_run(g["offsets"], w, s), ref.em_run(g["offsets"], w, s)
        assert np.abs(a["segs"] - b["segs"]).max() <= 1e-5    # endpoints pass through SegFitEM: two LM implementations
        assert np.array_equal(a["corrected"], b["corrected"]) and np.array_equal(a["anchor"], b["anchor"])
        assert a["backprop"] == b["backprop"] and b["backprop"][0] >= 0
        assert len(b["corrected"]) > 0 and len(b["anchor"]) > 0


@pytest.mark.parametrize("seed", range(6))
def test_seg_fit_em_is_the_references(oracle, ref, host, seed):
    rng = np.random.default_rng(seed)
    ang = rng.uniform(0.02, 1.5)
    n = int(rng.integers(8, 400))
    t = rng.uniform(-0.2, 2.2, n)
    data = np.stack([1 + t * np.cos(ang), 2 + t * np.sin(ang)], 1) + rng.normal(size=(n, 2)) * 0.01
    p1 = np.array([1.0, 2.0]) + rng.normal(size=2) * 0.03
    p2 = np.array([1 + 2 * np.cos(ang), 2 + 2 * np.sin(ang)]) + rng.normal(size=2) * 0.03
    want = ref.seg_fit(p1, p2, data)                          # reference SegFitEM over the stand-in LM
    assert np.abs(oracle.seg_fit(p1, p2, data) - want).max() <= 1e-5
    assert np.abs(host.seg_fit_em(p1, p2, data) - want).max() <= 1e-5


# ---- f3: explicit correction, constraint targets, back-propagation --------------------------------------------------
@pytest.mark.parametrize("ctype", [2, 4, 5, 6])
def test_explicit_correction_is_the_references(oracle, ref, host, ctype):
    rng = np.random.default_rng(200 + ctype)
    for trial in range(20):
        n = int(rng.integers(8, 120))
        poses = np.cumsum(rng.normal(size=(n, 3)) * [0.3, 0.3, 0.05], 0).astype(np.float32)
        a0 = rng.normal(size=2) * 5
        b0 = a0 + rng.normal(size=2) * 0.4
        da, db = rng.normal(size=2), rng.normal(size=2)
        if trial % 5 == 0:
            db = np.array([-da[1], da[0]])
        sel = np.array([a0, a0 + da, b0, b0 + db], np.float32)
        k = int(rng.integers(1, max(2, n // 3)))
        corrected = np.sort(rng.choice(np.arange(n // 2, n), min(k, n - n // 2), replace=False)).astype(np.int32)
        if trial % 3 == 0:
            corrected = (np.arange(n // 2, n // 2 + len(corrected))).astype(np.int32)
        anchor = np.sort(rng.choice(np.arange(0, n // 2), int(rng.integers(1, max(2, n // 4))), replace=False)).astype(np.int32)
        want_p, want_c, hc_i, hc_f = ref.app_exp_run(ctype, sel, poses, corrected, anchor)
        got_p, got_c = oracle.app_exp_corrections(ctype, sel, poses, corrected)
        assert same_bits(got_p, want_p) and same_bits(got_c, want_c), (ctype, trial)
        hp, hcorr = host.app_exp_correct(ctype, sel, poses, corrected)
        assert same_bits(hp, want_p) and same_bits(hcorr, want_c)
        # calculateConstraintTargets: |anchor| x |corrected| constraints, anchor-major, deltas frozen from the CORRECTED poses
        assert len(hc_i) == len(anchor) * len(corrected)
        assert np.array_equal(hc_i[:, 0], np.full(len(hc_i), ctype)) and np.array_equal(hc_i[:, 2], np.repeat(anchor, len(corrected)))
        assert np.array_equal(hc_i[:, 1], np.tile(corrected, len(anchor)))
        ti, tf = host.constraint_targets(ctype, sel, want_p, corrected, anchor)
        assert np.array_equal(ti, hc_i) and same_bits(tf, hc_f), (ctype, trial)


def test_back_propagation_is_the_references(oracle, ref):
    rng = np.random.default_rng(13)
    for trial in range(12):
        n = int(rng.integers(6, 400))
        poses = np.cumsum(rng.normal(size=(n, 3)) * [0.3, 0.3, 0.05], 0).astype(np.float32)
        cov = np.abs(rng.normal(size=(n, 9)) * 1e-3).astype(np.float32) + np.float32(1e-5)
        lo = int(rng.integers(0, n - 3))
        hi = int(rng.integers(lo + 2, n))
        c3 = (rng.normal(size=3) * [0.3, 0.3, 0.1]).astype(np.float32)
        want_p, want_c = ref.backprop(poses, cov, lo, hi, c3)
        got_p, got_c = oracle.backprop(poses, cov, lo, hi, c3)
        assert same_bits(got_p, want_p) and same_bits(got_c, want_c), trial


# ---- a20 + the chain: one whole correction on the reference's HitLSLAM::replayLog ------------------------------------
def _scipy_minimiser(oracle, poses32, hc_i, hc_f):
    """Independent solve (scipy LM to machine precision) of the odometry + human problem over the oracle's Jet functors."""
    from scipy.optimize import least_squares
    from scipy.sparse import lil_matrix
    n = len(poses32)
    consts = oracle.odometry_consts(poses32)
    blk_i, blk_d = oracle.human_blocks(poses32, hc_i, hc_f)
    nres_h = [{2: 3, 4: 2, 5: 1, 6: 1}[int(t)] for t in hc_i[:, 0]]
    x0 = poses32.astype(np.float64)

    def unpack(v):
        return np.concatenate([x0.reshape(-1)[:3], v]).reshape(n, 3)

    def res(v):
        x = unpack(v)
        r_o, _ = oracle.eval_odometry(consts, x, want_jac=False)
        r_h, _ = oracle.eval_human(blk_i, blk_d, x, want_jac=False)
        return np.concatenate([r_o.reshape(-1)] + [r_h[b, :k] for b, k in enumerate(nres_h)])

    def jac(v):
        x = unpack(v)
        _, J_o = oracle.eval_odometry(consts, x)
        _, J_h = oracle.eval_human(blk_i, blk_d, x)
        J = lil_matrix((3 * (n - 1) + sum(nres_h), 3 * n))
        for b in range(n - 1):
            J[3 * b:3 * b + 3, 3 * b:3 * b + 3] = J_o[b, 0]
            J[3 * b:3 * b + 3, 3 * b + 3:3 * b + 6] = J_o[b, 1]
        row = 3 * (n - 1)
        for b, k in enumerate(nres_h):
            p = int(blk_i[b, 1])
            J[row:row + k, 3 * p:3 * p + 3] = J_h[b, :k]
            row += k
        return J.tocsr()[:, 3:].toarray()
    sol = least_squares(res, x0.reshape(-1)[3:].copy(), jac=jac, method="lm", xtol=1e-15, ftol=1e-15, gtol=1e-15, max_nfev=4000)
    return unpack(sol.x), sol.cost


def test_whole_correction_chain_against_hitlslam(oracle, ref, host, maps, monkeypatch):
    """HitLSLAM::replayLog (verify -> EMInput -> AppExpCorrect -> Backprop -> angle wrap -> JointOpt::Run) on the reference's own
    code vs the same chain composed from the oracle's restated stages.  Everything up to the solve is bit-exact once both sides see
    the same stroke endpoints (those pass through SegFitEM: 1e-5).  The reference's solve is run to convergence (HITL_SHIM_LM_TIGHT)
    and compared with an independent minimiser over the oracle's functors at 2e-6 (the reference stores the result in float32)."""
    from hitl_slam_b200 import synth
    monkeypatch.setenv("HITL_SHIM_LM_TIGHT", "1")
    g = maps("small", **DRIFTY)
    segs = synth.pick_strokes(g, min_sep=0.045)
    cov = np.tile(np.array([1e-4, 0, 0, 0, 1e-4, 0, 0, 0, 1e-5], np.float32), (len(g["poses"]), 1))
    sess = ref.session(g["offsets"], g["pts"], g["nrm"], g["poses"], cov)
    assert sess.verify(4, segs) == 4
    assert sess.replay(4, segs) == 1                           # one group of human constraints was added: the correction was applied
    p_ref, cov_ref, w_ref = sess.state()
    hc_i, hc_f = sess.constraints(0)

    S = oracle.scans(g["offsets"], g["pts"], g["nrm"])
    em = oracle.em_run(g["offsets"], S.world_transform(g["poses"]), segs)
    em_ref = ref.em_run(g["offsets"], S.world_transform(g["poses"]), segs)
    assert np.array_equal(em["corrected"], em_ref["corrected"]) and np.array_equal(em["anchor"], em_ref["anchor"]) and em["backprop"] == em_ref["backprop"]
    assert np.abs(em["segs"] - em_ref["segs"]).max() <= 1e-5
    # from here on feed both sides the reference's refit endpoints, so that the remaining stages can be compared bit for bit
    p1, c3 = oracle.app_exp_corrections(4, em_ref["segs"], g["poses"], em["corrected"])
    ti, tf = host.constraint_targets(4, em_ref["segs"], p1, em["corrected"], em["anchor"])
    assert np.array_equal(ti, hc_i) and same_bits(tf, hc_f)
    p2, cov2 = oracle.backprop(p1, cov, em["backprop"][0], em["backprop"][1], c3)
    assert same_bits(cov2, cov_ref)
    p2[:, 2] = np.arctan2(np.sin(p2[:, 2]), np.cos(p2[:, 2]))            # float32 in, float32 out: HitLSLAM.cpp:365-369
    # JointOpt::Run on those poses, solved independently
    want, cost = _scipy_minimiser(oracle, p2, hc_i, hc_f)
    wrapped = want.copy()
    wrapped[:, 2] -= 2 * np.pi * np.rint(wrapped[:, 2] / (2 * np.pi))
    assert np.abs(p_ref - wrapped).max() <= 2e-6
    assert same_bits(S.world_transform(p_ref), w_ref)          # the session's world clouds are the transform of its final poses
    assert np.abs(p_ref - g["poses"]).max() > 1e-3             # and the correction did move the map


def test_sequential_corrections_against_hitlslam(oracle, ref, host, maps, monkeypatch):
    """BASELINE config 4 in miniature: several corrections replayed one after the other on the reference's HitLSLAM session, each drawn
    on the map as the previous ones left it and each adding a group of human constraints that the next joint optimisation keeps.  After
    every correction the oracle chain (fed with the reference's refit endpoints) reproduces the constraint targets and covariances bit
    for bit and an independent minimiser over ALL constraint groups so far reproduces the session's poses (1e-5)."""
    from hitl_slam_b200 import synth
    monkeypatch.setenv("HITL_SHIM_LM_TIGHT", "1")
    g = maps("c1", **DRIFTY)
    n = len(g["poses"])
    cov = np.tile(np.array([1e-4, 0, 0, 0, 1e-4, 0, 0, 0, 1e-5], np.float32), (n, 1))
    sess = ref.session(g["offsets"], g["pts"], g["nrm"], g["poses"], cov)
    S = oracle.scans(g["offsets"], g["pts"], g["nrm"])
    cur, start, groups, applied = dict(g), 0, [], 0
    for c in range(3):
        try:
            segs, start = synth.pick_strokes(cur, min_sep=0.045 if c == 0 else 0.01, start=start, return_next=True)
        except RuntimeError:
            break                                              # the remaining walls have no visible gap any more
        world = S.world_transform(cur["poses"])
        assert same_bits(world, sess.state()[2])               # the session's displayed clouds are the transform of its poses
        em = oracle.em_run(g["offsets"], world, segs)
        em_ref = ref.em_run(g["offsets"], world, segs)
        assert np.array_equal(em["corrected"], em_ref["corrected"]) and np.array_equal(em["anchor"], em_ref["anchor"]) and em["backprop"] == em_ref["backprop"]
        n_groups = sess.replay(4, segs)
        if not (em_ref["backprop"][0] >= 0 and em_ref["backprop"][1] >= 1):
            assert n_groups == len(groups)                     # rejected input (overlapping selections): nothing changes (HitLSLAM.cpp:332)
            continue
        assert n_groups == len(groups) + 1
        p_ref, cov_ref, _ = sess.state()
        hc_i, hc_f = sess.constraints(len(groups))
        p1, c3 = oracle.app_exp_corrections(4, em_ref["segs"], cur["poses"], em["corrected"])
        ti, tf = host.constraint_targets(4, em_ref["segs"], p1, em["corrected"], em["anchor"])
        assert np.array_equal(ti, hc_i) and same_bits(tf, hc_f), c
        p2, cov2 = oracle.backprop(p1, cov, em["backprop"][0], em["backprop"][1], c3)
        assert same_bits(cov2, cov_ref), c
        p2[:, 2] = np.arctan2(np.sin(p2[:, 2]), np.cos(p2[:, 2]))
        groups.append((hc_i, hc_f))
        all_i, all_f = np.concatenate([x[0] for x in groups]), np.concatenate([x[1] for x in groups])
        want, _ = _scipy_minimiser(oracle, p2, all_i, all_f)   # every group so far constrains the solve (JointOptimization.cpp:973-975)
        wrapped = want.copy()
        wrapped[:, 2] -= 2 * np.pi * np.rint(wrapped[:, 2] / (2 * np.pi))
        assert np.abs(p_ref - wrapped).max() <= 1e-5, c        # 500-pose chain, two solvers, float32 storage of the result
        cur = dict(cur); cur["poses"] = p_ref
        cov = cov_ref
        applied += 1
    assert applied >= 2


# ---- the committed golden fixtures are what the reference's own code produces ---------------------------------------
def test_golden_fixtures_equal_the_references_output(ref):
    """tests/golden/*.npz were generated from the oracle (tests/golden/make_golden.py); here the reference's own JointOpt /
    EMInput reproduce them from the fixture inputs, so the fixtures pin the CUDA path to the reference even on a box where
    oracle/_ref is absent."""
    import os
    from conftest import ROOT
    from hitl_slam_b200 import synth
    gd = os.path.join(ROOT, "tests", "golden")
    m = np.load(os.path.join(gd, "tiny_compensated.npz"))
    want = np.load(os.path.join(gd, "tiny_stf.npz"))
    J = ref.joint_opt(m["offsets"], m["pts"], m["nrm"], m["poses"])
    got = J.find_stf(m["poses"].astype(np.float64))
    for k in ("pair_i", "pair_j", "pair_off", "k", "idx"):
        assert np.array_equal(got[k], want[k]), k
    ev = np.load(os.path.join(gd, "tiny_eval.npz"))
    r_stf, J_stf, _ = J.eval_blocks(2, ev["x"], len(got["pair_i"]))
    r_odo, J_odo, _ = J.eval_blocks(0, ev["x"], len(m["poses"]))
    assert close(ev["r_stf"], r_stf) and close(ev["J_stf"].reshape(len(r_stf), -1), J_stf)
    assert close(ev["r_odo"], r_odo) and close(ev["J_odo"].reshape(len(r_odo), -1), J_odo)
    em = np.load(os.path.join(gd, "small_em.npz"))
    g = synth.generate("small")
    world = ref.joint_opt(g["offsets"], g["pts"], g["nrm"], g["poses"]).world_clouds()
    sets = ref.em_observation_sets(g["offsets"], world, em["strokes"])
    for f in range(2):
        assert np.array_equal(sets[f][0], em["set%d_pose" % f]) and np.array_equal(sets[f][1], em["set%d_off" % f]) and np.array_equal(sets[f][2], em["set%d_obs" % f])


# ---- drop-in demonstration: the reference's JointOpt bound to the product's C ABI -----------------------------------
def test_dropin_library_routes_the_hot_path_into_the_c_abi(maps):
    """oracle/_ref/libhitl_ref_dropin.so = the reference's JointOptimization.cpp with BuildKDTrees / FindSTFCorrespondences /
    FindVisualOdometryCorrespondences weakened and re-defined as C-ABI calls (oracle/ref_dropin_capi.cpp).  Without a GPU context the
    replaced BuildKDTrees must be the one that runs — and fail cleanly in hitl_set_scans, not fall back to the CPU trees."""
    from oracle.pyoracle import RefDropin
    if not RefDropin.available():
        pytest.skip("oracle/_ref/libhitl_ref_dropin.so not built")
    g = maps("tiny")
    with pytest.raises(RuntimeError, match="hitl_set_scans"):
        RefDropin().create(None, g["offsets"], g["pts"], g["nrm"], g["poses"])


def test_reference_post_human_optimization_runs_on_the_cpu_library(ref, maps):
    """The CPU side of the drop-in comparison (tests/test_gpu_vs_reference.py): PostHumanOptimization of the reference's own code."""
    g = maps("tiny")
    J = ref.joint_opt(g["offsets"], g["pts"], g["nrm"], g["poses"])
    r = J.post_human_optimization()
    want = J.find_stf(g["poses"].astype(np.float64))             # the search ran at the initial poses
    assert r["n_blocks"] == len(want["pair_i"]) and r["n_matches"] == len(want["k"]) and r["n_vo"] > 0
    assert len(r["gradient"]) == 3 * len(g["poses"]) and np.all(np.isfinite(r["pose_array"]))
    assert np.array_equal(r["pose_array"][0], g["poses"][0].astype(np.float64))      # pose 0 is held constant (:1197)
    assert np.abs(r["pose_array"] - g["poses"]).max() > 1e-4     # the solve moved the others


def test_reference_stf_problem_evaluation_matches_the_oracle(oracle, ref, maps):
    """Problem::Evaluate of the STF problem the reference builds (cost, residuals, gradient over poses) against the oracle's blocks: the CPU
    side of the cost-block drop-in comparison."""
    g = maps("tiny")
    x = jittered(g, 31)
    J = ref.joint_opt(g["offsets"], g["pts"], g["nrm"], g["poses"])
    cost, res, grad = J.evaluate_stf_problem(x, 4096)
    S = oracle.scans(g["offsets"], g["pts"], g["nrm"])
    corr = S.find_stf(x)
    r, Jb = S.eval_stf(x, corr)
    assert close(res, r) and abs(cost - 0.5 * (r ** 2).sum()) <= 1e-12 * cost
    want = np.zeros_like(x)
    for b in range(len(r)):
        i, j = int(corr["pair_i"][b]), int(corr["pair_j"][b])
        want[i] += Jb[b, 0].T @ r[b]
        want[j] += Jb[b, 1].T @ r[b]
    want[0] = 0.0                                              # pose 0 is constant (:1197)
    assert close(grad, want, 1e-10)
    from oracle.pyoracle import RefDropin
    if RefDropin.available(blocks=True):
        assert RefDropin(blocks=True).lib.dropin_has_gpu_blocks() == 1 and RefDropin().lib.dropin_has_gpu_blocks() == 0


def test_source_samples_of_a_full_map(oracle, ref, maps):
    """bench.py's reference arm times the reference's own FindSTFCorrespondences on a SAMPLE of source poses against all targets of the full
    map (RefJointOpt.restrict_sources parks the other poses' point_clouds_g_ entries; the loop itself is untouched).  The sampled search
    must return exactly the rows of the full search whose source pose is in the sample, and the oracle's strided search (which supplies
    the executed-query count) must agree with both."""
    g = maps("small")
    n = len(g["poses"])
    poses = g["poses"].astype(np.float64)
    J = ref.joint_opt(g["offsets"], g["pts"], g["nrm"], g["poses"])
    full = J.find_stf(poses)
    full = {k: np.array(v) for k, v in full.items()}
    S = oracle.scans(g["offsets"], g["pts"], g["nrm"])
    for stride, phase in ((7, 0), (16, 5), (1, 0)):
        ids = np.arange(phase, n, stride)
        J.restrict_sources(ids)
        part = J.find_stf(poses)
        keep = np.isin(full["pair_i"], ids)
        assert np.array_equal(part["pair_i"], full["pair_i"][keep]) and np.array_equal(part["pair_j"], full["pair_j"][keep])
        cnt = np.diff(full["pair_off"].astype(np.int64))
        rows = np.repeat(keep, cnt)
        assert np.array_equal(part["k"], full["k"][rows]) and np.array_equal(part["idx"], full["idx"][rows])
        port = S.find_stf(poses, src_lo=phase, src_stride=stride)
        assert_same_stf(port, part)
        per_pose = [S.find_stf(poses, src_lo=int(i), src_hi=int(i) + 1)["n_queries"] for i in ids[:5]]
        assert port["n_queries"] >= sum(per_pose)
    J.restrict_sources(None)
    assert_same_stf(J.find_stf(poses), full)
    whole = S.find_stf(poses)
    assert sum(S.find_stf(poses, src_lo=p, src_stride=4)["n_queries"] for p in range(4)) == whole["n_queries"]
