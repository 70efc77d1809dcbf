"""Pins the oracle's restated residual functors and E-step distance to the REFERENCE's own code:
residual_functors.h and eigen_helper.h compiled where they lie (oracle/_ref/libfunctors_ref.so,
built by oracle/Makefile when /root/reference is present; the prebuilt library travels to the GPU box)."""
import numpy as np
import pytest

from oracle.pyoracle import RefFunctors

pytestmark = pytest.mark.skipif(not RefFunctors.available(), reason="oracle/_ref/libfunctors_ref.so not built (needs /root/reference)")

TOL = 1e-12   # same Jet arithmetic on both sides; only the accumulation order inside a block may differ


def close(a, b, tol=TOL):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    s = max(np.abs(b).max() if b.size else 0.0, 1e-300)
    return np.abs(a - b).max() <= tol * max(s, 1.0) if a.size else True


@pytest.fixture(scope="module")
def ref():
    return RefFunctors()


def test_point_to_point_glob_functor(oracle, ref, maps):
    g = maps("tiny")
    S = oracle.scans(g["offsets"], g["pts"], g["nrm"])
    poses = g["poses"].astype(np.float64)
    corr = S.find_stf(poses)
    x = poses + np.random.default_rng(0).normal(size=poses.shape) * 0.01
    r, J = S.eval_stf(x, corr)
    off = g["offsets"].astype(np.int64)
    rng = np.random.default_rng(1)
    for b in rng.integers(0, len(corr["pair_i"]), 60):
        i, j = int(corr["pair_i"][b]), int(corr["pair_j"][b])
        m0, m1 = int(corr["pair_off"][b]), int(corr["pair_off"][b + 1])
        k, q = corr["k"][m0:m1].astype(np.int64), corr["idx"][m0:m1].astype(np.int64)
        rr, j0, j1, rp = ref.p2p_glob(g["pts"][off[i] + k], g["pts"][off[j] + q], g["nrm"][off[i] + k], g["nrm"][off[j] + q], 0.05, 1.0 / 40.0, x[i], x[j])
        assert close(r[b], rr) and close(J[b, 0], j0) and close(J[b, 1], j1)
        assert close(rr, rp)                             # T = double evaluation of the same functor


def test_point_to_point_glob_zero_sum_stays_zero(ref):
    # identical clouds and poses: every point residual is exactly 0 -> residual 0 with zero Jacobian (residual_functors.h:822-827)
    p = np.array([[1.0, 2.0], [3.0, -1.0]], np.float32)
    n = np.array([[1.0, 0.0], [0.0, 1.0]], np.float32)
    x = np.array([0.5, -0.25, 0.3])
    rr, j0, j1, _ = ref.p2p_glob(p, p, n, n, 0.05, 0.025, x, x)
    assert np.all(rr == 0) and np.all(j0 == 0) and np.all(j1 == 0)


def test_pose_constraint_functor(oracle, ref, maps):
    g = maps("small")
    consts = oracle.odometry_consts(g["poses"])
    x = g["poses"].astype(np.float64) + np.random.default_rng(2).normal(size=g["poses"].shape) * 0.02
    r, J = oracle.eval_odometry(consts, x)
    for b in range(0, len(consts), 5):
        rr, j0, j1 = ref.pose_constraint(consts[b], x[b], x[b + 1])
        assert close(r[b], rr) and close(J[b, 0], j0) and close(J[b, 1], j1)


def test_human_imposed_constraint_functors(oracle, ref, maps):
    g = maps("small")
    n = len(g["poses"])
    rng = np.random.default_rng(3)
    m = 40
    ids = np.stack([rng.choice([2, 4, 5, 6], m), rng.integers(0, n, m), rng.integers(0, n, m)], 1).astype(np.int32)
    deltas = rng.normal(size=(m, 4)).astype(np.float32)
    blk_i, blk_d = oracle.human_blocks(g["poses"], ids, deltas)
    x = g["poses"].astype(np.float64) + rng.normal(size=g["poses"].shape) * 0.05
    r, J = oracle.eval_human(blk_i, blk_d, x)
    for b in range(m):
        rr, jj = ref.human(blk_i[b, 0], blk_d[b], x[blk_i[b, 1]])
        k = len(rr)
        assert k == {2: 3, 4: 2, 5: 1, 6: 1}[int(blk_i[b, 0])]
        assert close(r[b, :k], rr) and close(J[b, :k], jj)


def test_point_to_line_functors(oracle, ref, maps):
    g = maps("tiny")
    n = len(g["poses"])
    rng = np.random.default_rng(4)
    x = g["poses"].astype(np.float64) + rng.normal(size=(n, 3)) * 0.01
    sizes = rng.integers(1, 60, 12)
    blk_off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    m = int(blk_off[-1])
    blk_pose = rng.integers(0, n, 12).astype(np.uint32)
    pts = rng.normal(size=(m, 2)).astype(np.float32) * 3
    ang = rng.uniform(0, 6.28, m)
    ln = np.stack([np.cos(ang), np.sin(ang)], 1).astype(np.float32)
    lo = rng.normal(size=m).astype(np.float32)
    valid = (rng.uniform(size=m) > 0.2).astype(np.uint8)
    rg, Jg = oracle.eval_p2l_glob(blk_pose, blk_off, pts, ln, lo, valid, 0.05, 1 / 50.0, x)
    for b in range(12):
        a, e = int(blk_off[b]), int(blk_off[b + 1])
        rr, jj = ref.p2l_glob(pts[a:e], ln[a:e], lo[a:e], valid[a:e], 0.05, 1 / 50.0, x[blk_pose[b]])
        assert close(rg[b], rr) and close(Jg[b], jj)
    pose_idx = rng.integers(0, n, m).astype(np.uint32)
    rs, Js = oracle.eval_p2l(pose_idx, pts, ln, lo, valid, 0.05, 1 / 50.0, x)
    for b in range(0, m, 7):
        rr, jj = ref.p2l(pts[b], ln[b], lo[b], valid[b], 0.05, 1 / 50.0, x[pose_idx[b]])
        assert close(rs[b], rr) and close(Js[b], jj)


def test_distance_to_line_segment_decides_the_same_inliers(oracle, ref, maps):
    """E-step: the oracle's inlier list equals thresholding the reference's own DistanceToLineSegment."""
    from hitl_slam_b200 import synth
    g = maps("small")
    S = oracle.scans(g["offsets"], g["pts"], g["nrm"])
    world = S.world_transform(g["poses"])
    strokes = synth.make_strokes(g)
    long_seg = np.array([[1.0, 1.5], [6.5, 1.52]], np.float32)            # > 1 m: the "t > 1.0 metres" quirk is exercised
    degenerate = np.array([[2.0, 1.5], [2.0, 1.5]], np.float32)           # zero-length stroke: normalized() returns the zero vector
    for seg in (strokes[:2], strokes[2:], long_seg, degenerate):
        d = ref.distance_to_line_segment(seg[0], seg[1], world)
        want = np.nonzero(d.astype(np.float64) < 0.03)[0]
        op, oi = oracle.em_inliers(g["offsets"], world, seg.reshape(-1))
        got = g["offsets"][op].astype(np.int64) + oi
        assert np.array_equal(got, want)
    assert len(want) >= 0
