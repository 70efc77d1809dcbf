"""GPU parity of the C++ host mirror (JointOpt / EMInput / GPU-backed Ceres cost blocks) against the
CPU oracle: block-by-block CostFunction::Evaluate, the search through the mirror, the EM stage, and
the optimised poses of both solves."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REL64 = 1e-9


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    s = np.abs(b).max() if b.size else 0.0
    return 0.0 if a.size == 0 else float(np.abs(a - b).max() / (s if s > 0 else 1.0))


@pytest.fixture()
def session(gpu, host):
    from hitl_slam_b200 import HostSession
    s = HostSession(gpu, host)
    yield s
    s.close()


def _constraints(n, poses=None):
    """type, constrained, anchor + deltas.  With poses: the relative pose the pair has now, nudged by a
    few centimetres / hundredths of a radian (what a real correction looks like); without: arbitrary."""
    ids = np.array([[2, n - 3, 1], [4, n - 4, 2], [5, n - 5, 3], [6, n - 6, 0], [4, n - 2, 5]], np.int32)
    deltas = np.array([[0.3, -0.2, 0.1, 0.0], [1.0, 0.5, -0.4, 1.2], [0, 0, 1.57, 0], [0, 0, 0.02, 0], [-0.7, 0.1, 3.0, -0.5]], np.float32)
    if poses is not None:
        rng = np.random.default_rng(11)
        for q, (_, c, a) in enumerate(ids):
            rel = poses[c, :2].astype(np.float64) - poses[a, :2]
            ca, sa = np.cos(poses[a, 2]), np.sin(poses[a, 2])
            d = poses[c, 2] - poses[a, 2]
            deltas[q] = [ca * rel[0] + sa * rel[1] + rng.normal() * 0.05, -sa * rel[0] + ca * rel[1] + rng.normal() * 0.05,
                         np.arctan2(np.sin(d), np.cos(d)) + rng.normal() * 0.02, 0.4 * q]
    return ids, deltas


def test_cost_functions_evaluate_like_ceres_would_call_them(session, oracle, maps):
    """Every residual block of the odometry + human + STF problem through CostFunction::Evaluate."""
    g = maps("tiny")
    n = len(g["poses"])
    session.set_map(g["poses"], g["offsets"], g["pts"], g["nrm"])
    ids, deltas = _constraints(n)
    session.add_constraints(ids, deltas)
    corr = session.find_stf()
    S = oracle.scans(g["offsets"], g["pts"], g["nrm"])
    ref = S.find_stf(g["poses"].astype(np.float64))
    for k in ("pair_i", "pair_j", "pair_off", "k", "idx"):
        assert np.array_equal(corr[k], ref[k]), k
    assert corr["n_queries"] == ref["n_queries"]
    x = g["poses"].astype(np.float64) + np.random.default_rng(2).normal(size=(n, 3)) * 0.01
    consts = oracle.odometry_consts(g["poses"])
    blk_i, blk_d = oracle.human_blocks(g["poses"], np.stack([ids[:, 0], ids[:, 1], ids[:, 2]], 1), deltas)
    r_odo, J_odo = oracle.eval_odometry(consts, x)
    r_hum, J_hum = oracle.eval_human(blk_i, blk_d, x)
    r_stf, J_stf = S.eval_stf(x, ref)
    n_odo, n_hum, n_stf = n - 1, len(ids), len(ref["pair_i"])
    nres_h = {2: 3, 4: 2, 5: 1, 6: 1}
    rng = np.random.default_rng(0)
    blocks = list(range(0, n_odo, 7)) + list(range(n_odo, n_odo + n_hum)) + [n_odo + n_hum + int(b) for b in rng.integers(0, n_stf, 40)]
    for b in blocks:
        r, j0, j1, total = session.evaluate_block(b, with_stf=True, pose_array=x)
        assert total == n_odo + n_hum + n_stf
        if b < n_odo:
            assert len(r) == 3 and rel_err(r, r_odo[b]) <= REL64 and rel_err(j0, J_odo[b, 0]) <= REL64 and rel_err(j1, J_odo[b, 1]) <= REL64
        elif b < n_odo + n_hum:
            h = b - n_odo
            k = nres_h[int(ids[h, 0])]
            assert len(r) == k and j1 is None
            assert np.abs(r - r_hum[h, :k]).max() <= 1e-12 and np.abs(j0 - J_hum[h, :k]).max() <= 1e-12
        else:
            s = b - n_odo - n_hum
            assert len(r) == 2 and rel_err(r, r_stf[s]) <= REL64 and rel_err(j0, J_stf[s, 0]) <= REL64 and rel_err(j1, J_stf[s, 1]) <= REL64


def _scipy_solve(fun, x0, n):
    from scipy.optimize import least_squares
    free0 = x0.reshape(-1)[3:].copy()                      # pose 0 constant

    def unpack(v):
        return np.concatenate([x0.reshape(-1)[:3], v]).reshape(n, 3)

    def res(v):
        return fun(unpack(v), False)

    def jac(v):
        return fun(unpack(v), True)
    sol = least_squares(res, free0, jac=lambda v: jac(v).toarray(), method="lm", xtol=1e-15, ftol=1e-15, gtol=1e-15, max_nfev=4000)
    return unpack(sol.x), sol


def test_solve_human_constraints_final_poses(session, oracle, maps):
    """SolveHumanConstraints (odometry + human blocks on the GPU, LM on the host) reaches the same
    minimiser as an independent solver over the oracle's Jet functors."""
    from scipy.sparse import lil_matrix
    g = maps("tiny")
    n = len(g["poses"])
    session.set_map(g["poses"], g["offsets"], g["pts"], g["nrm"])
    ids, deltas = _constraints(n, g["poses"])
    session.add_constraints(ids, deltas)
    session.solver_options(0, max_iterations=200, function_tolerance=1e-16, gradient_tolerance=1e-14, parameter_tolerance=1e-14)
    summ = session.solve(0)
    assert summ["final_cost"] < summ["initial_cost"] and summ["num_hc_residuals"] == 3 + 2 + 1 + 1 + 2
    _, got = session.poses()
    consts = oracle.odometry_consts(g["poses"])
    blk_i, blk_d = oracle.human_blocks(g["poses"], ids, deltas)
    nres_h = [{2: 3, 4: 2, 5: 1, 6: 1}[int(t)] for t in ids[:, 0]]

    def fun(x, want_jac):
        r_o, J_o = oracle.eval_odometry(consts, x, want_jac=True)
        r_h, J_h = oracle.eval_human(blk_i, blk_d, x, want_jac=True)
        r = np.concatenate([r_o.reshape(-1)] + [r_h[b, :k] for b, k in enumerate(nres_h)])
        if not want_jac:
            return r
        J = lil_matrix((len(r), 3 * n))
        for b in range(n - 1):
            J[3 * b:3 * b + 3, 3 * b:3 * b + 3] = J_o[b, 0]
            J[3 * b:3 * b + 3, 3 * b + 3:3 * b + 6] = J_o[b, 1]
        row = 3 * (n - 1)
        for b, k in enumerate(nres_h):
            p = int(blk_i[b, 1])
            J[row:row + k, 3 * p:3 * p + 3] = J_h[b, :k]
            row += k
        return J.tocsr()[:, 3:]
    want, sol = _scipy_solve(fun, g["poses"].astype(np.float64), n)
    assert abs(summ["final_cost"] - sol.cost) <= 1e-9 * max(sol.cost, 1e-12)
    assert np.abs(got - want).max() <= 1e-7               # optimised poses: tolerance stated on the double pose array
    # JointOpt::Run = the same solve + CopyParams (angle_mod, float narrowing)
    session.set_poses(g["poses"])
    session.joint_opt_run(post=False)
    pf, _ = session.poses()
    wrapped = want.copy()
    wrapped[:, 2] -= 2 * np.pi * np.rint(wrapped[:, 2] / (2 * np.pi))
    assert np.abs(pf - wrapped.astype(np.float32)).max() <= 2e-6


def _ceres_lm_numpy(fun, x0, n, max_iterations, ftol=1e-6, gtol=1e-10, ptol=1e-8):
    """Dense numpy restatement of Ceres' documented LM trust-region loop over
    the ORACLE's residuals/Jacobians: the pose-level checker for a solve that converges too slowly to
    compare minimisers (the STF residual is an RMS, so Gauss-Newton sees a rank-2 Hessian per block)."""
    x = x0.reshape(-1)[3:].copy()
    fixed = x0.reshape(-1)[:3]

    def unpack(v):
        return np.concatenate([fixed, v]).reshape(n, 3)
    r, J = fun(unpack(x), False), fun(unpack(x), True).toarray()
    cost = 0.5 * r @ r
    scale = 1.0 / (1.0 + np.sqrt((J * J).sum(0)))
    radius, dec, it_ok = 1e4, 2.0, 0
    costs = [cost]
    for _ in range(max_iterations):
        Js = J * scale
        H, g = Js.T @ Js, Js.T @ r
        lm = np.clip(np.diag(H), 1e-6, 1e32) / radius
        step = -np.linalg.solve(H + np.diag(lm), g)
        model = -(step @ (g + 0.5 * (H @ step)))
        if not model > 0:
            radius /= dec
            dec *= 2
            continue
        d = step * scale
        if np.linalg.norm(d) <= ptol * (np.linalg.norm(x) + ptol):
            break
        r_new = fun(unpack(x + d), False)
        cost_new = 0.5 * r_new @ r_new
        rho = (cost - cost_new) / model
        if rho > 1e-3:
            x, old, cost, r = x + d, cost, cost_new, r_new
            J = fun(unpack(x), True).toarray()
            costs.append(cost)
            it_ok += 1
            if abs(old - cost) <= ftol * old or np.abs(J.T @ r).max() <= gtol:
                break
            radius = min(1e16, radius / max(1.0 / 3.0, 1.0 - (2 * rho - 1) ** 3))
            dec = 2.0
        else:
            radius /= dec
            dec *= 2
    return unpack(x), cost, it_ok


def test_post_human_optimization_final_poses(session, oracle, maps):
    """PostHumanOptimization: search once, then LM over the STF blocks with pose 0 constant —
    optimised poses against the same LM loop run in numpy over the oracle's Jet functors."""
    from scipy.sparse import lil_matrix
    g = maps("tiny")
    n = len(g["poses"])
    session.set_map(g["poses"], g["offsets"], g["pts"], g["nrm"])
    S = oracle.scans(g["offsets"], g["pts"], g["nrm"])
    x0 = g["poses"].astype(np.float64)
    ref = S.find_stf(x0)
    nb = len(ref["pair_i"])

    def fun(x, want_jac):
        r, Jb = S.eval_stf(x, ref, want_jac=True)
        if not want_jac:
            return r.reshape(-1)
        J = lil_matrix((2 * nb, 3 * n))
        for b in range(nb):
            i, j = int(ref["pair_i"][b]), int(ref["pair_j"][b])
            J[2 * b:2 * b + 2, 3 * i:3 * i + 3] = Jb[b, 0]
            J[2 * b:2 * b + 2, 3 * j:3 * j + 3] = Jb[b, 1]
        return J.tocsr()[:, 3:]
    for iters in (5, 40):
        session.set_poses(g["poses"])
        session.solver_options(1, max_iterations=iters, function_tolerance=1e-6)      # the reference's options (:158, :1168 caps at 100)
        summ = session.solve(1)
        _, got = session.poses()
        want, want_cost, ok_steps = _ceres_lm_numpy(fun, x0, n, iters)
        r0 = fun(x0, False)
        assert abs(summ["initial_cost"] - 0.5 * (r0 ** 2).sum()) <= 1e-9 * 0.5 * (r0 ** 2).sum()
        assert summ["final_cost"] < summ["initial_cost"] and summ["successful_steps"] == ok_steps
        assert abs(summ["final_cost"] - want_cost) <= 1e-7 * want_cost
        assert np.abs(got - want).max() <= 1e-7               # optimised poses, FP64 mode (north_star: <= 1e-9 relative on ~10 m coordinates)
    grad, dims = session.gradient()                            # Problem::Evaluate at the solution (JointOptimization.cpp:1252)
    # constant blocks keep their (zero) columns, as in Ceres: the gradient has 3 entries per pose and pose 0's are zero
    assert len(grad) == 3 * n and dims[0] == 2 * nb and dims[1] == 3 * n and np.all(grad[:3] == 0.0)
    J = fun(got, True).toarray()
    assert np.abs(grad[3:] - J.T @ fun(got, False)).max() <= 1e-9 * max(1.0, np.abs(grad).max())


def test_fp32_mode_reaches_the_same_poses_within_1e5(session, maps):
    g = maps("tiny")
    session.set_map(g["poses"], g["offsets"], g["pts"], g["nrm"])
    session.solver_options(1, max_iterations=60, function_tolerance=1e-12, precision=0)
    session.solve(1)
    _, p64 = session.poses()
    session.set_poses(g["poses"])
    session.solver_options(1, precision=1)
    session.solve(1)
    _, p32 = session.poses()
    session.solver_options(1, precision=0)
    assert np.abs(p64 - p32).max() <= 1e-5 * max(1.0, np.abs(p64).max())


DRIFTY = dict(drift_xy=0.012, drift_th=0.004)   # enough odometry drift for a visible loop-closure error on the small maps


@pytest.mark.parametrize("name", ["small", "c1"])
def test_em_input_run_matches_oracle(session, gpu, oracle, maps, name):
    from hitl_slam_b200 import synth
    g = maps(name, **DRIFTY)
    session.set_map(g["poses"], g["offsets"], g["pts"], g["nrm"])
    session.world_transform(keep_host_copy=False)
    strokes = synth.pick_strokes(g, min_sep=0.045)
    S = oracle.scans(g["offsets"], g["pts"], g["nrm"])
    world = S.world_transform(g["poses"])
    want = oracle.em_run(g["offsets"], world, strokes)
    got = session.em_run(4, strokes)
    assert np.abs(got["segs"] - want["segs"]).max() <= 1e-5          # refit endpoints (two LM implementations)
    assert np.array_equal(got["corrected"], want["corrected"]) and np.array_equal(got["anchor"], want["anchor"])
    assert got["backprop"] == want["backprop"] and sum(got["rounds"]) == want["rounds"]
    assert len(got["corrected"]) > 0 and len(got["anchor"]) > 0
    # the same stage fed with host clouds (the reference's calling convention) gives the same answer
    session.world_transform(keep_host_copy=True)
    again = session.em_run(4, strokes)
    assert np.array_equal(again["segs"], got["segs"]) and np.array_equal(again["corrected"], got["corrected"])


@pytest.mark.parametrize("name", ["small", "c1"])
def test_device_m_step_matches_the_host_lm(session, gpu, host, maps, name):
    """hitl_em_refit (E-step + SegFitEM's LM in one cooperative kernel) against the host M-step on the SAME inliers: the fitted angle within
    1e-9, the same number of LM iterations, the float endpoints equal up to the last place; then EMInput::Run with the M-step on the device
    against the M-step on the host LM: same rounds, same pose lists, endpoints within 1e-6."""
    from hitl_slam_b200 import synth
    g = maps(name, **DRIFTY)
    session.set_map(g["poses"], g["offsets"], g["pts"], g["nrm"])
    session.world_transform(keep_host_copy=False)
    strokes = synth.pick_strokes(g, min_sep=0.045)
    rng = np.random.default_rng(7)
    for case in (strokes, strokes + rng.normal(size=(4, 2)).astype(np.float32) * 0.01, strokes[::-1].copy()):
        for seg in (case[:2], case[2:]):
            _, _, xy = gpu.em_inliers(seg.reshape(-1))
            want_seg, want_theta, want_it = host.seg_fit_em_theta(seg[0].astype(np.float64), seg[1].astype(np.float64), xy.astype(np.float64).reshape(-1))
            got_seg, info = gpu.em_refit(seg.reshape(-1))
            assert info["n_inliers"] == len(xy) > 0
            assert abs(info["theta"] - want_theta) <= 1e-9, (info, want_theta)
            assert info["iterations"] == want_it and info["final_cost"] <= info["initial_cost"] * (1 + 1e-12)
            assert np.abs(got_seg.reshape(2, 2) - want_seg).max() <= 2e-6
            again, info2 = gpu.em_refit(seg.reshape(-1))              # deterministic reductions: bit-identical on a second run
            assert np.array_equal(again, got_seg) and info2["theta"] == info["theta"]
    none_seg, none = gpu.em_refit(np.array([500, 500, 501, 500.5], np.float32))   # no inlier: the stroke is rebuilt from theta_0, as in the reference
    w0, t0, _ = host.seg_fit_em_theta(np.array([500.0, 500.0]), np.array([501.0, 500.5]), np.zeros(0))
    assert none["n_inliers"] == 0 and none["iterations"] == 0 and abs(none["theta"] - t0) <= 1e-15 and np.abs(none_seg.reshape(2, 2) - w0).max() <= 1e-4
    session.set_device_m_step(False)
    want = session.em_run(4, strokes)
    session.set_device_m_step(True)
    got = session.em_run(4, strokes)
    assert got["rounds"] == want["rounds"] and got["backprop"] == want["backprop"]
    assert np.array_equal(got["corrected"], want["corrected"]) and np.array_equal(got["anchor"], want["anchor"])
    assert np.abs(got["segs"] - want["segs"]).max() <= 1e-6
    # rounds chained on the device (both strokes, several rounds per host wait): the same bits as one wait per round
    for rounds in (1, 4):
        session.set_em_chain_rounds(rounds)
        other = session.em_run(4, strokes)
        assert np.array_equal(other["segs"].view(np.uint32), got["segs"].view(np.uint32)) and other["rounds"] == got["rounds"]
        assert np.array_equal(other["corrected"], got["corrected"]) and np.array_equal(other["anchor"], got["anchor"])
    session.set_em_chain_rounds(2)


def test_replayed_colinear_correction_end_to_end(session, host, oracle, maps, tmp_path):
    """BASELINE config 1: one colinear constraint replayed from a session log through EM -> constraint
    targets -> JointOpt::Run; the optimised poses satisfy the constraint and stay near odometry."""
    from hitl_slam_b200 import synth
    g = maps("c1", **DRIFTY)
    n = len(g["poses"])
    strokes = synth.pick_strokes(g, min_sep=0.045)
    log = str(tmp_path / "c1_logged.log")
    host.save_log(log, [(4, 0, strokes)])
    (ctype, undone, pts), = host.load_log(log)
    assert ctype == 4 and undone == 0
    session.set_map(g["poses"], g["offsets"], g["pts"], g["nrm"])
    session.world_transform()
    em = session.em_run(ctype, pts)
    assert em["backprop"][0] >= 0
    n_added = session.add_constraints_from_em()
    assert n_added == len(em["corrected"]) * len(em["anchor"])
    summ = session.joint_opt_run(post=False)
    assert summ["termination"] in (0, 1) and summ["num_hc_residuals"] == 2 * n_added
    # targets are frozen from the current poses, so the problem starts at its minimum (Appendix B.6/B.7)
    assert summ["final_cost"] <= summ["initial_cost"] + 1e-12
    p, _ = session.poses()
    assert np.abs(p - g["poses"]).max() <= 1e-3


def _bits_equal(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    return a.shape == b.shape and bool(((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))).all())


@pytest.mark.parametrize("n,lo,hi", [(12, 3, 4), (40, 0, 39), (300, 7, 250), (1500, 100, 1400), (5000, 0, 4999), (2600, 2599 - 1030, 2599)])
def test_backprop_bit_exact(gpu, host, oracle, n, lo, hi):
    """Backprop::BackPropagateError (Backprop.cpp:98-200): the O(L^2) pose update runs on the device (one CTA, systolic over
    the poses) and gives every pose the host loops' float operation sequence: poses and covariances bit for bit."""
    rng = np.random.default_rng(n + lo)
    poses = np.cumsum(rng.normal(size=(n, 3)) * [0.25, 0.25, 0.03], 0).astype(np.float32)
    cov = np.zeros((n, 9), np.float32)
    cov[:, 0] = cov[:, 4] = 1e-4 * (1 + np.arange(n) / 100) * rng.uniform(0.5, 1.5, n)
    cov[:, 8] = 1e-5 * (1 + np.arange(n) / 100)
    cov[:, [1, 2, 3, 5, 6, 7]] = rng.normal(size=(n, 6)) * 1e-6
    c3 = np.array([0.31, -0.22, 0.07], np.float32)
    want_p, want_c = oracle.backprop(poses, cov, lo, hi, c3)
    got_p, got_c, ms = host.backprop(gpu, poses, cov, lo, hi, c3)
    assert _bits_equal(got_p, want_p) and _bits_equal(got_c, want_c)
    assert np.abs(got_p[lo + 1:hi + 1] - poses[lo + 1:hi + 1]).max() > 1e-3         # something moved ...
    assert np.array_equal(got_p[:lo], poses[:lo]) and np.array_equal(got_p[hi + 1:], poses[hi + 1:])   # ... only inside the bounds
    # the last pose lands on the destination (that is what the weights are normalised for, up to the fused destination variance)
    dest = poses[hi, :2] + c3[:2]
    assert np.abs(got_p[hi, :2] - dest).max() < 0.5
    # bounds that leave nothing to do
    same_p, same_c, _ = host.backprop(gpu, poses, cov, hi, hi, c3)
    assert np.array_equal(same_p, poses) and np.array_equal(same_c, cov)


@pytest.mark.parametrize("name", ["small", "c1"])
def test_full_correction_chain_matches_oracle(session, gpu, oracle, maps, name):
    """HitLSLAM::Run's wiring (HitLSLAM.cpp:379-484) up to the joint optimisation: EM -> explicit correction -> back-propagation ->
    angle wrap, against the same chain of oracle stages.  The refit strokes differ by <= 1e-5 (two LM implementations), so the
    corrected poses are compared at 1e-4; fed with the mirror's own strokes the two later stages are bit-exact (tests above)."""
    from hitl_slam_b200 import synth
    g = maps(name, **DRIFTY)
    n = len(g["poses"])
    session.set_map(g["poses"], g["offsets"], g["pts"], g["nrm"])
    session.world_transform(keep_host_copy=False)
    strokes = synth.pick_strokes(g, min_sep=0.045)
    cov = np.ascontiguousarray(g["cov"], np.float32).reshape(n, 9).copy() if "cov" in g else np.tile(np.array([1e-4, 0, 0, 0, 1e-4, 0, 0, 0, 1e-5], np.float32), (n, 1))
    cov0 = cov.copy()
    out = session.correct(4, strokes, cov=cov, solve=False)
    assert out["applied"] and out["n_constraints"] == out["n_corrected"] * out["n_anchor"] > 0
    got, _ = session.poses()
    S = oracle.scans(g["offsets"], g["pts"], g["nrm"])
    em = oracle.em_run(g["offsets"], S.world_transform(g["poses"]), strokes)
    assert out["backprop"] == em["backprop"] and out["n_corrected"] == len(em["corrected"])
    # oracle chain on the MIRROR's refit strokes: everything after EM must agree bit for bit
    p1, c3 = oracle.app_exp_corrections(4, out["segs"], g["poses"], em["corrected"])
    p2, cov2 = oracle.backprop(p1, cov0, em["backprop"][0], em["backprop"][1], c3)
    wrapped = p2.copy()
    wrapped[:, 2] = np.arctan2(np.sin(p2[:, 2].astype(np.float32)), np.cos(p2[:, 2].astype(np.float32)))
    assert _bits_equal(got[:, :2], p2[:, :2]) and _bits_equal(cov, cov2)
    assert np.abs(got[:, 2] - wrapped[:, 2]).max() <= 1e-6
    # and on the oracle's own refit strokes within the LM tolerance
    q1, d3 = oracle.app_exp_corrections(4, em["segs"], g["poses"], em["corrected"])
    q2, _ = oracle.backprop(q1, cov0, em["backprop"][0], em["backprop"][1], d3)
    assert np.abs(got[:, :2] - q2[:, :2]).max() <= 1e-4
    # the correction did something: poses after the corrected range moved rigidly
    assert np.abs(got - g["poses"]).max() > 1e-3
    # the joint optimisation then starts from these poses with the new constraints
    summ = session.joint_opt_run(post=False)
    assert summ["termination"] in (0, 1) and summ["num_hc_residuals"] == 2 * out["n_constraints"]
