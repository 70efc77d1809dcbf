"""CPU suite (-m "not gpu"): pins the oracle, checks the host logic of the product and that the
C-ABI library loads and exports every symbol include/hitl_gpu.h declares."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, random_scans

from hitl_slam_b200 import ABI_SYMBOLS, HitlError, capi
from oracle.pyoracle import RefKDTree


# ---- C ABI ----------------------------------------------------------------------------------------
def test_library_exports_every_declared_symbol():
    lib = capi.load_gpu_library()
    header = open(os.path.join(ROOT, "include", "hitl_gpu.h")).read()
    declared = set(re.findall(r"\b(hitl_[a-z0-9_]+)\s*\(", header)) - {"hitl_stf_opts", "hitl_stf_info"}
    assert declared == set(ABI_SYMBOLS), declared ^ set(ABI_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    with pytest.raises(HitlError):
        capi.HitlGpu(0)


def test_product_does_not_touch_the_oracle():
    """The product package must not import / load / link anything under oracle/."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "hitl_slam_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "pyoracle" not in text and "liboracle" not in text and "hitl_oracle" not in text, f
    for so in ("libhitl_gpu.so", "libhitl_host.so"):
        needed = os.popen("objdump -p %s | grep NEEDED" % capi.lib_path(so)).read()
        assert "oracle" not in needed


# ---- libm parity of the shared sinf/cosf -----------------------------------------------------------
def _libm_uses_fma_variant():
    flags = open("/proc/cpuinfo").read()
    return " fma " in flags and " avx2 " in flags


def test_sincos_matches_libm(host):
    """hitl_math.h restates glibc 2.39's __sinf_fma/__cosf_fma; the reference calls the platform libm."""
    if not _libm_uses_fma_variant():
        pytest.skip("host libm does not select the FMA variant")
    stride = 1 if os.environ.get("HITL_EXHAUSTIVE") else 61
    assert host.sincos_mismatches(0, (1 << 32) // stride, stride) == 0
    # every float in [-8, 8] (all pose-angle differences live here)
    lo, hi = np.float32(0).view(np.uint32), np.float32(8).view(np.uint32)
    assert host.sincos_mismatches(int(lo), int(hi - lo) + 1, 1) == 0
    assert host.sincos_mismatches(int(lo) | 0x80000000, int(hi - lo) + 1, 1) == 0


def test_relative_pose_matches_oracle(host, oracle):
    rng = np.random.default_rng(5)
    poses = np.concatenate([rng.uniform(-50, 50, (64, 2)), rng.uniform(-7, 7, (64, 1))], 1)
    poses = poses.astype(np.float32).astype(np.float64).reshape(-1)
    out = np.zeros(6, np.float32)
    for a in range(0, 64, 3):
        for b in range(1, 64, 5):
            oracle.lib.orc_relative_pose(poses, a, b, out)
            assert np.array_equal(out.view(np.uint32), host.relative_pose(poses, a, b).view(np.uint32))


# ---- KD-tree: oracle vs the reference's own kdtree.cpp, product builder vs oracle -------------------
SIZES = [1, 2, 3, 4, 5, 7, 16, 17, 31, 32, 33, 100, 360, 721, 1080, 2160]


@pytest.mark.skipif(not RefKDTree.available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("n", SIZES)
def test_oracle_tree_matches_reference_kdtree(oracle, n):
    rng = np.random.default_rng(n)
    off, pts, nrm = random_scans(rng, 1, n, n)
    S, R = oracle.scans(off, pts, nrm), RefKDTree(pts, nrm)
    for x, y in zip(S.flatten(), R.flatten()):
        assert np.array_equal(x, y)
    q = (pts[rng.integers(0, n, 3000)] + rng.normal(size=(3000, 2)) * 0.1).astype(np.float32)
    q[:40] = pts[rng.integers(0, n, 40)]          # exact hits: the FLT_MIN early return
    q[40:80, 0] = pts[rng.integers(0, n, 40), 0]  # on a splitting plane: s == 0
    for mode in (0, 1):
        for thr in (0.05, 0.15, 1.0):
            d1, i1 = S.query(0, q, thr, mode)
            d2, i2 = R.query(q, thr, mode)
            assert np.array_equal(d1.view(np.uint32), d2.view(np.uint32)) and np.array_equal(i1, i2)
    for k in range(30):
        assert np.array_equal(S.radius(0, q[k, 0], q[k, 1], 0.3), R.radius(q[k, 0], q[k, 1], 0.3))


@pytest.mark.parametrize("n", SIZES)
def test_product_host_builder_matches_oracle(oracle, n):
    rng = np.random.default_rng(100 + n)
    off, pts, nrm = random_scans(rng, 1, n, n)
    nodes = capi.kdtree_build_host(pts, nrm)
    pn, idx, dim = oracle.scans(off, pts, nrm).flatten()
    assert np.array_equal(np.stack([nodes["px"], nodes["py"], nodes["nx"], nodes["ny"]], 1), pn)
    assert np.array_equal(nodes["index"], idx) and np.array_equal(nodes["dim"], dim)


def test_tree_invariants(oracle):
    rng = np.random.default_rng(3)
    off, pts, nrm = random_scans(rng, 1, 500, 500, ties=False)
    nodes = capi.kdtree_build_host(pts, nrm)
    assert sorted(nodes["index"]) == list(range(500))

    def check(pos, n):
        if n == 0:
            return
        d = nodes["dim"][pos]
        key = "px" if d == 0 else "py"
        nl, nr = n // 2, n - 1 - n // 2
        left, right = nodes[pos + 1:pos + 1 + nl], nodes[pos + 1 + nl:pos + 1 + nl + nr]
        assert (left[key] <= nodes[key][pos]).all() and (right[key] >= nodes[key][pos]).all()
        check(pos + 1, nl)
        check(pos + 1 + nl, nr)
    check(0, 500)


# ---- file format -------------------------------------------------------------------------------------
def test_loader_matches_oracle_and_golden(maps, oracle, host):
    for normals in ("compensated", "faithful"):
        g = maps("tiny", normals=normals, keep_file=True)
        og = oracle.load_pose_graph(g["path"])
        for k in ("poses", "cov", "offsets", "pts", "nrm"):
            assert np.array_equal(np.ascontiguousarray(og[k]).view(np.uint8), np.ascontiguousarray(g[k]).view(np.uint8)), k
    gold = np.load(os.path.join(ROOT, "tests", "golden", "tiny_compensated.npz"))
    g = maps("tiny", normals="compensated", keep_file=True)
    for k in ("poses", "offsets", "pts", "nrm"):
        assert np.array_equal(gold[k], g[k]), k
    # compensated normals load as unit vectors, faithful ones carry the loader's translation quirk
    assert abs(np.linalg.norm(g["nrm"], axis=1) - 1).max() < 1e-3
    gf = maps("tiny", normals="faithful", keep_file=True)
    assert np.linalg.norm(gf["nrm"], axis=1).max() > 1.5


def test_writer_format(tmp_path, host):
    p = str(tmp_path / "x.stfs.covars")
    host.save_stfs_covars(p, np.array([[1.23456, -2.5, 0.78539]], np.float32), np.arange(9, dtype=np.float32)[None] * 0.5,
                          np.array([0, 2], np.uint32), np.array([[3.00004, 4.5], [1, 2]], np.float32), np.array([[0, 1], [1, 0]], np.float32),
                          map_name="m", timestamp=12.5)
    lines = open(p).read().splitlines()
    assert lines[0] == "m" and lines[1] == "12.500000"
    assert lines[2] == "1.2346,-2.5000,0.7854,3.0000,4.5000, 0.0000,1.0000,0.000000, 0.500000, 1.000000, 1.500000, 2.000000, 2.500000, 3.000000, 3.500000, 4.000000"


# ---- oracle vs golden vectors / self checks ---------------------------------------------------------
def test_oracle_stf_matches_golden(maps, oracle):
    g = maps("tiny")
    gold = np.load(os.path.join(ROOT, "tests", "golden", "tiny_stf.npz"))
    r = oracle.scans(g["offsets"], g["pts"], g["nrm"]).find_stf(g["poses"].astype(np.float64))
    for k in ("pair_i", "pair_j", "pair_off", "k", "idx"):
        assert np.array_equal(gold[k], r[k]), k
    assert int(gold["n_queries"]) == r["n_queries"]


def test_oracle_matches_satisfy_both_gates(maps, oracle):
    """Every reported correspondence passes the distance and the normal gate (brute-force re-check)."""
    g = maps("small")
    poses = g["poses"].astype(np.float64)
    S = oracle.scans(g["offsets"], g["pts"], g["nrm"])
    r = S.find_stf(poses)
    off = g["offsets"]
    out = np.zeros(6, np.float32)
    min_cos = capi.default_min_cos()
    for b in range(0, len(r["pair_i"]), 37):
        i, j = int(r["pair_i"][b]), int(r["pair_j"][b])
        oracle.lib.orc_relative_pose(poses.reshape(-1), i, j, out)
        T = out.astype(np.float64)
        for m in range(int(r["pair_off"][b]), int(r["pair_off"][b + 1])):
            p = g["pts"][off[i] + r["k"][m]].astype(np.float64)
            q = np.array([T[0] * p[0] + T[1] * p[1] + T[4], T[2] * p[0] + T[3] * p[1] + T[5]])
            t = g["pts"][off[j] + r["idx"][m]].astype(np.float64)
            tn = g["nrm"][off[j] + r["idx"][m]].astype(np.float64)
            assert np.hypot(*(q - t)) < 0.15 + 1e-5
            assert abs(tn @ (q - t)) < 0.15 + 1e-5
            dth = poses[j, 2] - poses[i, 2]
            n = g["nrm"][off[i] + r["k"][m]].astype(np.float64)
            rn = np.array([np.cos(dth) * n[0] - np.sin(dth) * n[1], np.sin(dth) * n[0] + np.cos(dth) * n[1]])
            assert tn @ rn > min_cos - 1e-5
    counts = np.diff(r["pair_off"].astype(np.int64))
    assert (counts > 10).all()


def test_oracle_shards_concatenate(maps, oracle):
    g = maps("small")
    poses = g["poses"].astype(np.float64)
    S = oracle.scans(g["offsets"], g["pts"], g["nrm"])
    full = S.find_stf(poses)
    n = len(g["offsets"]) - 1
    parts = [S.find_stf(poses, src_lo=a, src_hi=b) for a, b in ((0, 50), (50, 51), (51, n))]
    assert np.array_equal(np.concatenate([p["pair_i"] for p in parts]), full["pair_i"])
    assert np.array_equal(np.concatenate([p["k"] for p in parts]), full["k"])
    assert sum(p["n_queries"] for p in parts) == full["n_queries"]


def test_oracle_jacobians_match_finite_differences(maps, oracle):
    g = maps("tiny")
    poses = g["poses"].astype(np.float64)
    S = oracle.scans(g["offsets"], g["pts"], g["nrm"])
    corr = S.find_stf(poses)
    assert len(corr["pair_i"]) > 0
    rng = np.random.default_rng(0)
    x = poses + rng.normal(size=poses.shape) * 0.01
    r0, J = S.eval_stf(x, corr)
    h = 1e-6
    for b in range(0, len(corr["pair_i"]), 11):
        for side, pose in ((0, int(corr["pair_i"][b])), (1, int(corr["pair_j"][b]))):
            for c in range(3):
                xp, xm = x.copy(), x.copy()
                xp[pose, c] += h
                xm[pose, c] -= h
                fd = (S.eval_stf(xp, corr, want_jac=False)[0][b] - S.eval_stf(xm, corr, want_jac=False)[0][b]) / (2 * h)
                assert np.allclose(J[b, side, :, c], fd, rtol=1e-5, atol=1e-7)
    consts = oracle.odometry_consts(g["poses"])
    r, Jo = oracle.eval_odometry(consts, x)
    for b in (0, 5, len(consts) - 1):
        for side in (0, 1):
            for c in range(3):
                xp, xm = x.copy(), x.copy()
                xp[b + side, c] += h
                xm[b + side, c] -= h
                fd = (oracle.eval_odometry(consts, xp, False)[0][b] - oracle.eval_odometry(consts, xm, False)[0][b]) / (2 * h)
                assert np.allclose(Jo[b, side, :, c], fd, rtol=1e-5, atol=1e-6)
    # at the build-time poses the odometry residuals vanish (Appendix B.7)
    r_at, _ = oracle.eval_odometry(consts, poses, False)
    assert np.abs(r_at).max() < 2e-2


def test_oracle_em_matches_golden(maps, oracle):
    from hitl_slam_b200 import synth
    g = maps("small")
    S = oracle.scans(g["offsets"], g["pts"], g["nrm"], build_trees=False)
    world = S.world_transform(g["poses"])
    strokes = synth.make_strokes(g)
    gold = np.load(os.path.join(ROOT, "tests", "golden", "small_em.npz"))
    op, oi = oracle.em_inliers(g["offsets"], world, strokes[:2].reshape(-1))
    assert np.array_equal(op, gold["inl_pose"]) and np.array_equal(oi, gold["inl_idx"])
    sets = oracle.em_assign(g["offsets"], world, strokes)
    for f in range(2):
        for a, name in zip(sets[f], ("pose", "off", "obs")):
            assert np.array_equal(a, gold["set%d_%s" % (f, name)])
    run = oracle.em_run(g["offsets"], world, strokes)
    assert np.allclose(run["segs"], gold["run_segs"], atol=1e-4)
    assert np.array_equal(run["corrected"], gold["run_corrected"]) and np.array_equal(run["anchor"], gold["run_anchor"])


def test_distance_helpers_edge_cases(oracle):
    seg = np.array([0, 0, 2, 0], np.float32)
    d = oracle.lib.orc_distance_to_line_segment
    # t is in metres and compared with 1.0 (Appendix B.8): beyond 1 m along a 2 m stroke counts as "past p1"
    assert d(seg, 0.5, 0.25) == pytest.approx(0.25)
    assert d(seg, 1.5, 0.0) == pytest.approx(0.5)          # distance to p1=(2,0), not 0
    assert d(seg, -1.0, 0.0) == pytest.approx(1.0)
    e = oracle.lib.orc_dist_to_line_seg
    assert e(seg, 1.5, 0.25) == pytest.approx(0.25)        # the assignment variant uses a true fraction
    assert e(seg, 3.0, 0.0) == pytest.approx(1.0)


def test_shard_ranges_balance():
    from hitl_slam_b200.sharding import shard_ranges
    off = np.concatenate([[0], np.cumsum(np.random.default_rng(0).integers(100, 800, 1000))]).astype(np.uint32)
    for w in (1, 2, 3, 8):
        r = shard_ranges(off, w)
        assert r[0][0] == 0 and r[-1][1] == 1000 and all(r[i][1] == r[i + 1][0] for i in range(w - 1))
        loads = [int(off[b]) - int(off[a]) for a, b in r]
        assert max(loads) - min(loads) <= 2 * 800


def test_stdsort_restatement_matches_std_sort(host):
    """csrc/stdsort_exact.h restates libstdc++'s std::sort (introsort + final insertion sort + heap-sort fallback); the GPU tree
    builder relies on it for segments with equal keys.  Compared with this image's std::sort — the one the host builder and
    the compiled reference kdtree.cpp use — on tie-heavy, sorted, reversed and adversarial inputs."""
    import ctypes as C
    lib = host.lib
    f32p = np.ctypeslib.ndpointer(np.float32, flags="C")
    lib.hitl_host_stdsort_mismatches.restype = C.c_uint64
    lib.hitl_host_stdsort_mismatches.argtypes = [f32p, C.c_uint32, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32)]
    rng = np.random.default_rng(0)
    heap = C.c_uint32()
    total_heap = 0

    def check(k):
        nonlocal total_heap
        k = np.ascontiguousarray(k, np.float32)
        assert lib.hitl_host_stdsort_mismatches(k, len(k), None, None, C.byref(heap)) == 0
        total_heap += heap.value

    for n in list(range(0, 40)) + [100, 257, 720, 1080, 2160, 65534]:
        for rep in range(3):
            check(rng.integers(0, max(2, n // (rep + 1) + 1), n))
            check(rng.normal(size=n))
            check(np.sort(rng.integers(0, 5, n)))
            check(np.sort(rng.integers(0, 5, n))[::-1])
        check(np.zeros(n))
    for n in (128, 1000, 4096, 30000):          # median-of-3 killer (Musser): reaches the depth limit
        k = n // 2
        a = np.zeros(n)
        for i in range(1, k + 1):
            if i % 2 == 1:
                a[i - 1] = i
                if i < k:
                    a[i] = k + i
            a[k + i - 1] = 2 * i
        check(a)
        check(a // 3)
    assert total_heap > 0
