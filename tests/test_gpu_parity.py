"""GPU parity suite (-m gpu): the sm_100a kernels, called through the C ABI, against the CPU
oracle on identical inputs.  Index sets and EM assignments bit-exact; residuals / Jacobians
within 1e-9 relative (FP64) and 1e-5 (FP32 mode), as BASELINE.json:north_star states."""
import os

import numpy as np
import pytest

from conftest import ROOT, assert_same_stf, random_scans

pytestmark = pytest.mark.gpu

REL64 = 1e-9     # FP64 tolerance (north_star)
REL32 = 1e-5     # FP32-mode tolerance (north_star)


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    scale = max(np.abs(b).max(), 1e-300) if b.size else 1.0
    return np.abs(a - b).max() / scale if b.size else 0.0


def load_map(gpu, g):
    gpu.set_scans(g["offsets"], g["pts"], g["nrm"])
    gpu.build_kdtrees()


# ---- device arithmetic -----------------------------------------------------------------------------
def test_device_sincos_bit_exact(gpu, host):
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-8, 8, 1 << 20), rng.uniform(-130, 130, 1 << 18), rng.normal(size=1 << 16) * 1e4,
                        np.array([0.0, -0.0, 1e-30, 2 ** -12, 0.785398, 0.785399, 119.99, 120.0, 1e9, 3e38])]).astype(np.float32)
    s, c = gpu.debug_sincos(x)
    hs = np.array([host.lib.hitl_host_sinf(float(v)) for v in x[:4096]], np.float32)
    assert np.array_equal(hs.view(np.uint32), s[:4096].view(np.uint32))
    # against the platform libm (what the reference calls) through numpy's float32 ufuncs is not
    # guaranteed to be libm; use the oracle's wrappers instead
    from oracle.pyoracle import Oracle
    o = Oracle()
    idx = rng.integers(0, len(x), 20000)
    ls = np.array([o.lib.orc_sinf(float(v)) for v in x[idx]], np.float32)
    lc = np.array([o.lib.orc_cosf(float(v)) for v in x[idx]], np.float32)
    assert np.array_equal(ls.view(np.uint32), s[idx].view(np.uint32))
    assert np.array_equal(lc.view(np.uint32), c[idx].view(np.uint32))


def test_device_relative_pose_bit_exact(gpu, oracle, maps):
    g = maps("small")
    load_map(gpu, g)
    poses = g["poses"].astype(np.float64)
    n = len(poses)
    rng = np.random.default_rng(1)
    src, dst = rng.integers(0, n, 2000).astype(np.uint32), rng.integers(0, n, 2000).astype(np.uint32)
    out = gpu.debug_relative_pose(poses, src, dst)
    ref = np.zeros(6, np.float32)
    for q in range(2000):
        oracle.lib.orc_relative_pose(poses.reshape(-1), int(src[q]), int(dst[q]), ref)
        assert np.array_equal(ref.view(np.uint32), out[q].view(np.uint32))


# ---- KD queries ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 2, 3, 17, 360, 1080, 2160])
def test_kd_queries_match_oracle(gpu, oracle, n):
    rng = np.random.default_rng(n)
    off, pts, nrm = random_scans(rng, 3, n, n, empty=(1,))
    gpu.set_scans(off, pts, nrm)
    gpu.build_kdtrees()
    S = oracle.scans(off, pts, nrm)
    nodes = gpu.get_kdtrees()
    pn, idx, dim = S.flatten()
    assert np.array_equal(nodes["index"], idx) and np.array_equal(nodes["dim"], dim) and np.array_equal(nodes["px"], pn[:, 0])
    for scan in (0, 1, 2):
        m = int(off[scan + 1] - off[scan])
        base = pts[off[scan]:off[scan + 1]] if m else np.zeros((1, 2), np.float32)
        q = (base[rng.integers(0, len(base), 5000)] + rng.normal(size=(5000, 2)) * 0.1).astype(np.float32)
        q[:50] = base[rng.integers(0, len(base), 50)]
        q[50:100, 0] = base[rng.integers(0, len(base), 50), 0]
        for mode in (0, 1):
            for thr in (0.05, 0.15, 1.0):
                d0, i0 = S.query(scan, q, thr, mode)
                d1, i1 = gpu.kd_query(scan, q, thr, mode)
                if m == 0:
                    assert (i1 == -1).all()
                    continue
                assert np.array_equal(i0, i1), (scan, mode, thr)
                assert np.array_equal(d0.view(np.uint32), d1.view(np.uint32)), (scan, mode, thr)
        if m:
            _, cnt = gpu.kd_query(scan, q[:200], 0.3, 2)
            want = [S.radius(scan, q[k, 0], q[k, 1], 0.3) for k in range(200)]
            ref = np.array([len(w) for w in want])
            assert np.array_equal(cnt, ref)
            # FindNeighborPoints' node LIST (kdtree.cpp:199-218): same nodes in the reference's push order; truncation keeps the head
            lists, cnt2 = gpu.kd_neighbors(scan, q[:200], 0.3, cap=int(ref.max()) + 1)
            assert np.array_equal(cnt2, ref)
            for k in range(200):
                assert np.array_equal(lists[k], np.asarray(want[k], np.int32)), (scan, k)
            short, cnt3 = gpu.kd_neighbors(scan, q[:50], 0.3, cap=2)
            assert np.array_equal(cnt3, ref[:50]) and all(np.array_equal(short[k], np.asarray(want[k][:2], np.int32)) for k in range(50))


def test_compact_tree_and_index_formats(gpu, oracle, maps):
    """hitl_set/get_kdtrees_compact (index | dim << 31 per node) reproduce the 24-byte node form bit for bit, a search over trees uploaded in the
    compact form returns the same lists, and hitl_get_stf16 returns the same indices in 16 bits."""
    g = maps("small")
    load_map(gpu, g)
    poses = g["poses"].astype(np.float64)
    nodes = gpu.get_kdtrees().copy()
    comp = gpu.get_kdtrees_compact().copy()
    assert np.array_equal(comp & 0x7FFFFFFF, nodes["index"].astype(np.uint32)) and np.array_equal(comp >> 31, nodes["dim"].astype(np.uint32))
    want = gpu.find_stf(poses)
    gpu.set_scans(g["offsets"], g["pts"], g["nrm"])
    gpu.set_kdtrees_compact(comp)
    again = gpu.get_kdtrees()
    for f in ("px", "py", "nx", "ny", "index", "dim"):
        assert np.array_equal(again[f].view(np.uint32), nodes[f].view(np.uint32)), f
    got = gpu.find_stf(poses, fetch=False)
    assert_same_stf(gpu.get_stf(got["n_pairs"], got["n_matches"]), want)
    g16 = gpu.get_stf16(got["n_pairs"], got["n_matches"])
    assert g16["k"].dtype == np.uint16 and g16["idx"].dtype == np.uint16
    assert_same_stf({k: np.asarray(v).astype(np.asarray(want[k]).dtype) for k, v in g16.items()}, want)
    assert_same_stf(oracle.scans(g["offsets"], g["pts"], g["nrm"]).find_stf(poses), want)
    bad = comp.copy(); bad[5] = 0x7FFFFFF0
    with pytest.raises(Exception):
        gpu.set_kdtrees_compact(bad)
    moved = nodes.copy(); moved["px"][7] += np.float32(0.25)          # a node that is not the scan point it names: refused (the culls are built from the scans)
    with pytest.raises(Exception):
        gpu.set_kdtrees(moved)
    turned = nodes.copy(); turned["nx"][9] = -turned["nx"][9] - np.float32(1.0)
    with pytest.raises(Exception):
        gpu.set_kdtrees(turned)
    load_map(gpu, g)                                     # leave the shared context in a sane state


def _both_builders(gpu, off, pts, nrm):
    gpu.set_scans(off, pts, nrm)
    gpu.debug_set_tree_builder(host=False)
    gpu.build_kdtrees()
    dev, exact = gpu.get_kdtrees().copy(), gpu.debug_tree_stats()
    gpu.debug_set_tree_builder(host=True)
    gpu.build_kdtrees()
    host = gpu.get_kdtrees().copy()
    gpu.debug_set_tree_builder(host=False)
    return dev, host, exact


def test_device_tree_build_matches_host_builder(gpu, oracle, maps):
    """The level-synchronous device builder (kdtree_gpu.cu) gives the host builder's trees node for node (the host builder
    is pinned to the reference's kdtree.cpp in tests/test_cpu_oracle.py), on maps and on random scans of awkward sizes."""
    for g in (maps("small"), maps("c1", normals="faithful")):
        dev, host, _ = _both_builders(gpu, g["offsets"], g["pts"], g["nrm"])
        assert dev.tobytes() == host.tobytes()
    rng = np.random.default_rng(3)
    sizes = [0, 1, 2, 3, 4, 5, 7, 8, 15, 16, 17, 31, 33, 64, 100, 719, 720, 1081, 2160, 0, 4099]
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint32)
    pts = (rng.normal(size=(off[-1], 2)) * 5).astype(np.float32)
    nrm = rng.normal(size=(off[-1], 2)).astype(np.float32)
    dev, host, _ = _both_builders(gpu, off, pts, nrm)
    assert dev.tobytes() == host.tobytes()
    S = oracle.scans(off, pts, nrm)
    pn, idx, dim = S.flatten()
    assert np.array_equal(dev["index"], idx) and np.array_equal(dev["dim"], dim) and np.array_equal(dev["px"], pn[:, 0])


def test_device_tree_build_with_equal_coordinates(gpu, oracle):
    """Equal keys: std::sort is not stable, the order (hence the tree) is whatever libstdc++'s introsort produces.  The
    device builder re-sorts such segments with the step-for-step restatement (stdsort_exact.h)."""
    rng = np.random.default_rng(11)
    scans = []
    scans.append(np.round(rng.normal(size=(1500, 2)) * 3, 1))                       # 0.1 lattice: many ties on both axes
    scans.append(np.stack([np.repeat(np.arange(40), 25), np.tile(np.arange(25), 40)], 1).astype(np.float64))   # exact grid
    dup = rng.normal(size=(300, 2)); scans.append(np.concatenate([dup, dup, dup[:77]]))                          # duplicated points
    z = np.zeros((257, 2)); z[::2, 0] = -0.0; z[:, 1] = rng.integers(0, 3, 257); scans.append(z)                   # -0 / +0 and 3 values
    scans.append(np.stack([np.linspace(-4, 4, 2000), np.zeros(2000)], 1))           # a wall on the x axis: y all equal
    scans.append(np.round(rng.normal(size=(17, 2)), 0))
    scans.append(np.ones((5000, 2)))                                                 # everything equal
    k = 4096; mus = np.zeros(k)                                                      # median-of-3 killer: drives introsort into its heap-sort fallback
    for i in range(1, k // 2 + 1):
        if i % 2 == 1:
            mus[i - 1] = i
            if i < k // 2:
                mus[i] = k // 2 + i
        mus[k // 2 + i - 1] = 2 * i
    scans.append(np.stack([mus // 3, np.zeros(k)], 1))
    off = np.concatenate([[0], np.cumsum([len(x) for x in scans])]).astype(np.uint32)
    pts = np.concatenate(scans).astype(np.float32)
    nrm = rng.normal(size=pts.shape).astype(np.float32)
    dev, host, exact = _both_builders(gpu, off, pts, nrm)
    assert exact > 100
    assert dev.tobytes() == host.tobytes()
    S = oracle.scans(off, pts, nrm)
    pn, idx, dim = S.flatten()
    assert np.array_equal(dev["index"], idx) and np.array_equal(dev["dim"], dim)


# ---- correspondence search -------------------------------------------------------------------------
@pytest.mark.parametrize("name,normals", [("tiny", "compensated"), ("tiny", "faithful"), ("small", "compensated"), ("small", "faithful")])
def test_find_stf_bit_exact(gpu, oracle, maps, name, normals):
    g = maps(name, normals=normals)
    load_map(gpu, g)
    poses = g["poses"].astype(np.float64)
    ref = oracle.scans(g["offsets"], g["pts"], g["nrm"]).find_stf(poses)
    for cull, fine in ((0, True), (0, False), (1, True)):
        gpu.debug_set_fine_occupancy(fine)
        out = gpu.find_stf(poses, opts=gpu.stf_opts(disable_culling=cull))
        assert_same_stf(out, ref)
        assert out["n_queries"] == ref["n_queries"]
    gpu.debug_set_fine_occupancy(True)
    assert len(ref["pair_i"]) > 0 or normals == "faithful"


def test_find_stf_golden_fixture(gpu):
    g = np.load(os.path.join(ROOT, "tests", "golden", "tiny_compensated.npz"))
    gold = np.load(os.path.join(ROOT, "tests", "golden", "tiny_stf.npz"))
    gpu.set_scans(g["offsets"], g["pts"], g["nrm"])
    gpu.build_kdtrees()
    out = gpu.find_stf(g["poses"].astype(np.float64))
    assert_same_stf(out, gold)
    assert out["n_queries"] == int(gold["n_queries"])


@pytest.mark.parametrize("opts", [dict(cap=1), dict(cap=3, skip=2), dict(cap=6, skip=5, min_corr=0), dict(thr=0.05, min_corr=3),
                                  dict(thr=0.4, cap=12, min_corr=25), dict(min_cos=-2.0), dict(cap=0)])
def test_find_stf_option_sweep(gpu, oracle, maps, opts):
    g = maps("small")
    load_map(gpu, g)
    poses = g["poses"].astype(np.float64)
    ref = oracle.scans(g["offsets"], g["pts"], g["nrm"]).find_stf(poses, **opts)
    out = gpu.find_stf(poses, opts=gpu.stf_opts(**opts))
    assert_same_stf(out, ref)
    if opts.get("cap", 6) > 0:
        assert out["n_queries"] == ref["n_queries"]


@pytest.mark.parametrize("rng_", [(0, 159), (10, 90), (37, 37), (50, 1000), (120, 60)])
def test_find_stf_pose_ranges(gpu, oracle, maps, rng_):
    g = maps("small")
    load_map(gpu, g)
    poses = g["poses"].astype(np.float64)
    ref = oracle.scans(g["offsets"], g["pts"], g["nrm"]).find_stf(poses, min_pose=rng_[0], max_pose=rng_[1])
    out = gpu.find_stf(poses, min_pose=rng_[0], max_pose=rng_[1])
    assert_same_stf(out, ref)
    assert out["n_queries"] == ref["n_queries"]


def test_find_stf_shards_concatenate(gpu, oracle, maps):
    from hitl_slam_b200.sharding import concat_stf, shard_ranges
    g = maps("small")
    load_map(gpu, g)
    poses = g["poses"].astype(np.float64)
    ref = oracle.scans(g["offsets"], g["pts"], g["nrm"]).find_stf(poses)
    for world in (2, 3, 8):
        parts = [gpu.find_stf(poses, src_lo=a, src_hi=b) for a, b in shard_ranges(g["offsets"], world)]
        cat = concat_stf(parts)
        assert_same_stf(cat, ref)
        assert cat["n_queries"] == ref["n_queries"]


def test_find_stf_is_independent_of_the_tile_schedule(gpu, oracle, maps):
    """The second and later calls hand tiles out heaviest-first (order measured by the previous call);
    results must not depend on it, also when the source range or the poses change in between."""
    g = maps("small")
    load_map(gpu, g)
    poses = g["poses"].astype(np.float64)
    S = oracle.scans(g["offsets"], g["pts"], g["nrm"])
    ref = S.find_stf(poses)
    for _ in range(3):
        assert_same_stf(gpu.find_stf(poses), ref)
    part = gpu.find_stf(poses, src_lo=30, src_hi=90)
    assert_same_stf(part, S.find_stf(poses, src_lo=30, src_hi=90))
    assert_same_stf(gpu.find_stf(poses, src_lo=30, src_hi=90), part)
    moved = poses + np.random.default_rng(4).normal(size=poses.shape) * 0.02
    assert_same_stf(gpu.find_stf(moved), S.find_stf(moved))
    assert_same_stf(gpu.find_stf(poses), ref)


@pytest.mark.parametrize("max_len", [32, 7, 4, 1])
def test_find_stf_is_independent_of_the_tiling(gpu, oracle, maps, max_len):
    """Tiles (the unit of work of the search) may be any partition of a scan into runs of <= 32 points."""
    g = maps("small")
    load_map(gpu, g)
    poses = g["poses"].astype(np.float64)
    S = oracle.scans(g["offsets"], g["pts"], g["nrm"])
    ref = S.find_stf(poses)
    try:
        gpu.debug_set_tiling(max_len, adaptive=False)
        out = gpu.find_stf(poses)
        assert_same_stf(out, ref)
        assert out["n_queries"] == ref["n_queries"]
        assert_same_stf(gpu.find_stf(poses, src_lo=20, src_hi=77), S.find_stf(poses, src_lo=20, src_hi=77))
    finally:
        gpu.debug_set_tiling(32, adaptive=True)


@pytest.mark.parametrize("max_len,parts", [(32, 2), (32, 5), (9, 16), (1, 3)])
def test_find_stf_target_axis_split_reapplies_the_cap(gpu, oracle, maps, max_len, parts):
    """A tile may be cut into consecutive TARGET ranges searched concurrently, each with a private cap state; the
    merge keeps each point's first `cap` matches in target order, which is the reference's sequential loop."""
    g = maps("small")
    load_map(gpu, g)
    poses = g["poses"].astype(np.float64)
    S = oracle.scans(g["offsets"], g["pts"], g["nrm"])
    ref = S.find_stf(poses)
    try:
        gpu.debug_set_tiling(max_len, adaptive=False, target_parts=parts)
        for cull in (0, 1):
            out = gpu.find_stf(poses, opts=gpu.stf_opts(disable_culling=cull))
            assert_same_stf(out, ref)
            assert out["n_queries"] == ref["n_queries"]
        assert_same_stf(gpu.find_stf(poses, src_lo=20, src_hi=77), S.find_stf(poses, src_lo=20, src_hi=77))
        for cap, skip, min_corr in ((2, 1, 10), (1, 3, 2), (11, 2, 0)):
            out = gpu.find_stf(poses, min_pose=7, max_pose=140, opts=gpu.stf_opts(cap=cap, skip=skip, min_corr=min_corr))
            want = S.find_stf(poses, min_pose=7, max_pose=140, cap=cap, skip=skip, min_corr=min_corr)
            assert_same_stf(out, want)
            assert out["n_queries"] == want["n_queries"]
    finally:
        gpu.debug_set_tiling(32, adaptive=True)


def test_find_stf_adaptive_tile_splitting_keeps_results(gpu, oracle, maps):
    """Heavy tiles are split after a search that was dominated by them (here: forced by a tiny shard);
    later searches over any range still return the reference's lists."""
    g = maps("c1")
    load_map(gpu, g)
    poses = g["poses"].astype(np.float64)
    S = oracle.scans(g["offsets"], g["pts"], g["nrm"])
    ref = S.find_stf(poses, src_lo=100, src_hi=140)
    for _ in range(4):                                   # each call may split further
        assert_same_stf(gpu.find_stf(poses, src_lo=100, src_hi=140), ref)
    full = S.find_stf(poses)
    for _ in range(2):
        out = gpu.find_stf(poses)
        assert_same_stf(out, full)
        assert out["n_queries"] == full["n_queries"]


def test_find_stf_work_feedback_covers_exactly_the_searched_sources(gpu, maps):
    g = maps("small")
    load_map(gpu, g)
    poses = g["poses"].astype(np.float64)
    gpu.find_stf(poses, src_lo=40, src_hi=100, fetch=False)
    w = gpu.stf_work()
    assert len(w) == len(poses) and (w[:40] == 0).all() and (w[100:] == 0).all() and (w[40:100] > 0).all()
    from hitl_slam_b200.sharding import shard_ranges_by_work
    gpu.find_stf(poses, fetch=False)
    cuts = shard_ranges_by_work(gpu.stf_work(), 4)
    assert cuts[0][0] == 0 and cuts[-1][1] == len(poses) and all(hi > lo for lo, hi in cuts)


def test_find_stf_ragged_and_empty_scans(gpu, oracle):
    rng = np.random.default_rng(11)
    # a strip of overlapping random scans, some empty, sizes from 1 to 300
    off, pts, nrm = random_scans(rng, 60, 1, 300, empty=(0, 7, 8, 59))
    pts = (pts * 0.5).astype(np.float32)
    poses = np.stack([np.linspace(0, 3, 60), rng.normal(size=60) * 0.1, rng.uniform(-0.3, 0.3, 60)], 1)
    gpu.set_scans(off, pts, nrm)
    gpu.build_kdtrees()
    S = oracle.scans(off, pts, nrm)
    for opts in (dict(), dict(min_corr=0, min_cos=-2.0), dict(thr=0.5, min_cos=-2.0, cap=4)):
        ref = S.find_stf(poses, **opts)
        out = gpu.find_stf(poses, opts=gpu.stf_opts(**opts))
        assert_same_stf(out, ref)
        assert out["n_queries"] == ref["n_queries"]
    assert len(S.find_stf(poses, min_corr=0, min_cos=-2.0)["k"]) > 0


@pytest.mark.parametrize("seed", [21, 22, 23])
def test_angle_gate_prefilter_with_arbitrary_normals(gpu, oracle, seed):
    """The direction prefilter must stay exact when normals are not unit vectors (the reference's loader leaves them un-normalised),
    zero, tiny, huge, pointing anywhere, with pose angles of many turns, for tight and loose gates: same lists as the oracle and as the
    search without any culling; poses beyond the filter's float-angle range (|theta| > 256 rad) simply bypass it."""
    rng = np.random.default_rng(seed)
    off, pts, nrm = random_scans(rng, 48, 20, 260, empty=(5,))
    pts = (pts * 0.4).astype(np.float32)
    m = len(nrm)
    mag = np.exp(rng.uniform(np.log(0.05), np.log(40.0), m)).astype(np.float32)
    mag[rng.random(m) < 0.05] = 0.0                                  # zero normals: the dot product is 0, never a match for min_cos > 0
    mag[rng.random(m) < 0.03] = 1e-30
    mag[rng.random(m) < 0.02] = 3e18
    nrm = (nrm * mag[:, None]).astype(np.float32)
    poses = np.stack([np.linspace(0, 2.5, 48), rng.normal(size=48) * 0.1, rng.uniform(-40.0, 40.0, 48)], 1)
    if seed == 23:
        poses[::5, 2] += 300.0                                       # beyond kDirMaxTheta: those pairs bypass the prefilter
    gpu.set_scans(off, pts, nrm)
    gpu.build_kdtrees()
    S = oracle.scans(off, pts, nrm)
    culled = 0
    for opts in (dict(min_corr=0), dict(min_corr=0, min_cos=0.999, thr=0.3), dict(min_corr=1, min_cos=0.2, thr=0.25, cap=3), dict(min_corr=0, min_cos=1e-6)):
        ref = S.find_stf(poses, **opts)
        out = gpu.find_stf(poses, opts=gpu.stf_opts(**opts))
        raw = gpu.find_stf(poses, opts=gpu.stf_opts(disable_culling=1, **opts))
        assert_same_stf(out, ref)
        assert_same_stf(raw, ref)
        assert out["n_queries"] == ref["n_queries"] == raw["n_queries"]
        culled += out["n_dir_culled"]
    assert culled > 0


def test_find_stf_empty_problem(gpu):
    gpu.set_scans(np.zeros(1, np.uint32), np.zeros((0, 2), np.float32), np.zeros((0, 2), np.float32))
    gpu.build_kdtrees()
    out = gpu.find_stf(np.zeros(0))
    assert out["n_pairs"] == 0 and out["n_matches"] == 0 and list(out["pair_off"]) == [0]


def test_find_stf_c1_full_size(gpu, oracle, maps):
    """BASELINE config 1 (500 poses x 360 points) end to end, bit-exact."""
    g = maps("c1")
    load_map(gpu, g)
    poses = g["poses"].astype(np.float64)
    ref = oracle.scans(g["offsets"], g["pts"], g["nrm"]).find_stf(poses)
    out = gpu.find_stf(poses)
    assert_same_stf(out, ref)
    assert out["n_queries"] == ref["n_queries"]
    assert out["n_pairs"] > 1000


def test_find_stf_culling_is_result_preserving_at_scale(gpu, maps):
    """Size-independent property: AABB culling never changes the result (2000 x 360, too big for the oracle to be quick)."""
    g = maps("c2", n_poses=2000, beams=360)
    load_map(gpu, g)
    poses = g["poses"].astype(np.float64)
    a = gpu.find_stf(poses)
    b = gpu.find_stf(poses, opts=gpu.stf_opts(disable_culling=1))
    assert_same_stf(a, b)
    assert a["n_queries"] == b["n_queries"] and a["n_traversals"] < b["n_traversals"]
    # the second (fine) occupancy level only removes walks that would find nothing
    gpu.debug_set_fine_occupancy(False)
    c = gpu.find_stf(poses)
    gpu.debug_set_fine_occupancy(True)
    assert_same_stf(a, c)
    assert a["n_queries"] == c["n_queries"] and a["n_traversals"] < c["n_traversals"] < b["n_traversals"]
    # the angle-gate direction prefilter only removes walks whose result would fail the normal gate
    gpu.debug_set_fine_occupancy(3)
    d = gpu.find_stf(poses)
    gpu.debug_set_fine_occupancy(True)
    assert_same_stf(a, d)
    assert a["n_queries"] == d["n_queries"] and d["n_dir_culled"] == 0 and a["n_dir_culled"] > 0
    assert a["n_traversals"] + a["n_dir_culled"] >= d["n_traversals"] * 0.98 and a["n_gate_fail"] < d["n_gate_fail"]
    # the tile-box vs scan cull (mip level of the coarse bitmap) only removes (tile, target) pairs whose every per-point test would fail
    gpu.debug_set_fine_occupancy(5)
    e = gpu.find_stf(poses)
    gpu.debug_set_fine_occupancy(True)
    assert_same_stf(a, e)
    assert a["n_queries"] == e["n_queries"] and a["n_tile_pairs"] < e["n_tile_pairs"] and a["n_coarse_pass"] <= e["n_coarse_pass"]
    counts = np.diff(a["pair_off"].astype(np.int64))
    assert (counts > 10).all()
    # per source point at most `cap` matches over all pairs
    key = a["pair_i"].astype(np.int64).repeat(counts) * (1 << 20) + a["k"]
    assert np.unique(key, return_counts=True)[1].max() <= 6
    # pairs sorted by (i, j), k ascending inside a pair
    pij = a["pair_i"].astype(np.int64) * (1 << 32) + a["pair_j"]
    assert (np.diff(pij) > 0).all()
    for bb in range(0, len(counts), 997):
        kk = a["k"][int(a["pair_off"][bb]):int(a["pair_off"][bb + 1])]
        assert (np.diff(kk.astype(np.int64)) > 0).all()


def test_find_vo_matches_oracle(gpu, oracle, maps):
    g = maps("small")
    load_map(gpu, g)
    poses = g["poses"].astype(np.float64)
    S = oracle.scans(g["offsets"], g["pts"], g["nrm"])
    for lo, hi in ((0, 159), (20, 60), (5, 5)):
        ref = S.find_vo(poses, lo, hi)
        out = gpu.find_vo(poses, lo, hi)
        for a, b in zip(ref, out):
            assert np.array_equal(a, b)
    assert len(S.find_vo(poses)[0]) > 0


# ---- world transform + EM ---------------------------------------------------------------------------
def test_world_transform_bit_exact(gpu, oracle, maps):
    g = maps("small")
    load_map(gpu, g)
    ref = oracle.scans(g["offsets"], g["pts"], g["nrm"], build_trees=False).world_transform(g["poses"])
    out = gpu.world_transform(g["poses"])
    assert np.array_equal(ref.view(np.uint32), out.view(np.uint32))


@pytest.mark.parametrize("name", ["small", "c1"])
def test_em_inliers_and_assign_bit_exact(gpu, oracle, maps, name):
    from hitl_slam_b200 import synth
    g = maps(name)
    load_map(gpu, g)
    world = gpu.world_transform(g["poses"])
    strokes = synth.make_strokes(g)
    rng = np.random.default_rng(2)
    cases = [strokes, strokes + rng.normal(size=(4, 2)).astype(np.float32) * 0.02,
             np.array([[0, 0], [30, 9], [1, 1], [1, 1]], np.float32),      # long diagonal stroke / degenerate stroke
             np.array([[-5, -5], [-4, -5], [100, 100], [101, 100]], np.float32)]   # nothing selected
    for s in cases:
        for seg in (s[:2], s[2:]):
            op, oi = oracle.em_inliers(g["offsets"], world, seg.reshape(-1))
            gp, gi, gxy = gpu.em_inliers(seg.reshape(-1))
            assert np.array_equal(op, gp) and np.array_equal(oi, gi)
            assert np.array_equal(gxy, world[g["offsets"][gp].astype(np.int64) + gi]) if len(gp) else True
            assert gpu.em_inliers(seg.reshape(-1), fetch=False) == len(op)
        ref = oracle.em_assign(g["offsets"], world, s)
        out = gpu.em_assign(s)
        for f in range(2):
            for a, b in zip(ref[f], out[f]):
                assert np.array_equal(a, b)
    gold = np.load(os.path.join(ROOT, "tests", "golden", "small_em.npz"))
    if name == "small":
        gp, gi, _ = gpu.em_inliers(gold["strokes"][:2].reshape(-1))
        assert np.array_equal(gp, gold["inl_pose"]) and np.array_equal(gi, gold["inl_idx"])


@pytest.mark.parametrize("name", ["small", "c1"])
def test_em_chunk_cull_is_result_preserving(gpu, oracle, maps, name):
    """E-steps after the first one on the same world clouds skip the 2048-point chunks whose bounding box is out of the stroke's
    reach.  The reach follows the reference's distance function, quirk included (the projection parameter is compared with 1.0 in
    METRES: a stroke shorter than 1 m collects the carrier line's points up to 1 m from its first endpoint, a longer one loses the
    points past 1 m unless they are near the far endpoint).  Oracle inliers == culled == unculled, for short / long / far strokes."""
    from hitl_slam_b200 import synth
    g = maps(name)
    load_map(gpu, g)
    world = gpu.world_transform(g["poses"])
    a, b = synth.make_strokes(g)[:2]
    d = (b - a) / max(float(np.linalg.norm(b - a)), 1e-6)
    rng = np.random.default_rng(11)
    segs = [np.concatenate([a, a + 0.25 * d]), np.concatenate([a, a + 0.999 * d]), np.concatenate([a, a + 1.001 * d]), np.concatenate([a, a + 4.0 * d]),
            np.concatenate([a + 4.0 * d, a]), np.concatenate([a, a - 0.5 * d]), np.array([1e4, 1e4, 1e4 + 1, 1e4], np.float32),
            np.array([np.nan, 0, 1, 1], np.float32), np.concatenate([a, a])]
    lo, hi = world.min(axis=0), world.max(axis=0)
    for _ in range(12):
        p = lo + rng.random(2) * (hi - lo)
        segs.append(np.concatenate([p, p + rng.normal(size=2) * rng.choice([0.2, 1.0, 5.0])]))
    for thr in (0.03, 0.4):
        for seg in segs:
            seg = np.asarray(seg, np.float32)
            op, oi = oracle.em_inliers(g["offsets"], world, seg, thr=thr)
            gpu.debug_set_em_cull(True)
            gpu.em_inliers(seg, thr=thr, fetch=False)                      # (re)records the chunk boxes
            cp, ci, cxy = gpu.em_inliers(seg, thr=thr)                     # culled pass
            gpu.debug_set_em_cull(False)
            up, ui, uxy = gpu.em_inliers(seg, thr=thr)
            assert np.array_equal(op, up) and np.array_equal(oi, ui), seg
            assert np.array_equal(cp, up) and np.array_equal(ci, ui) and np.array_equal(cxy, uxy), seg
    gpu.debug_set_em_cull(True)
    # new world clouds invalidate the boxes: the same strokes on shifted clouds
    shifted = world + np.float32(3.0)
    gpu.set_world_clouds(shifted)
    for seg in segs[:6]:
        seg = np.asarray(seg, np.float32)
        op, oi = oracle.em_inliers(g["offsets"], shifted, seg + np.float32(3.0))
        for _ in range(2):
            cp, ci, _ = gpu.em_inliers(seg + np.float32(3.0))
            assert np.array_equal(op, cp) and np.array_equal(oi, ci)


@pytest.mark.parametrize("name", ["small", "c1"])
def test_em_rounds_chained_on_the_device_equal_one_call_per_round(gpu, maps, name):
    """hitl_em_refit_chain: round r of a stroke reads the stroke round r-1 left in device memory.  Every returned round is bit for
    bit what hitl_em_refit returns when the host feeds the previous result back — for one stroke and for two side by side."""
    from hitl_slam_b200 import synth
    g = maps(name, drift_xy=0.012, drift_th=0.004)
    load_map(gpu, g)
    gpu.world_transform(g["poses"])
    strokes = synth.pick_strokes(g, min_sep=0.045).astype(np.float32).reshape(2, 4)
    rng = np.random.default_rng(5)
    for case in (strokes, strokes + rng.normal(size=(2, 4)).astype(np.float32) * 0.02, np.array([[500, 500, 501, 500.5], strokes[0]], np.float32)):
        want, winfo = [], []
        for seg in case:                                            # the one-call-per-round loop, 4 rounds regardless of convergence
            cur, rows, infos = seg.copy(), [], []
            for _ in range(4):
                cur, info = gpu.em_refit(cur)
                rows.append(cur.copy()); infos.append(info)
            want.append(rows); winfo.append(infos)
        for cull in (True, False):
            gpu.debug_set_em_cull(cull)
            for rounds in (1, 2, 4):
                segs, infos = gpu.em_refit_chain(case, rounds=rounds)
                for r in range(rounds):
                    for q in range(2):
                        assert np.array_equal(segs[r, q].view(np.uint32), want[q][r].view(np.uint32)), (cull, rounds, r, q)
                        for key in ("theta", "initial_cost", "final_cost", "n_inliers", "iterations", "evaluations", "termination"):
                            assert infos[r][q][key] == winfo[q][r][key], (key, cull, rounds, r, q)
            one, _ = gpu.em_refit_chain(case[1:], rounds=3)              # a single stroke
            assert all(np.array_equal(one[r, 0], want[1][r]) for r in range(3))
        gpu.debug_set_em_cull(True)
    with pytest.raises(Exception):
        gpu.em_refit_chain(case, rounds=5)
    with pytest.raises(Exception):
        gpu.em_refit_chain(np.zeros((3, 4), np.float32), rounds=1)


def test_em_with_empty_scans(gpu, oracle):
    rng = np.random.default_rng(4)
    off, pts, nrm = random_scans(rng, 40, 1, 90, empty=(0, 3, 4, 39))
    gpu.set_scans(off, pts, nrm)
    gpu.set_world_clouds(pts)
    seg = np.array([-1, -1, 1.5, 1.2, -2, 0.5, 2, 0.4], np.float32)
    op, oi = oracle.em_inliers(off, pts, seg[:4], thr=0.2)
    gp, gi, _ = gpu.em_inliers(seg[:4], thr=0.2)
    assert len(op) > 0 and np.array_equal(op, gp) and np.array_equal(oi, gi)
    ref, out = oracle.em_assign(off, pts, seg, thr=0.2, min_obs=1), gpu.em_assign(seg, thr=0.2, min_obs=1)
    for f in range(2):
        for a, b in zip(ref[f], out[f]):
            assert np.array_equal(a, b)


# ---- residuals / Jacobians / normal equations -----------------------------------------------------------
def _human_constraints(n):
    hc_i = np.array([[2, n - 3, 1], [4, n - 4, 2], [5, n - 5, 3], [6, n - 6, 0], [4, n - 2, 5]], np.int32)
    hc_f = np.array([[0.3, -0.2, 0.1, 0.0], [1.0, 0.5, -0.4, 1.2], [0, 0, 1.57, 0], [0, 0, 0.02, 0], [-0.7, 0.1, 3.0, -0.5]], np.float32)
    return hc_i, hc_f


@pytest.mark.parametrize("name", ["tiny", "small"])
def test_eval_matches_oracle(gpu, oracle, maps, name):
    g = maps(name)
    load_map(gpu, g)
    poses = g["poses"].astype(np.float64)
    n = len(poses)
    S = oracle.scans(g["offsets"], g["pts"], g["nrm"])
    corr = gpu.find_stf(poses)
    rng = np.random.default_rng(0)
    x = poses + rng.normal(size=poses.shape) * 0.01
    consts = oracle.odometry_consts(g["poses"])
    hc_i, hc_f = _human_constraints(n)
    blk_i, blk_d = oracle.human_blocks(g["poses"], hc_i, hc_f)
    gpu.set_odometry_blocks(consts)
    gpu.set_human_blocks(blk_i, blk_d)
    gpu.set_stf_blocks_from_search()
    r_stf, J_stf = S.eval_stf(x, corr)
    r_odo, J_odo = oracle.eval_odometry(consts, x)
    r_hum, J_hum = oracle.eval_human(blk_i, blk_d, x)
    for precision, tol in ((0, REL64), (1, REL32)):
        out = gpu.eval(x, precision=precision)
        assert rel_err(out["r_stf"], r_stf) <= tol and rel_err(out["J_stf"], J_stf) <= tol
        assert rel_err(out["r_odometry"], r_odo) <= tol and rel_err(out["J_odometry"], J_odo) <= tol
        assert rel_err(out["r_human"], r_hum) <= tol and rel_err(out["J_human"], J_hum) <= tol
    # explicit CSR upload gives the same numbers as the device-resident search result
    gpu.set_stf_blocks(corr)
    out2 = gpu.eval(x)
    assert np.array_equal(out2["r_stf"], gpu.eval(x)["r_stf"]) and rel_err(out2["J_stf"], J_stf) <= REL64
    if name == "tiny":
        gold = np.load(os.path.join(ROOT, "tests", "golden", "tiny_eval.npz"))
        o3 = gpu.eval(gold["x"])
        assert rel_err(o3["r_stf"], gold["r_stf"]) <= REL64 and rel_err(o3["J_stf"], gold["J_stf"]) <= REL64
        assert rel_err(o3["r_odometry"], gold["r_odo"]) <= REL64 and rel_err(o3["J_odometry"], gold["J_odo"]) <= REL64


def test_eval_point_to_line_matches_oracle(gpu, oracle, maps):
    g = maps("tiny")
    load_map(gpu, g)
    n = len(g["poses"])
    rng = np.random.default_rng(3)
    x = g["poses"].astype(np.float64) + rng.normal(size=(n, 3)) * 0.01
    sizes = rng.integers(1, 80, 25)
    blk_off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    m = int(blk_off[-1])
    blk_pose = rng.integers(0, n, 25).astype(np.uint32)
    pts = rng.normal(size=(m, 2)).astype(np.float32) * 3
    ang = rng.uniform(0, 6.28, m)
    ln = np.stack([np.cos(ang), np.sin(ang)], 1).astype(np.float32)
    lo = rng.normal(size=m).astype(np.float32)
    valid = (rng.uniform(size=m) > 0.2).astype(np.uint8)
    valid[blk_off[3]:blk_off[4]] = 0     # a block with no valid point: residual 0, Jacobian 0
    gpu.set_odometry_blocks(np.zeros((0, 9), np.float32))
    gpu.set_human_blocks(np.zeros((0, 2), np.int32), np.zeros((0, 4)))
    gpu.set_stf_blocks(dict(pair_i=np.zeros(0, np.uint32), pair_j=np.zeros(0, np.uint32), pair_off=np.zeros(1, np.uint64),
                            k=np.zeros(0, np.uint32), idx=np.zeros(0, np.uint32)))
    gpu.set_p2l_glob_blocks(blk_pose, blk_off, pts, ln, lo, valid, 0.05, 1 / 50.0)
    pose_idx = rng.integers(0, n, m).astype(np.uint32)
    gpu.set_p2l_blocks(pose_idx, pts, ln, lo, valid, 0.05, 1 / 50.0)
    rg, Jg = oracle.eval_p2l_glob(blk_pose, blk_off, pts, ln, lo, valid, 0.05, 1 / 50.0, x)
    rs, Js = oracle.eval_p2l(pose_idx, pts, ln, lo, valid, 0.05, 1 / 50.0, x)
    for precision, tol in ((0, REL64), (1, 2e-5)):
        out = gpu.eval(x, precision=precision)
        assert rel_err(out["r_p2l_glob"][:, 0], rg) <= tol and rel_err(out["J_p2l_glob"], Jg) <= tol
        assert rel_err(out["r_p2l"][:, 0], rs) <= tol and rel_err(out["J_p2l"], Js) <= tol
    assert out["r_p2l_glob"][3, 0] == 0 and (out["J_p2l_glob"][3] == 0).all()
    gpu.set_p2l_glob_blocks(np.zeros(0, np.uint32), np.zeros(1, np.uint64), np.zeros((0, 2), np.float32), np.zeros((0, 2), np.float32),
                            np.zeros(0, np.float32), np.zeros(0, np.uint8), 1.0, 1.0)
    gpu.set_p2l_blocks(np.zeros(0, np.uint32), np.zeros((0, 2), np.float32), np.zeros((0, 2), np.float32), np.zeros(0, np.float32),
                       np.zeros(0, np.uint8), 1.0, 1.0)


def test_normal_equations_match_oracle_jacobians(gpu, oracle, maps):
    g = maps("small")
    load_map(gpu, g)
    poses = g["poses"].astype(np.float64)
    n = len(poses)
    S = oracle.scans(g["offsets"], g["pts"], g["nrm"])
    corr = gpu.find_stf(poses)
    rng = np.random.default_rng(1)
    x = poses + rng.normal(size=poses.shape) * 0.01
    consts = oracle.odometry_consts(g["poses"])
    hc_i, hc_f = _human_constraints(n)
    blk_i, blk_d = oracle.human_blocks(g["poses"], hc_i, hc_f)
    gpu.set_odometry_blocks(consts)
    gpu.set_human_blocks(blk_i, blk_d)
    gpu.set_stf_blocks_from_search()
    r_stf, J_stf = S.eval_stf(x, corr)
    r_odo, J_odo = oracle.eval_odometry(consts, x)
    r_hum, J_hum = oracle.eval_human(blk_i, blk_d, x)
    H, gvec = np.zeros((n, 3, 3)), np.zeros((n, 3))
    Hoff = []
    for b in range(n - 1):
        for side, p in ((0, b), (1, b + 1)):
            H[p] += J_odo[b, side].T @ J_odo[b, side]
            gvec[p] += J_odo[b, side].T @ r_odo[b]
        Hoff.append(J_odo[b, 0].T @ J_odo[b, 1])
    for b in range(len(blk_i)):
        p = blk_i[b, 1]
        H[p] += J_hum[b].T @ J_hum[b]
        gvec[p] += J_hum[b].T @ r_hum[b]
    for b in range(len(corr["pair_i"])):
        for side, p in ((0, int(corr["pair_i"][b])), (1, int(corr["pair_j"][b]))):
            H[p] += J_stf[b, side].T @ J_stf[b, side]
            gvec[p] += J_stf[b, side].T @ r_stf[b]
        Hoff.append(J_stf[b, 0].T @ J_stf[b, 1])
    cost = 0.5 * ((r_odo ** 2).sum() + (r_hum ** 2).sum() + (r_stf ** 2).sum())
    out = gpu.normal_eq(x)
    assert rel_err(out["H_diag"], H) <= REL64 and rel_err(out["g"], gvec) <= REL64
    assert rel_err(out["H_off"], np.array(Hoff)) <= REL64
    assert abs(out["cost"] - cost) <= REL64 * cost
    ptr, nd = gpu.normal_eq_device()
    assert ptr and nd == 12 * n + 1
    # deterministic mode (hitl_set_deterministic): per-pose gather instead of FP64 atomics — the same blocks to rounding, and the SAME BITS
    # on every run, also after the blocks were registered again and with other work in between
    gpu.set_deterministic(True)
    try:
        d1 = gpu.normal_eq(x)
        assert rel_err(d1["H_diag"], H) <= REL64 and rel_err(d1["g"], gvec) <= REL64 and rel_err(d1["H_off"], np.array(Hoff)) <= REL64
        assert abs(d1["cost"] - cost) <= REL64 * cost
        assert rel_err(d1["H_diag"], out["H_diag"]) <= 1e-12 and rel_err(d1["g"], out["g"]) <= 1e-12
        gpu.eval(x + 0.5)
        d2 = gpu.normal_eq(x)
        gpu.set_odometry_blocks(consts)
        gpu.set_human_blocks(blk_i, blk_d)
        gpu.set_stf_blocks_from_search()
        d3 = gpu.normal_eq(x)
        for other in (d2, d3):
            for key in ("H_diag", "g", "H_off"):
                assert np.array_equal(other[key], d1[key]), key
            assert other["cost"] == d1["cost"]
        ga = gpu.gather_stf_blocks(0)                            # r and J of every block are resident after the deterministic pass
        assert np.array_equal(ga["r"], gpu.eval(x)["r_stf"])
    finally:
        gpu.set_deterministic(False)


# ---- error behaviour ------------------------------------------------------------------------------------
def test_error_paths(gpu):
    from hitl_slam_b200 import HitlError, HitlGpu
    fresh = HitlGpu(0)
    with pytest.raises(HitlError):
        fresh.find_stf(np.zeros(0))                      # scans not set (library state check)
    with pytest.raises(HitlError):
        fresh.find_stf(np.zeros(3))                      # wrapper: 3 values for a map of 0 poses
    with pytest.raises(HitlError):
        fresh.em_inliers(np.zeros(4, np.float32))        # world clouds not set
    fresh.set_scans(np.array([0, 2], np.uint32), np.zeros((2, 2), np.float32), np.zeros((2, 2), np.float32))
    with pytest.raises(HitlError):
        fresh.set_scans(np.array([0, 3, 2], np.uint32), np.zeros((3, 2), np.float32), np.zeros((3, 2), np.float32))
    with pytest.raises(HitlError):
        fresh.set_human_blocks(np.array([[3, 0]], np.int32), np.zeros((1, 4)))   # corner type unsupported, as in the reference
    # a rejected hitl_set_scans leaves the context as it was (validation precedes every state change) ...
    fresh.set_odometry_blocks(np.zeros((0, 9), np.float32))
    assert fresh.lib.hitl_set_scans(fresh.ctx, 2, np.array([0, 3, 2], np.uint32), np.zeros(6, np.float32), np.zeros(6, np.float32)) != 0
    fresh.n_poses, fresh.n_points = 1, 2
    fresh.build_kdtrees()
    assert fresh.find_stf(np.zeros(3))["n_pairs"] == 0
    with pytest.raises(HitlError):
        fresh.find_stf(np.zeros(6))                      # pose array of another map
    # ... and a successful one drops the residual blocks registered for the previous map (their indices were checked against it)
    fresh.set_scans(np.array([0, 2, 4, 6], np.uint32), np.zeros((6, 2), np.float32), np.zeros((6, 2), np.float32))
    fresh.set_odometry_blocks(np.zeros((2, 9), np.float32) + 1)
    assert fresh.layout().n_odometry == 2
    fresh.set_scans(np.array([0, 2], np.uint32), np.zeros((2, 2), np.float32), np.zeros((2, 2), np.float32))
    assert fresh.layout().n_odometry == 0
    with pytest.raises(HitlError):
        fresh.set_p2l_glob_blocks(np.zeros(2, np.uint32), np.array([0, 3, 2], np.uint64), np.zeros((3, 2), np.float32), np.zeros((3, 2), np.float32),
                                  np.zeros(3, np.float32), np.ones(3, np.uint8), 0.05, 0.025)   # blk_off not monotone
    with pytest.raises(HitlError):
        fresh.debug_relative_pose(np.zeros(3), [0], [5])                                        # pose index out of range
    fresh.build_kdtrees()
    assert fresh.find_stf(np.zeros(3), opts=fresh.stf_opts(min_corr=0xFFFFFFFF))["n_pairs"] == 0   # min_corr + 1 must not wrap to a zero divisor
    fresh.close()


# ---- the benchmarked workloads themselves (BASELINE configs 2 and 3 at full size) -------------------------
def _spread_chunks(n, n_chunks, width):
    """Source-pose chunks spread over the trajectory, the first and the LAST poses included (late poses see every earlier lap: heaviest tiles)."""
    return [(int(lo), int(lo) + width) for lo in np.linspace(0, n - width, n_chunks).astype(np.int64)]


@pytest.mark.parametrize("name,n_chunks", [("c2", 10), ("c3", 5)])
def test_find_stf_full_size_matches_oracle_on_chunks(gpu, oracle, maps, name, n_chunks):
    """c2 (5 000 x 720) and c3 (20 000 x 1080) exactly as bench.py runs them.  The whole map is too large for the oracle, so the oracle
    searches 16-pose source chunks against ALL targets and the rows of the GPU result whose source lies in the chunk must be identical
    (pairs, offsets, k, idx).  The GPU search is compared after >= 4 warm-up calls, i.e. with adaptive re-tiling, target-axis splits of
    heavy tiles and heaviest-first tickets live — the machinery that only engages at scale — first over the whole map (world = 1), then
    on the first / a middle / the last source shard of an 8-way cut (world = 8), each with its own warm-up."""
    from hitl_slam_b200.sharding import shard_ranges
    g = maps(name)
    load_map(gpu, g)
    poses = g["poses"].astype(np.float64)
    n = len(poses)
    S = oracle.scans(g["offsets"], g["pts"], g["nrm"])
    tiles0 = None
    for _ in range(4):
        info = gpu.find_stf(poses, fetch=False)
        tiles0 = tiles0 or info["n_tiles"]
    res = gpu.find_stf(poses)
    assert res["n_tiles"] >= tiles0
    checks = S.check_chunks(res, poses, _spread_chunks(n, n_chunks, 16))
    assert all(c[2] for c in checks), checks
    assert sum(c[4] for c in checks) > 1000
    # size-independent properties over the WHOLE result: pairs sorted by (i, j), > 10 matches per pair, <= cap matches per source point
    counts = np.diff(res["pair_off"].astype(np.int64))
    assert (counts > 10).all()
    assert (np.diff(res["pair_i"].astype(np.int64) * (1 << 32) + res["pair_j"]) > 0).all()
    per_point = np.bincount(g["offsets"].astype(np.int64)[res["pair_i"]].repeat(counts) + res["k"], minlength=int(g["offsets"][-1]))
    assert per_point.max() <= 6
    shards = shard_ranges(g["offsets"], 8)
    for r in (0, 3, 7):
        lo, hi = shards[r]
        for _ in range(4):
            gpu.find_stf(poses, src_lo=lo, src_hi=hi, fetch=False)
        part = gpu.find_stf(poses, src_lo=lo, src_hi=hi)
        assert len(part["pair_i"]) and part["pair_i"].min() >= lo and part["pair_i"].max() < hi
        checks = S.check_chunks(part, poses, [(hi - 16, hi)])
        assert all(c[2] for c in checks), (r, checks)
    if name == "c3":
        from conftest import _MAPS
        _MAPS.pop(("c3", ()), None)                      # 0.7 GB of host arrays: not kept for the rest of the session
