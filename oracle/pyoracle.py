"""ctypes front-end of the CPU oracle (oracle/hitl_oracle.hpp).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (hitl_slam_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C")
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C")


def build(ref=True):
    """Compile the oracle (and oracle/_ref when /root/reference is present)."""
    subprocess.run(["make", "-s", "-C", HERE] + ([] if ref else ["_build/liboracle.so"]), check=True)


def _load(name):
    path = os.path.join(HERE, "_build", name)
    if not os.path.exists(path):
        build()
    return C.CDLL(path)


class Oracle:
    """One loaded oracle library (parity build by default, `fast=True` for the timing build)."""

    def __init__(self, fast=False):
        self.lib = lib = _load("liboracle_fast.so" if fast else "liboracle.so")
        lib.orc_create.restype = C.c_void_p
        lib.orc_create.argtypes = [C.c_uint32, _u32p, _f32p, _f32p, C.c_int]
        lib.orc_destroy.argtypes = [C.c_void_p]
        lib.orc_flatten.argtypes = [C.c_void_p, _f32p, _i32p, _i32p]
        lib.orc_query.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, _f32p, C.c_float, C.c_int, _f32p, _i32p]
        lib.orc_radius.restype = C.c_uint32
        lib.orc_radius.argtypes = [C.c_void_p, C.c_uint32, C.c_float, C.c_float, C.c_float, _i32p, C.c_uint32]
        lib.orc_relative_pose.argtypes = [_f64p, C.c_uint32, C.c_uint32, _f32p]
        lib.orc_find_stf.argtypes = [C.c_void_p, _f64p, C.c_uint64, C.c_uint64, C.c_float, C.c_float, C.c_int,
                                     C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, _u64p]
        lib.orc_find_stf_strided.argtypes = [C.c_void_p, _f64p, C.c_uint64, C.c_uint64, C.c_float, C.c_float, C.c_int,
                                             C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64, _u64p]
        lib.orc_get_stf.argtypes = [C.c_void_p, _u32p, _u32p, _u64p, _u32p, _u32p]
        lib.orc_find_vo.restype = C.c_uint64
        lib.orc_find_vo.argtypes = [C.c_void_p, _f64p, C.c_int, C.c_int, C.c_float, C.c_float]
        lib.orc_get_vo.argtypes = [C.c_void_p, _u32p, _u32p, _u32p]
        lib.orc_world_transform.argtypes = [C.c_void_p, _f32p, _f32p]
        lib.orc_verify_input.restype = C.c_uint64
        lib.orc_verify_input.argtypes = [C.c_uint32, _u32p, _f32p, C.c_uint32, _f32p, C.c_float, _u32p]
        lib.orc_em_inliers.restype = C.c_uint64
        lib.orc_em_inliers.argtypes = [C.c_uint32, _u32p, _f32p, _f32p, C.c_double, _u32p, _u32p, C.c_uint64]
        lib.orc_em_assign.argtypes = [C.c_uint32, _u32p, _f32p, _f32p, C.c_double, C.c_uint32, _u32p,
                                      _u32p, _u64p, _u32p, _u32p, _u64p, _u32p]
        lib.orc_em_run.argtypes = [C.c_uint32, _u32p, _f32p, _f32p, _i32p, _i32p, _i32p]
        lib.orc_seg_fit.argtypes = [_f64p, _f64p, _f64p, C.c_int, _f32p]
        lib.orc_distance_to_line_segment.restype = C.c_float
        lib.orc_distance_to_line_segment.argtypes = [_f32p, C.c_float, C.c_float]
        lib.orc_dist_to_line_seg.restype = C.c_double
        lib.orc_dist_to_line_seg.argtypes = [_f32p, C.c_float, C.c_float]
        lib.orc_eval_stf.argtypes = [C.c_void_p, _f64p, C.c_uint64, _u32p, _u32p, _u64p, _u32p, _u32p, C.c_float,
                                     C.c_float, _f64p, C.c_void_p, C.c_int]
        lib.orc_odometry_consts.argtypes = [_f32p, C.c_uint32, _f32p]
        lib.orc_eval_odometry.argtypes = [_f32p, _f64p, C.c_uint32, _f64p, C.c_void_p]
        lib.orc_human_blocks.argtypes = [_f32p, C.c_uint32, _i32p, _f32p, _i32p, _f64p]
        lib.orc_eval_human.argtypes = [C.c_uint32, _i32p, _f64p, _f64p, _f64p, C.c_void_p]
        lib.orc_eval_p2l_glob.argtypes = [C.c_uint32, _u32p, _u64p, _f32p, _f32p, _f32p, _u8p, C.c_float, C.c_float,
                                          _f64p, _f64p, C.c_void_p]
        lib.orc_eval_p2l.argtypes = [C.c_uint64, _u32p, _f32p, _f32p, _f32p, _u8p, C.c_float, C.c_float, _f64p,
                                     _f64p, C.c_void_p]
        lib.orc_load_pose_graph.restype = C.c_void_p
        lib.orc_load_pose_graph.argtypes = [C.c_char_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        lib.orc_pose_graph_get.argtypes = [C.c_void_p, _f32p, _f32p, _u32p, _f32p, _f32p]
        lib.orc_pose_graph_free.argtypes = [C.c_void_p]
        lib.orc_sinf.restype = C.c_float
        lib.orc_sinf.argtypes = [C.c_float]
        lib.orc_cosf.restype = C.c_float
        lib.orc_cosf.argtypes = [C.c_float]
        lib.orc_num_threads.restype = C.c_int

    def num_threads(self):
        return self.lib.orc_num_threads()

    def scans(self, offsets, pts, nrm, build_trees=True):
        return OracleScans(self, offsets, pts, nrm, build_trees)

    def load_pose_graph(self, path):
        n, m = C.c_uint64(), C.c_uint64()
        h = self.lib.orc_load_pose_graph(path.encode(), C.byref(n), C.byref(m))
        if not h:
            raise IOError(path)
        poses = np.zeros(3 * n.value, np.float32)
        cov = np.zeros(9 * n.value, np.float32)
        off = np.zeros(n.value + 1, np.uint32)
        pts = np.zeros(2 * m.value, np.float32)
        nrm = np.zeros(2 * m.value, np.float32)
        self.lib.orc_pose_graph_get(h, poses, cov, off, pts, nrm)
        self.lib.orc_pose_graph_free(h)
        return dict(poses=poses.reshape(-1, 3), cov=cov.reshape(-1, 9), offsets=off, pts=pts.reshape(-1, 2),
                    nrm=nrm.reshape(-1, 2))

    def verify_input(self, offsets, world, sel, thr=0.05):
        """HitLSLAM::verifyUserInput: (points_verified, seen bit mask)."""
        offsets = np.ascontiguousarray(offsets, np.uint32)
        sel = np.ascontiguousarray(sel, np.float32).reshape(-1)
        mask = np.zeros(1, np.uint32)
        v = self.lib.orc_verify_input(len(offsets) - 1, offsets, np.ascontiguousarray(world, np.float32).reshape(-1), len(sel) // 2, sel, thr, mask)
        return int(v), int(mask[0])

    def em_inliers(self, offsets, world, seg, thr=0.03):
        offsets = np.ascontiguousarray(offsets, np.uint32)
        world = np.ascontiguousarray(world, np.float32).reshape(-1)
        cap = len(world) // 2
        op, oi = np.zeros(cap, np.uint32), np.zeros(cap, np.uint32)
        n = self.lib.orc_em_inliers(len(offsets) - 1, offsets, world, np.ascontiguousarray(seg, np.float32), thr, op, oi, cap)
        return op[:n].copy(), oi[:n].copy()

    def em_assign(self, offsets, world, segs, thr=0.03, min_obs=5):
        offsets = np.ascontiguousarray(offsets, np.uint32)
        world = np.ascontiguousarray(world, np.float32).reshape(-1)
        n, m = len(offsets) - 1, len(world) // 2
        ns = np.zeros(2, np.uint32)
        out = []
        bufs = [(np.zeros(n, np.uint32), np.zeros(n + 1, np.uint64), np.zeros(m, np.uint32)) for _ in range(2)]
        self.lib.orc_em_assign(n, offsets, world, np.ascontiguousarray(segs, np.float32).reshape(-1), thr, min_obs, ns,
                               bufs[0][0], bufs[0][1], bufs[0][2], bufs[1][0], bufs[1][1], bufs[1][2])
        for f in range(2):
            k = int(ns[f])
            off = bufs[f][1][:k + 1].copy()
            out.append((bufs[f][0][:k].copy(), off, bufs[f][2][:int(off[-1])].copy()))
        return out

    def em_run(self, offsets, world, segs):
        offsets = np.ascontiguousarray(offsets, np.uint32)
        world = np.ascontiguousarray(world, np.float32).reshape(-1)
        n = len(offsets) - 1
        s = np.ascontiguousarray(segs, np.float32).reshape(-1).copy()
        ret = np.zeros(6, np.int32)
        cor, anc = np.zeros(n, np.int32), np.zeros(n, np.int32)
        self.lib.orc_em_run(n, offsets, world, s, ret, cor, anc)
        return dict(segs=s.reshape(4, 2), corrected=cor[:ret[0]].copy(), anchor=anc[:ret[1]].copy(),
                    backprop=(int(ret[2]), int(ret[3])), swapped=bool(ret[4]), rounds=int(ret[5]))

    def seg_fit(self, p1, p2, data):
        out = np.zeros(4, np.float32)
        data = np.ascontiguousarray(data, np.float64).reshape(-1)
        self.lib.orc_seg_fit(np.ascontiguousarray(p1, np.float64), np.ascontiguousarray(p2, np.float64), data, len(data) // 2, out)
        return out.reshape(2, 2)

    def app_exp_corrections(self, ctype, sel, poses_f32, corrected):
        """AppExpCorrect::AppExpCorrections: returns (new poses [N,3] f32, correction C [3] f32 or None when nothing was applied)."""
        p = np.ascontiguousarray(poses_f32, np.float32).reshape(-1).copy()
        corr = np.ascontiguousarray(corrected, np.int32)
        c3 = np.zeros(3, np.float32)
        self.lib.orc_app_exp_corrections.argtypes = [C.c_int, _f32p, _f32p, C.c_uint32, _i32p, C.c_uint32, _f32p]
        ok = self.lib.orc_app_exp_corrections(int(ctype), np.ascontiguousarray(sel, np.float32).reshape(-1), p, len(p) // 3, corr, len(corr), c3)
        return p.reshape(-1, 3), (c3 if ok else None)

    def backprop(self, poses_f32, cov9, lo, hi, c3):
        """Backprop::Run: returns (new poses [N,3] f32, new covariances [N,9] f32)."""
        p = np.ascontiguousarray(poses_f32, np.float32).reshape(-1).copy()
        cov = np.ascontiguousarray(cov9, np.float32).reshape(-1).copy()
        self.lib.orc_backprop.argtypes = [_f32p, _f32p, C.c_uint32, C.c_int32, C.c_int32, _f32p]
        self.lib.orc_backprop.restype = None
        self.lib.orc_backprop(p, cov, len(p) // 3, int(lo), int(hi), np.ascontiguousarray(c3, np.float32))
        return p.reshape(-1, 3), cov.reshape(-1, 9)

    def odometry_consts(self, poses_f32):
        p = np.ascontiguousarray(poses_f32, np.float32).reshape(-1)
        n = len(p) // 3
        c = np.zeros(9 * (n - 1), np.float32)
        self.lib.orc_odometry_consts(p, n, c)
        return c.reshape(-1, 9)

    def eval_odometry(self, consts, poses_f64, want_jac=True):
        p = np.ascontiguousarray(poses_f64, np.float64).reshape(-1)
        n = len(p) // 3
        r = np.zeros(3 * (n - 1))
        J = np.zeros(18 * (n - 1)) if want_jac else None
        self.lib.orc_eval_odometry(np.ascontiguousarray(consts, np.float32).reshape(-1), p, n, r, J.ctypes.data if want_jac else None)
        return r.reshape(-1, 3), (J.reshape(-1, 2, 3, 3) if want_jac else None)

    def human_blocks(self, poses_f32, hc_i, hc_f):
        hc_i = np.ascontiguousarray(hc_i, np.int32).reshape(-1, 3)
        hc_f = np.ascontiguousarray(hc_f, np.float32).reshape(-1, 4)
        n = len(hc_i)
        bi, bd = np.zeros(2 * n, np.int32), np.zeros(4 * n)
        self.lib.orc_human_blocks(np.ascontiguousarray(poses_f32, np.float32).reshape(-1), n, hc_i.reshape(-1), hc_f.reshape(-1), bi, bd)
        return bi.reshape(-1, 2), bd.reshape(-1, 4)

    def eval_human(self, blk_i, blk_d, poses_f64, want_jac=True):
        n = len(blk_i)
        r = np.zeros(3 * n)
        J = np.zeros(9 * n) if want_jac else None
        self.lib.orc_eval_human(n, np.ascontiguousarray(blk_i, np.int32).reshape(-1), np.ascontiguousarray(blk_d, np.float64).reshape(-1),
                                np.ascontiguousarray(poses_f64, np.float64).reshape(-1), r, J.ctypes.data if want_jac else None)
        return r.reshape(-1, 3), (J.reshape(-1, 3, 3) if want_jac else None)

    def eval_p2l_glob(self, blk_pose, blk_off, pts, line_n, line_off, valid, std_dev, corr, poses_f64, want_jac=True):
        nb = len(blk_pose)
        r = np.zeros(nb)
        J = np.zeros(3 * nb) if want_jac else None
        self.lib.orc_eval_p2l_glob(nb, np.ascontiguousarray(blk_pose, np.uint32), np.ascontiguousarray(blk_off, np.uint64),
                                   np.ascontiguousarray(pts, np.float32).reshape(-1), np.ascontiguousarray(line_n, np.float32).reshape(-1),
                                   np.ascontiguousarray(line_off, np.float32), np.ascontiguousarray(valid, np.uint8), std_dev, corr,
                                   np.ascontiguousarray(poses_f64, np.float64).reshape(-1), r, J.ctypes.data if want_jac else None)
        return r, (J.reshape(-1, 3) if want_jac else None)

    def eval_p2l(self, pose_idx, pts, line_n, line_off, valid, std_dev, corr, poses_f64, want_jac=True):
        n = len(pose_idx)
        r = np.zeros(n)
        J = np.zeros(3 * n) if want_jac else None
        self.lib.orc_eval_p2l(n, np.ascontiguousarray(pose_idx, np.uint32), np.ascontiguousarray(pts, np.float32).reshape(-1),
                              np.ascontiguousarray(line_n, np.float32).reshape(-1), np.ascontiguousarray(line_off, np.float32),
                              np.ascontiguousarray(valid, np.uint8), std_dev, corr,
                              np.ascontiguousarray(poses_f64, np.float64).reshape(-1), r, J.ctypes.data if want_jac else None)
        return r, (J.reshape(-1, 3) if want_jac else None)


class OracleScans:
    def __init__(self, orc, offsets, pts, nrm, build_trees=True):
        self.orc, self.lib = orc, orc.lib
        self.offsets = np.ascontiguousarray(offsets, np.uint32)
        self.pts = np.ascontiguousarray(pts, np.float32).reshape(-1)
        self.nrm = np.ascontiguousarray(nrm, np.float32).reshape(-1)
        self.n = len(self.offsets) - 1
        self.h = self.lib.orc_create(self.n, self.offsets, self.pts, self.nrm, int(build_trees))

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.orc_destroy(self.h)
            self.h = None

    def flatten(self):
        m = int(self.offsets[-1])
        pn, idx, dim = np.zeros(4 * m, np.float32), np.zeros(m, np.int32), np.zeros(m, np.int32)
        self.lib.orc_flatten(self.h, pn, idx, dim)
        return pn.reshape(-1, 4), idx, dim

    def query(self, scan, q, thr, mode=0):
        q = np.ascontiguousarray(q, np.float32).reshape(-1)
        n = len(q) // 2
        d, i = np.zeros(n, np.float32), np.zeros(n, np.int32)
        self.lib.orc_query(self.h, scan, n, q, thr, mode, d, i)
        return d, i

    def radius(self, scan, qx, qy, thr):
        cap = int(self.offsets[scan + 1] - self.offsets[scan])
        idx = np.zeros(max(cap, 1), np.int32)
        n = self.lib.orc_radius(self.h, scan, qx, qy, thr, idx, cap)
        return idx[:n].copy()

    def find_stf(self, poses, min_pose=0, max_pose=None, thr=0.15, min_cos=None, cap=6, skip=1, min_corr=10,
                 src_lo=0, src_hi=None, src_stride=1):
        poses = np.ascontiguousarray(poses, np.float64).reshape(-1)
        if max_pose is None:
            max_pose = self.n - 1
        if min_cos is None:
            min_cos = default_min_cos()
        if src_hi is None:
            src_hi = 2 ** 62
        counts = np.zeros(3, np.uint64)
        if src_stride != 1:
            self.lib.orc_find_stf_strided(self.h, poses, min_pose, max_pose, thr, min_cos, cap, skip, min_corr, src_lo, src_hi, src_stride, counts)
        else:
            self.lib.orc_find_stf(self.h, poses, min_pose, max_pose, thr, min_cos, cap, skip, min_corr, src_lo, src_hi, counts)
        npairs, nm = int(counts[0]), int(counts[1])
        pi, pj = np.zeros(npairs, np.uint32), np.zeros(npairs, np.uint32)
        off = np.zeros(npairs + 1, np.uint64)
        k, idx = np.zeros(nm, np.uint32), np.zeros(nm, np.uint32)
        self.lib.orc_get_stf(self.h, pi, pj, off, k, idx)
        return dict(pair_i=pi, pair_j=pj, pair_off=off, k=k, idx=idx, n_queries=int(counts[2]))

    def check_chunks(self, result, poses, chunks, **opts):
        """Checker for searches too large to repeat on the CPU as a whole: for every [lo, hi) source-pose chunk, run the oracle's
        FindSTFCorrespondences over those sources against ALL targets and compare bit for bit with the rows of `result` (a CSR in
        the reference's (i, j, k) order, e.g. hitl_get_stf of a full-map or shard search) whose source pose lies in the chunk.
        Returns the list of (lo, hi, ok, n_pairs, n_matches)."""
        pi = np.asarray(result["pair_i"]); pj = np.asarray(result["pair_j"]); off = np.asarray(result["pair_off"]).astype(np.int64)
        rk = np.asarray(result["k"]); ridx = np.asarray(result["idx"])
        out = []
        for lo, hi in chunks:
            want = self.find_stf(poses, src_lo=lo, src_hi=hi, **opts)
            a, b = np.searchsorted(pi, lo, side="left"), np.searchsorted(pi, hi, side="left")
            m0, m1 = int(off[a]), int(off[b])
            ok = (np.array_equal(pi[a:b], want["pair_i"]) and np.array_equal(pj[a:b], want["pair_j"])
                  and np.array_equal(off[a:b + 1] - m0, want["pair_off"].astype(np.int64))
                  and np.array_equal(rk[m0:m1], want["k"]) and np.array_equal(ridx[m0:m1], want["idx"]))
            out.append((int(lo), int(hi), bool(ok), int(b - a), int(m1 - m0)))
        return out

    def find_vo(self, poses, min_pose=0, max_pose=None, thr=0.15, min_cos=None):
        poses = np.ascontiguousarray(poses, np.float64).reshape(-1)
        if max_pose is None:
            max_pose = self.n - 1
        if min_cos is None:
            min_cos = default_min_cos()
        n = self.lib.orc_find_vo(self.h, poses, min_pose, max_pose, thr, min_cos)
        sp, sk, tk = np.zeros(n, np.uint32), np.zeros(n, np.uint32), np.zeros(n, np.uint32)
        self.lib.orc_get_vo(self.h, sp, sk, tk)
        return sp, sk, tk

    def world_transform(self, poses_f32):
        out = np.zeros(len(self.pts), np.float32)
        self.lib.orc_world_transform(self.h, np.ascontiguousarray(poses_f32, np.float32).reshape(-1), out)
        return out.reshape(-1, 2)

    def eval_stf(self, poses, corr, std_dev=0.05, corr_factor=1.0 / 40.0, want_jac=True, parallel=False):
        poses = np.ascontiguousarray(poses, np.float64).reshape(-1)
        nb = len(corr["pair_i"])
        r = np.zeros(2 * nb)
        J = np.zeros(12 * nb) if want_jac else None
        self.lib.orc_eval_stf(self.h, poses, nb, corr["pair_i"], corr["pair_j"], corr["pair_off"], corr["k"], corr["idx"],
                              std_dev, corr_factor, r, J.ctypes.data if want_jac else None, int(parallel))
        return r.reshape(-1, 2), (J.reshape(-1, 2, 2, 3) if want_jac else None)


def default_min_cos():
    """cos(max_stf_angle_error = deg2rad(25)) as a float (config/non_markov_localization.cfg:48,
    JointOptimization.cpp:564); the angle option is a float, the cosine is stored to a float."""
    ang = np.float32(np.deg2rad(25.0))
    return float(np.float32(np.cos(np.float64(ang))))


class RefKDTree:
    """The reference's own KDTree<float,2> (oracle/_ref/libkdtree_ref.so), when it was built."""

    @staticmethod
    def available():
        return os.path.exists(os.path.join(HERE, "_ref", "libkdtree_ref.so"))

    def __init__(self, pts, nrm):
        self.lib = lib = C.CDLL(os.path.join(HERE, "_ref", "libkdtree_ref.so"))
        lib.ref_kd_create.restype = C.c_void_p
        lib.ref_kd_create.argtypes = [C.c_uint32, _f32p, _f32p]
        lib.ref_kd_destroy.argtypes = [C.c_void_p]
        lib.ref_kd_flatten.argtypes = [C.c_void_p, _f32p, _i32p, _i32p]
        lib.ref_kd_query.argtypes = [C.c_void_p, C.c_uint32, _f32p, C.c_float, C.c_int, _f32p, _i32p]
        lib.ref_kd_radius.restype = C.c_uint32
        lib.ref_kd_radius.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, _i32p, C.c_uint32]
        self.pts = np.ascontiguousarray(pts, np.float32).reshape(-1)
        self.nrm = np.ascontiguousarray(nrm, np.float32).reshape(-1)
        self.n = len(self.pts) // 2
        self.h = lib.ref_kd_create(self.n, self.pts, self.nrm)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.ref_kd_destroy(self.h)
            self.h = None

    def flatten(self):
        pn, idx, dim = np.zeros(4 * self.n, np.float32), np.zeros(self.n, np.int32), np.zeros(self.n, np.int32)
        self.lib.ref_kd_flatten(self.h, pn, idx, dim)
        return pn.reshape(-1, 4), idx, dim

    def query(self, q, thr, mode=0):
        q = np.ascontiguousarray(q, np.float32).reshape(-1)
        n = len(q) // 2
        d, i = np.zeros(n, np.float32), np.zeros(n, np.int32)
        self.lib.ref_kd_query(self.h, n, q, thr, mode, d, i)
        return d, i

    def radius(self, qx, qy, thr):
        idx = np.zeros(max(self.n, 1), np.int32)
        n = self.lib.ref_kd_radius(self.h, qx, qy, thr, idx, self.n)
        return idx[:n].copy()


class RefFunctors:
    """The reference's own residual functors and DistanceToLineSegment (oracle/_ref/libfunctors_ref.so:
    residual_functors.h + eigen_helper.h compiled where they lie against oracle/shim2), when built."""

    @staticmethod
    def available():
        return os.path.exists(os.path.join(HERE, "_ref", "libfunctors_ref.so"))

    def __init__(self):
        self.lib = lib = C.CDLL(os.path.join(HERE, "_ref", "libfunctors_ref.so"))
        lib.ref_p2p_glob.argtypes = [C.c_uint32, _f32p, _f32p, _f32p, _f32p, C.c_float, C.c_float, _f64p, _f64p, _f64p, _f64p, _f64p, _f64p]
        lib.ref_pose_constraint.argtypes = [_f32p, _f64p, _f64p, _f64p, _f64p, _f64p]
        lib.ref_human.restype = C.c_int
        lib.ref_human.argtypes = [C.c_int, _f64p, _f64p, _f64p, _f64p]
        lib.ref_p2l_glob.argtypes = [C.c_uint32, _f32p, _f32p, _f32p, _u8p, C.c_float, C.c_float, _f64p, _f64p, _f64p]
        lib.ref_p2l.argtypes = [_f32p, _f32p, C.c_float, C.c_int, C.c_float, C.c_float, _f64p, _f64p, _f64p]
        lib.ref_distance_to_line_segment.argtypes = [C.c_uint32, _f32p, _f32p, _f32p, _f32p]

    @staticmethod
    def _f32(a):
        return np.ascontiguousarray(a, np.float32).reshape(-1)

    def p2p_glob(self, p0, p1, n0, n1, std_dev, corr, x0, x1):
        r, j0, j1, rp = np.zeros(2), np.zeros(6), np.zeros(6), np.zeros(2)
        self.lib.ref_p2p_glob(len(self._f32(p0)) // 2, self._f32(p0), self._f32(p1), self._f32(n0), self._f32(n1), std_dev, corr,
                              np.ascontiguousarray(x0, np.float64), np.ascontiguousarray(x1, np.float64), r, j0, j1, rp)
        return r, j0.reshape(2, 3), j1.reshape(2, 3), rp

    def pose_constraint(self, consts9, x0, x1):
        r, j0, j1 = np.zeros(3), np.zeros(9), np.zeros(9)
        self.lib.ref_pose_constraint(self._f32(consts9), np.ascontiguousarray(x0, np.float64), np.ascontiguousarray(x1, np.float64), r, j0, j1)
        return r, j0.reshape(3, 3), j1.reshape(3, 3)

    def human(self, ctype, targets4, x):
        r, j = np.zeros(3), np.zeros(9)
        k = self.lib.ref_human(int(ctype), np.ascontiguousarray(targets4, np.float64), np.ascontiguousarray(x, np.float64), r, j)
        return r[:k], j[:3 * k].reshape(k, 3)

    def p2l_glob(self, pts, ln, lo, valid, std_dev, corr, x):
        r, j = np.zeros(1), np.zeros(3)
        self.lib.ref_p2l_glob(len(self._f32(pts)) // 2, self._f32(pts), self._f32(ln), self._f32(lo), np.ascontiguousarray(valid, np.uint8), std_dev, corr,
                              np.ascontiguousarray(x, np.float64), r, j)
        return r[0], j

    def p2l(self, pt, ln, lo, valid, std_dev, corr, x):
        r, j = np.zeros(1), np.zeros(3)
        self.lib.ref_p2l(self._f32(pt), self._f32(ln), float(lo), int(valid), std_dev, corr, np.ascontiguousarray(x, np.float64), r, j)
        return r[0], j

    def distance_to_line_segment(self, p0, p1, pts):
        pts = self._f32(pts)
        out = np.zeros(len(pts) // 2, np.float32)
        self.lib.ref_distance_to_line_segment(len(out), self._f32(p0), self._f32(p1), pts, out)
        return out


class RefBackend:
    """The reference's own back-end translation units (oracle/_ref/libhitl_ref.so: JointOptimization.cpp, EMinput.cpp,
    ApplyExplicitCorrection.cpp, Backprop.cpp, HitLSLAM.cpp, kdtree.cpp compiled where they lie against oracle/shim3,
    behind oracle/ref_hitl_capi.cpp), when built.  Pins the restated loops to the reference's own code."""

    @staticmethod
    def available(fast=False):
        return os.path.exists(os.path.join(HERE, "_ref", "libhitl_ref_fast.so" if fast else "libhitl_ref.so"))

    def __init__(self, fast=False):
        """fast=True loads the timing build (the reference's Release flags) instead of the parity build (-O2 -ffp-contract=off)."""
        self.lib = lib = C.CDLL(os.path.join(HERE, "_ref", "libhitl_ref_fast.so" if fast else "libhitl_ref.so"))
        vp = C.c_void_p
        lib.ref_jo_create.restype = vp
        lib.ref_jo_create.argtypes = [C.c_uint32, _u32p, _f32p, _f32p, _f32p]
        lib.ref_jo_destroy.argtypes = [vp]
        lib.ref_jo_set_options.argtypes = [vp, C.c_float, C.c_float, C.c_int, C.c_uint32, C.c_float, C.c_float]
        lib.ref_min_cos.restype = C.c_float
        lib.ref_min_cos.argtypes = [C.c_float]
        lib.ref_jo_set_pose_array.argtypes = [vp, _f64p]
        lib.ref_jo_get_pose_array.argtypes = [vp, _f64p]
        lib.ref_jo_set_poses.argtypes = [vp, _f32p]
        lib.ref_jo_world_clouds.argtypes = [vp, _f32p]
        lib.ref_jo_relative_pose.argtypes = [vp, C.c_uint32, _u32p, _u32p, _f32p]
        lib.ref_jo_find_stf.argtypes = [vp, C.c_uint64, C.c_uint64, _u64p]
        lib.ref_jo_restrict_sources.argtypes = [vp, C.c_uint32, _u32p]
        lib.ref_jo_get_stf.argtypes = [vp, _u32p, _u32p, _u64p, _u32p, _u32p, C.c_void_p]
        lib.ref_jo_find_vo.restype = C.c_uint64
        lib.ref_jo_find_vo.argtypes = [vp, C.c_int, C.c_int]
        lib.ref_jo_get_vo.argtypes = [vp, _u32p, _u32p, _u32p]
        lib.ref_jo_kd_query.argtypes = [vp, C.c_uint32, C.c_uint32, _f32p, C.c_float, C.c_int, _f32p, _i32p]
        lib.ref_jo_set_human_constraints.argtypes = [vp, C.c_uint32, _u32p, _i32p, _f32p]
        lib.ref_jo_eval_blocks.restype = C.c_int64
        lib.ref_jo_eval_blocks.argtypes = [vp, C.c_int, _f64p, C.c_uint64, C.c_int, C.c_int, _f64p, _f64p, _i32p]
        lib.ref_jo_run.argtypes = [vp, _f32p, _f64p]
        lib.ref_jo_post_human_optimization.restype = C.c_int
        lib.ref_jo_post_human_optimization.argtypes = [vp, _f64p, _u64p]
        lib.ref_jo_get_gradient.argtypes = [vp, _f64p]
        lib.ref_jo_evaluate_stf_problem.restype = C.c_int64
        lib.ref_jo_evaluate_stf_problem.argtypes = [vp, _f64p, _f64p, _f64p, C.c_uint64, _f64p]
        lib.ref_em_run.argtypes = [C.c_uint32, _u32p, _f32p, _f32p, C.c_int, _i32p, _i32p, _i32p]
        lib.ref_em_observation_sets.argtypes = [C.c_uint32, _u32p, _f32p, _f32p, _u32p, _u32p, _u64p, _u32p, _u32p, _u64p, _u32p]
        lib.ref_em_dist_to_line_seg.restype = C.c_double
        lib.ref_em_dist_to_line_seg.argtypes = [_f32p, _f32p, _f32p]
        lib.ref_em_seg_fit.argtypes = [_f64p, _f64p, _f64p, C.c_int, _f32p]
        lib.ref_app_exp_run.restype = C.c_uint32
        lib.ref_app_exp_run.argtypes = [C.c_int, _f32p, _f32p, C.c_uint32, _i32p, C.c_uint32, _i32p, C.c_uint32, _f32p, _i32p, _f32p]
        lib.ref_backprop_run.argtypes = [_f32p, _f32p, C.c_uint32, C.c_int, C.c_int, _f32p]
        lib.ref_session_create.restype = vp
        lib.ref_session_create.argtypes = [C.c_uint32, _u32p, _f32p, _f32p, _f32p, _f32p]
        lib.ref_session_destroy.argtypes = [vp]
        lib.ref_session_replay.restype = C.c_uint32
        lib.ref_session_replay.argtypes = [vp, C.c_int, _f32p]
        lib.ref_session_get.argtypes = [vp, _f32p, _f32p, _f32p]
        lib.ref_session_constraints.restype = C.c_uint32
        lib.ref_session_constraints.argtypes = [vp, C.c_uint32, C.c_void_p, C.c_void_p]
        lib.ref_session_verify.restype = C.c_size_t
        lib.ref_session_verify.argtypes = [vp, C.c_int, _f32p]

    @staticmethod
    def _f32(a):
        return np.ascontiguousarray(a, np.float32).reshape(-1)

    def min_cos(self, max_angle=None):
        if max_angle is None:
            max_angle = np.float32(np.deg2rad(25.0))
        return float(self.lib.ref_min_cos(float(max_angle)))

    def joint_opt(self, offsets, pts, nrm, poses_f32):
        return RefJointOpt(self, offsets, pts, nrm, poses_f32)

    def session(self, offsets, pts, nrm, poses_f32, cov9=None):
        return RefSession(self, offsets, pts, nrm, poses_f32, cov9)

    def em_run(self, offsets, world, segs, ctype=4):
        offsets = np.ascontiguousarray(offsets, np.uint32)
        n = len(offsets) - 1
        s = self._f32(segs).copy()
        ret = np.zeros(4, np.int32)
        cor, anc = np.zeros(n, np.int32), np.zeros(n, np.int32)
        self.lib.ref_em_run(n, offsets, self._f32(world), s, int(ctype), ret, cor, anc)
        return dict(segs=s.reshape(4, 2), corrected=cor[:ret[0]].copy(), anchor=anc[:ret[1]].copy(), backprop=(int(ret[2]), int(ret[3])))

    def em_observation_sets(self, offsets, world, segs):
        offsets = np.ascontiguousarray(offsets, np.uint32)
        n, m = len(offsets) - 1, int(offsets[-1])
        ns = np.zeros(2, np.uint32)
        bufs = [(np.zeros(n, np.uint32), np.zeros(n + 1, np.uint64), np.zeros(max(m, 1), np.uint32)) for _ in range(2)]
        self.lib.ref_em_observation_sets(n, offsets, self._f32(world), self._f32(segs), ns, bufs[0][0], bufs[0][1], bufs[0][2], bufs[1][0], bufs[1][1], bufs[1][2])
        out = []
        for f in range(2):
            k = int(ns[f])
            off = bufs[f][1][:k + 1].copy()
            out.append((bufs[f][0][:k].copy(), off, bufs[f][2][:int(off[-1])].copy()))
        return out

    def dist_to_line_seg(self, p1, p2, p):
        return float(self.lib.ref_em_dist_to_line_seg(self._f32(p1), self._f32(p2), self._f32(p)))

    def seg_fit(self, p1, p2, data):
        out = np.zeros(4, np.float32)
        data = np.ascontiguousarray(data, np.float64).reshape(-1)
        self.lib.ref_em_seg_fit(np.ascontiguousarray(p1, np.float64), np.ascontiguousarray(p2, np.float64), data, len(data) // 2, out)
        return out.reshape(2, 2)

    def app_exp_run(self, ctype, sel, poses_f32, corrected, anchor):
        """AppExpCorrect::Run: (poses [N,3] f32, C [3] f32, hc_i [B,3] i32, hc_f [B,4] f32)."""
        p = self._f32(poses_f32).copy()
        cor, anc = np.ascontiguousarray(corrected, np.int32), np.ascontiguousarray(anchor, np.int32)
        nb = max(len(cor) * len(anc), 1)
        c3, hi, hf = np.zeros(3, np.float32), np.zeros(3 * nb, np.int32), np.zeros(4 * nb, np.float32)
        n = self.lib.ref_app_exp_run(int(ctype), self._f32(sel), p, len(p) // 3, cor, len(cor), anc, len(anc), c3, hi, hf)
        return p.reshape(-1, 3), c3, hi[:3 * n].reshape(-1, 3).copy(), hf[:4 * n].reshape(-1, 4).copy()

    def backprop(self, poses_f32, cov9, lo, hi, c3):
        p, cov = self._f32(poses_f32).copy(), self._f32(cov9).copy()
        self.lib.ref_backprop_run(p, cov, len(p) // 3, int(lo), int(hi), self._f32(c3))
        return p.reshape(-1, 3), cov.reshape(-1, 9)


class RefJointOpt:
    """The reference's JointOpt over one map (trees built by its own BuildKDTrees)."""

    def __init__(self, ref, offsets, pts, nrm, poses_f32):
        self.ref, self.lib = ref, ref.lib
        self.offsets = np.ascontiguousarray(offsets, np.uint32)
        self.n, self.m = len(self.offsets) - 1, int(self.offsets[-1])
        self.h = self.lib.ref_jo_create(self.n, self.offsets, ref._f32(pts), ref._f32(nrm), ref._f32(poses_f32))
        self.set_options()

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.ref_jo_destroy(self.h)
            self.h = None

    def set_options(self, thr=0.15, max_angle=None, cap=6, skip=1, laser_std=0.05, corr=1.0 / 40.0):
        if max_angle is None:
            max_angle = np.float32(np.deg2rad(25.0))
        self.lib.ref_jo_set_options(self.h, thr, float(max_angle), cap, skip, laser_std, corr)

    def set_pose_array(self, poses_f64):
        self.lib.ref_jo_set_pose_array(self.h, np.ascontiguousarray(poses_f64, np.float64).reshape(-1))

    def pose_array(self):
        out = np.zeros(3 * self.n)
        self.lib.ref_jo_get_pose_array(self.h, out)
        return out.reshape(-1, 3)

    def set_poses(self, poses_f32):
        self.lib.ref_jo_set_poses(self.h, self.ref._f32(poses_f32))

    def world_clouds(self):
        out = np.zeros(2 * self.m, np.float32)
        self.lib.ref_jo_world_clouds(self.h, out)
        return out.reshape(-1, 2)

    def relative_pose(self, src, dst):
        src, dst = np.ascontiguousarray(src, np.uint32), np.ascontiguousarray(dst, np.uint32)
        out = np.zeros(6 * len(src), np.float32)
        self.lib.ref_jo_relative_pose(self.h, len(src), src, dst, out)
        return out.reshape(-1, 6)

    def restrict_sources(self, keep=None):
        """Timing samples of a full map: the reference's own FindSTFCorrespondences then searches only the kept SOURCE poses, against all
        targets (ref_jo_restrict_sources parks the other poses' point_clouds_g_ entries; None restores them)."""
        ids = np.ascontiguousarray(keep if keep is not None else [], np.uint32)
        self.lib.ref_jo_restrict_sources(self.h, len(ids), ids)

    def find_stf(self, poses_f64=None, min_pose=0, max_pose=None, with_points=False):
        if poses_f64 is not None:
            self.set_pose_array(poses_f64)
        if max_pose is None:
            max_pose = self.n - 1
        counts = np.zeros(2, np.uint64)
        self.lib.ref_jo_find_stf(self.h, int(min_pose), int(max_pose), counts)
        npairs, nm = int(counts[0]), int(counts[1])
        pi, pj = np.zeros(npairs, np.uint32), np.zeros(npairs, np.uint32)
        off = np.zeros(npairs + 1, np.uint64)
        k, idx = np.zeros(nm, np.uint32), np.zeros(nm, np.uint32)
        xy = np.zeros(8 * nm, np.float32) if with_points else None
        self.lib.ref_jo_get_stf(self.h, pi, pj, off, k, idx, xy.ctypes.data if with_points else None)
        out = dict(pair_i=pi, pair_j=pj, pair_off=off, k=k, idx=idx)
        if with_points:
            out["xy8"] = xy.reshape(-1, 4, 2)
        return out

    def find_vo(self, poses_f64=None, min_pose=0, max_pose=None):
        if poses_f64 is not None:
            self.set_pose_array(poses_f64)
        if max_pose is None:
            max_pose = self.n - 1
        n = int(self.lib.ref_jo_find_vo(self.h, int(min_pose), int(max_pose)))
        sp, sk, tk = np.zeros(n, np.uint32), np.zeros(n, np.uint32), np.zeros(n, np.uint32)
        self.lib.ref_jo_get_vo(self.h, sp, sk, tk)
        return sp, sk, tk

    def kd_query(self, scan, q, thr, mode=0):
        q = self.ref._f32(q)
        n = len(q) // 2
        d, i = np.zeros(n, np.float32), np.zeros(n, np.int32)
        self.lib.ref_jo_kd_query(self.h, int(scan), n, q, thr, mode, d, i)
        return d, i

    def set_human_constraints(self, groups):
        """groups: list of (hc_i [B,3] int32, hc_f [B,4] float32)."""
        off = np.concatenate([[0], np.cumsum([len(g[0]) for g in groups])]).astype(np.uint32)
        hi = np.concatenate([np.asarray(g[0], np.int32).reshape(-1, 3) for g in groups] or [np.zeros((0, 3), np.int32)])
        hf = np.concatenate([np.asarray(g[1], np.float32).reshape(-1, 4) for g in groups] or [np.zeros((0, 4), np.float32)])
        self.lib.ref_jo_set_human_constraints(self.h, len(groups), off, np.ascontiguousarray(hi).reshape(-1), np.ascontiguousarray(hf).reshape(-1))

    def eval_blocks(self, which, poses_f64, cap):
        """which: 0 odometry (r 3, J [2,3,3]), 1 human (r <= 3, J [nr,3]), 2 STF of the last find_stf (r 2, J [2,2,3])."""
        rs, js = {0: (3, 18), 1: (3, 9), 2: (2, 12)}[which]
        r, J, nr = np.zeros(rs * cap), np.zeros(js * cap), np.zeros(cap, np.int32)
        n = int(self.lib.ref_jo_eval_blocks(self.h, which, np.ascontiguousarray(poses_f64, np.float64).reshape(-1), cap, rs, js, r, J, nr))
        assert n >= 0, "eval_blocks: capacity too small"
        return r[:rs * n].reshape(n, rs), J[:js * n].reshape(n, js), nr[:n]

    def run(self):
        p, pa = np.zeros(3 * self.n, np.float32), np.zeros(3 * self.n)
        self.lib.ref_jo_run(self.h, p, pa)
        return p.reshape(-1, 3), pa.reshape(-1, 3)

    def evaluate_stf_problem(self, poses_f64, cap):
        """FindSTFCorrespondences + AddSTFConstraints + Problem::Evaluate on the reference's own code: (cost, residuals [B, 2], gradient [N, 3])."""
        x = np.ascontiguousarray(poses_f64, np.float64).reshape(-1)
        cost, res, grad = np.zeros(1), np.zeros(2 * cap), np.zeros(len(x))
        nb = int(self.lib.ref_jo_evaluate_stf_problem(self.h, x, cost, res, 2 * cap, grad))
        assert nb >= 0, "evaluate_stf_problem: capacity too small"
        return float(cost[0]), res[:2 * nb].reshape(-1, 2), grad.reshape(-1, 3)

    def post_human_optimization(self, poses_f64=None):
        """JointOpt::PostHumanOptimization on the reference's own CPU code (search + STF blocks + solve + Problem::Evaluate)."""
        if poses_f64 is not None:
            self.set_pose_array(poses_f64)
        out, counts = np.zeros(3 * self.n), np.zeros(4, np.uint64)
        t = int(self.lib.ref_jo_post_human_optimization(self.h, out, counts))
        grad = np.zeros(int(counts[3]))
        self.lib.ref_jo_get_gradient(self.h, grad)
        return dict(termination=t, pose_array=out.reshape(-1, 3), n_blocks=int(counts[0]), n_matches=int(counts[1]), n_vo=int(counts[2]), gradient=grad)


class RefSession:
    """The reference's HitLSLAM (init + replayLog): the whole correction chain on its own code."""

    def __init__(self, ref, offsets, pts, nrm, poses_f32, cov9=None):
        self.ref, self.lib = ref, ref.lib
        self.offsets = np.ascontiguousarray(offsets, np.uint32)
        self.n, self.m = len(self.offsets) - 1, int(self.offsets[-1])
        cov = np.zeros(9 * self.n, np.float32) if cov9 is None else ref._f32(cov9)
        self.h = self.lib.ref_session_create(self.n, self.offsets, ref._f32(pts), ref._f32(nrm), ref._f32(poses_f32), cov)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.ref_session_destroy(self.h)
            self.h = None

    def verify(self, ctype, sel):
        return int(self.lib.ref_session_verify(self.h, int(ctype), self.ref._f32(sel)))

    def replay(self, ctype, sel):
        return int(self.lib.ref_session_replay(self.h, int(ctype), self.ref._f32(sel)))

    def state(self):
        p, cov, w = np.zeros(3 * self.n, np.float32), np.zeros(9 * self.n, np.float32), np.zeros(2 * self.m, np.float32)
        self.lib.ref_session_get(self.h, p, cov, w)
        return p.reshape(-1, 3), cov.reshape(-1, 9), w.reshape(-1, 2)

    def constraints(self, g):
        n = int(self.lib.ref_session_constraints(self.h, g, None, None))
        hi, hf = np.zeros(3 * max(n, 1), np.int32), np.zeros(4 * max(n, 1), np.float32)
        self.lib.ref_session_constraints(self.h, g, hi.ctypes.data, hf.ctypes.data)
        return hi[:3 * n].reshape(-1, 3), hf[:4 * n].reshape(-1, 4)


class RefDropin:
    """The reference's JointOpt with BuildKDTrees / FindSTFCorrespondences / FindVisualOdometryCorrespondences re-bound to the product's
    C ABI (oracle/_ref/libhitl_ref_dropin.so, oracle/ref_dropin_capi.cpp): the drop-in demonstration.  Needs a hitl_ctx (a B200)."""

    @staticmethod
    def available(blocks=False):
        return os.path.exists(os.path.join(HERE, "_ref", "libhitl_ref_dropin_blocks.so" if blocks else "libhitl_ref_dropin.so"))

    def __init__(self, blocks=False):
        """blocks=True: the library that also replaces AddSTFConstraints with GPU-backed cost blocks (one batched hitl_eval per point)."""
        self.lib = lib = C.CDLL(os.path.join(HERE, "_ref", "libhitl_ref_dropin_blocks.so" if blocks else "libhitl_ref_dropin.so"))
        vp = C.c_void_p
        lib.dropin_evaluate_stf_problem.restype = C.c_int64
        lib.dropin_evaluate_stf_problem.argtypes = [vp, _f64p, _f64p, _f64p, C.c_uint64, _f64p, C.c_char_p, C.c_size_t]
        lib.dropin_has_gpu_blocks.restype = C.c_int
        lib.dropin_last_batches.restype = C.c_uint64
        lib.dropin_create.restype = vp
        lib.dropin_create.argtypes = [vp, C.c_uint32, _u32p, _f32p, _f32p, _f32p, C.c_char_p, C.c_size_t]
        lib.dropin_destroy.argtypes = [vp]
        lib.dropin_set_options.argtypes = [vp, C.c_float, C.c_float, C.c_int, C.c_uint32, C.c_float, C.c_float]
        lib.dropin_set_pose_array.argtypes = [vp, _f64p]
        lib.dropin_post_human_optimization.restype = C.c_int
        lib.dropin_post_human_optimization.argtypes = [vp, _f64p, _u64p, C.c_char_p, C.c_size_t]
        lib.dropin_get_gradient.argtypes = [vp, _f64p]

    def create(self, ctx, offsets, pts, nrm, poses_f32):
        """ctx: the c_void_p of a live hitl_ctx (HitlGpu.ctx) or None.  Raises RuntimeError with the library's message on failure."""
        offsets = np.ascontiguousarray(offsets, np.uint32)
        err = C.create_string_buffer(512)
        f32 = lambda a: np.ascontiguousarray(a, np.float32).reshape(-1)   # noqa: E731
        h = self.lib.dropin_create(ctx, len(offsets) - 1, offsets, f32(pts), f32(nrm), f32(poses_f32), err, 512)
        if not h:
            raise RuntimeError(err.value.decode() or "dropin_create failed")
        self.lib.dropin_set_options(h, 0.15, float(np.float32(np.deg2rad(25.0))), 6, 1, 0.05, 1.0 / 40.0)
        return h

    def destroy(self, h):
        self.lib.dropin_destroy(h)

    def evaluate_stf_problem(self, h, poses_f64, cap):
        """Search + AddSTFConstraints + Problem::Evaluate at poses_f64: (cost, residuals [B, 2], gradient [N, 3])."""
        x = np.ascontiguousarray(poses_f64, np.float64).reshape(-1)
        cost, res, grad, err = np.zeros(1), np.zeros(2 * cap), np.zeros(len(x)), C.create_string_buffer(512)
        nb = int(self.lib.dropin_evaluate_stf_problem(h, x, cost, res, 2 * cap, grad, err, 512))
        if nb < 0:
            raise RuntimeError(err.value.decode())
        return float(cost[0]), res[:2 * nb].reshape(-1, 2), grad.reshape(-1, 3)

    def post_human_optimization(self, h, n_poses, poses_f64=None):
        if poses_f64 is not None:
            self.lib.dropin_set_pose_array(h, np.ascontiguousarray(poses_f64, np.float64).reshape(-1))
        out, counts, err = np.zeros(3 * n_poses), np.zeros(4, np.uint64), C.create_string_buffer(512)
        t = self.lib.dropin_post_human_optimization(h, out, counts, err, 512)
        if t < 0:
            raise RuntimeError(err.value.decode())
        grad = np.zeros(int(counts[3]))
        self.lib.dropin_get_gradient(h, grad)
        return dict(termination=t, pose_array=out.reshape(-1, 3), n_blocks=int(counts[0]), n_matches=int(counts[1]), n_vo=int(counts[2]), gradient=grad)
