"""ctypes bindings of the C ABI (include/hitl_gpu.h) and of the host mirror library.

This is plumbing for tests/ and bench.py: every call goes straight through the C ABI of
libhitl_gpu.so.  There is no CPU fallback — if the library is missing or no CUDA device is
present, construction raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.path.join(HERE, "lib")

_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C")
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C")

KDNODE = np.dtype([("px", "<f4"), ("py", "<f4"), ("nx", "<f4"), ("ny", "<f4"), ("index", "<i4"), ("dim", "<i4")])

# every symbol include/hitl_gpu.h declares (tests check that the library exports all of them)
ABI_SYMBOLS = [
    "hitl_create", "hitl_destroy", "hitl_last_error", "hitl_stream", "hitl_launch_count", "hitl_sm_count", "hitl_last_kernel_ms",
    "hitl_host_alloc", "hitl_host_free",
    "hitl_set_scans", "hitl_build_kdtrees", "hitl_set_kdtrees", "hitl_get_kdtrees", "hitl_set_kdtrees_compact", "hitl_get_kdtrees_compact", "hitl_kd_query", "hitl_kd_neighbors",
    "hitl_find_stf", "hitl_get_stf", "hitl_get_stf16", "hitl_get_stf_work", "hitl_find_vo", "hitl_get_vo",
    "hitl_world_transform", "hitl_set_world_clouds", "hitl_verify_input", "hitl_em_inliers", "hitl_em_refit", "hitl_em_refit_chain", "hitl_em_assign",
    "hitl_set_stf_blocks_from_search", "hitl_set_stf_blocks", "hitl_set_odometry_blocks", "hitl_set_human_blocks",
    "hitl_set_p2l_glob_blocks", "hitl_set_p2l_blocks", "hitl_eval_layout_get", "hitl_eval", "hitl_normal_eq",
    "hitl_normal_eq_device", "hitl_set_deterministic", "hitl_comm_unique_id", "hitl_comm_init", "hitl_comm_destroy", "hitl_comm_info", "hitl_normal_eq_allreduce", "hitl_gather_stf_blocks", "hitl_set_scans_sharded", "hitl_set_kdtrees_sharded", "hitl_set_kdtrees_compact_sharded",
    "hitl_backprop_poses", "hitl_kdtree_build_host", "hitl_debug_sincos", "hitl_debug_relative_pose", "hitl_debug_tile_work", "hitl_debug_tile_desc", "hitl_debug_set_tiling",
    "hitl_debug_set_fine_occupancy", "hitl_debug_set_em_cull", "hitl_debug_set_search_variant", "hitl_debug_set_tree_builder", "hitl_debug_tree_stats",
]


class StfOpts(C.Structure):
    _fields_ = [("point_match_threshold", C.c_float), ("min_cosine_angle", C.c_float),
                ("max_correspondences_per_point", C.c_int32), ("num_skip_readings", C.c_uint32),
                ("min_inter_pose_correspondence", C.c_uint32), ("disable_culling", C.c_uint32)]


class StfInfo(C.Structure):
    _fields_ = [("n_pairs", C.c_uint64), ("n_matches", C.c_uint64), ("n_raw_matches", C.c_uint64),
                ("n_queries", C.c_uint64), ("n_traversals", C.c_uint64), ("n_tile_pairs", C.c_uint64), ("ms_search", C.c_float), ("ms_total", C.c_float),
                ("n_coarse_pass", C.c_uint64), ("n_in_radius", C.c_uint64), ("sum_tile_cycles", C.c_uint64), ("max_tile_cycles", C.c_uint64), ("n_tiles", C.c_uint32), ("n_tiles_next", C.c_uint32),
                ("n_gate_fail", C.c_uint64), ("n_over_cap", C.c_uint64), ("n_dir_culled", C.c_uint64)]


class EmFitInfo(C.Structure):
    _fields_ = [("theta", C.c_double), ("initial_cost", C.c_double), ("final_cost", C.c_double), ("n_inliers", C.c_uint64),
                ("iterations", C.c_int32), ("evaluations", C.c_int32), ("termination", C.c_int32), ("ms", C.c_float)]


class EvalLayout(C.Structure):
    _fields_ = [("n_odometry", C.c_uint64), ("n_human", C.c_uint64), ("n_stf", C.c_uint64), ("n_p2l_glob", C.c_uint64),
                ("n_p2l", C.c_uint64), ("n_residuals", C.c_uint64), ("n_jacobian", C.c_uint64)]


class HitlError(RuntimeError):
    pass


def lib_path(name="libhitl_gpu.so"):
    return os.path.join(LIB_DIR, name)


def load_gpu_library():
    path = lib_path()
    if not os.path.exists(path):
        raise HitlError("libhitl_gpu.so is not built (run `python -m hitl_slam_b200.build`); there is no CPU fallback")
    return C.CDLL(path)


def preload_nccl():
    """comm.cu binds NCCL at run time by soname (libnccl.so.2).  In a Python process that also imports torch, both must share ONE copy:
    glibc resolves a soname to whichever copy is already loaded, and torch's libtorch_cuda.so needs symbols of its own bundled NCCL
    (site-packages/nvidia/nccl), which an older system libnccl lacks.  So the pip-bundled copy is loaded first when it exists; a C++ host
    without Python simply gets the system library.  HITL_NCCL_LIBRARY overrides the choice."""
    path = os.environ.get("HITL_NCCL_LIBRARY")
    if not path:
        try:
            import importlib.util
            spec = importlib.util.find_spec("nvidia.nccl")
            for loc in (spec.submodule_search_locations if spec else []):
                cand = os.path.join(loc, "lib", "libnccl.so.2")
                if os.path.exists(cand):
                    path = cand
                    break
        except Exception:
            path = None
    if path:
        try:
            C.CDLL(path, mode=C.RTLD_GLOBAL)
        except OSError:
            pass


def default_min_cos():
    """cos(deg2rad(25)) stored to a float (config/non_markov_localization.cfg:48; JointOptimization.cpp:564)."""
    ang = np.float32(np.deg2rad(25.0))
    return float(np.float32(np.cos(np.float64(ang))))


class HitlGpu:
    """One context = one GPU.  Thin, argument-for-argument mirror of the C ABI."""

    def __init__(self, device=0):
        self.lib = lib = load_gpu_library()
        vp = C.c_void_p
        lib.hitl_create.argtypes = [C.POINTER(vp), C.c_int]
        lib.hitl_destroy.argtypes = [vp]
        lib.hitl_last_error.restype = C.c_char_p
        lib.hitl_last_error.argtypes = [vp]
        lib.hitl_stream.restype = vp
        lib.hitl_stream.argtypes = [vp]
        lib.hitl_launch_count.restype = C.c_uint64
        lib.hitl_launch_count.argtypes = [vp]
        lib.hitl_sm_count.argtypes = [vp]
        lib.hitl_last_kernel_ms.argtypes = [vp, C.c_int, C.POINTER(C.c_float)]
        lib.hitl_comm_unique_id.argtypes = [vp]
        lib.hitl_comm_init.argtypes = [vp, vp, C.c_int, C.c_int]
        lib.hitl_comm_destroy.argtypes = [vp]
        lib.hitl_comm_info.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.hitl_normal_eq_allreduce.argtypes = [vp, vp, vp, vp, vp, C.POINTER(C.c_float)]
        lib.hitl_gather_stf_blocks.argtypes = [vp, C.c_int, C.c_uint64, _u64p, vp, vp, vp, vp]
        lib.hitl_host_alloc.restype = vp
        lib.hitl_host_alloc.argtypes = [C.c_size_t]
        lib.hitl_host_free.argtypes = [vp]
        lib.hitl_set_scans.argtypes = [vp, C.c_uint32, _u32p, _f32p, _f32p]
        lib.hitl_build_kdtrees.argtypes = [vp]
        lib.hitl_set_kdtrees.argtypes = [vp, vp]
        lib.hitl_get_kdtrees.argtypes = [vp, vp]
        lib.hitl_kd_query.argtypes = [vp, C.c_uint32, C.c_uint32, _f32p, C.c_float, C.c_int, _f32p, _i32p]
        lib.hitl_kd_neighbors.argtypes = [vp, C.c_uint32, C.c_uint32, _f32p, C.c_float, C.c_uint32, vp, _u32p]
        lib.hitl_em_refit.argtypes = [vp, _f32p, C.c_double, C.c_int32, _f32p, C.POINTER(EmFitInfo)]
        lib.hitl_em_refit_chain.argtypes = [vp, C.c_uint32, _f32p, C.c_double, C.c_int32, C.c_uint32, _f32p, C.POINTER(EmFitInfo)]
        lib.hitl_find_stf.argtypes = [vp, _f64p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(StfOpts), C.POINTER(StfInfo)]
        lib.hitl_get_stf.argtypes = [vp, _u32p, _u32p, _u64p, _u32p, _u32p]
        lib.hitl_get_stf_work.argtypes = [vp, _u64p]
        lib.hitl_find_vo.argtypes = [vp, _f64p, C.c_int32, C.c_int32, C.POINTER(StfOpts), C.POINTER(C.c_uint64)]
        lib.hitl_get_vo.argtypes = [vp, _u32p, _u32p, _u32p]
        lib.hitl_world_transform.argtypes = [vp, _f32p, vp]
        lib.hitl_set_world_clouds.argtypes = [vp, _f32p]
        lib.hitl_em_inliers.argtypes = [vp, _f32p, C.c_double, C.c_uint64, vp, vp, vp, C.POINTER(C.c_uint64)]
        lib.hitl_em_assign.argtypes = [vp, _f32p, C.c_double, C.c_uint32, _u32p, _u32p, _u64p, _u32p, _u32p, _u64p, _u32p]
        lib.hitl_set_kdtrees_compact.argtypes = [vp, C.c_void_p]
        lib.hitl_get_kdtrees_compact.argtypes = [vp, C.c_void_p]
        lib.hitl_get_stf16.argtypes = [vp, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.hitl_verify_input.argtypes = [vp, C.c_uint32, _f32p, C.c_float, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        lib.hitl_set_stf_blocks_from_search.argtypes = [vp, C.c_float, C.c_float]
        lib.hitl_set_stf_blocks.argtypes = [vp, C.c_uint64, _u32p, _u32p, _u64p, _u32p, _u32p, C.c_float, C.c_float]
        lib.hitl_set_odometry_blocks.argtypes = [vp, C.c_uint32, _f32p]
        lib.hitl_set_human_blocks.argtypes = [vp, C.c_uint32, _i32p, _f64p]
        lib.hitl_set_p2l_glob_blocks.argtypes = [vp, C.c_uint32, _u32p, _u64p, _f32p, _f32p, _f32p, _u8p, C.c_float, C.c_float]
        lib.hitl_set_p2l_blocks.argtypes = [vp, C.c_uint64, _u32p, _f32p, _f32p, _f32p, _u8p, C.c_float, C.c_float]
        lib.hitl_eval_layout_get.argtypes = [vp, C.POINTER(EvalLayout)]
        lib.hitl_eval.argtypes = [vp, _f64p, C.c_int, vp, vp, C.POINTER(C.c_float)]
        lib.hitl_normal_eq.argtypes = [vp, _f64p, vp, vp, vp, vp, C.POINTER(C.c_float)]
        lib.hitl_normal_eq_device.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_uint64)]
        lib.hitl_debug_tile_work.argtypes = [vp, C.c_uint32, _u32p, C.POINTER(C.c_uint32)]
        lib.hitl_debug_tile_desc.argtypes = [vp, C.c_uint32, _u32p, _u32p, _u32p, _u32p, _u32p]
        lib.hitl_debug_set_tiling.argtypes = [vp, C.c_uint32, C.c_int, C.c_uint32]
        lib.hitl_debug_set_fine_occupancy.argtypes = [vp, C.c_int]
        lib.hitl_debug_set_em_cull.argtypes = [vp, C.c_int]
        lib.hitl_debug_set_search_variant.argtypes = [vp, C.c_int, C.c_int]
        lib.hitl_debug_set_tree_builder.argtypes = [vp, C.c_int]
        lib.hitl_debug_tree_stats.argtypes = [vp, C.POINTER(C.c_uint64)]
        lib.hitl_debug_sincos.argtypes = [vp, C.c_uint64, _f32p, _f32p, _f32p]
        lib.hitl_debug_relative_pose.argtypes = [vp, _f64p, C.c_uint32, _u32p, _u32p, _f32p]
        self.ctx = vp()
        rc = lib.hitl_create(C.byref(self.ctx), device)
        if rc != 0:
            self.ctx = None
            raise HitlError("hitl_create failed (status %d): no usable CUDA device %d — this library has no CPU fallback" % (rc, device))
        self.n_poses = 0
        self.n_points = 0
        self._pinned = []

    def pinned(self, shape, dtype):
        """numpy array backed by page-locked host memory from hitl_host_alloc (freed by close())."""
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) if np.ndim(shape) else int(shape)
        ptr = self.lib.hitl_host_alloc(max(n, 1) * dtype.itemsize)
        if not ptr:
            raise HitlError("hitl_host_alloc failed")
        self._pinned.append(ptr)
        buf = (C.c_char * (max(n, 1) * dtype.itemsize)).from_address(ptr)
        return np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)

    def pinned_copy(self, a):
        a = np.asarray(a)
        out = self.pinned(a.shape, a.dtype)
        out[...] = a
        return out

    def close(self):
        if getattr(self, "ctx", None):
            for ptr in self._pinned:
                self.lib.hitl_host_free(ptr)
            self._pinned = []
            self.lib.hitl_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise HitlError("status %d: %s" % (rc, self.lib.hitl_last_error(self.ctx).decode()))

    # ---- scans / trees ----
    def _poses(self, poses, dtype=np.float64, what="poses"):
        """The C ABI takes the pose array without a length and reads 3 * n_poses values: a short array is an error here, not a host over-read."""
        p = np.ascontiguousarray(poses, dtype).reshape(-1)
        if p.size != 3 * self.n_poses:
            raise HitlError("%s: expected %d values (3 per pose of the resident map), got %d" % (what, 3 * self.n_poses, p.size))
        return p

    def set_scans(self, offsets, pts, nrm, _fn="hitl_set_scans"):
        offsets = np.ascontiguousarray(offsets, np.uint32)
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1)
        nrm = np.ascontiguousarray(nrm, np.float32).reshape(-1)
        self.offsets = offsets
        self.n_poses = len(offsets) - 1
        self.n_points = int(offsets[-1]) if len(offsets) else 0
        if len(pts) != 2 * self.n_points or len(nrm) != 2 * self.n_points:
            raise HitlError("set_scans: clouds must hold 2 floats per point (%d points per the offsets)" % self.n_points)
        if len(pts) == 0:
            pts = np.zeros(2, np.float32)
            nrm = np.zeros(2, np.float32)
        fn = getattr(self.lib, _fn)
        fn.argtypes = self.lib.hitl_set_scans.argtypes
        self._ck(fn(self.ctx, self.n_poses, offsets, pts, nrm))

    def set_scans_sharded(self, offsets, pts, nrm):
        """hitl_set_scans_sharded: collective; every rank passes the same arrays and uploads 1/world of them."""
        self.set_scans(offsets, pts, nrm, _fn="hitl_set_scans_sharded")

    def set_kdtrees_compact_sharded(self, index_dim):
        index_dim = np.ascontiguousarray(index_dim, np.uint32)
        if index_dim.size != self.n_points:
            raise HitlError("set_kdtrees_compact_sharded: expected %d nodes (one per point), got %d" % (self.n_points, index_dim.size))
        self.lib.hitl_set_kdtrees_compact_sharded.argtypes = [C.c_void_p, C.c_void_p]
        self._ck(self.lib.hitl_set_kdtrees_compact_sharded(self.ctx, index_dim.ctypes.data_as(C.c_void_p)))

    def set_kdtrees_sharded(self, nodes):
        nodes = np.ascontiguousarray(nodes, KDNODE)
        if nodes.size != self.n_points:
            raise HitlError("set_kdtrees_sharded: expected %d nodes (one per point), got %d" % (self.n_points, nodes.size))
        self.lib.hitl_set_kdtrees_sharded.argtypes = [C.c_void_p, C.c_void_p]
        self._ck(self.lib.hitl_set_kdtrees_sharded(self.ctx, nodes.ctypes.data))

    def build_kdtrees(self):
        self._ck(self.lib.hitl_build_kdtrees(self.ctx))

    def set_kdtrees(self, nodes):
        nodes = np.ascontiguousarray(nodes, KDNODE)
        if nodes.size != self.n_points:
            raise HitlError("set_kdtrees: expected %d nodes (one per point), got %d" % (self.n_points, nodes.size))
        self._ck(self.lib.hitl_set_kdtrees(self.ctx, nodes.ctypes.data))

    def set_kdtrees_compact(self, index_dim):
        """Trees as one u32 per node (index | dim << 31, preorder): the points and normals are taken from the resident scans."""
        index_dim = np.ascontiguousarray(index_dim, np.uint32)
        if index_dim.size != self.n_points:
            raise HitlError("set_kdtrees_compact: expected %d nodes (one per point), got %d" % (self.n_points, index_dim.size))
        self._ck(self.lib.hitl_set_kdtrees_compact(self.ctx, index_dim.ctypes.data_as(C.c_void_p)))

    def get_kdtrees_compact(self, out=None):
        out = np.zeros(max(self.n_points, 1), np.uint32) if out is None else out
        self._ck(self.lib.hitl_get_kdtrees_compact(self.ctx, out.ctypes.data_as(C.c_void_p)))
        return out[:self.n_points]

    def get_stf16(self, n_pairs, n_matches, out=None):
        """hitl_get_stf with 16-bit point indices; out = (pair_i u32, pair_j u32, pair_off u64, k u16, idx u16) buffers or None."""
        if out is None:
            out = (np.zeros(n_pairs + 1, np.uint32), np.zeros(n_pairs + 1, np.uint32), np.zeros(n_pairs + 2, np.uint64), np.zeros(n_matches + 1, np.uint16), np.zeros(n_matches + 1, np.uint16))
        vp = C.c_void_p
        self._ck(self.lib.hitl_get_stf16(self.ctx, out[0].ctypes.data_as(vp), out[1].ctypes.data_as(vp), out[2].ctypes.data_as(vp), out[3].ctypes.data_as(vp), out[4].ctypes.data_as(vp)))
        return dict(pair_i=out[0][:n_pairs], pair_j=out[1][:n_pairs], pair_off=out[2][:n_pairs + 1], k=out[3][:n_matches], idx=out[4][:n_matches])

    def get_kdtrees(self):
        nodes = np.zeros(max(self.n_points, 1), KDNODE)
        self._ck(self.lib.hitl_get_kdtrees(self.ctx, nodes.ctypes.data))
        return nodes[:self.n_points]

    def kd_query(self, scan, q, thr, mode=0):
        q = np.ascontiguousarray(q, np.float32).reshape(-1)
        n = len(q) // 2
        d, i = np.zeros(max(n, 1), np.float32), np.zeros(max(n, 1), np.int32)
        self._ck(self.lib.hitl_kd_query(self.ctx, scan, n, q if n else np.zeros(2, np.float32), thr, mode, d, i))
        return d[:n], i[:n]

    def kd_neighbors(self, scan, q, thr, cap=64):
        """FindNeighborPoints: (lists of point indices in the reference's push order, truncated to cap; total counts)."""
        q = np.ascontiguousarray(q, np.float32).reshape(-1)
        n = len(q) // 2
        idx, cnt = np.full((max(n, 1), max(cap, 1)), -1, np.int32), np.zeros(max(n, 1), np.uint32)
        self._ck(self.lib.hitl_kd_neighbors(self.ctx, scan, n, q if n else np.zeros(2, np.float32), thr, cap, idx.ctypes.data if cap else None, cnt))
        return [idx[k, :min(int(cnt[k]), cap)].copy() for k in range(n)], cnt[:n]

    # ---- search ----
    @staticmethod
    def stf_opts(thr=0.15, min_cos=None, cap=6, skip=1, min_corr=10, disable_culling=0):
        return StfOpts(thr, default_min_cos() if min_cos is None else min_cos, cap, skip, min_corr, disable_culling)

    def find_stf(self, poses, min_pose=0, max_pose=None, src_lo=0, src_hi=0xFFFFFFFF, opts=None, fetch=True, out=None):
        poses = self._poses(poses, what="find_stf")
        if max_pose is None:
            max_pose = max(self.n_poses - 1, 0)
        opts = opts or self.stf_opts()
        info = StfInfo()
        self._ck(self.lib.hitl_find_stf(self.ctx, poses, min_pose, max_pose, src_lo, src_hi, C.byref(opts), C.byref(info)))
        res = dict(n_pairs=info.n_pairs, n_matches=info.n_matches, n_raw_matches=info.n_raw_matches, n_queries=info.n_queries,
                   n_traversals=info.n_traversals, n_tile_pairs=info.n_tile_pairs, n_coarse_pass=info.n_coarse_pass, n_in_radius=info.n_in_radius, sum_tile_cycles=info.sum_tile_cycles, max_tile_cycles=info.max_tile_cycles, ms_search=info.ms_search, ms_total=info.ms_total,
                   n_tiles=info.n_tiles, n_tiles_next=info.n_tiles_next, n_gate_fail=info.n_gate_fail, n_over_cap=info.n_over_cap, n_dir_culled=info.n_dir_culled)
        if fetch:
            res.update(self.get_stf(info.n_pairs, info.n_matches, out))
        return res

    def get_stf(self, n_pairs, n_matches, out=None):
        """out = (pair_i, pair_j, pair_off, k, idx) preallocated (e.g. pinned) arrays, or None."""
        if out is not None:
            pi, pj, off, k, idx = out
            if len(pi) < n_pairs or len(off) < n_pairs + 1 or len(k) < n_matches or len(idx) < n_matches:
                raise HitlError("get_stf: output buffers too small")
            off = off[:n_pairs + 1]
        else:
            pi, pj = np.zeros(max(n_pairs, 1), np.uint32), np.zeros(max(n_pairs, 1), np.uint32)
            off = np.zeros(n_pairs + 1, np.uint64)
            k, idx = np.zeros(max(n_matches, 1), np.uint32), np.zeros(max(n_matches, 1), np.uint32)
        self._ck(self.lib.hitl_get_stf(self.ctx, pi, pj, off, k, idx))
        return dict(pair_i=pi[:n_pairs], pair_j=pj[:n_pairs], pair_off=off, k=k[:n_matches], idx=idx[:n_matches])

    def stf_work(self):
        """SM cycles the last find_stf spent per source pose (shard-balancing feedback)."""
        w = np.zeros(max(self.n_poses, 1), np.uint64)
        self._ck(self.lib.hitl_get_stf_work(self.ctx, w))
        return w[:self.n_poses]

    def find_vo(self, poses, min_pose=0, max_pose=None, opts=None):
        poses = self._poses(poses, what="find_vo")
        if max_pose is None:
            max_pose = self.n_poses - 1
        opts = opts or self.stf_opts()
        n = C.c_uint64()
        self._ck(self.lib.hitl_find_vo(self.ctx, poses, min_pose, max_pose, C.byref(opts), C.byref(n)))
        m = max(n.value, 1)
        sp, sk, tk = np.zeros(m, np.uint32), np.zeros(m, np.uint32), np.zeros(m, np.uint32)
        self._ck(self.lib.hitl_get_vo(self.ctx, sp, sk, tk))
        return sp[:n.value], sk[:n.value], tk[:n.value]

    # ---- world / EM ----
    def world_transform(self, poses_f32, fetch=True):
        p = self._poses(poses_f32, np.float32, "world_transform")
        out = np.zeros(2 * max(self.n_points, 1), np.float32) if fetch else None
        self._ck(self.lib.hitl_world_transform(self.ctx, p, out.ctypes.data if fetch else None))
        return out[:2 * self.n_points].reshape(-1, 2) if fetch else None

    def set_world_clouds(self, world):
        w = np.ascontiguousarray(world, np.float32).reshape(-1)
        self._ck(self.lib.hitl_set_world_clouds(self.ctx, w if len(w) else np.zeros(2, np.float32)))

    def em_inliers(self, seg, thr=0.03, fetch=True, cap=None):
        seg = np.ascontiguousarray(seg, np.float32).reshape(-1)
        n = C.c_uint64()
        if not fetch:
            self._ck(self.lib.hitl_em_inliers(self.ctx, seg, thr, 0, None, None, None, C.byref(n)))
            return n.value
        cap = self.n_points if cap is None else cap
        op, oi, xy = np.zeros(max(cap, 1), np.uint32), np.zeros(max(cap, 1), np.uint32), np.zeros(2 * max(cap, 1), np.float32)
        self._ck(self.lib.hitl_em_inliers(self.ctx, seg, thr, cap, op.ctypes.data, oi.ctypes.data, xy.ctypes.data, C.byref(n)))
        return op[:n.value].copy(), oi[:n.value].copy(), xy[:2 * n.value].reshape(-1, 2).copy()

    def em_refit(self, seg, thr=0.03, max_iterations=25):
        """One EM round on the device (hitl_em_refit): E-step + SegFitEM's LM.  Returns (refit segment [4] f32, info dict)."""
        seg = np.ascontiguousarray(seg, np.float32).reshape(-1)
        out, info = np.zeros(4, np.float32), EmFitInfo()
        self._ck(self.lib.hitl_em_refit(self.ctx, seg, thr, max_iterations, out, C.byref(info)))
        return out, {k: getattr(info, k) for k, _ in EmFitInfo._fields_}

    def em_refit_chain(self, segs, rounds=2, thr=0.03, max_iterations=25):
        """`rounds` EM rounds of 1 or 2 strokes with one host wait (hitl_em_refit_chain).  segs: [n_strokes, 4].
        Returns (segments [rounds, n_strokes, 4] f32, infos [rounds][n_strokes] dicts)."""
        segs = np.ascontiguousarray(segs, np.float32).reshape(-1, 4)
        ns = len(segs)
        out, infos = np.zeros(4 * ns * rounds, np.float32), (EmFitInfo * (ns * rounds))()
        self._ck(self.lib.hitl_em_refit_chain(self.ctx, ns, segs.reshape(-1), thr, max_iterations, rounds, out, infos))
        return out.reshape(rounds, ns, 4), [[{k: getattr(infos[r * ns + q], k) for k, _ in EmFitInfo._fields_} for q in range(ns)] for r in range(rounds)]

    def debug_set_em_cull(self, on=True):
        """E-step chunk cull on / off (result-preserving)."""
        self._ck(self.lib.hitl_debug_set_em_cull(self.ctx, int(bool(on))))

    def verify_input(self, sel, thr=0.05):
        """HitLSLAM::verifyUserInput on the resident world clouds: (points_verified, seen bit mask)."""
        sel = np.ascontiguousarray(sel, np.float32).reshape(-1)
        v, m = C.c_uint32(), C.c_uint32()
        self._ck(self.lib.hitl_verify_input(self.ctx, len(sel) // 2, sel, C.c_float(thr), C.byref(v), C.byref(m)))
        return v.value, m.value

    def em_assign(self, segs, thr=0.03, min_obs=5):
        segs = np.ascontiguousarray(segs, np.float32).reshape(-1)
        n, m = max(self.n_poses, 1), max(self.n_points, 1)
        ns = np.zeros(2, np.uint32)
        bufs = [(np.zeros(n, np.uint32), np.zeros(n + 1, np.uint64), np.zeros(m, np.uint32)) for _ in range(2)]
        self._ck(self.lib.hitl_em_assign(self.ctx, segs, thr, min_obs, ns, bufs[0][0], bufs[0][1], bufs[0][2], bufs[1][0], bufs[1][1], bufs[1][2]))
        out = []
        for f in range(2):
            k = int(ns[f])
            off = bufs[f][1][:k + 1].copy()
            out.append((bufs[f][0][:k].copy(), off, bufs[f][2][:int(off[-1])].copy()))
        return out

    # ---- residual blocks ----
    def set_stf_blocks_from_search(self, std_dev=0.05, corr=1.0 / 40.0):
        self._ck(self.lib.hitl_set_stf_blocks_from_search(self.ctx, std_dev, corr))

    def set_stf_blocks(self, corr_set, std_dev=0.05, corr=1.0 / 40.0):
        n = len(corr_set["pair_i"])
        z32 = np.zeros(1, np.uint32)
        self._ck(self.lib.hitl_set_stf_blocks(self.ctx, n, np.ascontiguousarray(corr_set["pair_i"], np.uint32) if n else z32,
                                              np.ascontiguousarray(corr_set["pair_j"], np.uint32) if n else z32,
                                              np.ascontiguousarray(corr_set["pair_off"], np.uint64),
                                              np.ascontiguousarray(corr_set["k"], np.uint32) if n else z32,
                                              np.ascontiguousarray(corr_set["idx"], np.uint32) if n else z32, std_dev, corr))

    def set_odometry_blocks(self, consts9):
        c = np.ascontiguousarray(consts9, np.float32).reshape(-1)
        self._ck(self.lib.hitl_set_odometry_blocks(self.ctx, len(c) // 9, c if len(c) else np.zeros(9, np.float32)))

    def set_human_blocks(self, type_pose, targets4):
        tp = np.ascontiguousarray(type_pose, np.int32).reshape(-1)
        tg = np.ascontiguousarray(targets4, np.float64).reshape(-1)
        self._ck(self.lib.hitl_set_human_blocks(self.ctx, len(tp) // 2, tp if len(tp) else np.zeros(2, np.int32), tg if len(tg) else np.zeros(4)))

    def set_p2l_glob_blocks(self, blk_pose, blk_off, pts, line_n, line_off, valid, std_dev, corr):
        self._ck(self.lib.hitl_set_p2l_glob_blocks(self.ctx, len(blk_pose), np.ascontiguousarray(blk_pose, np.uint32), np.ascontiguousarray(blk_off, np.uint64),
                                                   np.ascontiguousarray(pts, np.float32).reshape(-1), np.ascontiguousarray(line_n, np.float32).reshape(-1),
                                                   np.ascontiguousarray(line_off, np.float32), np.ascontiguousarray(valid, np.uint8), std_dev, corr))

    def set_p2l_blocks(self, pose_idx, pts, line_n, line_off, valid, std_dev, corr):
        self._ck(self.lib.hitl_set_p2l_blocks(self.ctx, len(pose_idx), np.ascontiguousarray(pose_idx, np.uint32), np.ascontiguousarray(pts, np.float32).reshape(-1),
                                              np.ascontiguousarray(line_n, np.float32).reshape(-1), np.ascontiguousarray(line_off, np.float32),
                                              np.ascontiguousarray(valid, np.uint8), std_dev, corr))

    def layout(self):
        L = EvalLayout()
        self._ck(self.lib.hitl_eval_layout_get(self.ctx, C.byref(L)))
        return L

    def eval(self, poses, precision=0, want_jac=True, fetch=True, out=None):
        """out = (r, J) preallocated float64 arrays (e.g. pinned) at least as large as the layout."""
        poses = self._poses(poses, what="eval")
        L = self.layout()
        if out is not None:
            if len(out[0]) < L.n_residuals or (want_jac and len(out[1]) < L.n_jacobian):
                raise HitlError("eval: output buffers too small")
            r, J = out[0][:max(L.n_residuals, 1)], (out[1][:max(L.n_jacobian, 1)] if want_jac else None)
        else:
            r = np.zeros(max(L.n_residuals, 1)) if fetch else None
            J = np.zeros(max(L.n_jacobian, 1)) if (fetch and want_jac) else None
        ms = C.c_float()
        self._ck(self.lib.hitl_eval(self.ctx, poses, precision, r.ctypes.data if r is not None else None, J.ctypes.data if J is not None else None, C.byref(ms)))
        out = dict(ms=ms.value, layout=L)
        if fetch:
            out.update(split_eval(L, r, J))
        return out

    def normal_eq(self, poses, fetch=True):
        poses = self._poses(poses, what="normal_eq")
        n = self.n_poses
        L = self.layout()
        nbin = L.n_odometry + L.n_stf
        ms = C.c_float()
        if not fetch:
            self._ck(self.lib.hitl_normal_eq(self.ctx, poses, None, None, None, None, C.byref(ms)))
            return dict(ms=ms.value)
        H, g, Ho, cost = np.zeros(9 * n), np.zeros(3 * n), np.zeros(9 * max(nbin, 1)), np.zeros(1)
        self._ck(self.lib.hitl_normal_eq(self.ctx, poses, H.ctypes.data, g.ctypes.data, Ho.ctypes.data, cost.ctypes.data, C.byref(ms)))
        return dict(H_diag=H.reshape(n, 3, 3), g=g.reshape(n, 3), H_off=Ho[:9 * nbin].reshape(-1, 3, 3), cost=float(cost[0]), ms=ms.value)

    # ---- multi-GPU exchange (NCCL inside the C ABI) ----
    @staticmethod
    def comm_unique_id():
        """128-byte ncclUniqueId from hitl_comm_unique_id (one rank calls it; hand the bytes to the others out of band)."""
        preload_nccl()
        lib = load_gpu_library()
        lib.hitl_comm_unique_id.argtypes = [C.c_void_p]
        buf = (C.c_ubyte * 128)()
        if lib.hitl_comm_unique_id(buf) != 0:
            raise HitlError("hitl_comm_unique_id: NCCL is not available in this process")
        return bytes(buf)

    def comm_init(self, unique_id, rank, world):
        preload_nccl()
        buf = (C.c_ubyte * 128).from_buffer_copy(bytes(unique_id))
        self._ck(self.lib.hitl_comm_init(self.ctx, buf, rank, world))
        self.comm_rank, self.comm_world = rank, world

    def comm_destroy(self):
        self._ck(self.lib.hitl_comm_destroy(self.ctx))

    def comm_info(self):
        r, w, v = C.c_int(), C.c_int(), C.c_int()
        self._ck(self.lib.hitl_comm_info(self.ctx, C.byref(r), C.byref(w), C.byref(v)))
        return dict(rank=r.value, world=w.value, nccl_version=v.value)

    def normal_eq_allreduce(self, poses=None, fetch=True):
        """hitl_normal_eq_allreduce: (poses given) this rank's normal equations, then the in-library ncclAllReduce of [H_diag | g | cost]."""
        n = self.n_poses
        ms = C.c_float()
        pp = self._poses(poses, what="normal_eq_allreduce").ctypes.data if poses is not None else None
        if not fetch:
            self._ck(self.lib.hitl_normal_eq_allreduce(self.ctx, pp, None, None, None, C.byref(ms)))
            return dict(ms=ms.value)
        H, g, cost = np.zeros(9 * n), np.zeros(3 * n), np.zeros(1)
        self._ck(self.lib.hitl_normal_eq_allreduce(self.ctx, pp, H.ctypes.data, g.ctypes.data, cost.ctypes.data, C.byref(ms)))
        return dict(H_diag=H.reshape(n, 3, 3), g=g.reshape(n, 3), cost=float(cost[0]), ms=ms.value)

    def gather_stf_blocks(self, root=0, cap_blocks=None):
        """hitl_gather_stf_blocks after eval(): on root the STF blocks of all ranks in rank order (pair_i, pair_j, r [B,2], J [B,2,2,3])."""
        world = getattr(self, "comm_world", 1)
        rank = getattr(self, "comm_rank", 0)
        counts = np.zeros(world, np.uint64)
        if rank != root:
            self._ck(self.lib.hitl_gather_stf_blocks(self.ctx, root, 0, counts, None, None, None, None))
            return dict(counts=counts)
        if cap_blocks is None:                          # sizes first (the collective runs once more with buffers that fit)
            cap_blocks = int(self.layout().n_stf) if world == 1 else None
        if cap_blocks is None:
            raise HitlError("gather_stf_blocks: pass cap_blocks on the root of a multi-rank job (all ranks make ONE collective call)")
        pi, pj = np.zeros(max(cap_blocks, 1), np.uint32), np.zeros(max(cap_blocks, 1), np.uint32)
        r, J = np.zeros(2 * max(cap_blocks, 1)), np.zeros(12 * max(cap_blocks, 1))
        self._ck(self.lib.hitl_gather_stf_blocks(self.ctx, root, cap_blocks, counts, pi.ctypes.data, pj.ctypes.data, r.ctypes.data, J.ctypes.data))
        t = int(counts.sum())
        return dict(counts=counts, pair_i=pi[:t], pair_j=pj[:t], r=r[:2 * t].reshape(-1, 2), J=J[:12 * t].reshape(-1, 2, 2, 3))

    def set_deterministic(self, on=True):
        self.lib.hitl_set_deterministic.argtypes = [C.c_void_p, C.c_int]
        self._ck(self.lib.hitl_set_deterministic(self.ctx, int(bool(on))))

    def normal_eq_device(self):
        p, n = C.c_void_p(), C.c_uint64()
        self._ck(self.lib.hitl_normal_eq_device(self.ctx, C.byref(p), C.byref(n)))
        return p.value, n.value

    def debug_tile_work(self):
        n = C.c_uint32()
        self._ck(self.lib.hitl_debug_tile_work(self.ctx, 0, np.zeros(1, np.uint32), C.byref(n)))
        w = np.zeros(max(n.value, 1), np.uint32)
        self._ck(self.lib.hitl_debug_tile_work(self.ctx, n.value, w, C.byref(n)))
        return w[:n.value].astype(np.uint64) * 64

    def debug_tile_desc(self):
        n = len(self.debug_tile_work())
        a = [np.zeros(max(n, 1), np.uint32) for _ in range(5)]
        self._ck(self.lib.hitl_debug_tile_desc(self.ctx, n, *a))
        return dict(scan=a[0][:n], k0=a[1][:n] & 0xFFFF, len=a[1][:n] >> 16, jlo=a[2][:n], jhi=a[3][:n], open=a[4][:n])

    def debug_set_tiling(self, max_len=32, adaptive=True, target_parts=1):
        self._ck(self.lib.hitl_debug_set_tiling(self.ctx, max_len, int(adaptive), int(target_parts)))

    def debug_set_tree_builder(self, host=False):
        self._ck(self.lib.hitl_debug_set_tree_builder(self.ctx, int(host)))

    def debug_tree_stats(self):
        n = C.c_uint64()
        self._ck(self.lib.hitl_debug_tree_stats(self.ctx, C.byref(n)))
        return int(n.value)

    def debug_set_search_variant(self, variant=0, smem_carveout_pct=-1):
        self._ck(self.lib.hitl_debug_set_search_variant(self.ctx, int(variant), int(smem_carveout_pct)))

    def debug_set_fine_occupancy(self, on=True):
        self._ck(self.lib.hitl_debug_set_fine_occupancy(self.ctx, int(on)))

    def debug_sincos(self, x):
        x = np.ascontiguousarray(x, np.float32)
        s, c = np.zeros_like(x), np.zeros_like(x)
        self._ck(self.lib.hitl_debug_sincos(self.ctx, len(x), x, s, c))
        return s, c

    def debug_relative_pose(self, poses, src, dst):
        poses = np.ascontiguousarray(poses, np.float64).reshape(-1)
        src, dst = np.ascontiguousarray(src, np.uint32), np.ascontiguousarray(dst, np.uint32)
        out = np.zeros(6 * len(src), np.float32)
        self._ck(self.lib.hitl_debug_relative_pose(self.ctx, poses, len(src), src, dst, out))
        return out.reshape(-1, 6)

    KERNELS = {"stf_search_kernel": 0, "eval_stf_kernel": 1, "em_inliers_kernel": 2, "em_assign_kernel": 3, "world_transform_kernel": 4, "em_fit_kernel": 5, "allreduce": 6}

    def last_kernel_ms(self, name):
        """Duration of the last launch of one named kernel (CUDA events on the library's stream)."""
        ms = C.c_float()
        self._ck(self.lib.hitl_last_kernel_ms(self.ctx, self.KERNELS[name], C.byref(ms)))
        return ms.value

    def launch_count(self):
        return self.lib.hitl_launch_count(self.ctx)

    def sm_count(self):
        return self.lib.hitl_sm_count(self.ctx)


def split_eval(L, r, J):
    """Split the flat hitl_eval outputs into per-kind arrays (layout in include/hitl_gpu.h)."""
    out = {}
    ro = jo = 0
    for name, n, nr, nj, jshape in (("odometry", L.n_odometry, 3, 18, (2, 3, 3)), ("human", L.n_human, 3, 9, (3, 3)),
                                    ("stf", L.n_stf, 2, 12, (2, 2, 3)), ("p2l_glob", L.n_p2l_glob, 1, 3, (3,)), ("p2l", L.n_p2l, 1, 3, (3,))):
        out["r_" + name] = r[ro:ro + nr * n].reshape(n, nr)
        if J is not None:
            out["J_" + name] = J[jo:jo + nj * n].reshape((n,) + jshape)
        ro += nr * n
        jo += nj * n
    return out


def kdtree_build_host(pts, nrm):
    """Flat preorder KD-tree of one scan, built by the library's host builder (no GPU needed)."""
    lib = load_gpu_library()
    lib.hitl_kdtree_build_host.argtypes = [_f32p, _f32p, C.c_uint32, C.c_void_p]
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1)
    nrm = np.ascontiguousarray(nrm, np.float32).reshape(-1)
    n = len(pts) // 2
    nodes = np.zeros(max(n, 1), KDNODE)
    rc = lib.hitl_kdtree_build_host(pts if n else np.zeros(2, np.float32), nrm if n else np.zeros(2, np.float32), n, nodes.ctypes.data)
    if rc:
        raise HitlError("hitl_kdtree_build_host failed: %d" % rc)
    return nodes[:n]


# ---- host mirror library (file formats) ---------------------------------------------------------
class HostLib:
    def __init__(self):
        path = lib_path("libhitl_host.so")
        if not os.path.exists(path):
            raise HitlError("libhitl_host.so is not built (run `python -m hitl_slam_b200.build`)")
        self.lib = lib = C.CDLL(path)
        lib.hitl_host_load_pose_graph.restype = C.c_void_p
        lib.hitl_host_load_pose_graph.argtypes = [C.c_char_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        lib.hitl_host_pose_graph_get.argtypes = [C.c_void_p, _f32p, _f32p, _u32p, _f32p, _f32p]
        lib.hitl_host_pose_graph_free.argtypes = [C.c_void_p]
        lib.hitl_host_save_stfs_covars.argtypes = [C.c_char_p, C.c_char_p, C.c_double, C.c_uint32, _f32p, _f32p, _u32p, _f32p, _f32p]
        lib.hitl_host_save_poses.argtypes = [C.c_char_p, C.c_uint32, _f32p]
        lib.hitl_host_sinf.restype = C.c_float
        lib.hitl_host_sinf.argtypes = [C.c_float]
        lib.hitl_host_cosf.restype = C.c_float
        lib.hitl_host_cosf.argtypes = [C.c_float]
        lib.hitl_host_sincos_mismatches.restype = C.c_uint64
        lib.hitl_host_sincos_mismatches.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
        lib.hitl_host_relative_pose.argtypes = [_f64p, C.c_uint32, C.c_uint32, _f32p]

    def sincos_mismatches(self, first, count, stride):
        return int(self.lib.hitl_host_sincos_mismatches(first, count, stride))

    def relative_pose(self, poses, src, dst):
        out = np.zeros(6, np.float32)
        self.lib.hitl_host_relative_pose(np.ascontiguousarray(poses, np.float64).reshape(-1), src, dst, out)
        return out

    def load_pose_graph(self, path, cache=None):
        """cache: None = parse the text file; True = through path + ".hitlcache"; a string = through that cache file.
        With a cache the result carries from_cache (bool)."""
        n, m = C.c_uint64(), C.c_uint64()
        hit = C.c_int(0)
        if cache is None:
            h = self.lib.hitl_host_load_pose_graph(path.encode(), C.byref(n), C.byref(m))
        else:
            self.lib.hitl_host_load_pose_graph_cached.restype = C.c_void_p
            self.lib.hitl_host_load_pose_graph_cached.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_int)]
            h = self.lib.hitl_host_load_pose_graph_cached(path.encode(), None if cache is True else str(cache).encode(), C.byref(n), C.byref(m), C.byref(hit))
        if not h:
            raise IOError("cannot read pose graph " + path)
        poses, cov = np.zeros(3 * n.value, np.float32), np.zeros(9 * n.value, np.float32)
        off = np.zeros(n.value + 1, np.uint32)
        pts, nrm = np.zeros(2 * m.value, np.float32), np.zeros(2 * m.value, np.float32)
        self.lib.hitl_host_pose_graph_get(h, poses, cov, off, pts, nrm)
        self.lib.hitl_host_pose_graph_free(h)
        out = dict(poses=poses.reshape(-1, 3), cov=cov.reshape(-1, 9), offsets=off, pts=pts.reshape(-1, 2), nrm=nrm.reshape(-1, 2))
        if cache is not None:
            out["from_cache"] = bool(hit.value)
        return out

    def save_stfs_covars(self, path, poses, cov, offsets, obs_world, nrm_world, map_name="synthetic", timestamp=0.0):
        poses = np.ascontiguousarray(poses, np.float32).reshape(-1)
        rc = self.lib.hitl_host_save_stfs_covars(path.encode(), map_name.encode(), timestamp, len(poses) // 3, poses,
                                                 np.ascontiguousarray(cov, np.float32).reshape(-1), np.ascontiguousarray(offsets, np.uint32),
                                                 np.ascontiguousarray(obs_world, np.float32).reshape(-1),
                                                 np.ascontiguousarray(nrm_world, np.float32).reshape(-1))
        if rc:
            raise IOError("cannot write " + path)

    # ---- device-free pieces of the C++ mirror ----
    def _bind_mirror(self):
        lib = self.lib
        if getattr(lib, "_mirror_bound", False):
            return
        vp = C.c_void_p
        lib.hitl_host_seg_fit_em.argtypes = [_f64p, _f64p, _f64p, C.c_int, _f32p]
        lib.hitl_host_app_exp_correct.argtypes = [C.c_int, _f32p, C.c_uint32, _f32p, C.c_uint32, _i32p, _f32p]
        lib.hitl_host_constraint_targets.argtypes = [C.c_int, _f32p, C.c_uint32, _f32p, C.c_uint32, _i32p, C.c_uint32, _i32p, _i32p, _f32p]
        lib.hitl_host_backprop.argtypes = [vp, C.c_uint32, _f32p, _f32p, C.c_int32, C.c_int32, _f32p, C.POINTER(C.c_float)]
        lib.hitl_host_session_correct.argtypes = [vp, C.c_int, _f32p, vp, C.c_int, _i32p, _f64p, _f64p]
        lib.hitl_host_odometry_consts.argtypes = [_f32p, C.c_uint32, _f32p]
        lib.hitl_host_human_targets.argtypes = [_f32p, C.c_uint32, C.c_uint32, _i32p, _f32p, _f64p]
        lib.hitl_host_solver_selftest.argtypes = [_f64p, C.c_int, C.c_int, C.c_int, _f64p]
        lib.hitl_host_load_log.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, _i32p, _i32p, _i32p, _f32p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        lib.hitl_host_save_log.argtypes = [C.c_char_p, C.c_uint32, _i32p, _i32p, _i32p, _f32p]
        lib.hitl_host_session_create.restype = vp
        lib.hitl_host_session_create.argtypes = [vp]
        lib.hitl_host_session_destroy.argtypes = [vp]
        lib.hitl_host_session_error.restype = C.c_char_p
        lib.hitl_host_session_error.argtypes = [vp]
        lib.hitl_host_session_set_map.argtypes = [vp, C.c_uint32, _f32p, _u32p, _f32p, _f32p]
        lib.hitl_host_session_set_poses.argtypes = [vp, _f32p]
        lib.hitl_host_session_get_poses.argtypes = [vp, _f32p, _f64p]
        lib.hitl_host_session_world_transform.argtypes = [vp, C.c_int]
        lib.hitl_host_session_em_run.argtypes = [vp, C.c_int, _f32p, _i32p]
        lib.hitl_host_session_em_poses.argtypes = [vp, _i32p, _i32p]
        lib.hitl_host_session_add_constraints_from_em.argtypes = [vp, C.POINTER(C.c_uint32)]
        lib.hitl_host_session_add_constraints.argtypes = [vp, C.c_uint32, _i32p, _f32p]
        lib.hitl_host_session_clear_constraints.argtypes = [vp]
        lib.hitl_host_session_verify_input.argtypes = [vp, _f32p, C.POINTER(C.c_uint32)]
        lib.hitl_host_session_solver_options.argtypes = [vp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int]
        lib.hitl_host_session_joint_opt_run.argtypes = [vp, C.c_int, _f64p]
        lib.hitl_host_session_solve.argtypes = [vp, C.c_int, _f64p]
        lib.hitl_host_session_copy_params.argtypes = [vp]
        lib.hitl_host_session_find_stf.argtypes = [vp, C.c_uint64, C.c_uint64, _u64p]
        lib.hitl_host_session_get_stf.argtypes = [vp, _u32p, _u32p, _u64p, _u32p, _u32p]
        lib.hitl_host_session_gradient.argtypes = [vp, C.c_uint64, _f64p, C.POINTER(C.c_uint64), _u64p]
        lib.hitl_host_session_evaluate_block.argtypes = [vp, C.c_int, C.c_uint64, vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32), _f64p, _f64p, _f64p, C.POINTER(C.c_uint64)]
        lib._mirror_bound = True

    def seg_fit_em_theta(self, p1, p2, data):
        """The host M-step (FitSegmentAngle on the host LM): (endpoints [2,2] f32, theta, LM iterations)."""
        self._bind_mirror()
        self.lib.hitl_host_seg_fit_em_theta.argtypes = [_f64p, _f64p, _f64p, C.c_int, _f32p, C.POINTER(C.c_double), C.POINTER(C.c_int)]
        data = np.ascontiguousarray(data, np.float64).reshape(-1)
        out, th, it = np.zeros(4, np.float32), C.c_double(), C.c_int()
        rc = self.lib.hitl_host_seg_fit_em_theta(np.ascontiguousarray(p1, np.float64), np.ascontiguousarray(p2, np.float64), data if len(data) else np.zeros(2), len(data) // 2, out,
                                                 C.byref(th), C.byref(it))
        if rc != 0:
            raise HitlError("hitl_host_seg_fit_em_theta failed")
        return out.reshape(2, 2), th.value, it.value

    def seg_fit_em(self, p1, p2, data):
        """EMInput::SegFitEM of the C++ mirror (host only)."""
        self._bind_mirror()
        data = np.ascontiguousarray(data, np.float64).reshape(-1)
        out = np.zeros(4, np.float32)
        rc = self.lib.hitl_host_seg_fit_em(np.ascontiguousarray(p1, np.float64), np.ascontiguousarray(p2, np.float64), data if len(data) else np.zeros(2), len(data) // 2, out)
        if rc:
            raise HitlError("hitl_host_seg_fit_em failed")
        return out.reshape(2, 2)

    def odometry_consts(self, poses_f32):
        """PoseConstraint constants as JointOpt::AddOdometryConstraints of the C++ mirror computes them."""
        self._bind_mirror()
        p = np.ascontiguousarray(poses_f32, np.float32).reshape(-1)
        n = len(p) // 3
        out = np.zeros(9 * max(n - 1, 1), np.float32)
        self.lib.hitl_host_odometry_consts(p, n, out)
        return out[:9 * max(n - 1, 0)].reshape(-1, 9)

    def human_targets(self, poses_f32, ids3, deltas4):
        self._bind_mirror()
        p = np.ascontiguousarray(poses_f32, np.float32).reshape(-1)
        ids3 = np.ascontiguousarray(ids3, np.int32).reshape(-1)
        out = np.zeros(4 * (len(ids3) // 3))
        self.lib.hitl_host_human_targets(p, len(p) // 3, len(ids3) // 3, ids3, np.ascontiguousarray(deltas4, np.float32).reshape(-1), out)
        return out.reshape(-1, 4)

    def solver_selftest(self, x0, max_iterations=200, hold_x1=False, force_cg=False):
        self._bind_mirror()
        x = np.ascontiguousarray(x0, np.float64).copy()
        out = np.zeros(4)
        self.lib.hitl_host_solver_selftest(x, max_iterations, int(hold_x1), int(force_cg), out)
        return x, dict(initial_cost=out[0], final_cost=out[1], iterations=int(out[2]), termination=int(out[3]))

    def evaluate_selftest(self, x0, hold_x1=False):
        """Problem::Evaluate of Powell's function: (gradient, (rows, cols, nnz) of the CRS Jacobian, cost)."""
        self._bind_mirror()
        self.lib.hitl_host_evaluate_selftest.argtypes = [_f64p, C.c_int, _f64p, _u64p, C.POINTER(C.c_double)]
        g, dims, cost = np.zeros(4), np.zeros(3, np.uint64), C.c_double()
        n = self.lib.hitl_host_evaluate_selftest(np.ascontiguousarray(x0, np.float64), int(hold_x1), g, dims, C.byref(cost))
        return g[:max(n, 0)], tuple(int(v) for v in dims), cost.value

    def save_log(self, path, entries):
        """entries: list of (type, undone, [[x, y], ...]) in the reference's session-log format."""
        self._bind_mirror()
        types = np.array([e[0] for e in entries], np.int32)
        undone = np.array([e[1] for e in entries], np.int32)
        npts = np.array([len(e[2]) for e in entries], np.int32)
        pts = np.concatenate([np.asarray(e[2], np.float32).reshape(-1, 2) for e in entries]).reshape(-1) if entries else np.zeros(2, np.float32)
        z = np.zeros(1, np.int32)
        if self.lib.hitl_host_save_log(path.encode(), len(entries), types if len(entries) else z, undone if len(entries) else z, npts if len(entries) else z,
                                       np.ascontiguousarray(pts, np.float32)):
            raise IOError("cannot write " + path)

    def load_log(self, path, cap_entries=4096):
        self._bind_mirror()
        types, undone, npts = np.zeros(cap_entries, np.int32), np.zeros(cap_entries, np.int32), np.zeros(cap_entries, np.int32)
        pts = np.zeros(2 * 8 * cap_entries, np.float32)
        ne, npnt = C.c_uint32(), C.c_uint32()
        rc = self.lib.hitl_host_load_log(path.encode(), cap_entries, 8 * cap_entries, types, undone, npts, pts, C.byref(ne), C.byref(npnt))
        if rc:
            raise IOError("cannot read session log %s (status %d)" % (path, rc))
        out, o = [], 0
        for e in range(ne.value):
            m = int(npts[e])
            out.append((int(types[e]), int(undone[e]), pts[2 * o:2 * (o + m)].reshape(-1, 2).copy()))
            o += m
        return out


def _hostlib_app_exp_correct(self, ctype, sel, poses_f32, corrected):
    """AppExpCorrect::Run on explicit inputs: returns (poses [N,3] f32, C [3] f32 or None if no group was applied)."""
    self._bind_mirror()
    p = np.ascontiguousarray(poses_f32, np.float32).reshape(-1).copy()
    corr = np.ascontiguousarray(corrected, np.int32)
    c3 = np.zeros(3, np.float32)
    rc = self.lib.hitl_host_app_exp_correct(int(ctype), np.ascontiguousarray(sel, np.float32).reshape(-1), len(p) // 3, p, len(corr), corr, c3)
    if rc < 0:
        raise HitlError("hitl_host_app_exp_correct failed")
    return p.reshape(-1, 3), (c3 if rc == 1 else None)


def _hostlib_backprop(self, gpu, poses_f32, cov9, lo, hi, c3):
    """Backprop::Run (pose update on the GPU): returns (poses [N,3] f32, cov [N,9] f32, device ms)."""
    self._bind_mirror()
    p = np.ascontiguousarray(poses_f32, np.float32).reshape(-1).copy()
    cov = np.ascontiguousarray(cov9, np.float32).reshape(-1).copy()
    ms = C.c_float()
    rc = self.lib.hitl_host_backprop(gpu.ctx, len(p) // 3, p, cov, int(lo), int(hi), np.ascontiguousarray(c3, np.float32), C.byref(ms))
    if rc != 0:
        raise HitlError("hitl_host_backprop failed: " + gpu.lib.hitl_last_error(gpu.ctx).decode())
    return p.reshape(-1, 3), cov.reshape(-1, 9), ms.value


def _hostlib_constraint_targets(self, ctype, sel, poses_f32, corrected, anchor):
    """AppExpCorrect::calculateConstraintTargets: (ids3 [B,3] i32 = type, constrained, anchor; deltas4 [B,4] f32), anchor-major."""
    self._bind_mirror()
    p = np.ascontiguousarray(poses_f32, np.float32).reshape(-1)
    cor, anc = np.ascontiguousarray(corrected, np.int32), np.ascontiguousarray(anchor, np.int32)
    nb = max(len(cor) * len(anc), 1)
    ids, dl = np.zeros(3 * nb, np.int32), np.zeros(4 * nb, np.float32)
    n = self.lib.hitl_host_constraint_targets(int(ctype), np.ascontiguousarray(sel, np.float32).reshape(-1), len(p) // 3, p, len(cor), cor, len(anc), anc, ids, dl)
    if n < 0:
        raise HitlError("hitl_host_constraint_targets failed")
    return ids[:3 * n].reshape(-1, 3).copy(), dl[:4 * n].reshape(-1, 4).copy()


HostLib.app_exp_correct = _hostlib_app_exp_correct
HostLib.constraint_targets = _hostlib_constraint_targets
HostLib.backprop = _hostlib_backprop


class HostSession:
    """JointOpt + EMInput of the C++ host mirror on one GPU context (host_capi.cpp)."""

    def __init__(self, gpu, host=None):
        self.gpu = gpu
        self.host = host or HostLib()
        self.host._bind_mirror()
        self.lib = self.host.lib
        self.s = self.lib.hitl_host_session_create(gpu.ctx)
        if not self.s:
            raise HitlError("hitl_host_session_create failed")
        self.n_poses = 0

    def close(self):
        if getattr(self, "s", None):
            self.lib.hitl_host_session_destroy(self.s)
            self.s = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise HitlError("host session: " + self.lib.hitl_host_session_error(self.s).decode())

    def set_map(self, poses, offsets, pts, nrm):
        poses = np.ascontiguousarray(poses, np.float32).reshape(-1)
        self.n_poses = len(poses) // 3
        self.offsets = np.ascontiguousarray(offsets, np.uint32)
        self.n_points = int(self.offsets[-1])
        self._ck(self.lib.hitl_host_session_set_map(self.s, self.n_poses, poses, self.offsets, np.ascontiguousarray(pts, np.float32).reshape(-1),
                                                    np.ascontiguousarray(nrm, np.float32).reshape(-1)))
        self.gpu.n_poses, self.gpu.n_points, self.gpu.offsets = self.n_poses, self.n_points, self.offsets

    def set_poses(self, poses):
        self._ck(self.lib.hitl_host_session_set_poses(self.s, np.ascontiguousarray(poses, np.float32).reshape(-1)))

    def poses(self):
        p, a = np.zeros(3 * self.n_poses, np.float32), np.zeros(3 * self.n_poses)
        self.lib.hitl_host_session_get_poses(self.s, p, a)
        return p.reshape(-1, 3), a.reshape(-1, 3)

    def world_transform(self, keep_host_copy=False):
        self._ck(self.lib.hitl_host_session_world_transform(self.s, int(keep_host_copy)))

    def use_shard_contexts(self, extra_gpus):
        """JointOpt::UseShardContexts: HitlGpu objects on OTHER devices that share the search and the STF blocks (call set_map afterwards)."""
        self._extra = list(extra_gpus)                    # keep them alive
        arr = (C.c_void_p * max(len(self._extra), 1))(*[g.ctx for g in self._extra])
        self.lib.hitl_host_session_use_shard_contexts.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p)]
        self._ck(self.lib.hitl_host_session_use_shard_contexts(self.s, len(self._extra), arr))

    def shard_info(self):
        out, n = np.zeros(3 * 64, np.uint64), C.c_uint32()
        self.lib.hitl_host_session_shard_info.argtypes = [C.c_void_p, C.c_uint32, _u64p, C.POINTER(C.c_uint32)]
        self.lib.hitl_host_session_shard_info(self.s, 64, out, C.byref(n))
        return [tuple(int(x) for x in out[3 * r:3 * r + 3]) for r in range(n.value)]

    def set_device_m_step(self, on=True):
        """Where EMInput's M-step runs: device (hitl_em_refit, default) or the host LM (the checker)."""
        self.lib.hitl_host_session_set_device_m_step.argtypes = [C.c_void_p, C.c_int]
        self.lib.hitl_host_session_set_device_m_step(self.s, int(bool(on)))

    def set_em_chain_rounds(self, rounds=2):
        """Device M-step: EM rounds of both strokes enqueued per host wait (hitl_em_refit_chain); 1 = a wait per round."""
        self.lib.hitl_host_session_set_em_chain_rounds.argtypes = [C.c_void_p, C.c_int]
        self.lib.hitl_host_session_set_em_chain_rounds(self.s, int(rounds))

    def em_run(self, correction_type, selected_points):
        sel = np.ascontiguousarray(selected_points, np.float32).reshape(-1).copy()
        info = np.zeros(6, np.int32)
        self._ck(self.lib.hitl_host_session_em_run(self.s, int(correction_type), sel, info))
        cor, anc = np.zeros(max(int(info[0]), 1), np.int32), np.zeros(max(int(info[1]), 1), np.int32)
        self.lib.hitl_host_session_em_poses(self.s, cor, anc)
        return dict(segs=sel.reshape(4, 2), corrected=cor[:info[0]].copy(), anchor=anc[:info[1]].copy(), backprop=(int(info[2]), int(info[3])),
                    rounds=(int(info[4]), int(info[5])))

    def verify_input(self, selected_points):
        """HitLSLAM::verifyUserInput on the session's resident world clouds: number of verified points (4 = go on)."""
        v = C.c_uint32()
        self._ck(self.lib.hitl_host_session_verify_input(self.s, np.ascontiguousarray(selected_points, np.float32).reshape(-1), C.byref(v)))
        return v.value

    def correct(self, correction_type, selected_points, cov=None, solve=True, verify=False):
        """One full human correction as HitLSLAM::Run wires it: (input verification ->) EM -> explicit correction -> back-propagation ->
        constraints (-> JointOpt::Run).  cov [N,9] f32 is updated in place when given.  With verify=True an input that fails
        verifyUserInput is dropped, as in HitLSLAM::replayLog."""
        sel = np.ascontiguousarray(selected_points, np.float32).reshape(-1).copy()
        if verify and self.verify_input(sel) != 4:
            return dict(segs=sel.reshape(4, 2), n_corrected=0, n_anchor=0, backprop=(0, 0), rounds=(0, 0), n_constraints=0, applied=False, verified=False, ms={})
        info, ms, summ = np.zeros(8, np.int32), np.zeros(5), np.zeros(6)
        if cov is not None:
            assert cov.dtype == np.float32 and cov.flags["C_CONTIGUOUS"]
        self._ck(self.lib.hitl_host_session_correct(self.s, int(correction_type), sel, cov.ctypes.data if cov is not None else None, int(solve), info, ms, summ))
        return dict(segs=sel.reshape(4, 2), n_corrected=int(info[0]), n_anchor=int(info[1]), backprop=(int(info[2]), int(info[3])), rounds=(int(info[4]), int(info[5])),
                    n_constraints=int(info[6]), applied=bool(info[7]),
                    ms=dict(em=ms[0], explicit=ms[1], backprop=ms[2], backprop_device=ms[3], joint_opt=ms[4]),
                    initial_cost=summ[0], final_cost=summ[1], successful_steps=int(summ[2]), unsuccessful_steps=int(summ[3]))

    def add_constraints_from_em(self):
        n = C.c_uint32()
        self._ck(self.lib.hitl_host_session_add_constraints_from_em(self.s, C.byref(n)))
        return n.value

    def add_constraints(self, ids3, deltas4):
        ids3 = np.ascontiguousarray(ids3, np.int32).reshape(-1)
        self._ck(self.lib.hitl_host_session_add_constraints(self.s, len(ids3) // 3, ids3, np.ascontiguousarray(deltas4, np.float32).reshape(-1)))

    def clear_constraints(self):
        self.lib.hitl_host_session_clear_constraints(self.s)

    def solver_options(self, which=0, max_iterations=-1, function_tolerance=-1.0, gradient_tolerance=-1.0, parameter_tolerance=-1.0, precision=-1, verbose=-1):
        self.lib.hitl_host_session_solver_options(self.s, which, max_iterations, function_tolerance, gradient_tolerance, parameter_tolerance, precision, verbose)

    @staticmethod
    def _summary(v):
        return dict(initial_cost=v[0], final_cost=v[1], successful_steps=int(v[2]), unsuccessful_steps=int(v[3]), termination=int(v[4]), num_hc_residuals=int(v[5]))

    def joint_opt_run(self, post=False):
        v = np.zeros(6)
        self._ck(self.lib.hitl_host_session_joint_opt_run(self.s, int(post), v))
        return self._summary(v)

    def solve(self, mode=0):
        v = np.zeros(6)
        self._ck(self.lib.hitl_host_session_solve(self.s, mode, v))
        return self._summary(v)

    def copy_params(self):
        self.lib.hitl_host_session_copy_params(self.s)

    def find_stf(self, min_pose=0, max_pose=None):
        c = np.zeros(3, np.uint64)
        self._ck(self.lib.hitl_host_session_find_stf(self.s, min_pose, self.n_poses - 1 if max_pose is None else max_pose, c))
        npair, nm = int(c[0]), int(c[1])
        pi, pj, off = np.zeros(max(npair, 1), np.uint32), np.zeros(max(npair, 1), np.uint32), np.zeros(npair + 1, np.uint64)
        k, idx = np.zeros(max(nm, 1), np.uint32), np.zeros(max(nm, 1), np.uint32)
        self.lib.hitl_host_session_get_stf(self.s, pi, pj, off, k, idx)
        return dict(pair_i=pi[:npair], pair_j=pj[:npair], pair_off=off, k=k[:nm], idx=idx[:nm], n_queries=int(c[2]), n_pairs=npair, n_matches=nm)

    def gradient(self):
        n, dims = C.c_uint64(), np.zeros(3, np.uint64)
        g = np.zeros(3 * self.n_poses)
        self.lib.hitl_host_session_gradient(self.s, len(g), g, C.byref(n), dims)
        return g[:n.value], tuple(int(x) for x in dims)

    def evaluate_block(self, block, with_stf=False, pose_array=None):
        """One residual block through CostFunction::Evaluate, the way Ceres calls it."""
        nres, nblk, total = C.c_int32(), C.c_int32(), C.c_uint64()
        r, j0, j1 = np.zeros(3), np.zeros(9), np.zeros(9)
        pa = np.ascontiguousarray(pose_array, np.float64).reshape(-1) if pose_array is not None else None
        self._ck(self.lib.hitl_host_session_evaluate_block(self.s, int(with_stf), block, pa.ctypes.data if pa is not None else None, C.byref(nres), C.byref(nblk),
                                                           r, j0, j1, C.byref(total)))
        k = nres.value
        return r[:k].copy(), j0[:3 * k].reshape(k, 3).copy(), (j1[:3 * k].reshape(k, 3).copy() if nblk.value == 2 else None), total.value
