// search.cu — KD-tree queries and the scan-to-scan correspondence search on sm_100a.
//
// Replaces (paths relative to HitL-SLAM/src/):
//   KDTree<float,2>::FindNearestPointNormal / FindNearestPoint / FindNeighborPoints
//       perception_tools/kdtree.cpp:141-197, :220-273, :199-218
//   JointOpt::RelativePoseTransform          human_in_the_loop_slam/JointOptimization.cpp:296-305
//   JointOpt::FindSTFCorrespondences         JointOptimization.cpp:561-642
//   JointOpt::FindVisualOdometryCorrespondences  JointOptimization.cpp:432-468
//
// Compiled with --fmad=false: every float expression below is one IEEE operation at a time, in
// the order the reference (Eigen 3) evaluates it, so index sets are bit-exact.
//
// Work decomposition of the search:
//   * a warp owns a "tile" = 32 consecutive points of one source scan i and walks the target
//     poses j in ascending order by itself (the per-point cap makes j a sequential loop per
//     point, but tiles are independent of each other);
//   * 32 target poses are culled per step, one per lane, by a world-frame AABB overlap test
//     (tile box vs scan box, both conservatively inflated) -> ballot of candidate j;
//   * per candidate j the warp forms T_ij, transforms its points, culls each point exactly
//     against scan j's robot-frame AABB inflated by thr, and only the survivors walk the tree;
//   * the tree walk is the reference's recursion unrolled into a frame machine with a
//     16-byte-per-level stack in shared memory; nodes come through the read-only L1 path;
//   * matches are appended per tile in (j, k) order by ballot/popc; three small kernels then
//     merge the tiles of each pose into the reference's (i, j, k) order, apply the
//     "> 10 matches per pair" rule and emit the CSR arrays.
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <time.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include <cub/device/device_radix_sort.cuh>
#include "hitl_internal.h"
#include "hitl_math.h"

namespace hitl {

// ------------------------------------------------------------------------------------------------
// Tree walk: FindNearestPointNormal as a frame machine.
// The recursion's pending work lives on a stack of 8-byte entries, at most two per level:
//   gate entry {gate, far child (pos | n << 16)}   far.n >= 1, so the second word is >= 0x10000
//   best entry {best, bestpos}                     bestpos <= 0xFFFF
// `gate` folds the reference's three far-side cases into one compare at return time, `gate < min(best, thr)`:
//   |s|   far side is entered only if |s| < min(best, thr)   (pushed only when |s| < thr, else it never can be)
//   -inf  s == 0: left, then right unconditionally
//   (no entry)  no far side: missing child, or |s| >= thr
// A node in the radius pushes its own distance as a best entry ABOVE its gate entry: on the way back the child's result
// is first merged with it (the node wins ties, as `child < best` in the reference), the merged value bounds the gate,
// and when the far side is entered the merged value is parked as a best entry so that it wins ties over the far result.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kNoPos = 0xFFFFu;

struct TreeRef {
  const float4* __restrict__ pm;     // px, py, bits(index | dim << 31), 0 : everything a node visit needs in one 16 B load
  const float2* __restrict__ nn;     // nx, ny : read only when the node is inside the radius
};
__device__ __forceinline__ uint32_t node_meta(const float4& nd) { return __float_as_uint(nd.z); }

// stack[level * stride + tid]
__device__ __forceinline__ void nearest_point_normal(const TreeRef t, uint32_t n_nodes, float qx, float qy, float thr,
                                                     uint2* stack, uint32_t stride, float* out_best, uint32_t* out_pos) {
  const float thr2 = fmul(thr, thr);
  const float kNegInf = __int_as_float(0xff800000);
  uint32_t pos = 0, n = n_nodes, sp = 0;
  float ret_best;
  uint32_t ret_pos;
  for (;;) {
    // ---- call(pos, n) ----
    for (;;) {
      const float4 nd = __ldg(t.pm + pos);
      float best = FLT_MAX;
      uint32_t bestpos = kNoPos;
      // (node - q)^2 == (q - node)^2 bit for bit, so one pair of differences serves the radius test, the normal
      // distance and the splitting-plane offset s
      const float dx = fsub(qx, nd.x), dy = fsub(qy, nd.y);
      bool exact = false;
      if (fadd(fmul(dx, dx), fmul(dy, dy)) < thr2) {
        const float2 nv = __ldg(t.nn + pos);
        bestpos = pos;
        best = fabsf(fadd(fmul(nv.x, dx), fmul(nv.y, dy)));
        if (best < FLT_MIN) { best = 0.0f; exact = true; }
      }
      const float s = (node_meta(nd) >> 31) ? dy : dx;
      const uint32_t nl = n >> 1, nr = n - 1 - nl;
      const bool go_right = s > 0.0f;
      const uint32_t near_n = go_right ? nr : nl, far_n = go_right ? nl : nr;
      // leaf (nl == 0 implies nr == 0), exact hit, or s > 0 with no right child (`other` stays NULL): the call returns its own node
      if (exact || near_n == 0) { ret_best = best; ret_pos = bestpos; break; }
      const uint32_t lpos = pos + 1, rpos = lpos + nl;
      const float as = fabsf(s);
      const bool both = !go_right && !(s < 0.0f);                   // s == 0: no third visit, both children in order
      if (far_n != 0 && (both || as < thr)) {
        stack[sp * stride] = make_uint2(__float_as_uint(both ? kNegInf : as), (go_right ? lpos : rpos) | (far_n << 16));
        ++sp;
      }
      if (bestpos != kNoPos) { stack[sp * stride] = make_uint2(__float_as_uint(best), bestpos); ++sp; }
      // (neither pushed = tail call: the near child's result is this call's result)
      pos = go_right ? rpos : lpos; n = near_n;
    }
    // ---- return(ret_best, ret_pos) ----
    for (;;) {
      if (sp == 0) { *out_best = ret_best; *out_pos = ret_pos; return; }
      --sp;
      const uint2 e = stack[sp * stride];
      if (e.y < 0x10000u) {                            // best entry: wins ties over what came back from below
        if (!(ret_best < __uint_as_float(e.x))) { ret_best = __uint_as_float(e.x); ret_pos = e.y; }
        continue;
      }
      const float bound = ret_best < thr ? ret_best : thr;
      if (__uint_as_float(e.x) < bound) {              // call(far); the merged result so far waits as a best entry
        stack[sp * stride] = make_uint2(__float_as_uint(ret_best), ret_pos);
        ++sp;
        pos = e.y & 0xFFFFu; n = e.y >> 16;
        break;
      }
    }
  }
}

enum { FAR_NONE = 0, FAR_COND = 1, FAR_UNCOND = 2 };
// FindNearestPoint (kdtree.cpp:220-273): Euclidean, the bound min(best, thr) is handed down.
// Only used by the consecutive-pose matcher and hitl_kd_query: explicit local stack.
__device__ void nearest_point(const TreeRef t, uint32_t n_nodes, float qx, float qy, float thr0, float* out_best,
                              uint32_t* out_pos) {
  struct Frame { float best, abs_s, thr; uint32_t far, bestpos, state; };
  Frame st[17];
  uint32_t pos = 0, n = n_nodes, level = 0;
  float thr = thr0;
  float ret_best; uint32_t ret_pos;
  for (;;) {
    for (;;) {
      const float4 nd = __ldg(t.pm + pos);
      const int dim = (int)(node_meta(nd) >> 31);
      const float ex = fsub(nd.x, qx), ey = fsub(nd.y, qy);
      float best = sqrtf(fadd(fmul(ex, ex), fmul(ey, ey)));
      uint32_t bestpos = pos;
      const float s = dim ? fsub(qy, nd.y) : fsub(qx, nd.x);
      const uint32_t nl = n >> 1, nr = n - 1 - nl;
      if (best < FLT_MIN) { ret_best = 0.0f; ret_pos = pos; break; }
      if (nl == 0) { ret_best = best; ret_pos = bestpos; break; }
      uint32_t far, state, near_pos, near_n;
      if (s < 0.0f) { near_pos = pos + 1; near_n = nl; far = (pos + 1 + nl) | (nr << 16); state = nr ? FAR_COND : FAR_NONE; }
      else if (s > 0.0f) {
        if (nr == 0) { ret_best = best; ret_pos = bestpos; break; }
        near_pos = pos + 1 + nl; near_n = nr; far = (pos + 1) | (nl << 16); state = FAR_COND;
      } else { near_pos = pos + 1; near_n = nl; far = (pos + 1 + nl) | (nr << 16); state = nr ? FAR_UNCOND : FAR_NONE; }
      Frame f; f.best = best; f.abs_s = fabsf(s); f.thr = thr; f.far = far; f.bestpos = bestpos; f.state = state;
      st[level++] = f;
      thr = best < thr ? best : thr;        // min(current_best_dist, threshold) handed to the child
      pos = near_pos; n = near_n;
    }
    for (;;) {
      if (level == 0) { *out_best = ret_best; *out_pos = ret_pos; return; }
      Frame& f = st[level - 1];
      if (ret_best < f.best) { f.best = ret_best; f.bestpos = ret_pos; }
      const float bound = f.best < f.thr ? f.best : f.thr;
      if (f.state == FAR_UNCOND || (f.state == FAR_COND && f.abs_s < bound)) {
        f.state = FAR_NONE;
        pos = f.far & 0xFFFFu; n = f.far >> 16; thr = bound;
        break;
      }
      --level;
      ret_best = f.best; ret_pos = f.bestpos;
    }
  }
}

// FindNeighborPoints (kdtree.cpp:199-218): number of nodes with |p - q| < thr (visit order is
// irrelevant for a count; the reference has no caller).
__device__ uint32_t neighbor_count(const TreeRef t, uint32_t n_nodes, float qx, float qy, float thr) {
  uint32_t stack_pos[34], stack_n[34];
  int sp = 0; uint32_t count = 0;
  stack_pos[0] = 0; stack_n[0] = n_nodes; sp = 1;
  while (sp) {
    --sp;
    const uint32_t pos = stack_pos[sp], n = stack_n[sp];
    const float4 nd = __ldg(t.pm + pos);
    const int dim = (int)(node_meta(nd) >> 31);
    const float ex = fsub(nd.x, qx), ey = fsub(nd.y, qy);
    if (sqrtf(fadd(fmul(ex, ex), fmul(ey, ey))) < thr) ++count;
    const float s = dim ? fsub(qy, nd.y) : fsub(qx, nd.x);
    const uint32_t nl = n >> 1, nr = n - 1 - nl;
    if (s > -thr && nr) { stack_pos[sp] = pos + 1 + nl; stack_n[sp] = nr; ++sp; }
    if (s < thr && nl) { stack_pos[sp] = pos + 1; stack_n[sp] = nl; ++sp; }
  }
  return count;
}

// FindNeighborPoints with its result list: the point indices of the nodes within thr, in the reference's push order (node, then the
// left subtree, then the right subtree — kdtree.cpp:204-217).  Each query owns `cap` output slots; the count is always complete.
__global__ void kd_neighbors_kernel(const float4* pm, uint32_t tree_off, uint32_t n_nodes, uint32_t nq, const float2* q, float thr, uint32_t cap,
                                    int32_t* __restrict__ index_out, uint32_t* __restrict__ count_out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  uint32_t count = 0;
  if (n_nodes) {
    const float4* t = pm + tree_off;
    const float2 p = q[i];
    uint32_t stack_pos[34], stack_n[34];
    int sp = 1;
    stack_pos[0] = 0; stack_n[0] = n_nodes;
    while (sp) {
      --sp;
      const uint32_t pos = stack_pos[sp], n = stack_n[sp];
      const float4 nd = __ldg(t + pos);
      const int dim = (int)(node_meta(nd) >> 31);
      const float ex = fsub(nd.x, p.x), ey = fsub(nd.y, p.y);
      if (sqrtf(fadd(fmul(ex, ex), fmul(ey, ey))) < thr) {
        if (count < cap) index_out[(size_t)i * cap + count] = (int32_t)(node_meta(nd) & 0x7FFFFFFFu);
        ++count;
      }
      const float s = dim ? fsub(p.y, nd.y) : fsub(p.x, nd.x);
      const uint32_t nl = n >> 1, nr = n - 1 - nl;
      if (s > -thr && nr) { stack_pos[sp] = pos + 1 + nl; stack_n[sp] = nr; ++sp; }   // popped after the left subtree
      if (s < thr && nl) { stack_pos[sp] = pos + 1; stack_n[sp] = nl; ++sp; }
    }
  }
  count_out[i] = count;
}

__global__ void kd_query_kernel(const float4* pm, const float2* nn, uint32_t tree_off, uint32_t n_nodes, uint32_t nq,
                                const float2* q, float thr, int mode, float* dist, int32_t* index) {
  extern __shared__ uint2 smem_stack[];
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  TreeRef t; t.pm = pm + tree_off; t.nn = nn + tree_off;
  if (n_nodes == 0) { if (mode != 2) dist[i] = FLT_MAX; index[i] = mode == 2 ? 0 : -1; return; }
  const float2 p = q[i];
  if (mode == 2) { index[i] = (int32_t)neighbor_count(t, n_nodes, p.x, p.y, thr); return; }
  float best; uint32_t pos;
  if (mode == 0) nearest_point_normal(t, n_nodes, p.x, p.y, thr, smem_stack + threadIdx.x, blockDim.x, &best, &pos);
  else nearest_point(t, n_nodes, p.x, p.y, thr, &best, &pos);
  dist[i] = best;
  index[i] = pos == kNoPos ? -1 : (int32_t)(node_meta(__ldg(t.pm + pos)) & 0x7FFFFFFFu);
}

// ------------------------------------------------------------------------------------------------
// Per-scan robot-frame AABB (once per hitl_set_scans). One warp per scan.
// ------------------------------------------------------------------------------------------------
__global__ void scan_aabb_kernel(const float2* __restrict__ pts, const uint32_t* __restrict__ off, uint32_t n_poses,
                                 float4* __restrict__ aabb) {
  const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n_poses) return;
  float x0 = FLT_MAX, y0 = FLT_MAX, x1 = -FLT_MAX, y1 = -FLT_MAX;
  for (uint32_t k = off[w] + lane; k < off[w + 1]; k += 32) {
    const float2 p = pts[k];
    x0 = fminf(x0, p.x); y0 = fminf(y0, p.y); x1 = fmaxf(x1, p.x); y1 = fmaxf(y1, p.y);
  }
  for (int o = 16; o; o >>= 1) {
    x0 = fminf(x0, __shfl_xor_sync(0xffffffffu, x0, o)); y0 = fminf(y0, __shfl_xor_sync(0xffffffffu, y0, o));
    x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, o)); y1 = fmaxf(y1, __shfl_xor_sync(0xffffffffu, y1, o));
  }
  if (lane == 0) aabb[w] = make_float4(x0, y0, x1, y1);
}

// ------------------------------------------------------------------------------------------------
// K0: per-pose records for one search call.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float cull_margin(float a, float b, float c, float d) {
  // Absorbs float rounding of the reference's transforms: 5 mm + 2^-18 of the largest coordinate.
  const float m = fmaxf(fmaxf(fabsf(a), fabsf(b)), fmaxf(fabsf(c), fabsf(d)));
  return 0.005f + m * 3.8146973e-6f;
}

__global__ void pose_prep_kernel(const double* __restrict__ pose, const float4* __restrict__ aabb,
                                 const uint32_t* __restrict__ off, const GridRec* __restrict__ grid, const float* __restrict__ nmax, uint32_t n_poses,
                                 PoseRec* __restrict__ rec, float4* __restrict__ src, float4* __restrict__ wbox) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_poses) return;
  const Aff2 a = pose_affine(pose[3 * i], pose[3 * i + 1], pose[3 * i + 2]);
  const Aff2 inv = affine_inverse(a);
  PoseRec r;
  r.theta = (float)pose[3 * i + 2]; r.nmax = nmax[i];
  r.i00 = inv.m00; r.i01 = inv.m01; r.i10 = inv.m10; r.i11 = inv.m11; r.itx = inv.tx; r.ity = inv.ty;
  const float4 b = aabb[i];
  r.off = off[i]; r.n = off[i + 1] - off[i];
  const GridRec g = grid[i];
  r.gx0 = g.gx0; r.gy0 = g.gy0; r.ginv = g.ginv; r.gdim = g.gdim; r.goff = g.goff; r.foff = g.foff;
  if (r.n == 0) {
    wbox[i] = make_float4(FLT_MAX, FLT_MAX, -FLT_MAX, -FLT_MAX);
  } else {
    float x[4], y[4];
    affine_apply(a, b.x, b.y, &x[0], &y[0]); affine_apply(a, b.z, b.y, &x[1], &y[1]);
    affine_apply(a, b.x, b.w, &x[2], &y[2]); affine_apply(a, b.z, b.w, &x[3], &y[3]);
    const float wx0 = fminf(fminf(x[0], x[1]), fminf(x[2], x[3])), wx1 = fmaxf(fmaxf(x[0], x[1]), fmaxf(x[2], x[3]));
    const float wy0 = fminf(fminf(y[0], y[1]), fminf(y[2], y[3])), wy1 = fmaxf(fmaxf(y[0], y[1]), fmaxf(y[2], y[3]));
    const float m = cull_margin(wx0, wy0, wx1, wy1);
    wbox[i] = make_float4(wx0 - m, wy0 - m, wx1 + m, wy1 + m);
  }
  rec[i] = r;
  src[i] = make_float4(a.m00, a.m10, a.tx, a.ty);
}

// Union of the (inflated) world boxes of every 32 consecutive poses: the first level of the target sweep.
__global__ void pose_group_box_kernel(const float4* __restrict__ wbox, uint32_t n_poses, float4* __restrict__ gbox) {
  const uint32_t g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (g * 32 >= n_poses) return;
  const uint32_t j = g * 32 + lane;
  float4 b = j < n_poses ? wbox[j] : make_float4(FLT_MAX, FLT_MAX, -FLT_MAX, -FLT_MAX);
  for (int o = 16; o; o >>= 1) {
    b.x = fminf(b.x, __shfl_xor_sync(0xffffffffu, b.x, o)); b.y = fminf(b.y, __shfl_xor_sync(0xffffffffu, b.y, o));
    b.z = fmaxf(b.z, __shfl_xor_sync(0xffffffffu, b.z, o)); b.w = fmaxf(b.w, __shfl_xor_sync(0xffffffffu, b.w, o));
  }
  if (lane == 0) gbox[g] = b;
}

// ------------------------------------------------------------------------------------------------
// Occupancy bitmaps: scan j's grid has cell size >= thr * (1 + 2^-9) and every point marks the
// 3x3 block of cells around its own cell.  A query q with an unmarked (or out-of-grid) cell has
// no node with |node - q|^2 < thr^2 as the reference evaluates it in float: |qx - px| < thr puts
// q's cell index within +-1 of p's (the 2^-9 slack dominates the rounding of the cell formula,
// which is the same instruction sequence for points and queries).  Skipping the tree walk for
// such a query is therefore exact: the walk would return FLT_MAX and change no state.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool grid_cell_dims(const float gx0, const float gy0, const float ginv, const uint32_t nxi, const uint32_t nyi, float x,
                                               float y, uint32_t* cx, uint32_t* cy) {
  const float fx = fmul(fsub(x, gx0), ginv), fy = fmul(fsub(y, gy0), ginv);
  const float nx = (float)nxi, ny = (float)nyi;
  if (!(fx >= 0.0f && fy >= 0.0f && fx < nx && fy < ny)) return false;
  *cx = (uint32_t)fx; *cy = (uint32_t)fy;
  return true;
}
__device__ __forceinline__ bool grid_cell(const float gx0, const float gy0, const float ginv, const uint32_t gdim, float x, float y,
                                          uint32_t* cx, uint32_t* cy) {
  return grid_cell_dims(gx0, gy0, ginv, gdim & 0xFFFFu, gdim >> 16, x, y, cx, cy);
}
// Second level: the same grid with kFineCells x kFineCells cells per coarse cell (same origin; 1/cell scaled by a power
// of two, so exact).  thr = kFineCells * cell / (1 + 2^-9): |qx - px| < thr puts q's fine cell index within
// +-kFineCells of p's (the 2^-9 slack, 0.0078 fine cells, dominates the rounding of the cell formula, < 1e-3 cells up to
// 16k cells per axis), and when |dx| = kFineCells the real |qx - px| exceeds (kFineCells - 1) cells, which caps |dy| at
// kFineCells - 1.  Every point marks that footprint (9 x 9 cells minus the four corners); a query with an unmarked fine
// cell has no in-radius node, exactly as for the coarse level.
__device__ __forceinline__ bool fine_cell(const float gx0, const float gy0, const float ginv, const uint32_t gdim, float x, float y,
                                          uint32_t* cx, uint32_t* cy) {
  return grid_cell_dims(gx0, gy0, fmul(ginv, (float)kFineCells), (gdim & 0xFFFFu) * kFineCells, (gdim >> 16) * kFineCells, x, y, cx, cy);
}

__global__ void occupancy_build_kernel(const float2* __restrict__ pts, const uint32_t* __restrict__ off, const uint32_t* __restrict__ tile_scan,
                                       const uint32_t* __restrict__ tile_k0, uint32_t n_tiles, const GridRec* __restrict__ grid,
                                       uint32_t* __restrict__ occ, uint32_t* __restrict__ occ_mip, const uint32_t* __restrict__ moff) {
  const uint32_t tile = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (tile >= n_tiles) return;                                  // warp-uniform
  const uint32_t i = tile_scan[tile], kl = tile_k0[tile], k = (kl & 0xFFFFu) + lane;
  const GridRec g = grid[i];
  uint32_t cx = 0, cy = 0;
  bool ok = lane < (kl >> 16) && k < off[i + 1] - off[i];
  if (ok) { const float2 p = pts[off[i] + k]; ok = grid_cell(g.gx0, g.gy0, g.ginv, g.gdim, p.x, p.y, &cx, &cy); }   // false cannot happen: the grid covers the AABB
  // consecutive beams fall into the same cell (cell 15 cm, beam spacing millimetres): one lane per distinct cell does the marking
  const uint32_t same = __match_any_sync(0xffffffffu, ok ? (cy << 16 | cx) : (0xFFFF0000u | lane));
  if (!ok || (uint32_t)(__ffs(same) - 1) != lane) return;
  const uint32_t nx = g.gdim & 0xFFFFu, ny = g.gdim >> 16;
  for (int dy = -1; dy <= 1; ++dy)
    for (int dx = -1; dx <= 1; ++dx) {
      const int x = (int)cx + dx, y = (int)cy + dy;
      if (x < 0 || y < 0 || x >= (int)nx || y >= (int)ny) continue;
      const uint32_t bit = (uint32_t)y * nx + (uint32_t)x;
      atomicOr(occ + g.goff + (bit >> 5), 1u << (bit & 31));
      if (occ_mip) {                                            // the same mark one level up: mip cell = kMipCells x kMipCells coarse cells
        const uint32_t mnx = (nx + kMipCells - 1) / kMipCells, mbit = ((uint32_t)y / kMipCells) * mnx + (uint32_t)x / kMipCells;
        atomicOr(occ_mip + moff[i] + (mbit >> 5), 1u << (mbit & 31));
      }
    }
}

__global__ void occupancy_fine_build_kernel(const float2* __restrict__ pts, const uint32_t* __restrict__ off, const uint32_t* __restrict__ tile_scan,
                                            const uint32_t* __restrict__ tile_k0, uint32_t n_tiles, const GridRec* __restrict__ grid,
                                            uint32_t* __restrict__ occ) {
  const uint32_t tile = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (tile >= n_tiles) return;                                  // warp-uniform
  const uint32_t i = tile_scan[tile], kl = tile_k0[tile], k = (kl & 0xFFFFu) + lane;
  const GridRec g = grid[i];
  if (g.foff == kNoFine) return;                                // warp-uniform (one scan per tile)
  uint32_t cx = 0, cy = 0;
  bool ok = lane < (kl >> 16) && k < off[i + 1] - off[i];
  if (ok) { const float2 p = pts[off[i] + k]; ok = fine_cell(g.gx0, g.gy0, g.ginv, g.gdim, p.x, p.y, &cx, &cy); }   // false cannot happen: the grid covers the AABB
  const unsigned long long key = ok ? ((unsigned long long)cy << 32 | cx) : (0xFFFFFFFF00000000ull | lane);
  const uint32_t same = __match_any_sync(0xffffffffu, key);     // one lane per distinct fine cell
  if (!ok || (uint32_t)(__ffs(same) - 1) != lane) return;
  const int nx = (int)((g.gdim & 0xFFFFu) * kFineCells), ny = (int)((g.gdim >> 16) * kFineCells);
  for (int dy = -kFineCells; dy <= kFineCells; ++dy) {
    const int y = (int)cy + dy;
    if (y < 0 || y >= ny) continue;
    const int r = (dy == -kFineCells || dy == kFineCells) ? kFineCells - 1 : kFineCells;
    const int xa = max((int)cx - r, 0), xb = min((int)cx + r, nx - 1);
    if (xa > xb) continue;
    const uint64_t b0 = (uint64_t)y * (uint32_t)nx + (uint32_t)xa, b1 = b0 + (uint32_t)(xb - xa);
    const uint64_t w0 = b0 >> 5, w1 = b1 >> 5;                                  // a row of <= 9 bits spans at most two words
    const uint32_t lo_mask = 0xFFFFFFFFu << (b0 & 31), hi_mask = 0xFFFFFFFFu >> (31 - (b1 & 31));
    if (w0 == w1) atomicOr(occ + g.foff + w0, lo_mask & hi_mask);
    else { atomicOr(occ + g.foff + w0, lo_mask); atomicOr(occ + g.foff + w1, hi_mask); }
  }
}

// ------------------------------------------------------------------------------------------------
// Direction occupancy: the angle gate's prefilter.  43 % of the tree walks of a dense map end with an in-radius node whose
// normal fails `nb . R n > min_cos` (points near corners and wall ends see, in most overlapping scans, only the OTHER wall):
// executed queries without a match, each a full walk.  Per coarse cell of scan j (same 3 x 3 dilation as the coarse bitmap, so
// the cell of a query covers every node within thr of it) a 16-bit mask records which of 16 direction bins (22.5 degrees)
// hold the normal of some NODE near the cell.  The gate can only pass for nodes whose normal lies within
// alpha = acos(min_cos / (|nb| |R n|)) of the rotated source normal; if no bin that intersects that window (widened by a float
// margin) is set, every in-radius node fails the gate, the query cannot produce a match whatever node the walk would return,
// and skipping the walk is exact (the query still counts as executed: that accounting does not depend on walks).
// Angles here are plain floats with kDirMargin of slack; inputs whose float angles could exceed the slack (|theta| > 256 rad,
// zero / non-finite normals) simply bypass the filter.
// ------------------------------------------------------------------------------------------------
constexpr float kDirBinsPerRad = 16.0f / 6.28318530717958647692f;
constexpr float kDirMargin = 4e-3f;                 // rad: dominates atan2f / float-angle rounding (< 1e-4 for |theta| <= 256)
constexpr float kDirMaxTheta = 256.0f;
// bins [floor(u_lo), floor(u_hi)] of a window in bin units, as a 16-bit mask
__device__ __forceinline__ uint32_t dir_window_mask(float u_lo, float u_hi) {
  const int lo = (int)floorf(u_lo), hi = (int)floorf(u_hi);
  const int cnt = hi - lo + 1;
  if (cnt >= 16) return 0xFFFFu;
  const uint32_t m = (1u << cnt) - 1u, sh = (uint32_t)lo & 15u;
  const uint32_t r = m << sh;
  return (r | (r >> 16)) & 0xFFFFu;
}

// one thread per node: marks its direction bin(s) in the 3 x 3 coarse cells around the node's cell; per-scan max |normal|
__global__ void occupancy_dir_build_kernel(const float4* __restrict__ node_pm, const float2* __restrict__ node_nn, const uint32_t* __restrict__ off,
                                           uint32_t n_poses, uint64_t n_nodes, const GridRec* __restrict__ grid, uint32_t* __restrict__ occ_dir,
                                           float* __restrict__ nmax) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_nodes) return;
  uint32_t lo = 0, hi = n_poses;
  while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (off[mid] <= t) lo = mid; else hi = mid; }
  const uint32_t i = lo;
  const GridRec g = grid[i];
  const float4 nd = node_pm[t];
  const float2 nv = node_nn[t];
  const float len = sqrtf(nv.x * nv.x + nv.y * nv.y);
  uint32_t bins;
  float bound = len;                                            // this node's contribution to the scan's bound on |normal|
  if (!(len > 0.0f) || !(len < 1e30f)) {
    // zero normal: the dot product is 0, the gate passes only for min_cos < 0 (the filter is off then); non-finite or huge: every bin
    bins = (len > 0.0f || len != len) ? 0xFFFFu : 0u;
    bound = (len != len || len >= 1e30f) ? 1e30f : 0.0f;
  } else {
    const float u = atan2f(nv.y, nv.x) * kDirBinsPerRad;      // (-8, 8]
    bins = dir_window_mask(u - kDirMargin * kDirBinsPerRad, u + kDirMargin * kDirBinsPerRad);
  }
  uint32_t cx = 0, cy = 0;
  const bool ok = bins != 0 && grid_cell(g.gx0, g.gy0, g.ginv, g.gdim, nd.x, nd.y, &cx, &cy);   // false for bins != 0 cannot happen: the grid covers the AABB
  // one atomic per (scan, cell) and per scan in this warp: lanes of the same group OR their bins / take the max bound
  const uint32_t act = __activemask();
  const uint32_t scan_grp = __match_any_sync(act, i);
  const int gmax = __reduce_max_sync(scan_grp, __float_as_int(bound));        // non-negative floats order like ints
  if ((uint32_t)(__ffs(scan_grp) - 1) == (threadIdx.x & 31u) && gmax > 0) atomicMax(reinterpret_cast<int*>(nmax + i), gmax);
  const unsigned long long key = ok ? ((unsigned long long)i << 32 | cy << 16 | cx) : (0xFFFFFFFF00000000ull | (threadIdx.x & 31u));
  const uint32_t same = __match_any_sync(act, key);
  bins = __reduce_or_sync(same, bins);
  if (!ok || (uint32_t)(__ffs(same) - 1) != (threadIdx.x & 31u)) return;
  const uint32_t nx = g.gdim & 0xFFFFu, ny = g.gdim >> 16;
  for (int dy = -1; dy <= 1; ++dy)
    for (int dx = -1; dx <= 1; ++dx) {
      const int x = (int)cx + dx, y = (int)cy + dy;
      if (x < 0 || y < 0 || x >= (int)nx || y >= (int)ny) continue;
      const uint64_t cell = (uint64_t)g.goff * 32 + (uint32_t)y * nx + (uint32_t)x;   // cells of scan i start at bit goff * 32 of the coarse bitmap
      atomicOr(occ_dir + (cell >> 1), bins << (16 * (cell & 1)));
    }
}

// ------------------------------------------------------------------------------------------------
// K1: the search.  Persistent warps; each warp pulls 32-point source tiles from a global counter.
//
// Per tile, the ascending target loop is split in two decoupled stages so that the expensive part
// (the tree walk) always runs with full lanes:
//   stage 1 (cull)  for a block of 32 target poses: lane-per-pose world-box test -> candidate
//                   mask; the candidates' records are staged in shared memory; for every candidate
//                   each lane transforms its own point and tests the target's occupancy bitmap
//                   (independent loads, pipelined) -> 32-bit "needs a walk" mask per lane;
//   stage 2 (walk)  (lane, j) items are appended to a per-warp queue in (j, lane) order; whenever
//                   32 items are queued, ALL 32 lanes take one item each (point and normal of the
//                   owning lane come by shuffle), walk tree j and evaluate both gates; results
//                   are committed in queue order: an item counts only while its owner's match
//                   count is below the cap (exactly the reference's `continue`), so walking an
//                   item speculatively past the cap is harmless and is simply dropped.
// Committed matches are appended in queue order, so a tile's records come out sorted by (j, k).
// ------------------------------------------------------------------------------------------------
struct SearchParams {
  const float2* __restrict__ pts; const float2* __restrict__ nrm;
  const float4* __restrict__ node_pm; const float2* __restrict__ node_nn;
  const PoseRec* __restrict__ rec; const float4* __restrict__ src; const float4* __restrict__ wbox; const double* __restrict__ pose;
  const float4* __restrict__ gbox; uint32_t n_groups;   // union of the world boxes of every 32 consecutive poses
  const uint32_t* __restrict__ occ; const uint32_t* __restrict__ occ_fine; const uint32_t* __restrict__ occ_dir;   // occ_dir = null: direction prefilter off
  const uint32_t* __restrict__ tile_scan; const uint32_t* __restrict__ tile_k0;
  const uint2* __restrict__ tile_j; const uint32_t* __restrict__ tile_slot;   // target range of the unit, record slot
  uint32_t tile_lo, tile_hi;          // tiles of the source shard
  uint32_t jmin, jmax;                // inclusive target range
  float thr, min_cos; int cap; uint32_t skip; uint32_t no_cull;
  const uint32_t* __restrict__ occ_mip; const uint32_t* __restrict__ moff;   // occ_mip = null: tile-box cull off
  float dir_alpha_unit;               // acos(min_cos / 1.0006): half-width of the angle gate for unit normals (host-computed)
  uint32_t* __restrict__ raw_j; uint32_t* __restrict__ raw_k; uint32_t* __restrict__ raw_idx; uint32_t* __restrict__ tile_cnt;
  unsigned long long* __restrict__ counters;   // [6] = tile ticket
  unsigned long long* __restrict__ pose_work;  // SM cycles spent on the tiles of each source pose (load-balancing feedback)
  const uint32_t* __restrict__ tile_order;     // tiles of the shard, most expensive first (from the previous call), or null
  uint32_t* __restrict__ tile_work;            // cycles / 64 per tile of this call
  uint32_t* __restrict__ tile_open;            // points of the tile that ended below the cap (they kept it alive through all its targets)
  uint32_t* __restrict__ tile_end;             // first target pose the tile did NOT need to look at any more (all its points were capped before it)
};

constexpr int kSearchThreads = 128;
constexpr int kSearchWarps = kSearchThreads / 32;
constexpr int kSearchMinBlocks = 16;   // 48 warps / SM: the walk is latency-bound, more resident warps hide the node loads
constexpr uint32_t kQueueCap = 64;
constexpr uint32_t kNotCapped = 0xFFFFFFFFu;
constexpr uint32_t kSparseMaxPoints = 24;   // coarse stage: lane-per-candidate loop over the active points when at most this many points are below the cap

struct __align__(16) WarpShared {
  uint4 rec[16][3];        // per candidate of the current half block of 16 target poses: T_ij, scan size, occupancy grid descriptor
  uint32_t cj[64];         // stage 1: compacted candidate targets (passed the world-box test), ascending j, <= 63 waiting
  uint32_t queue[kQueueCap];   // stage A: items that passed the coarse occupancy level; item = owner lane | j << 5
  uint32_t wq[kQueueCap];      // stage B: items that also passed the fine level
  uint32_t cnt[32];        // matches per lane's point
  uint32_t exec_last[32];  // j that filled the cap (kNotCapped otherwise)
};

// Tile-box vs target-scan cull (stage 1a, one lane per target pose j).  Every point of the tile lies in the robot-frame box
// [rx0, rx1] x [ry0, ry1] of source scan i; T_ij maps that box to a parallelogram whose axis-aligned hull in frame j (the four
// transformed corners, widened by a margin that dominates the float rounding of T_ij * p) contains every T_ij * p the later
// stages compute.  The coarse cell of any such point therefore lies in the hull's cell range (the cell formula is monotone in
// the coordinate), so if NO coarse cell of scan j in that range is marked — tested one level up, on the kMipCells x kMipCells
// reduction of the coarse bitmap — every per-point coarse test of this (tile, j) pair would fail: skipping the pair is exact.
__device__ __forceinline__ bool tile_box_may_touch(const SearchParams& P, uint32_t j, const Aff2& src, float bcx, float bcy, float bhx, float bhy) {
  const PoseRec rj = P.rec[j];
  if (rj.n == 0) return false;                                  // empty scan: no point can have a neighbour (the coarse stage drops it too)
  Aff2 inv; inv.m00 = rj.i00; inv.m01 = rj.i01; inv.m10 = rj.i10; inv.m11 = rj.i11; inv.tx = rj.itx; inv.ty = rj.ity;
  const Aff2 T = affine_mul(inv, src);
  // hull of the transformed box: centre +- |L| * half extents (the box is given by its centre (bcx, bcy) and half extents (bhx, bhy))
  float cx, cy;
  affine_apply(T, bcx, bcy, &cx, &cy);
  const float ex = fadd(fmul(fabsf(T.m00), bhx), fmul(fabsf(T.m01), bhy)), ey = fadd(fmul(fabsf(T.m10), bhx), fmul(fabsf(T.m11), bhy));
  const float m = cull_margin(fabsf(cx) + ex, fabsf(cy) + ey, 0.0f, 0.0f);
  const float lox = cx - ex - m, hix = cx + ex + m, loy = cy - ey - m, hiy = cy + ey + m;
  // cell range with the cell formula of grid_cell_dims (same instruction sequence: monotone in the coordinate)
  const float nxf = (float)(rj.gdim & 0xFFFFu), nyf = (float)(rj.gdim >> 16);
  const float fx0 = fmul(fsub(lox, rj.gx0), rj.ginv), fx1 = fmul(fsub(hix, rj.gx0), rj.ginv);
  const float fy0 = fmul(fsub(loy, rj.gy0), rj.ginv), fy1 = fmul(fsub(hiy, rj.gy0), rj.ginv);
  if (!(fx1 >= 0.0f && fy1 >= 0.0f && fx0 < nxf && fy0 < nyf)) return !(fx0 == fx0 && fx1 == fx1 && fy0 == fy0 && fy1 == fy1);   // the hull misses the grid: every point test fails (NaN: keep the pair)
  const uint32_t cx0 = (uint32_t)fmaxf(fx0, 0.0f), cy0 = (uint32_t)fmaxf(fy0, 0.0f);
  const uint32_t cx1 = (uint32_t)fminf(fx1, nxf - 1.0f), cy1 = (uint32_t)fminf(fy1, nyf - 1.0f);
  const uint32_t mx0 = cx0 / kMipCells, mx1 = cx1 / kMipCells, my0 = cy0 / kMipCells, my1 = cy1 / kMipCells;
  if (mx1 - mx0 >= 32u || my1 - my0 >= 12u) return true;         // an oversized tile (its points straddle a depth jump): not worth the scan
  const uint32_t mnx = ((rj.gdim & 0xFFFFu) + kMipCells - 1) / kMipCells;
  const uint32_t* __restrict__ mip = P.occ_mip + __ldg(P.moff + j);
  // one mip row at a time: its cells are consecutive bits, a span of <= 32 of them lies in at most two words
  const uint32_t span = (mx1 - mx0 == 31u) ? 0xFFFFFFFFu : ((1u << (mx1 - mx0 + 1)) - 1u);
  for (uint32_t my = my0; my <= my1; ++my) {
    const uint32_t b0 = my * mnx + mx0, w0 = b0 >> 5, sh = b0 & 31u;
    uint32_t bits = __ldg(mip + w0) >> sh;
    if (sh && ((b0 + (mx1 - mx0)) >> 5) != w0) bits |= __ldg(mip + w0 + 1) << (32u - sh);
    if (bits & span) return true;
  }
  return false;
}

template <int MINB>
__global__ void __launch_bounds__(kSearchThreads, MINB) stf_search_kernel(const SearchParams P) {
  __shared__ WarpShared s_warp[kSearchWarps];
  WarpShared& W = s_warp[threadIdx.x >> 5];
  const uint32_t lane = threadIdx.x & 31, lt = (1u << lane) - 1u;
  uint2 stack[34];                                             // walk entries (local memory, 8 B each, at most 2 per tree level)
  unsigned long long n_trav = 0, n_cand = 0;
  uint32_t n_coarse = 0, n_inrad = 0, n_gatefail = 0, n_overcap = 0, n_dirskip = 0;   // per-thread diagnostics (flushed per kernel)

  for (;;) {
    uint32_t tile = 0;
    if (lane == 0) {
      const uint32_t ticket = (uint32_t)atomicAdd(P.counters + 6, 1ull);
      // longest-processing-time-first: heavy tiles start early so the last wave is made of light ones
      tile = ticket >= P.tile_hi - P.tile_lo ? 0xFFFFFFFFu : (P.tile_order ? __ldg(P.tile_order + ticket) : P.tile_lo + ticket);
    }
    tile = __shfl_sync(0xffffffffu, tile, 0);
    if (tile == 0xFFFFFFFFu) break;
    const long long t_begin = clock64();
    const uint32_t i = P.tile_scan[tile], kl = P.tile_k0[tile], k0 = kl & 0xFFFFu;
    const uint32_t i_off = P.rec[i].off, i_n = P.rec[i].n;
    const uint32_t k = k0 + lane;
    const bool valid = lane < (kl >> 16) && k < i_n && (k % P.skip) == 0;
    float2 p = make_float2(0.f, 0.f), nv = make_float2(0.f, 0.f);
    if (valid) { p = P.pts[i_off + k]; nv = P.nrm[i_off + k]; }
    const double theta_i = P.pose[3 * i + 2];
    // direction prefilter: this lane's normal as (angle in bin units, length); usable only for finite non-zero normals and moderate pose angles
    const float nlen = sqrtf(nv.x * nv.x + nv.y * nv.y);
    const float thf_i = (float)theta_i;
    const bool dir_ok = P.occ_dir != nullptr && valid && nlen > 0.0f && nlen < 1e30f && fabsf(thf_i) <= kDirMaxTheta;
    const float nang = dir_ok ? atan2f(nv.y, nv.x) : 0.0f;
    const float4 si = P.src[i];
    Aff2 src; src.m00 = si.x; src.m01 = -si.y; src.m10 = si.y; src.m11 = si.x; src.tx = si.z; src.ty = si.w;

    // world-frame box of this tile, inflated by thr + margin (conservative pair cull)
    float bx0 = FLT_MAX, by0 = FLT_MAX, bx1 = -FLT_MAX, by1 = -FLT_MAX;
    if (valid) { float wx, wy; affine_apply(src, p.x, p.y, &wx, &wy); bx0 = bx1 = wx; by0 = by1 = wy; }
    for (int o = 16; o; o >>= 1) {
      bx0 = fminf(bx0, __shfl_xor_sync(0xffffffffu, bx0, o)); by0 = fminf(by0, __shfl_xor_sync(0xffffffffu, by0, o));
      bx1 = fmaxf(bx1, __shfl_xor_sync(0xffffffffu, bx1, o)); by1 = fmaxf(by1, __shfl_xor_sync(0xffffffffu, by1, o));
    }
    {
      const float m = P.thr + cull_margin(bx0, by0, bx1, by1);
      bx0 -= m; by0 -= m; bx1 += m; by1 += m;
    }
    // robot-frame box of the tile's points as centre + half extents (tile-box vs target-scan cull, stage 1a)
    float rx0 = FLT_MAX, ry0 = FLT_MAX, rx1 = -FLT_MAX, ry1 = -FLT_MAX;
    if (valid) { rx0 = rx1 = p.x; ry0 = ry1 = p.y; }
    for (int o = 16; o; o >>= 1) {
      rx0 = fminf(rx0, __shfl_xor_sync(0xffffffffu, rx0, o)); ry0 = fminf(ry0, __shfl_xor_sync(0xffffffffu, ry0, o));
      rx1 = fmaxf(rx1, __shfl_xor_sync(0xffffffffu, rx1, o)); ry1 = fmaxf(ry1, __shfl_xor_sync(0xffffffffu, ry1, o));
    }
    // the half extents are rounded UP so that centre +- extent contains [rx0, rx1] x [ry0, ry1] whatever the rounding of the centre
    const float bcx = 0.5f * (rx0 + rx1), bcy = 0.5f * (ry0 + ry1);
    const float bhx = fmaxf(bcx - rx0, rx1 - bcx) * 1.000001f + 1e-6f, bhy = fmaxf(bcy - ry0, ry1 - bcy) * 1.000001f + 1e-6f;
    const uint32_t out_base = P.tile_slot[tile] * (uint32_t)P.cap;   // this unit's private record region
    const uint2 tj = P.tile_j[tile];
    const bool is_unit = tj.x != 0 || tj.y != kFullRange;        // one of several target ranges of a split tile
    const uint32_t jlo = max(P.jmin, tj.x), jhi = min(P.jmax, tj.y);
    uint32_t wcount = 0;                                        // records written by the warp
    uint32_t qn = 0, qw = 0;                                    // queued items: stage A (coarse-passed), stage B (to walk)
    bool active = valid;
    W.cnt[lane] = 0; W.exec_last[lane] = kNotCapped;
    __syncwarp();

    // Stage A -> B: all lanes take one coarse-passed item each, form T_ij and the query point in frame j and test the
    // FINE occupancy level; survivors are appended to the walk queue in the same (j, lane) order.
    auto filter = [&](uint32_t nitems) {
      const bool item = lane < nitems;
      const uint32_t it = item ? W.queue[lane] : 0u;
      const uint32_t o = it & 31u, j = it >> 5;
      const float opx = __shfl_sync(0xffffffffu, p.x, o), opy = __shfl_sync(0xffffffffu, p.y, o);
      const float o_nang = __shfl_sync(0xffffffffu, nang, o), o_nlen = __shfl_sync(0xffffffffu, nlen, o);
      const bool o_dir = __shfl_sync(0xffffffffu, (int)dir_ok, o) != 0;
      bool pass = false; float qx = 0.f, qy = 0.f;
      if (item && W.cnt[o] < (uint32_t)P.cap) {
        ++n_coarse;
        const PoseRec rj = P.rec[j];
        Aff2 inv; inv.m00 = rj.i00; inv.m01 = rj.i01; inv.m10 = rj.i10; inv.m11 = rj.i11; inv.tx = rj.itx; inv.ty = rj.ity;
        const Aff2 T = affine_mul(inv, src);                    // T_ij = target^-1 * source (JointOptimization.cpp:304)
        affine_apply(T, opx, opy, &qx, &qy);
        pass = true;
        if (!P.no_cull && rj.foff != kNoFine) {
          uint32_t cx, cy;
          pass = fine_cell(rj.gx0, rj.gy0, rj.ginv, rj.gdim, qx, qy, &cx, &cy);
          if (pass) {
            const uint64_t bit = (uint64_t)cy * ((rj.gdim & 0xFFFFu) * kFineCells) + cx;
            pass = (__ldg(P.occ_fine + rj.foff + (bit >> 5)) >> (bit & 31)) & 1u;
          }
        }
        if (pass && o_dir && !P.no_cull && fabsf(rj.theta) <= kDirMaxTheta) {
          // Angle gate prefilter (see "Direction occupancy" above).  The gate rotates the source normal by (float)(theta_j - theta_i)
          // (JointOptimization.cpp:604-606); in float angles that is o_nang + (theta_j - theta_i) up to < 1e-4 rad.
          const float bound = fmul(fmul(rj.nmax, o_nlen), 1.0002f);        // >= |nb| |R n| for every node of scan j
          const float ca = P.min_cos / bound;
          if (ca < 1.0f) {                                                    // else no node can pass at all; leave that to the walk
            const float alpha = (bound <= 1.0005f ? P.dir_alpha_unit : acosf(fmaxf(ca, 0.0f))) + kDirMargin;
            const float u = (o_nang + (rj.theta - thf_i)) * kDirBinsPerRad, a16 = alpha * kDirBinsPerRad;
            uint32_t ccx, ccy;
            if (grid_cell(rj.gx0, rj.gy0, rj.ginv, rj.gdim, qx, qy, &ccx, &ccy)) {
              const uint64_t cell = (uint64_t)rj.goff * 32 + ccy * (rj.gdim & 0xFFFFu) + ccx;
              const uint32_t have = (__ldg(P.occ_dir + (cell >> 1)) >> (16 * (cell & 1))) & 0xFFFFu;
              if ((have & dir_window_mask(u - a16, u + a16)) == 0) { pass = false; ++n_dirskip; }
            }
          }
        }
      }
      const uint32_t pm = __ballot_sync(0xffffffffu, pass);
      if (pass) W.wq[qw + __popc(pm & lt)] = it;
      qw += __popc(pm);
      // shift the remaining stage-A items to the front
      const uint32_t rest = qn - nitems;
      uint32_t a = 0, b = 0;
      if (lane < rest) a = W.queue[nitems + lane];
      if (lane + 32 < rest) b = W.queue[nitems + lane + 32];
      __syncwarp();
      if (lane < rest) W.queue[lane] = a;
      if (lane + 32 < rest) W.queue[lane + 32] = b;
      qn = rest;
      __syncwarp();
    };

    // Walks one batch of up to 32 stage-B items (all lanes busy) and commits it in queue order.
    auto drain = [&](uint32_t nitems) {
      const bool item = lane < nitems;
      const uint32_t it = item ? W.wq[lane] : 0u;
      const uint32_t o = it & 31u, j = it >> 5;
      // point / normal of the owning lane
      const float opx = __shfl_sync(0xffffffffu, p.x, o), opy = __shfl_sync(0xffffffffu, p.y, o);
      const float onx = __shfl_sync(0xffffffffu, nv.x, o), ony = __shfl_sync(0xffffffffu, nv.y, o);
      bool ok = false; uint32_t tgt = 0;
      const uint32_t cnt_o = W.cnt[o];
      if (item && cnt_o < (uint32_t)P.cap) {
        const float4 ra = *reinterpret_cast<const float4*>(&P.rec[j].i00);
        const float4 rb = *reinterpret_cast<const float4*>(&P.rec[j].itx);   // itx, ity, scan offset, size
        Aff2 inv; inv.m00 = ra.x; inv.m01 = ra.y; inv.m10 = ra.z; inv.m11 = ra.w; inv.tx = rb.x; inv.ty = rb.y;
        const Aff2 T = affine_mul(inv, src);                    // T_ij = target^-1 * source (JointOptimization.cpp:304)
        float qx, qy;
        affine_apply(T, opx, opy, &qx, &qy);
        const uint32_t toff = __float_as_uint(rb.z), tn = __float_as_uint(rb.w);
        TreeRef t; t.pm = P.node_pm + toff; t.nn = P.node_nn + toff;
        float best; uint32_t bpos;
        nearest_point_normal(t, tn, qx, qy, P.thr, stack, 1, &best, &bpos);
        ++n_trav;
        if (best < P.thr) {                                     // implies bpos valid: best < FLT_MAX only via an in-radius node
          ++n_inrad;
          // Rotation2Df(theta_j - theta_i) * normal  (JointOptimization.cpp:604-606)
          const float dth = (float)(P.pose[3 * j + 2] - theta_i);
          const float sn = sinf_rn(dth), cs = cosf_rn(dth);
          float rnx, rny; rot_apply(cs, sn, onx, ony, &rnx, &rny);
          const float2 nb = __ldg(t.nn + bpos);
          ok = fadd(fmul(nb.x, rnx), fmul(nb.y, rny)) > P.min_cos;
          tgt = node_meta(__ldg(t.pm + bpos)) & 0x7FFFFFFFu;
          n_gatefail += ok ? 0u : 1u;
        }
      }
      // commit in queue order: an item counts only while its owner is below the cap
      const uint32_t same = __match_any_sync(0xffffffffu, item ? o : (32u + lane));
      const uint32_t okmask = __ballot_sync(0xffffffffu, ok);
      const uint32_t earlier = __popc(same & okmask & lt);
      const bool commit = ok && (cnt_o + earlier < (uint32_t)P.cap);
      n_overcap += (ok && !commit) ? 1u : 0u;
      const uint32_t cm = __ballot_sync(0xffffffffu, commit);
      if (commit) {
        const uint32_t dst = out_base + wcount + __popc(cm & lt);
        P.raw_j[dst] = j; P.raw_k[dst] = k0 + o; P.raw_idx[dst] = tgt;
        atomicAdd(&W.cnt[o], 1u);
        if (cnt_o + earlier + 1 == (uint32_t)P.cap) W.exec_last[o] = j;
      }
      wcount += __popc(cm);
      __syncwarp();
      // shift the remaining items to the front
      const uint32_t rest = qw - nitems;
      uint32_t a = 0;
      if (lane < rest) a = W.wq[nitems + lane];
      __syncwarp();
      if (lane < rest) W.wq[lane] = a;
      qw = rest;
      active = valid && W.cnt[lane] < (uint32_t)P.cap;
      __syncwarp();
    };

    uint32_t j_reached = jlo;                                   // sweep position when the tile ended (scheduling feedback only)
    if (jlo <= jhi && __any_sync(0xffffffffu, active)) {
      bool all_done = false;
      uint32_t nc = 0;                                          // candidates waiting in W.cj (ascending j)

      // ---- stage 1b: the next m <= 32 candidates (one per lane / slot), coarse occupancy level, queueing, walks ----
      // Candidates come compacted: whatever the density of overlapping targets along the trajectory, every pass over this
      // stage works on a full set of 32 (the tail of a tile excepted).
      auto process = [&](uint32_t m) {
        const uint32_t slots = m >= 32 ? 0xFFFFFFFFu : ((1u << m) - 1u);
        const bool have = lane < m;
        const uint32_t jl = have ? W.cj[lane] : 0u;
        const uint32_t act_mask = __ballot_sync(0xffffffffu, active);
        uint32_t need = 0, todo = slots;
        if (!P.no_cull && __popc(act_mask) <= (int)kSparseMaxPoints) {
          // sparse form: few points of the tile are still below the cap (stragglers keep a tile alive through all its
          // targets), so the loop runs over the active POINTS with one candidate target per lane: each lane keeps its own
          // T_ij and tests the owner's point against its own target's coarse bitmap.
          Aff2 T; uint32_t goff = 0, gdim = 0; float gx0 = 0.f, gy0 = 0.f, ginv = 0.f;
          bool live = false;
          if (have) {
            const PoseRec rj = P.rec[jl];
            Aff2 inv; inv.m00 = rj.i00; inv.m01 = rj.i01; inv.m10 = rj.i10; inv.m11 = rj.i11; inv.tx = rj.itx; inv.ty = rj.ity;
            T = affine_mul(inv, src);                             // T_ij = target^-1 * source (JointOptimization.cpp:304)
            goff = rj.goff; gdim = rj.gdim; gx0 = rj.gx0; gy0 = rj.gy0; ginv = rj.ginv;
            live = rj.n != 0;
          }
          for (uint32_t am = act_mask; am; am &= am - 1) {
            const uint32_t o = __ffs(am) - 1;
            const float opx = __shfl_sync(0xffffffffu, p.x, o), opy = __shfl_sync(0xffffffffu, p.y, o);
            bool in = live;
            if (in) {
              float qx, qy;
              affine_apply(T, opx, opy, &qx, &qy);
              uint32_t cx, cy;
              in = grid_cell(gx0, gy0, ginv, gdim, qx, qy, &cx, &cy);
              if (in) {
                const uint32_t bit = cy * (gdim & 0xFFFFu) + cx;
                in = (__ldg(P.occ + goff + (bit >> 5)) >> (bit & 31)) & 1u;
              }
            }
            const uint32_t mm = __ballot_sync(0xffffffffu, in);   // bit c = candidate slot c passes for owner o
            if (lane == o) need = mm;
          }
          todo = __reduce_or_sync(0xffffffffu, need);
        } else {
          // dense form: the candidates' T_ij and grid descriptors are staged in shared memory, 16 at a time, and every lane
          // tests its own point against each
          for (uint32_t half = 0; half < 2; ++half) {
            const uint32_t hm = slots & (0xFFFFu << (16 * half));
            if (hm == 0) continue;
            if (have && (lane >> 4) == half) {
              const PoseRec rj = P.rec[jl];
              Aff2 inv; inv.m00 = rj.i00; inv.m01 = rj.i01; inv.m10 = rj.i10; inv.m11 = rj.i11; inv.tx = rj.itx; inv.ty = rj.ity;
              const Aff2 T = affine_mul(inv, src);
              uint4* const r = W.rec[lane & 15];
              r[0] = make_uint4(__float_as_uint(T.m00), __float_as_uint(T.m01), __float_as_uint(T.m10), __float_as_uint(T.m11));
              r[1] = make_uint4(__float_as_uint(T.tx), __float_as_uint(T.ty), rj.n, rj.goff);
              r[2] = make_uint4(__float_as_uint(rj.gx0), __float_as_uint(rj.gy0), __float_as_uint(rj.ginv), rj.gdim);
            }
            __syncwarp();
            for (uint32_t cm = hm; cm; cm &= cm - 1) {
              const uint32_t c = __ffs(cm) - 1;
              const uint4 r0 = W.rec[c & 15][0], r1 = W.rec[c & 15][1];
              bool in = active && r1.z != 0;                        // r1.z = n
              if (in && !P.no_cull) {
                Aff2 T; T.m00 = __uint_as_float(r0.x); T.m01 = __uint_as_float(r0.y); T.m10 = __uint_as_float(r0.z); T.m11 = __uint_as_float(r0.w);
                T.tx = __uint_as_float(r1.x); T.ty = __uint_as_float(r1.y);
                float qx, qy;
                affine_apply(T, p.x, p.y, &qx, &qy);
                const uint4 r2 = W.rec[c & 15][2];
                uint32_t cx, cy;
                in = grid_cell(__uint_as_float(r2.x), __uint_as_float(r2.y), __uint_as_float(r2.z), r2.w, qx, qy, &cx, &cy);
                if (in) {
                  const uint32_t bit = cy * (r2.w & 0xFFFFu) + cx;
                  in = (__ldg(P.occ + r1.w + (bit >> 5)) >> (bit & 31)) & 1u;
                }
              }
              need |= (uint32_t)in << c;
            }
            __syncwarp();
          }
        }
        // ---- stage 2: queue (j, lane) items in order; walk whenever a full batch is available ----
        for (uint32_t cm = todo & slots; cm && !all_done; cm &= cm - 1) {
          const uint32_t c = __ffs(cm) - 1;
          const bool want = ((need >> c) & 1u) && active;
          const uint32_t mm = __ballot_sync(0xffffffffu, want);
          if (mm == 0) continue;
          if (want) W.queue[qn + __popc(mm & lt)] = lane | (W.cj[c] << 5);
          qn += __popc(mm);
          __syncwarp();
          if (qn >= 32) {
            filter(32);
            if (qw >= 32) {
              drain(32);
              if (!__any_sync(0xffffffffu, active)) all_done = true;
            }
          }
        }
        // drop the consumed candidates
        const uint32_t rest = nc - m;
        uint32_t keep = 0;
        if (lane < rest) keep = W.cj[m + lane];
        __syncwarp();
        if (lane < rest) W.cj[lane] = keep;
        nc = rest;
        __syncwarp();
      };

      // ---- stage 1a': GROUPS of 32 consecutive target poses, one group per lane: the union of their world boxes (consecutive
      // poses of a trajectory are neighbours in space) against the tile box.  Only blocks of groups that overlap are looked at
      // pose by pose, so the sweep over the targets costs N / 1024 steps plus the overlapping part, not N / 32 (100 k-pose maps). ----
      uint32_t gmask = 0, gbase = 0xFFFFFFFFu;
      for (uint32_t jb = jlo & ~31u; jb <= jhi && !all_done; jb += 32) {
        j_reached = jb + 32;
        if (!P.no_cull) {
          const uint32_t grp = jb >> 5;
          if ((grp >> 5) != gbase) {
            gbase = grp >> 5;
            const uint32_t gl = (gbase << 5) + lane;
            bool ghit = false;
            if (gl < P.n_groups) { const float4 gb = __ldg(P.gbox + gl); ghit = !(gb.x > bx1 || gb.z < bx0 || gb.y > by1 || gb.w < by0); }
            gmask = __ballot_sync(0xffffffffu, ghit);
          }
          const uint32_t rest = gmask >> (grp & 31u);
          if (rest == 0) { jb = (((gbase + 1) << 10)) - 32; continue; }     // no overlapping group left in this span of 1024 poses
          if (!(rest & 1u)) { jb += ((__ffs(rest) - 1) << 5) - 32; continue; }   // jump to the next overlapping group
        }
        // ---- stage 1a: world-box test, one target pose per lane; survivors are appended to the candidate list ----
        const uint32_t jl = jb + lane;
        bool hit = jl >= jlo && jl <= jhi && jl != i;
        if (hit && !P.no_cull) {
          const float4 wb = __ldg(P.wbox + jl);
          hit = !(wb.x > bx1 || wb.z < bx0 || wb.y > by1 || wb.w < by0);
          if (hit && P.occ_mip) hit = tile_box_may_touch(P, jl, src, bcx, bcy, bhx, bhy);
        }
        const uint32_t cand = __ballot_sync(0xffffffffu, hit);
        if (cand == 0) continue;
        n_cand += __popc(cand);
        if (hit) W.cj[nc + __popc(cand & lt)] = jl;
        nc += __popc(cand);
        __syncwarp();
        if (nc >= 32) process(32);
      }
      if (nc && !all_done) process(nc);
      while ((qn || qw) && !all_done) {
        if (qn && qw < 32) { filter(qn < 32 ? qn : 32); if ((qn && qw < 32) || qw == 0) continue; }
        drain(qw < 32 ? qw : 32);
        if (!__any_sync(0xffffffffu, active)) all_done = true;
      }
    }
    const uint32_t open_pts = __popc(__ballot_sync(0xffffffffu, valid && W.cnt[lane] < (uint32_t)P.cap));
    if (lane == 0) { P.tile_cnt[tile] = wcount; P.tile_open[tile] = open_pts; P.tile_end[tile] = j_reached; }
    // queries the reference semantics execute for this point: every j != i in range up to and
    // including the one that filled the cap (JointOptimization.cpp:597-600)
    unsigned long long exec = 0;
    if (valid && !is_unit) {                                    // split tiles: stf_split_merge_kernel counts after re-applying the cap
      const bool i_in_range = i >= P.jmin && i <= P.jmax;
      const uint32_t el = W.exec_last[lane];
      if (el == kNotCapped) exec = (unsigned long long)(P.jmax - P.jmin + 1) - (i_in_range ? 1 : 0);
      else exec = (unsigned long long)(el - P.jmin + 1) - ((i_in_range && i <= el) ? 1 : 0);
    }
    for (int o = 16; o; o >>= 1) exec += __shfl_xor_sync(0xffffffffu, exec, o);
    if (lane == 0) {
      atomicAdd(P.counters + 0, exec); atomicAdd(P.counters + 2, (unsigned long long)wcount);
      const unsigned long long dt = (unsigned long long)(clock64() - t_begin);
      atomicAdd(P.pose_work + i, dt);
      P.tile_work[tile] = (uint32_t)min(dt >> 6, 0xFFFFFFFFull);
      atomicAdd(P.counters + 7, dt >> 6);
      atomicMax(P.counters + 8, dt >> 6);
    }
    __syncwarp();
  }
  unsigned long long n_co = n_coarse, n_ir = n_inrad, n_gf = n_gatefail, n_oc = n_overcap, n_ds = n_dirskip;
  for (int o = 16; o; o >>= 1) {
    n_trav += __shfl_xor_sync(0xffffffffu, n_trav, o);
    n_co += __shfl_xor_sync(0xffffffffu, n_co, o); n_ir += __shfl_xor_sync(0xffffffffu, n_ir, o);
    n_gf += __shfl_xor_sync(0xffffffffu, n_gf, o); n_oc += __shfl_xor_sync(0xffffffffu, n_oc, o); n_ds += __shfl_xor_sync(0xffffffffu, n_ds, o);
  }
  if (lane == 0) {
    atomicAdd(P.counters + 1, n_trav); atomicAdd(P.counters + 5, n_cand); atomicAdd(P.counters + 9, n_co); atomicAdd(P.counters + 10, n_ir);
    atomicAdd(P.counters + 11, n_gf); atomicAdd(P.counters + 12, n_oc); atomicAdd(P.counters + 13, n_ds);
  }
}

// ------------------------------------------------------------------------------------------------
// K1b': tiles split along the target axis.  Every unit searched its own target range with a private cap state, so a
// point may hold up to `cap` matches PER UNIT.  The reference's sequential loop keeps the first `cap` matches of a
// point in ascending target order and stops there: concatenating the units in range order and truncating every
// point at `cap` gives exactly that (a unit's own first `cap` contain every match that can survive).  One warp per
// group walks its units in order, drops the surplus records in place (lists stay sorted by (j, k)), rewrites the
// unit counts, and accounts the queries the reference semantics execute (up to the target that filled the cap).
// ------------------------------------------------------------------------------------------------
struct MergeParams {
  const uint2* __restrict__ groups; uint32_t n_groups, tile_lo, tile_hi;
  const uint32_t* __restrict__ tile_scan; const uint32_t* __restrict__ tile_k0; const uint32_t* __restrict__ tile_slot;
  const uint32_t* __restrict__ off;
  uint32_t* raw_j; uint32_t* raw_k; uint32_t* raw_idx; uint32_t* tile_cnt;
  uint32_t jmin, jmax, skip; int cap;
  unsigned long long* counters;
};

__global__ void __launch_bounds__(128) stf_split_merge_kernel(const MergeParams P) {
  __shared__ uint32_t s_cnt[4][32], s_last[4][32];
  const uint32_t w = threadIdx.x >> 5, lane = threadIdx.x & 31, lt = (1u << lane) - 1u;
  const uint32_t g = blockIdx.x * 4 + w;
  if (g >= P.n_groups) return;
  const uint2 grp = P.groups[g];
  if (grp.x < P.tile_lo || grp.x >= P.tile_hi) return;            // group of another shard
  const uint32_t i = P.tile_scan[grp.x], kl = P.tile_k0[grp.x], k0 = kl & 0xFFFFu, len = kl >> 16;
  s_cnt[w][lane] = 0; s_last[w][lane] = kNotCapped;
  __syncwarp();
  unsigned long long dropped = 0;
  for (uint32_t u = 0; u < grp.y; ++u) {
    const uint32_t t = grp.x + u, c = P.tile_cnt[t], base = P.tile_slot[t] * (uint32_t)P.cap;
    uint32_t wr = 0;
    for (uint32_t r0 = 0; r0 < c; r0 += 32) {
      const uint32_t r = r0 + lane;
      const bool have = r < c;
      uint32_t j = 0, k = 0, idx = 0;
      if (have) { j = P.raw_j[base + r]; k = P.raw_k[base + r]; idx = P.raw_idx[base + r]; }
      const uint32_t o = have ? k - k0 : 32u + lane;
      const uint32_t same = __match_any_sync(0xffffffffu, o);
      const uint32_t earlier = __popc(same & lt);
      const uint32_t prior = have ? s_cnt[w][o] : 0u;
      const bool keep = have && prior + earlier < (uint32_t)P.cap;
      const uint32_t km = __ballot_sync(0xffffffffu, keep);
      __syncwarp();
      if (keep) {
        const uint32_t dst = base + wr + __popc(km & lt);           // dst <= base + r: in-place compaction
        P.raw_j[dst] = j; P.raw_k[dst] = k; P.raw_idx[dst] = idx;
        atomicAdd(&s_cnt[w][o], 1u);
        if (prior + earlier + 1 == (uint32_t)P.cap) s_last[w][o] = j;
      }
      wr += __popc(km);
      __syncwarp();
    }
    dropped += c - wr;
    if (lane == 0) P.tile_cnt[t] = wr;
  }
  // queries the reference semantics execute for each point of the tile (as in the search kernel for unsplit tiles)
  const uint32_t i_n = P.off[i + 1] - P.off[i], k = k0 + lane;
  unsigned long long exec = 0;
  if (lane < len && k < i_n && (k % P.skip) == 0) {
    const bool i_in_range = i >= P.jmin && i <= P.jmax;
    const uint32_t el = s_last[w][lane];
    if (el == kNotCapped) exec = (unsigned long long)(P.jmax - P.jmin + 1) - (i_in_range ? 1 : 0);
    else exec = (unsigned long long)(el - P.jmin + 1) - ((i_in_range && i <= el) ? 1 : 0);
  }
  for (int o = 16; o; o >>= 1) exec += __shfl_xor_sync(0xffffffffu, exec, o);
  if (lane == 0) { atomicAdd(P.counters + 0, exec); atomicAdd(P.counters + 2, 0ull - dropped); }
}

// ------------------------------------------------------------------------------------------------
// K1c: order the per-tile records of each source pose by (j, k), drop pairs with <= min_corr
// matches, emit CSR.  Each CTA owns a private u32[n_poses] scratch slice indexed by j.
//   pass 0 (count): per pose -> kept matches / kept pairs
//   (host-launched scan over poses)
//   pass 1 (fill):  per pose -> pair_i, pair_j, pair_off, k, idx at the scanned offsets
// ------------------------------------------------------------------------------------------------
constexpr int kOrderThreads = 512;   // the record sweeps are chains of dependent loads per thread: many threads, few iterations each
constexpr uint32_t kDropped = 0xFFFFFFFFu;
constexpr uint32_t kMaxTileRecords = 32 * 64;   // a tile holds <= 32 points x cap (<= 64) records
constexpr uint32_t kOrderSlots = 1024;          // kept pairs of one source pose whose fill position lives in shared memory
constexpr uint32_t kOrderTileBatch = 256;       // tile descriptors fetched at once
constexpr uint32_t kOrderMaskWords = 6144;      // shared bitmap of the fast placement path (pairs x ceil(points / 32) words)
constexpr uint32_t kSlotFlag = 0x80000000u;     // cntj[j] = kSlotFlag | slot (offsets inside a pose's segment stay below 2^31)

struct OrderParams {
  const uint32_t* __restrict__ raw_j; const uint32_t* __restrict__ raw_k; const uint32_t* __restrict__ raw_idx;
  const uint32_t* __restrict__ tile_cnt; const uint32_t* __restrict__ tile_begin; const uint32_t* __restrict__ tile_slot; const uint32_t* __restrict__ off;
  uint32_t src_lo, src_hi, n_poses; int cap; uint32_t min_corr;
  uint32_t* scratch;                  // gridDim.x * n_poses, zero on entry and on exit
  unsigned long long* pose_cnt;       // 2 per source pose of the shard: matches, pairs (pass 1: exclusive offsets)
  // single-pass variant (PASS == 2): poses are handed out by ticket in ascending order; the (matches, pairs) offsets of a pose come from a
  // decoupled look-back over the states its predecessors publish: 4 words per pose = {aggregate matches | flag, aggregate pairs,
  // inclusive-prefix matches | flag, inclusive-prefix pairs}, zero on entry
  unsigned long long* state; unsigned long long* ticket; unsigned long long* counters;
  uint32_t* pair_i; uint32_t* pair_j; unsigned long long* pair_off; uint32_t* out_k; uint32_t* out_idx;
};

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total, uint32_t* sm /* 33 words */) {
  const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t x = v;
  for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= (uint32_t)o) x += y; }
  if (lane == 31) sm[w] = x;
  __syncthreads();
  if (w == 0) {
    uint32_t s = lane < (blockDim.x >> 5) ? sm[lane] : 0;
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= (uint32_t)o) s += y; }
    sm[lane] = s;
  }
  __syncthreads();
  const uint32_t base = w ? sm[w - 1] : 0;
  *total = sm[(blockDim.x >> 5) - 1];
  __syncthreads();
  return base + x - v;
}

// Exclusive scan of a (matches, pairs) couple packed as matches | pairs << 40 (matches per pose < 2^40).
__device__ __forceinline__ unsigned long long block_exclusive_scan64(unsigned long long v, unsigned long long* total, unsigned long long* sm /* 33 words */) {
  const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  unsigned long long x = v;
  for (int o = 1; o < 32; o <<= 1) { const unsigned long long y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= (uint32_t)o) x += y; }
  if (lane == 31) sm[w] = x;
  __syncthreads();
  unsigned long long base = 0, tot = 0;
  for (uint32_t q = 0; q < (blockDim.x >> 5); ++q) { const unsigned long long t = sm[q]; if (q < w) base += t; tot += t; }
  *total = tot;
  __syncthreads();
  return base + x - v;
}

template <int PASS>
__global__ void __launch_bounds__(kOrderThreads) stf_order_kernel(const OrderParams P) {
  __shared__ unsigned long long sm[kOrderThreads / 32 + 1];
  __shared__ uint32_t s_jlo, s_jhi;
  __shared__ unsigned long long s_off_m, s_off_p;
  __shared__ uint32_t s_hcnt[kOrderTileBatch], s_hbase[kOrderTileBatch], s_htotal;   // tile descriptors of the pose (prefix of counts, record bases)
  __shared__ uint32_t s_next[PASS >= 1 ? kOrderSlots : 1];      // next free slot of the pose's first kOrderSlots kept pairs
  __shared__ uint32_t s_tcnt[PASS >= 1 ? kOrderTileBatch : 1], s_tbase[PASS >= 1 ? kOrderTileBatch : 1];
  __shared__ uint32_t s_mask[PASS >= 1 ? kOrderMaskWords : 1];  // fast path: one bit row per kept pair over the pose's points
  uint32_t* const s_j = s_mask;                                 // fallback path: the j column of the tile being placed (<= kMaxTileRecords)
  __shared__ uint16_t s_pref[PASS >= 1 ? kOrderMaskWords : 1];  // ... and the popcount prefix of each row word
  __shared__ uint32_t s_total;
  uint32_t* const cntj = P.scratch + (size_t)blockIdx.x * P.n_poses;
  __shared__ uint32_t s_ticket;
  for (uint32_t it = 0;; ++it) {
    uint32_t i;
    if (PASS == 2) {                                            // ascending tickets: a pose only ever waits for poses that are already being processed
      if (threadIdx.x == 0) s_ticket = (uint32_t)atomicAdd(P.ticket, 1ull);
      __syncthreads();
      i = P.src_lo + s_ticket;
    } else {
      i = P.src_lo + blockIdx.x + it * gridDim.x;
    }
    if (i >= P.src_hi) break;
    const uint32_t tb = P.tile_begin[i], te = P.tile_begin[i + 1];
    if (threadIdx.x == 0) { s_jlo = 0xFFFFFFFFu; s_jhi = 0; }
    __syncthreads();
    // -- histogram over j.  Tile descriptors are fetched in one sweep and the records are addressed by a flat index, so all
    //    loads of a pose are independent of each other (no per-tile chain of dependent round trips). --
    uint32_t jlo = 0xFFFFFFFFu, jhi = 0;
    const uint32_t ntile_h = te - tb;
    if (ntile_h <= kOrderTileBatch) {
      for (uint32_t q = threadIdx.x; q < ntile_h; q += kOrderThreads) { s_hcnt[q] = P.tile_cnt[tb + q]; s_hbase[q] = P.tile_slot[tb + q] * (uint32_t)P.cap; }
      __syncthreads();
      if (threadIdx.x == 0) { uint32_t acc = 0; for (uint32_t q = 0; q < ntile_h; ++q) { const uint32_t c = s_hcnt[q]; s_hcnt[q] = acc; acc += c; } s_htotal = acc; }
      __syncthreads();
      const uint32_t R = s_htotal;
      for (uint32_t g = threadIdx.x; g < R; g += kOrderThreads) {
        uint32_t lo = 0, hi = ntile_h;
        while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (s_hcnt[mid] <= g) lo = mid; else hi = mid; }
        const uint32_t j = P.raw_j[s_hbase[lo] + (g - s_hcnt[lo])];
        atomicAdd(&cntj[j], 1u);
        jlo = min(jlo, j); jhi = max(jhi, j);
      }
    } else {
      for (uint32_t t = tb; t < te; ++t) {
        const uint32_t c = P.tile_cnt[t], base = P.tile_slot[t] * (uint32_t)P.cap;
        for (uint32_t u = threadIdx.x; u < c; u += kOrderThreads) {
          const uint32_t j = P.raw_j[base + u];
          atomicAdd(&cntj[j], 1u);
          jlo = min(jlo, j); jhi = max(jhi, j);
        }
      }
    }
    if (jlo != 0xFFFFFFFFu) { atomicMin(&s_jlo, jlo); atomicMax(&s_jhi, jhi); }
    __syncthreads();
    jlo = s_jlo; jhi = s_jhi;
    unsigned long long tot_m = 0, tot_p = 0;
    if (PASS == 1 && threadIdx.x == 0) { s_off_m = P.pose_cnt[2 * (i - P.src_lo)]; s_off_p = P.pose_cnt[2 * (i - P.src_lo) + 1]; }
    __syncthreads();
    if (PASS == 2) {
      // -- totals of this pose (the counting scan), then its offsets by decoupled look-back over the preceding poses --
      uint32_t run_m = 0, run_p = 0;
      if (jlo != 0xFFFFFFFFu)
        for (uint32_t j0 = jlo; j0 <= jhi; j0 += 4 * kOrderThreads) {
          uint32_t lm = 0, lp = 0;
          for (int q = 0; q < 4; ++q) {
            const uint32_t j = j0 + 4 * threadIdx.x + q;
            const uint32_t c = j <= jhi ? cntj[j] : 0;
            if (c > P.min_corr) { lm += c; lp += 1u; }
          }
          unsigned long long tot2;
          block_exclusive_scan64((unsigned long long)lm | ((unsigned long long)lp << 40), &tot2, sm);
          run_m += (uint32_t)(tot2 & 0xFFFFFFFFFFull); run_p += (uint32_t)(tot2 >> 40);
        }
      if (threadIdx.x < 32) {
        constexpr unsigned long long kFlag = 1ull << 63;
        const uint32_t lane = threadIdx.x, q = i - P.src_lo;
        unsigned long long* const st = P.state;
        if (lane == 0) { st[4 * (size_t)q + 1] = run_p; __threadfence(); atomicExch(&st[4 * (size_t)q], kFlag | run_m); }
        unsigned long long pm = 0, pp = 0;
        long long base = (long long)q - 1;
        bool done = q == 0;
        while (!done) {
          const long long c = base - lane;
          unsigned long long vm = 0, vp = 0;
          bool is_prefix = c < 0;                                 // before the first pose: an empty inclusive prefix
          if (c >= 0) {
            for (;;) {
              const unsigned long long v2 = atomicAdd(&st[4 * (size_t)c + 2], 0ull);
              if (v2 & kFlag) { __threadfence(); vp = *(volatile unsigned long long*)&st[4 * (size_t)c + 3]; vm = v2 & ~kFlag; is_prefix = true; break; }
              const unsigned long long v0 = atomicAdd(&st[4 * (size_t)c], 0ull);
              if (v0 & kFlag) { __threadfence(); vp = *(volatile unsigned long long*)&st[4 * (size_t)c + 1]; vm = v0 & ~kFlag; break; }
            }
          }
          const uint32_t pre = __ballot_sync(0xffffffffu, is_prefix);
          const uint32_t first = pre ? (uint32_t)(__ffs(pre) - 1) : 31u;       // nearest predecessor that already knows its inclusive prefix
          if (lane > first) { vm = 0; vp = 0; }
          for (int o = 16; o; o >>= 1) { vm += __shfl_xor_sync(0xffffffffu, vm, o); vp += __shfl_xor_sync(0xffffffffu, vp, o); }
          pm += vm; pp += vp;
          if (pre) done = true; else base -= 32;
        }
        if (lane == 0) {
          st[4 * (size_t)q + 3] = pp + run_p; __threadfence(); atomicExch(&st[4 * (size_t)q + 2], kFlag | (pm + run_m));
          s_off_m = pm; s_off_p = pp;
          if (i + 1 == P.src_hi) { P.counters[3] = pp + run_p; P.counters[4] = pm + run_m; P.pair_off[pp + run_p] = pm + run_m; }
        }
      }
      __syncthreads();
    }
    if (jlo != 0xFFFFFFFFu) {
      // -- scan over j in [jlo, jhi]: kept-match prefix and kept-pair prefix --
      uint32_t run_m = 0, run_p = 0;
      for (uint32_t j0 = jlo; j0 <= jhi; j0 += 4 * kOrderThreads) {
        uint32_t keep[4], lm = 0, lp = 0;
        for (int q = 0; q < 4; ++q) {
          const uint32_t j = j0 + 4 * threadIdx.x + q;
          const uint32_t c = j <= jhi ? cntj[j] : 0;
          keep[q] = c > P.min_corr ? c : 0;
          lm += keep[q]; lp += keep[q] ? 1u : 0u;
        }
        unsigned long long tot2;
        const unsigned long long ex2 = block_exclusive_scan64((unsigned long long)lm | ((unsigned long long)lp << 40), &tot2, sm);
        uint32_t em = (uint32_t)(ex2 & 0xFFFFFFFFFFull), ep = (uint32_t)(ex2 >> 40);
        const uint32_t tm = (uint32_t)(tot2 & 0xFFFFFFFFFFull), tp = (uint32_t)(tot2 >> 40);
        if (PASS >= 1) {
          for (int q = 0; q < 4; ++q) {
            const uint32_t j = j0 + 4 * threadIdx.x + q;
            if (j > jhi) break;
            if (keep[q]) {
              const uint32_t slot = run_p + ep;       // index of pair (i, j) among the pose's kept pairs
              const unsigned long long po = s_off_p + slot;
              P.pair_i[po] = i; P.pair_j[po] = j; P.pair_off[po] = s_off_m + run_m + em;
              // next free slot of the pair inside the pose's kept segment: in shared memory for the first kOrderSlots pairs
              // (cntj[j] then names the slot), in cntj[j] itself beyond
              if (slot < kOrderSlots) { s_next[slot] = run_m + em; cntj[j] = kSlotFlag | slot; }
              else cntj[j] = run_m + em;
              em += keep[q]; ep += 1;
            } else {
              cntj[j] = kDropped;
            }
          }
        }
        run_m += tm; run_p += tp;
      }
      tot_m = run_m; tot_p = run_p;
      if (PASS >= 1) {
        // -- placement, tile by tile: every tile list is sorted by (j, k) and the tiles of a pose come in ascending k
        //    (units of a split tile: ascending target range), so pair (i, j) is the concatenation over tiles of each
        //    tile's run of j.  A record's slot = the pair's next free slot + its rank inside the run (bisection of the
        //    tile's j column in shared memory); the last record of a run then advances the pair's next free slot.
        //    Tile descriptors are fetched once per pose, so the per-tile chain touches shared memory only. --
        // Fast path (the usual case): the pose's kept pairs x ceil(points / 32) words fit the shared bitmap.  Within pair (i, j)
        // the final order is ascending source point k and a point appears at most once, so a record's rank is the number of
        // the pair's points below k: every record sets bit k of its pair's row, rows are prefix-popcounted, every record
        // reads its rank back.  Three sweeps over the records with independent loads — no per-tile chain.
        const uint32_t n_pts = P.off[i + 1] - P.off[i], W = (n_pts + 31) >> 5, ntile = te - tb;
        const bool fast = tot_p <= kOrderSlots && (uint64_t)tot_p * W <= kOrderMaskWords && ntile <= kOrderTileBatch;
        if (fast) {
          __syncthreads();
          for (uint32_t q = threadIdx.x; q < ntile; q += kOrderThreads) { s_tcnt[q] = P.tile_cnt[tb + q]; s_tbase[q] = P.tile_slot[tb + q] * (uint32_t)P.cap; }
          for (uint32_t q = threadIdx.x; q < tot_p * W; q += kOrderThreads) s_mask[q] = 0;
          __syncthreads();
          if (threadIdx.x == 0) { uint32_t acc = 0; for (uint32_t q = 0; q < ntile; ++q) { const uint32_t c = s_tcnt[q]; s_tcnt[q] = acc; acc += c; } s_total = acc; }
          __syncthreads();
          const uint32_t R = s_total;
          auto locate = [&](uint32_t g) -> uint32_t {                     // flat record index -> address in the raw arrays
            uint32_t lo = 0, hi = ntile;                                   // last tile with prefix <= g
            while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (s_tcnt[mid] <= g) lo = mid; else hi = mid; }
            return s_tbase[lo] + (g - s_tcnt[lo]);
          };
          for (uint32_t g = threadIdx.x; g < R; g += kOrderThreads) {
            const uint32_t a = locate(g), v = cntj[P.raw_j[a]];
            if (v == kDropped) continue;
            const uint32_t k = P.raw_k[a];
            atomicOr(&s_mask[(v & ~kSlotFlag) * W + (k >> 5)], 1u << (k & 31));
          }
          __syncthreads();
          for (uint32_t e = threadIdx.x; e < tot_p; e += kOrderThreads) {
            uint32_t acc = 0;
            for (uint32_t w = 0; w < W; ++w) { s_pref[e * W + w] = (uint16_t)acc; acc += __popc(s_mask[e * W + w]); }
          }
          __syncthreads();
          for (uint32_t g = threadIdx.x; g < R; g += kOrderThreads) {
            const uint32_t a = locate(g), v = cntj[P.raw_j[a]];
            if (v == kDropped) continue;
            const uint32_t k = P.raw_k[a], e = v & ~kSlotFlag;
            const uint32_t rank = s_pref[e * W + (k >> 5)] + __popc(s_mask[e * W + (k >> 5)] & ((1u << (k & 31)) - 1u));
            const unsigned long long dst = s_off_m + s_next[e] + rank;
            P.out_k[dst] = k; P.out_idx[dst] = P.raw_idx[a];
          }
        } else
        for (uint32_t t0 = tb; t0 < te; t0 += kOrderTileBatch) {
          const uint32_t nt = min(kOrderTileBatch, te - t0);
          __syncthreads();
          for (uint32_t q = threadIdx.x; q < nt; q += kOrderThreads) { s_tcnt[q] = P.tile_cnt[t0 + q]; s_tbase[q] = P.tile_slot[t0 + q] * (uint32_t)P.cap; }
          __syncthreads();
          for (uint32_t q = 0; q < nt; ++q) {
            const uint32_t c = s_tcnt[q], base = s_tbase[q];
            if (c == 0) continue;
            __syncthreads();
            for (uint32_t u = threadIdx.x; u < c; u += kOrderThreads) s_j[u] = P.raw_j[base + u];
            __syncthreads();
            for (uint32_t u = threadIdx.x; u < c; u += kOrderThreads) {
              const uint32_t j = s_j[u], v = cntj[j];
              if (v == kDropped) continue;
              uint32_t lo = 0, hi = u;                                   // first record of the run of j
              while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (s_j[mid] < j) lo = mid + 1; else hi = mid; }
              const uint32_t st = (v & kSlotFlag) ? s_next[v & ~kSlotFlag] : v;
              const unsigned long long dst = s_off_m + st + (u - lo);
              P.out_k[dst] = P.raw_k[base + u]; P.out_idx[dst] = P.raw_idx[base + u];
            }
            __syncthreads();
            for (uint32_t u = threadIdx.x; u < c; u += kOrderThreads) {
              const uint32_t j = s_j[u];
              if (u + 1 != c && s_j[u + 1] == j) continue;               // only the last record of a run
              const uint32_t v = cntj[j];
              if (v == kDropped) continue;
              uint32_t lo = 0, hi = u;
              while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (s_j[mid] < j) lo = mid + 1; else hi = mid; }
              if (v & kSlotFlag) s_next[v & ~kSlotFlag] += u - lo + 1; else cntj[j] = v + (u - lo + 1);
            }
          }
        }
      }
      // -- clear the touched scratch entries --
      __syncthreads();
      for (uint32_t j = jlo + threadIdx.x; j <= jhi; j += kOrderThreads) cntj[j] = 0;
    }
    if (PASS == 0 && threadIdx.x == 0) { P.pose_cnt[2 * (i - P.src_lo)] = tot_m; P.pose_cnt[2 * (i - P.src_lo) + 1] = tot_p; }
    __syncthreads();
  }
}

// Exclusive scan over the per-pose (matches, pairs) counts; totals to counters[3], [4] and the terminating CSR offset
// pair_off[n_pairs] = n_matches. Single CTA.
__global__ void pose_cnt_scan_kernel(unsigned long long* pose_cnt, uint32_t n, unsigned long long* counters, unsigned long long* pair_off) {
  __shared__ unsigned long long sm_m[32], sm_p[32];
  __shared__ unsigned long long carry_m, carry_p;
  if (threadIdx.x == 0) { carry_m = 0; carry_p = 0; }
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (uint32_t b = 0; b < n; b += blockDim.x) {
    const uint32_t i = b + threadIdx.x;
    const unsigned long long vm = i < n ? pose_cnt[2 * i] : 0, vp = i < n ? pose_cnt[2 * i + 1] : 0;
    unsigned long long xm = vm, xp = vp;
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long ym = __shfl_up_sync(0xffffffffu, xm, o), yp = __shfl_up_sync(0xffffffffu, xp, o);
      if (lane >= (uint32_t)o) { xm += ym; xp += yp; }
    }
    if (lane == 31) { sm_m[w] = xm; sm_p[w] = xp; }
    __syncthreads();
    if (w == 0) {
      unsigned long long sm_ = lane < nw ? sm_m[lane] : 0, sp_ = lane < nw ? sm_p[lane] : 0;
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long ym = __shfl_up_sync(0xffffffffu, sm_, o), yp = __shfl_up_sync(0xffffffffu, sp_, o);
        if (lane >= (uint32_t)o) { sm_ += ym; sp_ += yp; }
      }
      sm_m[lane] = sm_; sm_p[lane] = sp_;
    }
    __syncthreads();
    const unsigned long long bm = carry_m + (w ? sm_m[w - 1] : 0), bp = carry_p + (w ? sm_p[w - 1] : 0);
    if (i < n) { pose_cnt[2 * i] = bm + xm - vm; pose_cnt[2 * i + 1] = bp + xp - vp; }
    __syncthreads();
    if (threadIdx.x == 0) { carry_m += sm_m[nw - 1]; carry_p += sm_p[nw - 1]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { counters[3] = carry_p; counters[4] = carry_m; pair_off[carry_p] = carry_m; }
}

// ------------------------------------------------------------------------------------------------
// Consecutive-pose matcher (FindVisualOdometryCorrespondences). One warp per source scan,
// ordered output through a per-scan slot region, compacted on the host side of the call
// (the result is never consumed downstream in the reference; kept simple).
// ------------------------------------------------------------------------------------------------
__global__ void vo_search_kernel(const float2* __restrict__ pts, const float2* __restrict__ nrm, const float4* __restrict__ node_pm,
                                 const float2* __restrict__ node_nn, const PoseRec* __restrict__ rec, const float4* __restrict__ srcs,
                                 const double* __restrict__ pose, uint32_t i_lo, uint32_t i_hi, float thr, float min_cos,
                                 uint32_t* __restrict__ out_tk, uint32_t* __restrict__ scan_cnt) {
  const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const uint32_t i = i_lo + w;
  if (i >= i_hi) return;
  const PoseRec ri = rec[i], rj = rec[i + 1];
  const float4 si = srcs[i];
  Aff2 src; src.m00 = si.x; src.m01 = -si.y; src.m10 = si.y; src.m11 = si.x; src.tx = si.z; src.ty = si.w;
  Aff2 inv; inv.m00 = rj.i00; inv.m01 = rj.i01; inv.m10 = rj.i10; inv.m11 = rj.i11; inv.tx = rj.itx; inv.ty = rj.ity;
  const Aff2 T = affine_mul(inv, src);
  const float dth = (float)(pose[3 * i + 5] - pose[3 * i + 2]);
  const float sn = sinf_rn(dth), cs = cosf_rn(dth);
  TreeRef t; t.pm = node_pm + rj.off; t.nn = node_nn + rj.off;
  uint32_t total = 0;
  for (uint32_t k0 = 0; k0 < ri.n; k0 += 32) {
    const uint32_t k = k0 + lane;
    bool ok = false; uint32_t tgt = 0;
    if (k < ri.n && rj.n) {
      const float2 p = pts[ri.off + k], nv = nrm[ri.off + k];
      float qx, qy; affine_apply(T, p.x, p.y, &qx, &qy);
      float best; uint32_t bpos;
      nearest_point(t, rj.n, qx, qy, thr, &best, &bpos);
      float rnx, rny; rot_apply(cs, sn, nv.x, nv.y, &rnx, &rny);
      const float2 nb = __ldg(t.nn + bpos);
      ok = best < thr && fadd(fmul(nb.x, rnx), fmul(nb.y, rny)) > min_cos;
      tgt = node_meta(__ldg(t.pm + bpos)) & 0x7FFFFFFFu;
    }
    const uint32_t m = __ballot_sync(0xffffffffu, ok);
    // slot per source point: target index or 0xFFFFFFFF
    if (k < ri.n) out_tk[ri.off + k] = ok ? tgt : 0xFFFFFFFFu;
    total += __popc(m);
  }
  if (lane == 0) scan_cnt[i] = total;
}

}  // namespace hitl

// ================================================================================================
// Host side of the C ABI for this file.
// ================================================================================================
using namespace hitl;

static uint32_t stack_levels(uint32_t max_scan) {
  uint32_t d = 1;
  while ((1u << d) <= max_scan) ++d;   // floor(log2(max_scan)) + 1 >= number of non-leaf levels + 1
  return d;
}

extern "C" int hitl_kd_query(hitl_ctx* ctx, uint32_t scan, uint32_t nq, const float* q_xy, float threshold, int mode,
                             float* dist_out, int32_t* index_out) {
  if (!ctx) return HITL_ERR_ARG;
  HITL_DEVICE(ctx);
  if (!ctx->have_trees) return fail(ctx, HITL_ERR_STATE, "hitl_kd_query: trees not built");
  if (scan >= ctx->n_poses || mode < 0 || mode > 2 || (nq && (!q_xy || !index_out))) return fail(ctx, HITL_ERR_ARG, "hitl_kd_query: bad argument");
  if (nq == 0) return HITL_OK;
  TmpBuf<float2> dq; TmpBuf<float> dd; TmpBuf<int32_t> di;
  HITL_CUDA(dq.ensure(nq)); HITL_CUDA(dd.ensure(nq)); HITL_CUDA(di.ensure(nq));
  HITL_CUDA(cudaMemcpyAsync(dq.p, q_xy, sizeof(float2) * nq, cudaMemcpyHostToDevice, ctx->stream));
  const uint32_t toff = ctx->h_off[scan], tn = ctx->h_off[scan + 1] - toff;
  const int threads = 128;
  const size_t smem = (size_t)stack_levels(ctx->max_scan) * 2 * threads * sizeof(uint2);
  kd_query_kernel<<<(nq + threads - 1) / threads, threads, smem, ctx->stream>>>(ctx->d_node_pm.p, ctx->d_node_nn.p, toff, tn, nq, dq.p,
                                                                              threshold, mode, dd.p, di.p);
  HITL_LAUNCH_CHECK("kd_query_kernel");
  if (dist_out && mode != 2) HITL_CUDA(cudaMemcpyAsync(dist_out, dd.p, sizeof(float) * nq, cudaMemcpyDeviceToHost, ctx->stream));
  HITL_CUDA(cudaMemcpyAsync(index_out, di.p, sizeof(int32_t) * nq, cudaMemcpyDeviceToHost, ctx->stream));
  HITL_CUDA(cudaStreamSynchronize(ctx->stream));
  return HITL_OK;
}

extern "C" int hitl_kd_neighbors(hitl_ctx* ctx, uint32_t scan, uint32_t nq, const float* q_xy, float threshold, uint32_t cap, int32_t* index_out,
                                 uint32_t* count_out) {
  if (!ctx) return HITL_ERR_ARG;
  HITL_DEVICE(ctx);
  if (!ctx->have_trees) return fail(ctx, HITL_ERR_STATE, "hitl_kd_neighbors: trees not built");
  if (scan >= ctx->n_poses || (nq && (!q_xy || !count_out || (cap && !index_out)))) return fail(ctx, HITL_ERR_ARG, "hitl_kd_neighbors: bad argument");
  if (nq == 0) return HITL_OK;
  TmpBuf<float2> dq; TmpBuf<int32_t> di; TmpBuf<uint32_t> dc;
  HITL_CUDA(dq.ensure(nq)); HITL_CUDA(di.ensure((size_t)nq * cap)); HITL_CUDA(dc.ensure(nq));
  HITL_CUDA(cudaMemcpyAsync(dq.p, q_xy, sizeof(float2) * nq, cudaMemcpyHostToDevice, ctx->stream));
  if (cap) HITL_CUDA(cudaMemsetAsync(di.p, 0xFF, sizeof(int32_t) * (size_t)nq * cap, ctx->stream));   // unused slots read -1
  const uint32_t toff = ctx->h_off[scan], tn = ctx->h_off[scan + 1] - toff;
  kd_neighbors_kernel<<<(nq + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_node_pm.p, toff, tn, nq, dq.p, threshold, cap, di.p, dc.p);
  HITL_LAUNCH_CHECK("kd_neighbors_kernel");
  if (cap) HITL_CUDA(cudaMemcpyAsync(index_out, di.p, sizeof(int32_t) * (size_t)nq * cap, cudaMemcpyDeviceToHost, ctx->stream));
  HITL_CUDA(cudaMemcpyAsync(count_out, dc.p, sizeof(uint32_t) * nq, cudaMemcpyDeviceToHost, ctx->stream));
  HITL_CUDA(cudaStreamSynchronize(ctx->stream));
  return HITL_OK;
}

namespace hitl {
// (Re)build the per-scan occupancy bitmaps for threshold thr (cached until the scans or thr change).
int ensure_occupancy(hitl_ctx* ctx, float thr) {
  if (ctx->grid_valid && ctx->grid_thr == thr) return HITL_OK;
  struct timespec ts0; clock_gettime(CLOCK_MONOTONIC, &ts0);
  const uint32_t n = ctx->n_poses;
  std::vector<GridRec> tab(n);
  std::vector<uint64_t> fine_words(n, 0);
  const float c0 = thr * (1.0f + 1.0f / 512.0f);
  uint64_t words = 0;
  for (uint32_t i = 0; i < n; ++i) {
    GridRec g; g.gx0 = g.gy0 = 0.f; g.ginv = 0.f; g.gdim = 0; g.goff = 0; g.foff = kNoFine;
    if (ctx->h_off[i + 1] > ctx->h_off[i]) {
      const float* b = &ctx->h_aabb[4 * (size_t)i];
      const float ext = std::max(b[2] - b[0], b[3] - b[1]);
      float c = std::max(c0, ext / 4000.0f);
      if (!(c > 0.0f)) c = 1.0f;
      g.gx0 = b[0] - 1.5f * c; g.gy0 = b[1] - 1.5f * c; g.ginv = 1.0f / c;
      const uint32_t nx = (uint32_t)((b[2] - g.gx0) * g.ginv) + 4, ny = (uint32_t)((b[3] - g.gy0) * g.ginv) + 4;
      g.gdim = nx | (ny << 16);
      if (words >= 0xFFFFFFFFull) return fail(ctx, HITL_ERR_ARG, "occupancy bitmaps exceed 2^32 words");
      g.goff = (uint32_t)words;
      words += ((uint64_t)nx * ny + 31) / 32;
      g.foff = 0;
      fine_words[i] = ((uint64_t)nx * kFineCells * ny * kFineCells + 31) / 32;
    }
    tab[i] = g;
  }
  // second level (cell / kFineCells): offsets in words; dropped as a whole when it would not fit 32-bit word offsets
  uint64_t fwords = 0;
  for (uint32_t i = 0; i < n; ++i) fwords += fine_words[i];
  const bool fine = ctx->fine_occupancy && fwords < 0xFFFFFFFFull;
  fwords = 0;
  for (uint32_t i = 0; i < n; ++i) {
    tab[i].foff = (fine && fine_words[i]) ? (uint32_t)fwords : kNoFine;
    if (fine) fwords += fine_words[i];
  }
  // mip level: the coarse bitmap reduced kMipCells x kMipCells
  std::vector<uint32_t> moff(n ? n : 1, 0);
  uint64_t mwords = 0;
  for (uint32_t i = 0; i < n; ++i) {
    moff[i] = (uint32_t)mwords;
    const uint32_t nx = tab[i].gdim & 0xFFFFu, ny = tab[i].gdim >> 16;
    mwords += ((uint64_t)((nx + kMipCells - 1) / kMipCells) * ((ny + kMipCells - 1) / kMipCells) + 31) / 32;
  }
  HITL_CUDA(ctx->d_occ_mip.ensure(mwords)); HITL_CUDA(ctx->d_moff.ensure(n));
  if (n) HITL_CUDA(cudaMemcpyAsync(ctx->d_moff.p, moff.data(), 4 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
  HITL_CUDA(cudaMemsetAsync(ctx->d_occ_mip.p, 0, 4 * (mwords ? mwords : 1), ctx->stream));
  HITL_CUDA(ctx->d_grid.ensure(n)); HITL_CUDA(ctx->d_occ.ensure(words)); HITL_CUDA(ctx->d_occ_fine.ensure(fwords));
  if (n) HITL_CUDA(cudaMemcpyAsync(ctx->d_grid.p, tab.data(), sizeof(GridRec) * n, cudaMemcpyHostToDevice, ctx->stream));
  HITL_CUDA(cudaMemsetAsync(ctx->d_occ.p, 0, 4 * (words ? words : 1), ctx->stream));
  HITL_CUDA(cudaMemsetAsync(ctx->d_occ_fine.p, 0, 4 * (fwords ? fwords : 1), ctx->stream));
  if (ctx->n_tiles) {
    const int threads = 128;
    const uint32_t blocks = (uint32_t)(((size_t)ctx->n_tiles * 32 + threads - 1) / threads);
    occupancy_build_kernel<<<blocks, threads, 0, ctx->stream>>>(
        ctx->d_pts.p, ctx->d_off.p, ctx->d_tile_scan.p, ctx->d_tile_k0.p, ctx->n_tiles, ctx->d_grid.p, ctx->d_occ.p, ctx->d_occ_mip.p, ctx->d_moff.p);
    HITL_LAUNCH_CHECK("occupancy_build_kernel");
    if (fine && fwords) {
      occupancy_fine_build_kernel<<<blocks, threads, 0, ctx->stream>>>(
          ctx->d_pts.p, ctx->d_off.p, ctx->d_tile_scan.p, ctx->d_tile_k0.p, ctx->n_tiles, ctx->d_grid.p, ctx->d_occ_fine.p);
      HITL_LAUNCH_CHECK("occupancy_fine_build_kernel");
    }
  }
  // direction masks of the node normals (two 16-bit cells per word) and the per-scan bound on |normal|
  HITL_CUDA(ctx->d_nmax.ensure(n));
  HITL_CUDA(cudaMemsetAsync(ctx->d_nmax.p, 0, 4 * (size_t)(n ? n : 1), ctx->stream));
  if (ctx->dir_occupancy && ctx->n_points) {
    const uint64_t cells = words * 32;
    HITL_CUDA(ctx->d_occ_dir.ensure((cells + 1) / 2));
    HITL_CUDA(cudaMemsetAsync(ctx->d_occ_dir.p, 0, 4 * ((cells + 1) / 2), ctx->stream));
    occupancy_dir_build_kernel<<<(uint32_t)((ctx->n_points + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_node_pm.p, ctx->d_node_nn.p, ctx->d_off.p, n, ctx->n_points,
                                                                                                 ctx->d_grid.p, ctx->d_occ_dir.p, ctx->d_nmax.p);
    HITL_LAUNCH_CHECK("occupancy_dir_build_kernel");
  }
  HITL_CUDA(cudaStreamSynchronize(ctx->stream));   // tab is a local
  if (getenv("HITL_STF_TIMING")) {
    struct timespec ts1; clock_gettime(CLOCK_MONOTONIC, &ts1);
    fprintf(stderr, "occupancy rebuild (coarse + mip + fine + direction levels): %.3f ms host wall\n", (ts1.tv_sec - ts0.tv_sec) * 1e3 + (ts1.tv_nsec - ts0.tv_nsec) * 1e-6);
  }
  ctx->grid_valid = true; ctx->grid_thr = thr;
  return HITL_OK;
}

int upload_poses_and_prep(hitl_ctx* ctx, const double* pose_array, float thr) {
  int rc = ensure_occupancy(ctx, thr);
  if (rc) return rc;
  HITL_CUDA(ctx->d_pose.ensure(3 * (size_t)ctx->n_poses));
  HITL_CUDA(ctx->d_rec.ensure(ctx->n_poses)); HITL_CUDA(ctx->d_src.ensure(ctx->n_poses));
  HITL_CUDA(ctx->d_wbox.ensure(ctx->n_poses));
  HITL_CUDA(cudaMemcpyAsync(ctx->d_pose.p, pose_array, sizeof(double) * 3 * ctx->n_poses, cudaMemcpyHostToDevice, ctx->stream));
  pose_prep_kernel<<<(ctx->n_poses + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_pose.p, ctx->d_aabb.p, ctx->d_off.p, ctx->d_grid.p, ctx->d_nmax.p, ctx->n_poses,
                                                                        ctx->d_rec.p, ctx->d_src.p, ctx->d_wbox.p);
  HITL_LAUNCH_CHECK("pose_prep_kernel");
  const uint32_t n_groups = (ctx->n_poses + 31) / 32;
  HITL_CUDA(ctx->d_gbox.ensure(n_groups));
  if (n_groups) {
    pose_group_box_kernel<<<(n_groups * 32 + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_wbox.p, ctx->n_poses, ctx->d_gbox.p);
    HITL_LAUNCH_CHECK("pose_group_box_kernel");
  }
  return HITL_OK;
}
int launch_scan_aabb(hitl_ctx* ctx) {
  HITL_CUDA(ctx->d_aabb.ensure(ctx->n_poses));
  if (ctx->n_poses == 0) return HITL_OK;
  const int threads = 128;
  scan_aabb_kernel<<<((size_t)ctx->n_poses * 32 + threads - 1) / threads, threads, 0, ctx->stream>>>(ctx->d_pts.p, ctx->d_off.p, ctx->n_poses,
                                                                                                   ctx->d_aabb.p);
  HITL_LAUNCH_CHECK("scan_aabb_kernel");
  return HITL_OK;
}
}  // namespace hitl

extern "C" int hitl_find_stf(hitl_ctx* ctx, const double* pose_array, uint32_t min_pose, uint32_t max_pose, uint32_t src_lo,
                             uint32_t src_hi, const hitl_stf_opts* o, hitl_stf_info* info) {
  if (!ctx) return HITL_ERR_ARG;
  HITL_DEVICE(ctx);
  if (!pose_array || !o) return fail(ctx, HITL_ERR_ARG, "hitl_find_stf: null argument");
  if (!ctx->have_trees) return fail(ctx, HITL_ERR_STATE, "hitl_find_stf: scans/trees not set");
  if (o->num_skip_readings == 0) return fail(ctx, HITL_ERR_ARG, "hitl_find_stf: num_skip_readings must be >= 1");
  if (o->max_correspondences_per_point > 64) return fail(ctx, HITL_ERR_ARG, "hitl_find_stf: max_correspondences_per_point > 64 unsupported");
  hitl_stf_info inf; memset(&inf, 0, sizeof(inf));
  ctx->have_stf = false; ctx->n_pairs = ctx->n_matches = 0;
  if (ctx->stf_from_search) ctx->eval_valid = ctx->neq_valid = false; ctx->inc_valid = false;
  // poses_end = min(max_poses + 1, n)  (JointOptimization.cpp:566-567)
  const uint32_t n = ctx->n_poses;
  const uint64_t poses_end = std::min<uint64_t>((uint64_t)max_pose + 1, n);
  const int cap = o->max_correspondences_per_point;
  uint32_t lo = std::max(min_pose, src_lo), hi = (uint32_t)std::min<uint64_t>(poses_end, src_hi);
  if (poses_end <= min_pose || lo >= hi || cap <= 0 || ctx->n_points == 0) {
    if (cap <= 0 && poses_end > min_pose && lo < hi) inf.n_queries = 0;
    ctx->have_stf = true;
    if (info) *info = inf;
    return HITL_OK;
  }
  const uint32_t jmin = min_pose, jmax = (uint32_t)poses_end - 1;
  const size_t rec_cap = (size_t)ctx->n_points * cap, raw_cap = (size_t)ctx->n_slots * cap;
  if (raw_cap >= 0xFFFFFFFFull) return fail(ctx, HITL_ERR_ARG, "hitl_find_stf: n_points * cap exceeds 2^32 records");
  HITL_CUDA(ctx->d_raw_j.ensure(raw_cap)); HITL_CUDA(ctx->d_raw_k.ensure(raw_cap)); HITL_CUDA(ctx->d_raw_idx.ensure(raw_cap));
  HITL_CUDA(ctx->d_tile_cnt.ensure(ctx->n_tiles));
  HITL_CUDA(ctx->d_counters.ensure(16));
  HITL_CUDA(ctx->d_pose_cnt.ensure(2 * (size_t)n + 2));
  HITL_CUDA(ctx->d_pose_work.ensure(n));
  HITL_CUDA(ctx->d_k.ensure(rec_cap)); HITL_CUDA(ctx->d_idx.ensure(rec_cap));
  const size_t pair_cap = rec_cap / ((size_t)o->min_inter_pose_correspondence + 1) + 1;   // + 1 in 64 bits: UINT32_MAX must not wrap to a zero divisor
  HITL_CUDA(ctx->d_pair_i.ensure(pair_cap)); HITL_CUDA(ctx->d_pair_j.ensure(pair_cap)); HITL_CUDA(ctx->d_pair_off.ensure(pair_cap + 1));

  HITL_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
  HITL_CUDA(cudaMemsetAsync(ctx->d_counters.p, 0, 16 * sizeof(uint64_t), ctx->stream));
  HITL_CUDA(cudaMemsetAsync(ctx->d_pose_work.p, 0, sizeof(uint64_t) * n, ctx->stream));
  int rc = upload_poses_and_prep(ctx, pose_array, o->point_match_threshold);
  if (rc) return rc;

  HITL_CUDA(cudaEventRecord(ctx->evx[0], ctx->stream));   // after pose upload + prep
  SearchParams P;
  P.pts = ctx->d_pts.p; P.nrm = ctx->d_nrm.p; P.node_pm = ctx->d_node_pm.p; P.node_nn = ctx->d_node_nn.p;
  P.occ_dir = (ctx->dir_occupancy && o->min_cosine_angle > 0.0f && o->min_cosine_angle < 1.0f) ? ctx->d_occ_dir.p : nullptr;
  P.gbox = ctx->d_gbox.p; P.n_groups = (n + 31) / 32;
  P.occ_mip = ctx->mip_occupancy ? ctx->d_occ_mip.p : nullptr; P.moff = ctx->d_moff.p;
  P.dir_alpha_unit = acosf(std::min(1.0f, std::max(0.0f, o->min_cosine_angle / 1.0006f)));
  P.rec = ctx->d_rec.p; P.src = ctx->d_src.p; P.occ = ctx->d_occ.p; P.occ_fine = ctx->d_occ_fine.p; P.wbox = ctx->d_wbox.p; P.pose = ctx->d_pose.p; P.tile_scan = ctx->d_tile_scan.p; P.tile_k0 = ctx->d_tile_k0.p;
  P.tile_j = ctx->d_tile_j.p; P.tile_slot = ctx->d_tile_slot.p;
  P.tile_lo = ctx->h_tile_begin[lo]; P.tile_hi = ctx->h_tile_begin[hi];
  P.jmin = jmin; P.jmax = jmax; P.thr = o->point_match_threshold; P.min_cos = o->min_cosine_angle; P.cap = cap;
  P.skip = o->num_skip_readings; P.no_cull = o->disable_culling;
  P.raw_j = ctx->d_raw_j.p; P.raw_k = ctx->d_raw_k.p; P.raw_idx = ctx->d_raw_idx.p; P.tile_cnt = ctx->d_tile_cnt.p;
  P.counters = (unsigned long long*)ctx->d_counters.p;
  P.pose_work = (unsigned long long*)ctx->d_pose_work.p;
  P.tile_work = ctx->d_tile_work.p;
  HITL_CUDA(ctx->d_tile_open.ensure(ctx->n_tiles)); HITL_CUDA(ctx->d_tile_end.ensure(ctx->n_tiles));
  P.tile_open = ctx->d_tile_open.p; P.tile_end = ctx->d_tile_end.p;
  P.tile_order = nullptr;
  const uint32_t n_tiles = P.tile_hi - P.tile_lo;
  const uint32_t wpb = kSearchThreads / 32;
  if (n_tiles > 1 && !o->disable_culling) {
    // Heaviest-first ticket order from the latest per-tile cost estimates (measured by the previous search,
    // inherited by the children of split tiles, 0 for tiles never searched): a scheduling hint only.
    size_t tmp_bytes = 0;
    HITL_CUDA(cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp_bytes, ctx->d_tile_work.p + P.tile_lo, ctx->d_tile_keys.p, ctx->d_tile_iota.p + P.tile_lo,
                                                        ctx->d_tile_order.p, (int)n_tiles, 0, 32, ctx->stream));
    HITL_CUDA(ctx->d_sort_tmp.ensure(tmp_bytes));
    HITL_CUDA(cub::DeviceRadixSort::SortPairsDescending(ctx->d_sort_tmp.p, tmp_bytes, ctx->d_tile_work.p + P.tile_lo, ctx->d_tile_keys.p, ctx->d_tile_iota.p + P.tile_lo,
                                                        ctx->d_tile_order.p, (int)n_tiles, 0, 32, ctx->stream));
    ctx->launches += 1;
    P.tile_order = ctx->d_tile_order.p;
  }
  HITL_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
  if (n_tiles) {
    // persistent grid: one wave of CTAs (a multiple of the SM count), warps pull tiles from a ticket
    int per_sm = 0;
    void (*kernel)(const SearchParams) = ctx->search_variant == 1 ? stf_search_kernel<12> : (ctx->search_variant == 2 ? stf_search_kernel<10> : stf_search_kernel<kSearchMinBlocks>);
    if (ctx->search_carveout >= 0) HITL_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, ctx->search_carveout));
    HITL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kSearchThreads, 0));
    if (per_sm < 1) per_sm = 1;
    const uint32_t grid = std::min<uint32_t>((n_tiles + wpb - 1) / wpb, (uint32_t)(ctx->sm_count * per_sm));
    HITL_KERNEL_BEGIN(HITL_K_STF_SEARCH);
    kernel<<<grid, kSearchThreads, 0, ctx->stream>>>(P);
    HITL_KERNEL_END(HITL_K_STF_SEARCH);
    HITL_LAUNCH_CHECK("stf_search_kernel");
  }
  HITL_CUDA(cudaEventRecord(ctx->ev[2], ctx->stream));
  if (ctx->n_groups && n_tiles) {
    MergeParams M;
    M.groups = ctx->d_groups.p; M.n_groups = ctx->n_groups; M.tile_lo = P.tile_lo; M.tile_hi = P.tile_hi;
    M.tile_scan = ctx->d_tile_scan.p; M.tile_k0 = ctx->d_tile_k0.p; M.tile_slot = ctx->d_tile_slot.p; M.off = ctx->d_off.p;
    M.raw_j = ctx->d_raw_j.p; M.raw_k = ctx->d_raw_k.p; M.raw_idx = ctx->d_raw_idx.p; M.tile_cnt = ctx->d_tile_cnt.p;
    M.jmin = jmin; M.jmax = jmax; M.skip = o->num_skip_readings; M.cap = cap; M.counters = (unsigned long long*)ctx->d_counters.p;
    stf_split_merge_kernel<<<(ctx->n_groups + 3) / 4, 128, 0, ctx->stream>>>(M);
    HITL_LAUNCH_CHECK("stf_split_merge_kernel");
  }

  HITL_CUDA(cudaEventRecord(ctx->evx[1], ctx->stream));   // after the split merge
  OrderParams Q;
  Q.raw_j = ctx->d_raw_j.p; Q.raw_k = ctx->d_raw_k.p; Q.raw_idx = ctx->d_raw_idx.p; Q.tile_cnt = ctx->d_tile_cnt.p;
  Q.tile_begin = ctx->d_tile_begin.p; Q.tile_slot = ctx->d_tile_slot.p; Q.off = ctx->d_off.p; Q.src_lo = lo; Q.src_hi = hi; Q.n_poses = n; Q.cap = cap;
  Q.min_corr = o->min_inter_pose_correspondence;
  int order_per_sm = 0;
  HITL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&order_per_sm, ctx->order_two_pass ? stf_order_kernel<1> : stf_order_kernel<2>, kOrderThreads, 0));
  if (order_per_sm < 1) order_per_sm = 1;
  const uint32_t order_grid = std::min<uint32_t>(hi - lo, (uint32_t)(ctx->sm_count * order_per_sm));   // one resident wave: a CTA keeps an n_poses-sized scratch
  HITL_CUDA(ctx->d_srt_j.ensure((size_t)order_grid * n));   // reused as the j-indexed scratch
  HITL_CUDA(cudaMemsetAsync(ctx->d_srt_j.p, 0, sizeof(uint32_t) * (size_t)order_grid * n, ctx->stream));
  Q.scratch = ctx->d_srt_j.p; Q.pose_cnt = (unsigned long long*)ctx->d_pose_cnt.p;
  Q.pair_i = ctx->d_pair_i.p; Q.pair_j = ctx->d_pair_j.p; Q.pair_off = (unsigned long long*)ctx->d_pair_off.p;
  Q.out_k = ctx->d_k.p; Q.out_idx = ctx->d_idx.p;
  if (!ctx->order_two_pass) {
    // single pass: count, look back for the pose's offsets, place (one resident wave of CTAs, poses by ascending ticket)
    HITL_CUDA(ctx->d_order_state.ensure(4 * (size_t)(hi - lo)));
    HITL_CUDA(cudaMemsetAsync(ctx->d_order_state.p, 0, 32 * (size_t)(hi - lo), ctx->stream));
    Q.state = (unsigned long long*)ctx->d_order_state.p; Q.ticket = (unsigned long long*)ctx->d_counters.p + 14; Q.counters = (unsigned long long*)ctx->d_counters.p;
    HITL_CUDA(cudaEventRecord(ctx->evx[2], ctx->stream));
    HITL_CUDA(cudaEventRecord(ctx->evx[3], ctx->stream));
    stf_order_kernel<2><<<order_grid, kOrderThreads, 0, ctx->stream>>>(Q);
    HITL_LAUNCH_CHECK("stf_order_kernel<2>");
  } else {
  stf_order_kernel<0><<<order_grid, kOrderThreads, 0, ctx->stream>>>(Q);
  HITL_LAUNCH_CHECK("stf_order_kernel<0>");
  HITL_CUDA(cudaEventRecord(ctx->evx[2], ctx->stream));   // after order<0>
  pose_cnt_scan_kernel<<<1, 1024, 0, ctx->stream>>>((unsigned long long*)ctx->d_pose_cnt.p, hi - lo, (unsigned long long*)ctx->d_counters.p, (unsigned long long*)ctx->d_pair_off.p);
  HITL_LAUNCH_CHECK("pose_cnt_scan_kernel");
  HITL_CUDA(cudaEventRecord(ctx->evx[3], ctx->stream));   // after the scan
  stf_order_kernel<1><<<order_grid, kOrderThreads, 0, ctx->stream>>>(Q);
  HITL_LAUNCH_CHECK("stf_order_kernel<1>");
  }
  HITL_CUDA(cudaEventRecord(ctx->ev[3], ctx->stream));
  HITL_CUDA(cudaMemcpyAsync(ctx->h_pinned, ctx->d_counters.p, 16 * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
  HITL_CUDA(cudaStreamSynchronize(ctx->stream));
  inf.n_queries = ctx->h_pinned[0]; inf.n_traversals = ctx->h_pinned[1]; inf.n_raw_matches = ctx->h_pinned[2];
  inf.n_pairs = ctx->h_pinned[3]; inf.n_matches = ctx->h_pinned[4]; inf.n_tile_pairs = ctx->h_pinned[5]; inf.n_coarse_pass = ctx->h_pinned[9]; inf.n_in_radius = ctx->h_pinned[10];
  inf.sum_tile_cycles = ctx->h_pinned[7] << 6; inf.max_tile_cycles = ctx->h_pinned[8] << 6;
  inf.n_gate_fail = ctx->h_pinned[11]; inf.n_over_cap = ctx->h_pinned[12]; inf.n_dir_culled = ctx->h_pinned[13];
  const uint64_t work_sum = ctx->h_pinned[7];
  HITL_CUDA(cudaEventElapsedTime(&inf.ms_search, ctx->ev[1], ctx->ev[2]));
  HITL_CUDA(cudaEventElapsedTime(&inf.ms_total, ctx->ev[0], ctx->ev[3]));
  if (getenv("HITL_STF_TIMING")) {
    float a = 0, b = 0, c = 0, d = 0, e = 0, f = 0;
    cudaEventElapsedTime(&a, ctx->ev[0], ctx->evx[0]); cudaEventElapsedTime(&b, ctx->evx[0], ctx->ev[1]); cudaEventElapsedTime(&c, ctx->ev[2], ctx->evx[1]);
    cudaEventElapsedTime(&d, ctx->evx[1], ctx->evx[2]); cudaEventElapsedTime(&e, ctx->evx[2], ctx->evx[3]); cudaEventElapsedTime(&f, ctx->evx[3], ctx->ev[3]);
    fprintf(stderr, "find_stf phases (ms): prep %.3f | schedule sort %.3f | search %.3f | split merge %.3f | order<0> %.3f | scan %.3f | order<1> %.3f\n", a, b, inf.ms_search, c, d, e, f);
  }
  ctx->n_pairs = inf.n_pairs; ctx->n_matches = inf.n_matches; ctx->have_stf = true;
  inf.n_tiles = n_tiles; inf.n_tiles_next = n_tiles;
  if (info) *info = inf;
  // Adaptive tiling for the NEXT call: a tile is a sequential loop over the target poses, so the heaviest
  // tile bounds the kernel from below; when it exceeds its fair share of this shard by far (small maps,
  // shards of a multi-GPU run) it is split into shorter tiles.  Scheduling only: results do not depend on it.
  if (ctx->adaptive_tiling && n_tiles > 1 && !o->disable_culling) {
    const uint64_t slots = (uint64_t)ctx->sm_count * 64;
    const uint64_t fair = work_sum / slots + 1;                              // cycles/64 per resident warp if perfectly packed
    const uint64_t limit = std::max<uint64_t>(fair / std::max(1u, ctx->split_limit_div), 4096);   // never split tiles cheaper than ~0.13 ms
    const uint64_t h_max = ctx->h_pinned[8];                                 // heaviest tile of this call
    if (ctx->split_lo != lo || ctx->split_hi != hi) { ctx->split_lo = lo; ctx->split_hi = hi; ctx->split_rounds = 0; }
    // Re-tiling is a SETUP cost (milliseconds of host work): it may only follow the first max_split_rounds CALLS on a source range.
    // Counting calls, not splits, matters: a heaviest tile that hovers around the threshold would otherwise keep re-tiling now and
    // then for ever (seen on one rank of an 8-rank run: 5-10 ms of host time inside steady-state steps).
    const bool may_split = ctx->split_rounds < ctx->max_split_rounds;
    if (may_split) ++ctx->split_rounds;
    if (may_split && h_max > 2 * limit) {
      std::vector<uint32_t> h_work(n_tiles), h_open(n_tiles), h_end(n_tiles), est;
      HITL_CUDA(cudaMemcpy(h_work.data(), ctx->d_tile_work.p + P.tile_lo, 4 * (size_t)n_tiles, cudaMemcpyDeviceToHost));
      HITL_CUDA(cudaMemcpy(h_open.data(), ctx->d_tile_open.p + P.tile_lo, 4 * (size_t)n_tiles, cudaMemcpyDeviceToHost));
      HITL_CUDA(cudaMemcpy(h_end.data(), ctx->d_tile_end.p + P.tile_lo, 4 * (size_t)n_tiles, cudaMemcpyDeviceToHost));
      if (split_heavy_tiles(ctx, h_work, h_open, h_end, P.tile_lo, P.tile_hi, limit, &est)) {
        rc = upload_tiling(ctx);
        if (rc) return rc;
        const uint32_t new_lo = ctx->h_tile_begin[lo], new_hi = ctx->h_tile_begin[hi], nn = new_hi - new_lo;
        if (info) info->n_tiles_next = nn;
        // the heaviest-first schedule survives the re-tiling: children inherit their parent's share of the work
        HITL_CUDA(cudaMemcpyAsync(ctx->d_tile_work.p, est.data(), 4 * (size_t)ctx->n_tiles, cudaMemcpyHostToDevice, ctx->stream));
        HITL_CUDA(cudaStreamSynchronize(ctx->stream));   // est is a local
      }
    }
  }
  return HITL_OK;
}

extern "C" int hitl_debug_set_tiling(hitl_ctx* ctx, uint32_t max_len, int adaptive, uint32_t target_parts) {
  if (!ctx) return HITL_ERR_ARG;
  HITL_DEVICE(ctx);
  if (ctx->h_off.empty()) return fail(ctx, HITL_ERR_STATE, "hitl_debug_set_tiling: scans not set");
  ctx->adaptive_tiling = adaptive != 0;
  ctx->target_splitting = adaptive == 1;
  ctx->split_rounds = 0;
  int rc = build_tiling(ctx, max_len);
  if (rc || target_parts <= 1 || ctx->n_poses < 2) return rc;
  // forced uniform cut of EVERY tile into target_parts consecutive target ranges (parity tests of the cap merge)
  const uint32_t parts = std::min(target_parts, ctx->n_poses);
  std::vector<uint32_t> scan, kl, jlo, jhi, begin(ctx->n_poses + 1, 0);
  uint32_t pose = 0;
  for (uint32_t t = 0; t < (uint32_t)ctx->h_tile_scan.size(); ++t) {
    const uint32_t i = ctx->h_tile_scan[t];
    while (pose <= i) begin[pose++] = (uint32_t)scan.size();
    for (uint32_t q = 0; q < parts; ++q) {
      scan.push_back(i); kl.push_back(ctx->h_tile_kl[t]);
      jlo.push_back((uint32_t)((uint64_t)ctx->n_poses * q / parts));
      jhi.push_back(q + 1 == parts ? kFullRange : (uint32_t)((uint64_t)ctx->n_poses * (q + 1) / parts) - 1);
    }
  }
  while (pose <= ctx->n_poses) begin[pose++] = (uint32_t)scan.size();
  ctx->h_tile_scan.swap(scan); ctx->h_tile_kl.swap(kl); ctx->h_tile_jlo.swap(jlo); ctx->h_tile_jhi.swap(jhi); ctx->h_tile_begin.swap(begin);
  return upload_tiling(ctx);
}

extern "C" int hitl_debug_set_search_variant(hitl_ctx* ctx, int variant, int smem_carveout_pct) {
  if (!ctx || variant < 0 || variant > 2 || smem_carveout_pct > 100) return HITL_ERR_ARG;
  ctx->search_variant = variant;
  ctx->search_carveout = smem_carveout_pct;
  return HITL_OK;
}

extern "C" int hitl_debug_set_fine_occupancy(hitl_ctx* ctx, int on) {
  if (!ctx) return HITL_ERR_ARG;
  HITL_DEVICE(ctx);
  // bit 0: fine level, bit 1 set: direction prefilter OFF (on = 1 keeps both culls, on = 0 drops the fine level only, on = 2 / 3 drop the prefilter)
  const int fine = on & 1, dir = (on & 2) ? 0 : 1;
  ctx->mip_occupancy = (on & 4) ? 0 : 1;                       // bit 2 set: tile-box cull OFF (no rebuild needed: the kernel just ignores the level)
  if (ctx->fine_occupancy != fine || ctx->dir_occupancy != dir) { ctx->fine_occupancy = fine; ctx->dir_occupancy = dir; ctx->grid_valid = false; }
  return HITL_OK;
}

extern "C" int hitl_get_stf(hitl_ctx* ctx, uint32_t* pair_i, uint32_t* pair_j, uint64_t* pair_off, uint32_t* k, uint32_t* idx) {
  if (!ctx) return HITL_ERR_ARG;
  HITL_DEVICE(ctx);
  if (!ctx->have_stf) return fail(ctx, HITL_ERR_STATE, "hitl_get_stf: no search result");
  const uint64_t np = ctx->n_pairs, nm = ctx->n_matches;
  if (pair_off && np == 0) pair_off[0] = 0;
  if (np) {
    if (pair_i) HITL_CUDA(cudaMemcpyAsync(pair_i, ctx->d_pair_i.p, 4 * np, cudaMemcpyDeviceToHost, ctx->stream));
    if (pair_j) HITL_CUDA(cudaMemcpyAsync(pair_j, ctx->d_pair_j.p, 4 * np, cudaMemcpyDeviceToHost, ctx->stream));
    if (pair_off) HITL_CUDA(cudaMemcpyAsync(pair_off, ctx->d_pair_off.p, 8 * (np + 1), cudaMemcpyDeviceToHost, ctx->stream));
  }
  if (nm) {
    if (k) HITL_CUDA(cudaMemcpyAsync(k, ctx->d_k.p, 4 * nm, cudaMemcpyDeviceToHost, ctx->stream));
    if (idx) HITL_CUDA(cudaMemcpyAsync(idx, ctx->d_idx.p, 4 * nm, cudaMemcpyDeviceToHost, ctx->stream));
  }
  HITL_CUDA(cudaStreamSynchronize(ctx->stream));
  return HITL_OK;
}

// Point indices as 16-bit values (scans hold at most 65534 points, hitl_set_scans): half the bytes of the two largest result arrays.
namespace hitl {
__global__ void pack_u16_kernel(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, uint64_t n, uint32_t* __restrict__ out_a, uint32_t* __restrict__ out_b) {
  // one thread packs two consecutive entries of each array into one 32-bit word
  const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, i = 2 * w;
  if (i >= n) return;
  const uint32_t a0 = a[i], b0 = b[i], a1 = i + 1 < n ? a[i + 1] : 0u, b1 = i + 1 < n ? b[i + 1] : 0u;
  out_a[w] = (a0 & 0xFFFFu) | (a1 << 16);
  out_b[w] = (b0 & 0xFFFFu) | (b1 << 16);
}
}  // namespace hitl
extern "C" int hitl_get_stf16(hitl_ctx* ctx, uint32_t* pair_i, uint32_t* pair_j, uint64_t* pair_off, uint16_t* k, uint16_t* idx) {
  if (!ctx) return HITL_ERR_ARG;
  HITL_DEVICE(ctx);
  if (!ctx->have_stf) return fail(ctx, HITL_ERR_STATE, "hitl_get_stf16: no search result");
  const uint64_t np = ctx->n_pairs, nm = ctx->n_matches;
  if (pair_off && np == 0) pair_off[0] = 0;
  if (np) {
    if (pair_i) HITL_CUDA(cudaMemcpyAsync(pair_i, ctx->d_pair_i.p, 4 * np, cudaMemcpyDeviceToHost, ctx->stream));
    if (pair_j) HITL_CUDA(cudaMemcpyAsync(pair_j, ctx->d_pair_j.p, 4 * np, cudaMemcpyDeviceToHost, ctx->stream));
    if (pair_off) HITL_CUDA(cudaMemcpyAsync(pair_off, ctx->d_pair_off.p, 8 * (np + 1), cudaMemcpyDeviceToHost, ctx->stream));
  }
  if (nm && (k || idx)) {
    const uint64_t words = (nm + 1) / 2;
    HITL_CUDA(ctx->d_pack_k.ensure(words)); HITL_CUDA(ctx->d_pack_idx.ensure(words));   // dedicated staging (no aliasing of search scratch)
    pack_u16_kernel<<<(uint32_t)((words + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_k.p, ctx->d_idx.p, nm, ctx->d_pack_k.p, ctx->d_pack_idx.p);
    HITL_LAUNCH_CHECK("pack_u16_kernel");
    if (k) HITL_CUDA(cudaMemcpyAsync(k, ctx->d_pack_k.p, 2 * nm, cudaMemcpyDeviceToHost, ctx->stream));
    if (idx) HITL_CUDA(cudaMemcpyAsync(idx, ctx->d_pack_idx.p, 2 * nm, cudaMemcpyDeviceToHost, ctx->stream));
  }
  HITL_CUDA(cudaStreamSynchronize(ctx->stream));
  return HITL_OK;
}

extern "C" int hitl_get_stf_work(hitl_ctx* ctx, uint64_t* work_per_pose) {
  if (!ctx) return HITL_ERR_ARG;
  HITL_DEVICE(ctx);
  if (!ctx->have_stf || !ctx->d_pose_work.p) return fail(ctx, HITL_ERR_STATE, "hitl_get_stf_work: no search result");
  if (!work_per_pose && ctx->n_poses) return fail(ctx, HITL_ERR_ARG, "hitl_get_stf_work: null output");
  if (ctx->n_poses) HITL_CUDA(cudaMemcpyAsync(work_per_pose, ctx->d_pose_work.p, 8 * (size_t)ctx->n_poses, cudaMemcpyDeviceToHost, ctx->stream));
  HITL_CUDA(cudaStreamSynchronize(ctx->stream));
  return HITL_OK;
}

extern "C" int hitl_debug_tile_work(hitl_ctx* ctx, uint32_t cap, uint32_t* work_out, uint32_t* n_tiles_out) {
  if (!ctx) return HITL_ERR_ARG;
  HITL_DEVICE(ctx);
  if (!ctx->have_stf || !ctx->d_tile_work.p) return fail(ctx, HITL_ERR_STATE, "hitl_debug_tile_work: no search result");
  if (n_tiles_out) *n_tiles_out = ctx->n_tiles;
  const uint32_t n = std::min(cap, ctx->n_tiles);
  if (n && work_out) HITL_CUDA(cudaMemcpyAsync(work_out, ctx->d_tile_work.p, 4 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  HITL_CUDA(cudaStreamSynchronize(ctx->stream));
  return HITL_OK;
}

// Tile descriptors of the current tiling (host tables) and the "open points" count the last search measured per tile.
extern "C" int hitl_debug_tile_desc(hitl_ctx* ctx, uint32_t cap, uint32_t* scan, uint32_t* k0_len, uint32_t* jlo, uint32_t* jhi, uint32_t* open) {
  if (!ctx) return HITL_ERR_ARG;
  HITL_DEVICE(ctx);
  const uint32_t n = std::min<uint32_t>(cap, (uint32_t)ctx->h_tile_scan.size());
  for (uint32_t t = 0; t < n; ++t) {
    if (scan) scan[t] = ctx->h_tile_scan[t];
    if (k0_len) k0_len[t] = ctx->h_tile_kl[t];
    if (jlo) jlo[t] = ctx->h_tile_jlo[t];
    if (jhi) jhi[t] = ctx->h_tile_jhi[t];
  }
  if (open && n && ctx->d_tile_open.p) {
    HITL_CUDA(cudaMemcpyAsync(open, ctx->d_tile_open.p, 4 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    HITL_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return HITL_OK;
}

extern "C" int hitl_find_vo(hitl_ctx* ctx, const double* pose_array, int32_t min_pose, int32_t max_pose, const hitl_stf_opts* o,
                            uint64_t* n_out) {
  if (!ctx) return HITL_ERR_ARG;
  HITL_DEVICE(ctx);
  if (!pose_array || !o || min_pose < 0) return fail(ctx, HITL_ERR_ARG, "hitl_find_vo: bad argument");
  if (!ctx->have_trees) return fail(ctx, HITL_ERR_STATE, "hitl_find_vo: scans/trees not set");
  ctx->n_vo = 0;
  if (n_out) *n_out = 0;
  const uint64_t poses_end = std::min<uint64_t>((uint64_t)((int64_t)max_pose + 1 < 0 ? 0 : (int64_t)max_pose + 1), ctx->n_poses);
  if ((int64_t)poses_end < (int64_t)min_pose + 1 || poses_end < 2 || (uint64_t)min_pose + 1 >= poses_end) return HITL_OK;
  int rc = upload_poses_and_prep(ctx, pose_array, o->point_match_threshold);
  if (rc) return rc;
  const uint32_t i_lo = (uint32_t)min_pose, i_hi = (uint32_t)poses_end - 1;   // sources i in [i_lo, i_hi)
  HITL_CUDA(ctx->d_vo_tk.ensure(ctx->n_points)); HITL_CUDA(ctx->d_tile_cnt.ensure(std::max<size_t>(ctx->n_tiles, ctx->n_poses)));
  const int threads = 128;
  vo_search_kernel<<<((size_t)(i_hi - i_lo) * 32 + threads - 1) / threads, threads, 0, ctx->stream>>>(
      ctx->d_pts.p, ctx->d_nrm.p, ctx->d_node_pm.p, ctx->d_node_nn.p, ctx->d_rec.p, ctx->d_src.p, ctx->d_pose.p, i_lo, i_hi, o->point_match_threshold,
      o->min_cosine_angle, ctx->d_vo_tk.p, ctx->d_tile_cnt.p);
  HITL_LAUNCH_CHECK("vo_search_kernel");
  // compaction on the host side of the call (result is (N-1)*P slots; the reference never consumes it)
  const uint32_t p_lo = ctx->h_off[i_lo], p_hi = ctx->h_off[i_hi];
  std::vector<uint32_t> slots(p_hi - p_lo);
  if (p_hi > p_lo) HITL_CUDA(cudaMemcpyAsync(slots.data(), ctx->d_vo_tk.p + p_lo, 4 * (size_t)(p_hi - p_lo), cudaMemcpyDeviceToHost, ctx->stream));
  HITL_CUDA(cudaStreamSynchronize(ctx->stream));
  std::vector<uint32_t> sp, sk, tk;
  for (uint32_t i = i_lo; i < i_hi; ++i)
    for (uint32_t k = ctx->h_off[i]; k < ctx->h_off[i + 1]; ++k)
      if (slots[k - p_lo] != 0xFFFFFFFFu) { sp.push_back(i); sk.push_back(k - ctx->h_off[i]); tk.push_back(slots[k - p_lo]); }
  ctx->n_vo = sp.size();
  HITL_CUDA(ctx->d_vo_sp.ensure(sp.size())); HITL_CUDA(ctx->d_vo_sk.ensure(sp.size())); HITL_CUDA(ctx->d_vo_tk.ensure(std::max<size_t>(sp.size(), ctx->n_points)));
  if (!sp.empty()) {
    HITL_CUDA(cudaMemcpy(ctx->d_vo_sp.p, sp.data(), 4 * sp.size(), cudaMemcpyHostToDevice));
    HITL_CUDA(cudaMemcpy(ctx->d_vo_sk.p, sk.data(), 4 * sp.size(), cudaMemcpyHostToDevice));
    HITL_CUDA(cudaMemcpy(ctx->d_vo_tk.p, tk.data(), 4 * sp.size(), cudaMemcpyHostToDevice));
  }
  if (n_out) *n_out = ctx->n_vo;
  return HITL_OK;
}

extern "C" int hitl_get_vo(hitl_ctx* ctx, uint32_t* source_pose, uint32_t* source_point, uint32_t* target_point) {
  if (!ctx) return HITL_ERR_ARG;
  HITL_DEVICE(ctx);
  if (ctx->n_vo == 0) return HITL_OK;
  if (source_pose) HITL_CUDA(cudaMemcpy(source_pose, ctx->d_vo_sp.p, 4 * ctx->n_vo, cudaMemcpyDeviceToHost));
  if (source_point) HITL_CUDA(cudaMemcpy(source_point, ctx->d_vo_sk.p, 4 * ctx->n_vo, cudaMemcpyDeviceToHost));
  if (target_point) HITL_CUDA(cudaMemcpy(target_point, ctx->d_vo_tk.p, 4 * ctx->n_vo, cudaMemcpyDeviceToHost));
  return HITL_OK;
}
