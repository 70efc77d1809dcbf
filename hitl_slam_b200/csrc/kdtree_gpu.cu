// kdtree_gpu.cu — KD-tree construction on the device, all scans at once, with the reference's tree shape.
//
// Replaces BuildKDTrees / KDTree<float,2>::BuildKDTree / GetSplittingPlane
//   (human_in_the_loop_slam/JointOptimization.cpp:514-537, perception_tools/kdtree.cpp:37-69, :106-139),
// which the reference runs serially over the scans on the host once per session.
//
// Shape contract (the search is a pruned, lossy walk whose answer depends on the exact tree): per node
//   split dimension = the larger of the two SEQUENTIAL float sums of squared deviations from the sequential float
//                     mean, taken over the node's points in their current order (dimension 0 on ties);
//   std::sort of the node's points on that coordinate with a strict `<` (not stable);
//   node = element n/2, children = the two sorted halves in their sorted order, emitted in preorder.
//
// Level-synchronous build.  A "segment" is the point range of one pending node; all segments of one tree level, of all
// scans, are processed together:
//   kd_dim_kernel    one thread per segment: the four sequential sums (this is the only part that must be serial —
//                    float addition is not associative — and a level costs one pass over each segment, so the whole
//                    build is ~2 x the largest scan of dependent additions);
//   kd_key_kernel    one thread per point: its segment (bisection) and its sort key;
//   cub::DeviceSegmentedSort  all segments at once;
//   kd_tie_kernel    one thread per point: flags segments that contain two equal keys.  With distinct keys every
//                    correct sort returns the same order; with equal keys libstdc++'s introsort decides, so
//   kd_emit_kernel   (one thread per segment) re-sorts flagged segments from their pre-sort order with the step-for-step
//                    restatement of std::sort in stdsort_exact.h, then emits the node and the two child segments;
//   a scan compacts the children into the next level's segment list (order preserved, so bisection keeps working).
// Compiled with --fmad=false.
#include <cub/device/device_scan.cuh>
#include <cub/device/device_segmented_sort.cuh>
#include "hitl_internal.h"
#include "hitl_math.h"
#include "stdsort_exact.h"
#include <stdlib.h>
#include <time.h>

namespace hitl {
namespace {

struct SegArrays { uint32_t* b; uint32_t* e; uint32_t* gpos; uint32_t* base; };

__global__ void kd_init_kernel(const uint32_t* __restrict__ off, uint32_t n_poses, uint64_t m, SegArrays s, uint32_t* __restrict__ valid, uint32_t* __restrict__ ids) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < m) ids[t] = (uint32_t)t;
  if (t < n_poses) {
    s.b[t] = off[t]; s.e[t] = off[t + 1]; s.gpos[t] = off[t]; s.base[t] = off[t];
    valid[t] = off[t + 1] > off[t];
  }
}

constexpr uint32_t kWarpSegment = 32;   // segments of at least this many points are summed by a whole warp (coalesced loads)

// Large segments: one warp per segment.  The lanes fetch 32 points at a time; every lane then performs the SAME sequential
// additions in point order (values come by shuffle), so the float sums are those of the serial loop.
__global__ void kd_dim_warp_kernel(SegArrays s, uint32_t nseg, const uint32_t* __restrict__ ids, const float2* __restrict__ pts, uint8_t* __restrict__ dims) {
  const uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (t >= nseg) return;
  const uint32_t b = s.b[t], e = s.e[t];
  if (e - b < kWarpSegment) return;
  float mx = 0.0f, my = 0.0f;
  for (uint32_t k0 = b; k0 < e; k0 += 32) {
    float2 p = make_float2(0.f, 0.f);
    if (k0 + lane < e) p = __ldg(pts + ids[k0 + lane]);
    const uint32_t cnt = min(32u, e - k0);
    for (uint32_t j = 0; j < cnt; ++j) { mx = fadd(mx, __shfl_sync(0xffffffffu, p.x, j)); my = fadd(my, __shfl_sync(0xffffffffu, p.y, j)); }
  }
  const float n = (float)(e - b);
  mx = __fdiv_rn(mx, n); my = __fdiv_rn(my, n);
  float dx = 0.0f, dy = 0.0f;
  for (uint32_t k0 = b; k0 < e; k0 += 32) {
    float2 p = make_float2(0.f, 0.f);
    if (k0 + lane < e) p = __ldg(pts + ids[k0 + lane]);
    const float ax = fsub(p.x, mx), ay = fsub(p.y, my);
    const float sx = fmul(ax, ax), sy = fmul(ay, ay);
    const uint32_t cnt = min(32u, e - k0);
    for (uint32_t j = 0; j < cnt; ++j) { dx = fadd(dx, __shfl_sync(0xffffffffu, sx, j)); dy = fadd(dy, __shfl_sync(0xffffffffu, sy, j)); }
  }
  int dim = 0;
  float best = 0.0f;
  if (dx > best) { dim = 0; best = dx; }
  if (dy > best) { dim = 1; best = dy; }
  if (lane == 0) dims[t] = (uint8_t)dim;
}

__global__ void kd_dim_kernel(SegArrays s, uint32_t nseg, const uint32_t* __restrict__ ids, const float2* __restrict__ pts, uint8_t* __restrict__ dims) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nseg) return;
  const uint32_t b = s.b[t], e = s.e[t];
  if (e - b >= kWarpSegment) return;                 // kd_dim_warp_kernel's
  float mx = 0.0f, my = 0.0f;
#pragma unroll 4
  for (uint32_t k = b; k < e; ++k) { const float2 p = __ldg(pts + ids[k]); mx = fadd(mx, p.x); my = fadd(my, p.y); }
  const float n = (float)(e - b);
  mx = __fdiv_rn(mx, n); my = __fdiv_rn(my, n);
  float dx = 0.0f, dy = 0.0f;
#pragma unroll 4
  for (uint32_t k = b; k < e; ++k) {
    const float2 p = __ldg(pts + ids[k]);
    const float ax = fsub(p.x, mx), ay = fsub(p.y, my);
    dx = fadd(dx, fmul(ax, ax)); dy = fadd(dy, fmul(ay, ay));
  }
  int dim = 0;
  float best = 0.0f;
  if (dx > best) { dim = 0; best = dx; }
  if (dy > best) { dim = 1; best = dy; }
  dims[t] = (uint8_t)dim;
}

__device__ __forceinline__ uint32_t find_segment(const uint32_t* __restrict__ seg_b, uint32_t nseg, uint32_t k) {
  uint32_t lo = 0, hi = nseg;                       // last segment with b <= k (segments are sorted and disjoint)
  while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (seg_b[mid] <= k) lo = mid; else hi = mid; }
  return lo;
}

__global__ void kd_key_kernel(uint64_t m, SegArrays s, uint32_t nseg, const uint32_t* __restrict__ ids, const float2* __restrict__ pts,
                              const uint8_t* __restrict__ dims, float* __restrict__ keys, uint32_t* __restrict__ seg_of) {
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= m) return;
  const uint32_t sg = find_segment(s.b, nseg, (uint32_t)k);
  if (k < s.b[sg] || k >= s.e[sg]) { seg_of[k] = 0xFFFFFFFFu; keys[k] = 0.0f; return; }   // already a node
  const float2 p = __ldg(pts + ids[k]);
  keys[k] = dims[sg] ? p.y : p.x;
  seg_of[k] = sg;
}

__global__ void kd_tie_kernel(uint64_t m, const float* __restrict__ keys_sorted, const uint32_t* __restrict__ seg_of, uint32_t* __restrict__ tie) {
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k + 1 >= m) return;
  const uint32_t sg = seg_of[k];
  if (sg != 0xFFFFFFFFu && seg_of[k + 1] == sg && keys_sorted[k] == keys_sorted[k + 1]) tie[sg] = 1u;   // also -0 == +0
}

struct CoordKey {
  const float2* pts; int dim;
  __host__ __device__ __forceinline__ float operator()(uint32_t id) const { const float2 p = pts[id]; return dim ? p.y : p.x; }
};

__global__ void kd_emit_kernel(SegArrays s, uint32_t nseg, const uint32_t* __restrict__ ids_before, uint32_t* __restrict__ ids_sorted,
                               const float2* __restrict__ pts, const float2* __restrict__ nrm, const uint8_t* __restrict__ dims,
                               const uint32_t* __restrict__ tie, float4* __restrict__ node_pm, float2* __restrict__ node_nn, SegArrays c,
                               uint32_t* __restrict__ child_valid, unsigned long long* __restrict__ n_exact) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nseg) return;
  const uint32_t b = s.b[t], e = s.e[t], n = e - b, gpos = s.gpos[t], base = s.base[t];
  const int dim = dims[t];
  if (tie[t]) {
    // equal keys: the order is whatever libstdc++'s introsort makes of the PRE-sort order
    for (uint32_t k = b; k < e; ++k) ids_sorted[k] = ids_before[k];
    CoordKey key{pts, dim};
    stdsort::sort_ids(ids_sorted + b, (int64_t)n, key);
    atomicAdd(n_exact, 1ull);
  }
  const uint32_t half = n / 2, mid = b + half;
  const uint32_t id = ids_sorted[mid];
  const float2 p = pts[id];
  node_pm[gpos] = make_float4(p.x, p.y, __uint_as_float(((id - base) & 0x7FFFFFFFu) | (dim ? 0x80000000u : 0u)), 0.0f);
  node_nn[gpos] = nrm[id];
  const bool has_l = half > 0, has_r = mid + 1 < e;
  c.b[2 * t] = b; c.e[2 * t] = mid; c.gpos[2 * t] = gpos + 1; c.base[2 * t] = base; child_valid[2 * t] = has_l;
  c.b[2 * t + 1] = mid + 1; c.e[2 * t + 1] = e; c.gpos[2 * t + 1] = gpos + 1 + half; c.base[2 * t + 1] = base; child_valid[2 * t + 1] = has_r;
}

__global__ void kd_compact_kernel(SegArrays in, const uint32_t* __restrict__ valid, const uint32_t* __restrict__ pos, uint32_t n_in, SegArrays out,
                                  uint32_t* __restrict__ total) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_in) return;
  if (valid[t]) { const uint32_t o = pos[t]; out.b[o] = in.b[t]; out.e[o] = in.e[t]; out.gpos[o] = in.gpos[t]; out.base[o] = in.base[t]; }
  if (t == n_in - 1) *total = pos[t] + valid[t];
}

// One stream-ordered allocation (cudaMallocAsync / cudaFreeAsync: no device-wide synchronisation) carved into the work buffers.
struct Slab {
  cudaStream_t st = nullptr; char* base = nullptr; size_t used = 0, cap = 0;
  ~Slab() { if (base) cudaFreeAsync(base, st); }
  template <typename T> T* take(size_t n) { used = (used + 255) & ~(size_t)255; T* p = reinterpret_cast<T*>(base + used); used += n * sizeof(T); return p; }
};

}  // namespace

// Builds every scan's tree straight into the resident SoA node arrays.  n_exact_out (optional) = number of segments that
// needed the exact std::sort emulation (segments with equal keys).
int build_kdtrees_device(hitl_ctx* ctx, uint64_t* n_exact_out) {
  const uint64_t m = ctx->n_points;
  const uint32_t np = ctx->n_poses;
  HITL_CUDA(ctx->d_node_pm.ensure(m)); HITL_CUDA(ctx->d_node_nn.ensure(m));
  if (n_exact_out) *n_exact_out = 0;
  if (m == 0) return HITL_OK;
  if (m >= 0x7FFFFFFFull) return fail(ctx, HITL_ERR_ARG, "hitl_build_kdtrees: more than 2^31 points");
  // the number of segments of a level never exceeds the number of points still unplaced, nor 2 x the previous level
  const size_t seg_cap = (size_t)std::max<uint64_t>(m, np) + 2;
  // all work buffers come from ONE stream-ordered allocation (cudaMallocAsync / cudaFreeAsync: no device-wide sync)
  Slab slab;
  slab.st = ctx->stream;
  size_t scan_bytes = 0, sort_bytes = 0;
  HITL_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)(2 * seg_cap), ctx->stream));
  HITL_CUDA(cub::DeviceSegmentedSort::SortPairs(nullptr, sort_bytes, (const float*)nullptr, (float*)nullptr, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)m,
                                                (int)std::min<uint64_t>(seg_cap, m), (uint32_t*)nullptr, (uint32_t*)nullptr, ctx->stream));
  const size_t tmp_bytes = std::max(scan_bytes, sort_bytes) + 256;
  slab.cap = 4 * 4 * seg_cap + 4 * 8 * seg_cap + 2 * 8 * seg_cap + 5 * 4 * m + 4 * seg_cap + seg_cap + tmp_bytes + 64 * 256;
  HITL_CUDA(cudaMallocAsync((void**)&slab.base, slab.cap, ctx->stream));
  struct { uint32_t* p; } segA[4], segB[4], valid, pos, ids_a, ids_b, seg_of, tie, total;
  struct { float* p; } key_in, key_out;
  struct { uint8_t* p; } dims, tmp;
  struct { unsigned long long* p; } n_exact;
  for (int i = 0; i < 4; ++i) { segA[i].p = slab.take<uint32_t>(seg_cap); segB[i].p = slab.take<uint32_t>(2 * seg_cap); }
  valid.p = slab.take<uint32_t>(2 * seg_cap); pos.p = slab.take<uint32_t>(2 * seg_cap);
  ids_a.p = slab.take<uint32_t>(m); ids_b.p = slab.take<uint32_t>(m); seg_of.p = slab.take<uint32_t>(m);
  key_in.p = slab.take<float>(m); key_out.p = slab.take<float>(m);
  tie.p = slab.take<uint32_t>(seg_cap); dims.p = slab.take<uint8_t>(seg_cap); total.p = slab.take<uint32_t>(1);
  n_exact.p = slab.take<unsigned long long>(1); tmp.p = slab.take<uint8_t>(tmp_bytes);
  if (slab.used > slab.cap) return fail(ctx, HITL_ERR_STATE, "hitl_build_kdtrees: work slab too small");
  HITL_CUDA(cudaMemsetAsync(n_exact.p, 0, 8, ctx->stream));
  SegArrays cur{segA[0].p, segA[1].p, segA[2].p, segA[3].p}, nxt{segB[0].p, segB[1].p, segB[2].p, segB[3].p};
  const int T = 256;
  auto blocks = [&](uint64_t n) { return (uint32_t)((n + T - 1) / T); };
  // level 0: one segment per non-empty scan (compacted from the per-scan list)
  kd_init_kernel<<<blocks(std::max<uint64_t>(m, np)), T, 0, ctx->stream>>>(ctx->d_off.p, np, m, nxt, valid.p, ids_a.p);
  HITL_LAUNCH_CHECK("kd_init_kernel");
  uint32_t n_in = np, nseg = 0;
  uint32_t* ids_cur = ids_a.p; uint32_t* ids_oth = ids_b.p;
  for (int level = 0; level < 40; ++level) {
    // compact the candidate list (level 0: scans; later: children) into `cur`
    size_t sb = tmp_bytes;
    HITL_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, sb, valid.p, pos.p, (int)n_in, ctx->stream));
    kd_compact_kernel<<<blocks(n_in), T, 0, ctx->stream>>>(nxt, valid.p, pos.p, n_in, cur, total.p);
    HITL_LAUNCH_CHECK("kd_compact_kernel");
    HITL_CUDA(cudaMemcpyAsync(&nseg, total.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    HITL_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->launches += 1;
    if (getenv("HITL_KD_DEBUG")) {
      static double t_prev = 0;
      struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
      const double now = ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
      fprintf(stderr, "kd level %d: n_in %u nseg %u  (+%.2f ms)\n", level, n_in, nseg, t_prev ? now - t_prev : 0.0);
      t_prev = now;
    }
    if (nseg == 0) break;
    if (nseg > seg_cap) return fail(ctx, HITL_ERR_STATE, "hitl_build_kdtrees: segment list overflow");
    if ((ctx->max_scan >> level) + 1 >= kWarpSegment) {          // a level-L segment holds at most max_scan / 2^L points
      kd_dim_warp_kernel<<<(uint32_t)(((uint64_t)nseg * 32 + 127) / 128), 128, 0, ctx->stream>>>(cur, nseg, ids_cur, ctx->d_pts.p, dims.p);
      HITL_LAUNCH_CHECK("kd_dim_warp_kernel");
    }
    kd_dim_kernel<<<(nseg + 63) / 64, 64, 0, ctx->stream>>>(cur, nseg, ids_cur, ctx->d_pts.p, dims.p);
    HITL_LAUNCH_CHECK("kd_dim_kernel");
    kd_key_kernel<<<blocks(m), T, 0, ctx->stream>>>(m, cur, nseg, ids_cur, ctx->d_pts.p, dims.p, key_in.p, seg_of.p);
    HITL_LAUNCH_CHECK("kd_key_kernel");
    size_t need = 0;
    HITL_CUDA(cub::DeviceSegmentedSort::SortPairs(nullptr, need, key_in.p, key_out.p, ids_cur, ids_oth, (int)m, (int)nseg, cur.b, cur.e, ctx->stream));
    if (need > tmp_bytes) return fail(ctx, HITL_ERR_STATE, "hitl_build_kdtrees: sort workspace too small");
    need = tmp_bytes;
    HITL_CUDA(cub::DeviceSegmentedSort::SortPairs(tmp.p, need, key_in.p, key_out.p, ids_cur, ids_oth, (int)m, (int)nseg, cur.b, cur.e, ctx->stream));
    ctx->launches += 1;
    HITL_CUDA(cudaMemsetAsync(tie.p, 0, 4 * (size_t)nseg, ctx->stream));
    kd_tie_kernel<<<blocks(m), T, 0, ctx->stream>>>(m, key_out.p, seg_of.p, tie.p);
    HITL_LAUNCH_CHECK("kd_tie_kernel");
    kd_emit_kernel<<<(nseg + 63) / 64, 64, 0, ctx->stream>>>(cur, nseg, ids_cur, ids_oth, ctx->d_pts.p, ctx->d_nrm.p, dims.p, tie.p, ctx->d_node_pm.p,
                                                       ctx->d_node_nn.p, nxt, valid.p, n_exact.p);
    HITL_LAUNCH_CHECK("kd_emit_kernel");
    n_in = 2 * nseg;
    std::swap(ids_cur, ids_oth);
  }
  if (nseg != 0) return fail(ctx, HITL_ERR_STATE, "hitl_build_kdtrees: tree deeper than 40 levels");
  unsigned long long ne = 0;
  HITL_CUDA(cudaMemcpyAsync(&ne, n_exact.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
  HITL_CUDA(cudaStreamSynchronize(ctx->stream));
  if (n_exact_out) *n_exact_out = ne;
  return HITL_OK;
}

}  // namespace hitl
