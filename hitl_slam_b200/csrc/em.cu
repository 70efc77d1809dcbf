// em.cu — world-frame clouds and the EM point-to-feature assignment on sm_100a.
//
// Replaces (paths relative to HitL-SLAM/src/):
//   HitLSLAM::transformPointCloudsToWorldFrame    human_in_the_loop_slam/HitLSLAM.cpp:245-254
//   E-step of EMInput::AutomaticEndpointAdjustment  human_in_the_loop_slam/EMinput.cpp:207-218
//       with Eigen::DistanceToLineSegment         shared/math/eigen_helper.h:66-81
//   EMInput::EstablishObservationSets             EMinput.cpp:281-323  with distToLineSeg :269-279
//
// Both are HBM streams: 8 B read per point, a few bytes written per inlier.  Compiled with
// --fmad=false; float expressions follow Eigen's evaluation order so flags are bit-exact.
#include <float.h>
#include <string.h>
#include <algorithm>
#include "hitl_internal.h"
#include "hitl_math.h"

namespace hitl {

// ---- K4: world = Rotation2Df(theta) * p + t ------------------------------------------------
__global__ void world_transform_kernel(const float2* __restrict__ pts, const uint32_t* __restrict__ off, const float* __restrict__ poses,
                                       uint32_t n_poses, float2* __restrict__ world) {
  // one warp per scan: the pose's sin/cos is computed once per lane (uniform), points stream coalesced
  const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n_poses) return;
  const float th = poses[3 * w + 2];
  const float s = sinf_rn(th), c = cosf_rn(th);
  const float tx = poses[3 * w], ty = poses[3 * w + 1];
  for (uint32_t k = off[w] + lane; k < off[w + 1]; k += 32) {
    const float2 p = pts[k];
    float2 o;
    o.x = fadd(dot2(c, p.x, -s, p.y), tx);
    o.y = fadd(dot2(s, p.x, c, p.y), ty);
    world[k] = o;
  }
}

// ---- distance helpers ------------------------------------------------------------------------
struct Seg { float p0x, p0y, p1x, p1y, dirx, diry; };   // dir = (p1 - p0).normalized()

// eigen_helper.h:66-81 (t in metres compared with 1.0 — kept)
__device__ __forceinline__ float distance_to_line_segment(const Seg& s, float px, float py) {
  const float ax = fsub(px, s.p0x), ay = fsub(py, s.p0y);
  const float t = dot2(ax, s.dirx, ay, s.diry);
  if (t < 0.0f) return sqrtf(dot2(ax, ax, ay, ay));
  if (t > 1.0f) { const float bx = fsub(px, s.p1x), by = fsub(py, s.p1y); return sqrtf(dot2(bx, bx, by, by)); }
  return fabsf(dot2(-s.diry, ax, s.dirx, ay));
}
// EMinput.cpp:269-279
struct Seg2 { float p1x, p1y, p2x, p2y, dx, dy, dd; };   // d = p2 - p1, dd = d.d
__device__ __forceinline__ float dist_to_line_seg(const Seg2& s, float px, float py) {
  const float ax = fsub(px, s.p1x), ay = fsub(py, s.p1y);
  const float t = dot2(ax, s.dx, ay, s.dy) / s.dd;
  if (t < 0.0f) return sqrtf(dot2(ax, ax, ay, ay));
  if (t > 1.0f) { const float bx = fsub(px, s.p2x), by = fsub(py, s.p2y); return sqrtf(dot2(bx, bx, by, by)); }
  const float qx = fadd(s.p1x, fmul(t, s.dx)), qy = fadd(s.p1y, fmul(t, s.dy));
  const float cx = fsub(px, qx), cy = fsub(py, qy);
  return sqrtf(dot2(cx, cx, cy, cy));
}

// ---- K2a: E-step inliers with single-pass ordered compaction (decoupled look-back) ------------
constexpr int kEmThreads = 256;
constexpr int kEmPerThread = 8;                       // 4 x float4 loads = 8 points per thread
constexpr int kEmChunk = kEmThreads * kEmPerThread;   // 2048 points per CTA step
// chunk state word: bits 63..62 = status (0 none, 1 aggregate, 2 inclusive prefix), low 62 bits = value
constexpr unsigned long long kStAgg = 1ull << 62, kStPre = 2ull << 62, kStMask = 3ull << 62;

// Stroke of a chained round (hitl_em_refit_chain): the four floats the previous round's fit left on the device.  Same operations,
// same order and same roundings as the host's make_seg (no contraction on either side), so a chained round sees the bits a host
// round trip would have produced.
__device__ __forceinline__ Seg seg_from_device(const float* __restrict__ s) {
  Seg o;
  o.p0x = s[0]; o.p0y = s[1]; o.p1x = s[2]; o.p1y = s[3];
  const float dx = fsub(s[2], s[0]), dy = fsub(s[3], s[1]);
  const float z = dot2(dx, dx, dy, dy);
  if (z > 0.0f) { const float n = __fsqrt_rn(z); o.dirx = __fdiv_rn(dx, n); o.diry = __fdiv_rn(dy, n); }
  else { o.dirx = dx; o.diry = dy; }
  return o;
}

// The E-step's reach around a stroke, as a box: distance_to_line_segment() measures from p0 for t < 0, from p1 for t > 1 (t in
// METRES along the unit direction — the reference's quirk, kept) and from the carrier line in between, so every point closer than
// thr lies within thr of p0, of p1 or of the piece p0 .. p0 + 1 m * dir.  A chunk whose points' bounding box misses this box
// (inflated by thr and a float-rounding margin far above the few ulps the distance evaluation can lose) has no inlier.
// ok = false (degenerate or non-finite stroke: with dir = 0 EVERY point is at distance 0) disables the cull.
struct ReachBox { float x0, y0, x1, y1; bool ok; };
__device__ __forceinline__ ReachBox stroke_reach(const Seg& s, double thr) {
  ReachBox b;
  const float qx = s.p0x + s.dirx, qy = s.p0y + s.diry;
  const float len2 = s.dirx * s.dirx + s.diry * s.diry;
  const float lox = fminf(fminf(s.p0x, s.p1x), qx), hix = fmaxf(fmaxf(s.p0x, s.p1x), qx);
  const float loy = fminf(fminf(s.p0y, s.p1y), qy), hiy = fmaxf(fmaxf(s.p0y, s.p1y), qy);
  const float big = fmaxf(fmaxf(fabsf(lox), fabsf(hix)), fmaxf(fabsf(loy), fabsf(hiy)));
  const float m = (float)thr + 1e-3f + 1e-5f * big;
  b.x0 = lox - m; b.y0 = loy - m; b.x1 = hix + m; b.y1 = hiy + m;
  b.ok = len2 > 0.5f && len2 < 2.0f && thr >= 0.0 && thr < 1e6 && isfinite(b.x0) && isfinite(b.y0) && isfinite(b.x1) && isfinite(b.y1);
  return b;
}

// Chunk look-back of the ordered compaction, by the 32 lanes of warp 0 together: publishes this chunk's count, returns (on every
// lane) the number of inliers before the chunk.  All chunks of a wave publish their aggregate at about the same moment, so a
// single thread walking back one predecessor per L2 round trip needs ~sqrt(2 * chunk) trips before it meets an inclusive prefix;
// the warp inspects 32 predecessors per trip.  Chunks are taken by ticket, so every predecessor is already running: no deadlock.
__device__ __forceinline__ unsigned long long em_chunk_prefix(unsigned long long* state, uint32_t chunk, uint32_t block_total, uint32_t lane) {
  if (chunk == 0) {
    if (lane == 0) atomicExch(&state[0], kStPre | block_total);
    return 0;
  }
  if (lane == 0) atomicExch(&state[chunk], kStAgg | block_total);
  unsigned long long prefix = 0;
  for (int64_t hi = (int64_t)chunk - 1;; hi -= 32) {
    const int64_t c = hi - (int64_t)lane;
    unsigned long long v = kStPre;                                       // before chunk 0: an empty inclusive prefix
    if (c >= 0) do { v = atomicAdd(&state[c], 0ull); } while ((v & kStMask) == 0);
    const uint32_t pre = __ballot_sync(0xffffffffu, (v & kStMask) == kStPre);
    const uint32_t first = pre ? (uint32_t)(__ffs(pre) - 1) : 31u;       // nearest predecessor that already knows its inclusive prefix
    unsigned long long val = lane <= first ? (v & ~kStMask) : 0ull;
    for (int o = 16; o; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
    prefix += val;
    if (pre) break;
  }
  if (lane == 0) atomicExch(&state[chunk], kStPre | (prefix + block_total));
  return prefix;
}

// box_mode: 0 = plain pass; 1 = plain pass that also records every chunk's bounding box in `boxes`; 2 = chunks whose recorded box
// misses the stroke's reach are counted as empty without being read (the boxes belong to the resident world clouds).
template <int box_mode>
__global__ void __launch_bounds__(kEmThreads) em_inliers_kernel(const float2* __restrict__ world, const uint32_t* __restrict__ off,
                                                                uint32_t n_poses, uint64_t n_points, Seg seg_arg, const float* __restrict__ seg_dev,
                                                                double thr, unsigned long long* state, uint32_t* ticket, uint64_t cap,
                                                                uint32_t* __restrict__ out_pose, uint32_t* __restrict__ out_idx,
                                                                float2* __restrict__ out_xy, unsigned long long* total, float4* boxes) {
  __shared__ uint32_t s_chunk;
  __shared__ uint32_t s_warp[kEmThreads / 32];
  __shared__ float s_box[4][kEmThreads / 32];
  __shared__ unsigned long long s_prefix;
  const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t n_chunks = (uint32_t)((n_points + kEmChunk - 1) / kEmChunk);
  const Seg seg = seg_dev ? seg_from_device(seg_dev) : seg_arg;
  ReachBox reach;
  reach.ok = false;
  if (box_mode == 2) reach = stroke_reach(seg, thr);
  for (;;) {
    if (threadIdx.x == 0) s_chunk = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t chunk = s_chunk;
    if (chunk >= n_chunks) return;
    if (reach.ok) {
      const float4 b = boxes[chunk];                       // (min x, min y, max x, max y) of the chunk's points
      if (b.x > reach.x1 || b.z < reach.x0 || b.y > reach.y1 || b.w < reach.y0) {
        if (w == 0) {
          const unsigned long long prefix = em_chunk_prefix(state, chunk, 0u, lane);
          if (lane == 0 && chunk == n_chunks - 1) *total = prefix;
        }
        __syncthreads();   // s_chunk reuse
        continue;
      }
    }
    // thread owns 8 consecutive points: keeps (pose, idx) order inside the thread, lanes ascending
    const uint64_t base = (uint64_t)chunk * kEmChunk + (uint64_t)threadIdx.x * kEmPerThread;
    float2 p[kEmPerThread];
    if (base + kEmPerThread <= n_points) {
      const float4* src = reinterpret_cast<const float4*>(world + base);
#pragma unroll
      for (int q = 0; q < kEmPerThread / 2; ++q) { const float4 v = __ldg(src + q); p[2 * q] = make_float2(v.x, v.y); p[2 * q + 1] = make_float2(v.z, v.w); }
    } else {
#pragma unroll
      for (int q = 0; q < kEmPerThread; ++q) p[q] = base + q < n_points ? world[base + q] : make_float2(FLT_MAX, FLT_MAX);
    }
    uint32_t mask = 0;
#pragma unroll
    for (int q = 0; q < kEmPerThread; ++q) {
      const bool in = base + q < n_points && (double)distance_to_line_segment(seg, p[q].x, p[q].y) < thr;
      mask |= (uint32_t)in << q;
    }
    const uint32_t cnt = __popc(mask);
    // block exclusive scan of cnt
    uint32_t x = cnt;
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= (uint32_t)o) x += y; }
    if (lane == 31) s_warp[w] = x;
    if (box_mode == 1) {
      // bounding box of the chunk's real points (the tail padding is excluded; fminf / fmaxf drop NaNs, which are never inliers)
      float bx0 = FLT_MAX, by0 = FLT_MAX, bx1 = -FLT_MAX, by1 = -FLT_MAX;
#pragma unroll
      for (int q = 0; q < kEmPerThread; ++q)
        if (base + q < n_points) { bx0 = fminf(bx0, p[q].x); by0 = fminf(by0, p[q].y); bx1 = fmaxf(bx1, p[q].x); by1 = fmaxf(by1, p[q].y); }
      for (int o = 16; o; o >>= 1) {
        bx0 = fminf(bx0, __shfl_xor_sync(0xffffffffu, bx0, o)); by0 = fminf(by0, __shfl_xor_sync(0xffffffffu, by0, o));
        bx1 = fmaxf(bx1, __shfl_xor_sync(0xffffffffu, bx1, o)); by1 = fmaxf(by1, __shfl_xor_sync(0xffffffffu, by1, o));
      }
      if (lane == 0) { s_box[0][w] = bx0; s_box[1][w] = by0; s_box[2][w] = bx1; s_box[3][w] = by1; }
    }
    __syncthreads();
    uint32_t wbase = 0, block_total = 0;
    for (int q = 0; q < kEmThreads / 32; ++q) { const uint32_t v = s_warp[q]; if (q < (int)w) wbase += v; block_total += v; }
    const uint32_t excl = wbase + x - cnt;
    // publish aggregate, look back for the exclusive prefix of this chunk
    if (w == 0) {
      const unsigned long long prefix = em_chunk_prefix(state, chunk, block_total, lane);
      if (lane == 0) {
        s_prefix = prefix;
        if (chunk == n_chunks - 1) *total = prefix + block_total;
        if (box_mode == 1) {
          float4 b = make_float4(s_box[0][0], s_box[1][0], s_box[2][0], s_box[3][0]);
          for (int q = 1; q < kEmThreads / 32; ++q) { b.x = fminf(b.x, s_box[0][q]); b.y = fminf(b.y, s_box[1][q]); b.z = fmaxf(b.z, s_box[2][q]); b.w = fmaxf(b.w, s_box[3][q]); }
          boxes[chunk] = b;
        }
      }
    }
    __syncthreads();
    unsigned long long o = s_prefix + excl;
    if (mask && out_pose) {
#pragma unroll
      for (int q = 0; q < kEmPerThread; ++q) {
        if (mask & (1u << q)) {
          if (o < cap) {
            const uint32_t g = (uint32_t)(base + q);
            // pose of point g: last scan with off[pose] <= g (upper_bound - 1 skips empty scans correctly)
            uint32_t lo = 0, hi = n_poses;
            while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (off[mid] <= g) lo = mid; else hi = mid; }
            out_pose[o] = lo; out_idx[o] = g - off[lo];
            if (out_xy) out_xy[o] = p[q];
          }
          ++o;
        }
      }
    }
    __syncthreads();   // s_chunk / s_warp reuse
  }
}

// ---- K2c: M-step on the device — SegFitEM's one-parameter Levenberg-Marquardt over the resident inliers ----------------
// EMinput.cpp:107-191: the stroke keeps its midpoint and length; the only unknown is its direction angle theta, fitted to the
// E-step inliers by least squares on the point-to-segment distance (segDistResidualEM, three cases on the projection
// parameter t) with Ceres' LM at its defaults (<= 25 iterations, theta_0 = acos(|dx| / length)).  The whole loop runs in ONE
// cooperative launch: every evaluation is a grid-wide reduction of sum r^2, sum r dr/dtheta, sum (dr/dtheta)^2 in a fixed
// order (per-thread strided sums -> shuffle tree -> per-CTA partials -> one grid barrier -> every CTA adds the partials in
// CTA order), and every CTA runs the same scalar trust-region logic on the same sums, so no CTA waits for a host round trip
// or for a designated leader.  The control flow restates host/ceres_solver.cpp::Solve for one scalar parameter (Jacobi
// scaling from the first Jacobian, (H + clamp(H)/radius) step, model_change / rho acceptance, radius update
// radius / max(1/3, 1 - (2 rho - 1)^3), halving 2, 4, 8 ... on rejection, function / gradient / parameter tolerances).
// The derivative is analytic where the reference differentiates a Jet: same value up to rounding (test: theta within 1e-9).
struct FitSums { double rr, jr, jj; };
struct FitResult { double theta, cost0, cost; unsigned long long n; int iterations, evaluations, termination; float seg[4]; };
constexpr int kFitThreads = 256;
constexpr int kFitMaxBlocks = 64;

__device__ __forceinline__ void seg_angle_residual(double px, double py, double cmx, double cmy, double len, double ax, double ay, double* r, double* dr) {
  // e1 = cm + len a, e2 = cm - len a, d = e2 - e1; t = (p - e1).d / d.d
  const double e1x = cmx + len * ax, e1y = cmy + len * ay, e2x = cmx - len * ax, e2y = cmy - len * ay;
  const double dx = e2x - e1x, dy = e2y - e1y;
  const double t = ((px - e1x) * dx + (py - e1y) * dy) / (dx * dx + dy * dy);
  const double apx = -ay, apy = ax;                       // da / dtheta of the unit direction
  double wx, wy, k;                                       // r = |w|, dr/dtheta = k (w . a') / r
  if (t < 0.0) { wx = px - e1x; wy = py - e1y; k = -len; }
  else if (t > 1.0) { wx = px - e2x; wy = py - e2y; k = len; }
  else { wx = px - (e1x + t * dx); wy = py - (e1y + t * dy); k = -len * (1.0 - 2.0 * t); }
  const double rr = sqrt(wx * wx + wy * wy);
  *r = rr;
  *dr = rr > 0.0 ? k * (wx * apx + wy * apy) / rr : 0.0;
}

__global__ void __launch_bounds__(kFitThreads) em_fit_kernel(const float2* __restrict__ xy, const unsigned long long* __restrict__ n_ptr, double p1x, double p1y,
                                                             double p2x, double p2y, const float* __restrict__ seg_dev /* chained round: the stroke, else NULL */,
                                                             int max_iterations, double* partial /* 2 x gridDim.x x 3 */,
                                                             unsigned int* barrier /* 2 words, zero on entry */, FitResult* out) {
  __shared__ double s_red[3][kFitThreads / 32];
  __shared__ FitSums s_tot;
  if (seg_dev) { p1x = (double)seg_dev[0]; p1y = (double)seg_dev[1]; p2x = (double)seg_dev[2]; p2y = (double)seg_dev[3]; }
  const unsigned long long n = *n_ptr;
  const double cmx = (p1x + p2x) / 2.0, cmy = (p1y + p2y) / 2.0;
  const double hy = sqrt((p1x - p2x) * (p1x - p2x) + (p1y - p2y) * (p1y - p2y));
  const double len = hy / 2.0;
  // CTAs that take part: one per 1024 inliers (4 per thread), at most the grid.  A stroke on one wall collects a few thousand
  // points; with every CTA of the grid at the barrier and in the partial sums, each evaluation would pay for 64 arrivals.
  // The rest leave before touching the barrier; n is the same on every CTA, so all agree.
  const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t G = (uint32_t)min((unsigned long long)gridDim.x, max(1ull, (n + 1023ull) / 1024ull));
  if (blockIdx.x >= G) return;
  uint32_t n_eval = 0, barrier_goal = 0;

  // grid barrier (all CTAs are co-resident: the launch is cooperative and G <= #SMs): arrival counter, monotone goal
  auto grid_sync = [&]() {
    __syncthreads();
    if (threadIdx.x == 0) {
      barrier_goal += G;
      __threadfence();
      atomicAdd(barrier, 1u);
      while (*(volatile unsigned int*)barrier < barrier_goal) { }
      __threadfence();
    }
    __syncthreads();
  };
  auto evaluate = [&](double theta) -> FitSums {
    double ax = cos(theta), ay = sin(theta);
    const double inv = sqrt(ax * ax + ay * ay);             // alpha.normalize()
    ax /= inv; ay /= inv;
    double rr = 0.0, jr = 0.0, jj = 0.0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * kFitThreads + threadIdx.x; i < n; i += (unsigned long long)G * kFitThreads) {
      const float2 p = xy[i];
      double r, dr;
      seg_angle_residual((double)p.x, (double)p.y, cmx, cmy, len, ax, ay, &r, &dr);
      rr += r * r; jr += dr * r; jj += dr * dr;
    }
    for (int o = 16; o; o >>= 1) { rr += __shfl_xor_sync(0xffffffffu, rr, o); jr += __shfl_xor_sync(0xffffffffu, jr, o); jj += __shfl_xor_sync(0xffffffffu, jj, o); }
    if (lane == 0) { s_red[0][w] = rr; s_red[1][w] = jr; s_red[2][w] = jj; }
    __syncthreads();
    if (G == 1) {
      // one CTA holds every inlier (<= 1024 of them, the usual stroke): its own sums are the totals.  No trip through global memory
      // and no grid barrier — four dependent L2 round trips per evaluation otherwise; same additions in the same order.
      if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0, c = 0.0;
        for (int q = 0; q < kFitThreads / 32; ++q) { a += s_red[0][q]; b += s_red[1][q]; c += s_red[2][q]; }
        s_tot.rr = a; s_tot.jr = b; s_tot.jj = c;
      }
      __syncthreads();
      const FitSums t1 = s_tot;
      ++n_eval;
      __syncthreads();
      return t1;
    }
    double* mine = partial + ((size_t)(n_eval & 1u) * G + blockIdx.x) * 3;
    if (threadIdx.x == 0) {
      double a = 0.0, b = 0.0, c = 0.0;
      for (int q = 0; q < kFitThreads / 32; ++q) { a += s_red[0][q]; b += s_red[1][q]; c += s_red[2][q]; }
      mine[0] = a; mine[1] = b; mine[2] = c;
    }
    grid_sync();
    if (w == 0) {
      // the same fixed order on every CTA: lane q adds the partials q, q + 32, ..., then one butterfly
      const volatile double* all = partial + (size_t)(n_eval & 1u) * G * 3;
      double a = 0.0, b = 0.0, c = 0.0;
      for (uint32_t q = lane; q < G; q += 32) { a += all[3 * q]; b += all[3 * q + 1]; c += all[3 * q + 2]; }
      for (int o = 16; o; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); c += __shfl_xor_sync(0xffffffffu, c, o); }
      if (lane == 0) { s_tot.rr = a; s_tot.jr = b; s_tot.jj = c; }
    }
    __syncthreads();
    const FitSums t = s_tot;
    ++n_eval;
    __syncthreads();
    return t;
  };

  double x = acos(fabs(p1x - p2x) / hy);                    // in [0, pi/2]: the sign of the slope is dropped, as in the reference
  int iterations = 0, termination = 0;                      // 0 no convergence (iteration cap), 1 convergence
  double cost0 = 0.0, cost = 0.0;
  if (n > 0) {
    const double kFtol = 1e-6, kGtol = 1e-10, kPtol = 1e-8, kMinRel = 1e-3, kMinDiag = 1e-6, kMaxDiag = 1e32, kMinRadius = 1e-32, kMaxRadius = 1e16;
    FitSums S = evaluate(x);
    cost0 = cost = 0.5 * S.rr;
    const double scale = 1.0 / (1.0 + sqrt(S.jj));
    double g = S.jr * scale, H = S.jj * scale * scale;
    if (fabs(g / scale) <= kGtol) termination = 1;
    double radius = 1e4, decrease = 2.0;
    for (int iter = 1; iter <= max_iterations && termination == 0; ++iter) {
      iterations = iter;
      const double lm2 = fmin(fmax(H, kMinDiag), kMaxDiag) / radius;
      const double A = H + lm2;
      bool solved = A > 0.0;
      double step = 0.0, model_change = 0.0;
      if (solved) {
        const double d = sqrt(A);
        step = -((g / d) / d);
        model_change -= step * (g + 0.5 * (H * step));
      }
      if (!solved || !(model_change > 0.0)) {
        radius /= decrease; decrease *= 2.0;
        if (radius < kMinRadius) termination = 1;
        continue;
      }
      const double dx = step * scale, x_new = x + dx;
      if (sqrt(dx * dx) <= kPtol * (sqrt(x * x) + kPtol)) { termination = 1; break; }
      const FitSums Sn = evaluate(x_new);
      const double new_cost = 0.5 * Sn.rr, cost_change = cost - new_cost, rho = cost_change / model_change;
      if (rho > kMinRel) {
        x = x_new;
        const double old_cost = cost;
        cost = new_cost;
        g = Sn.jr * scale; H = Sn.jj * scale * scale;
        if (fabs(cost_change) <= kFtol * old_cost) { termination = 1; break; }
        if (fabs(g / scale) <= kGtol) { termination = 1; break; }
        const double t = 2.0 * rho - 1.0;
        radius = fmin(kMaxRadius, radius / fmax(1.0 / 3.0, 1.0 - t * t * t));
        decrease = 2.0;
      } else {
        radius /= decrease; decrease *= 2.0;
        if (radius < kMinRadius) termination = 1;
      }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double ax = cos(x), ay = sin(x);
    const double l2 = sqrt(ax * ax + ay * ay);
    ax /= l2; ay /= l2;
    FitResult R;
    R.theta = x; R.cost0 = cost0; R.cost = cost; R.n = n; R.iterations = iterations; R.evaluations = (int)n_eval; R.termination = termination;
    R.seg[0] = (float)(cmx + len * ax); R.seg[1] = (float)(cmy + len * ay); R.seg[2] = (float)(cmx - len * ax); R.seg[3] = (float)(cmy - len * ay);
    *out = R;
  }
}

// ---- K2b: observation sets of both strokes -----------------------------------------------------
// One warp per scan; indices written in order into a per-scan slot region (same offsets as the
// scan), counts per scan.  Poses with <= min_obs indices are dropped by the compaction below.
__global__ void em_assign_kernel(const float2* __restrict__ world, const uint32_t* __restrict__ off, uint32_t n_poses, Seg2 sa, Seg2 sb,
                                 double thr, uint32_t* __restrict__ obs_a, uint32_t* __restrict__ obs_b, uint32_t* __restrict__ cnt_a,
                                 uint32_t* __restrict__ cnt_b) {
  const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n_poses) return;
  const uint32_t o0 = off[w], o1 = off[w + 1];
  uint32_t na = 0, nb = 0;
  for (uint32_t k0 = o0; k0 < o1; k0 += 32) {
    const uint32_t k = k0 + lane;
    bool fa = false, fb = false;
    if (k < o1) {
      const float2 p = __ldg(world + k);
      fa = (double)dist_to_line_seg(sa, p.x, p.y) < thr;
      fb = (double)dist_to_line_seg(sb, p.x, p.y) < thr;
    }
    const uint32_t ma = __ballot_sync(0xffffffffu, fa), mb = __ballot_sync(0xffffffffu, fb);
    const uint32_t lt = (1u << lane) - 1u;
    if (fa) obs_a[o0 + na + __popc(ma & lt)] = k - o0;
    if (fb) obs_b[o0 + nb + __popc(mb & lt)] = k - o0;
    na += __popc(ma); nb += __popc(mb);
  }
  if (lane == 0) { cnt_a[w] = na; cnt_b[w] = nb; }
}

// Single-CTA ordered compaction of the kept poses: set_pose / set_off from per-scan counts.
__global__ void em_sets_scan_kernel(const uint32_t* __restrict__ cnt, uint32_t n_poses, uint32_t min_obs, uint32_t* __restrict__ set_pose,
                                    unsigned long long* __restrict__ set_off, uint32_t* __restrict__ slot_of_pose, unsigned long long* totals) {
  __shared__ unsigned long long sm_c[32]; __shared__ uint32_t sm_s[32];
  __shared__ unsigned long long carry_c; __shared__ uint32_t carry_s;
  if (threadIdx.x == 0) { carry_c = 0; carry_s = 0; }
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (uint32_t b = 0; b < n_poses; b += blockDim.x) {
    const uint32_t i = b + threadIdx.x;
    const uint32_t c = i < n_poses ? cnt[i] : 0;
    const bool keep = c > min_obs;
    unsigned long long xc = keep ? c : 0; uint32_t xs = keep ? 1u : 0u;
    const unsigned long long vc = xc; const uint32_t vs = xs;
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long yc = __shfl_up_sync(0xffffffffu, xc, o); const uint32_t ys = __shfl_up_sync(0xffffffffu, xs, o);
      if (lane >= (uint32_t)o) { xc += yc; xs += ys; }
    }
    if (lane == 31) { sm_c[w] = xc; sm_s[w] = xs; }
    __syncthreads();
    if (w == 0) {
      unsigned long long a = lane < nw ? sm_c[lane] : 0; uint32_t s = lane < nw ? sm_s[lane] : 0;
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long ya = __shfl_up_sync(0xffffffffu, a, o); const uint32_t ys = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= (uint32_t)o) { a += ya; s += ys; }
      }
      sm_c[lane] = a; sm_s[lane] = s;
    }
    __syncthreads();
    const unsigned long long ec = carry_c + (w ? sm_c[w - 1] : 0) + xc - vc;
    const uint32_t es = carry_s + (w ? sm_s[w - 1] : 0) + xs - vs;
    if (i < n_poses) {
      slot_of_pose[i] = keep ? es : 0xFFFFFFFFu;
      if (keep) { set_pose[es] = i; set_off[es] = ec; }
    }
    __syncthreads();
    if (threadIdx.x == 0) { carry_c += sm_c[nw - 1]; carry_s += sm_s[nw - 1]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { set_off[carry_s] = carry_c; totals[0] = carry_s; totals[1] = carry_c; }
}

// Gather the kept scans' index lists into the dense CSR payload. One warp per scan.
__global__ void em_sets_gather_kernel(const uint32_t* __restrict__ obs_slots, const uint32_t* __restrict__ off, const uint32_t* __restrict__ cnt,
                                      const uint32_t* __restrict__ slot_of_pose, const unsigned long long* __restrict__ set_off, uint32_t n_poses,
                                      uint32_t* __restrict__ obs_out) {
  const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n_poses) return;
  const uint32_t s = slot_of_pose[w];
  if (s == 0xFFFFFFFFu) return;
  const unsigned long long dst = set_off[s];
  const uint32_t c = cnt[w], src = off[w];
  for (uint32_t k = lane; k < c; k += 32) obs_out[dst + k] = obs_slots[src + k];
}

// HitLSLAM::verifyUserInput (HitLSLAM.cpp:218-243): which of the selected points have ANY world-frame point closer than the
// selection threshold ((w - s).norm() < thr, float).  The reference scans the clouds per selected point and stops at the first
// hit; existence does not depend on the order, so this is one HBM stream over the resident clouds (8 B per point, two points
// per 16 B load), flags OR-reduced per warp and merged with one atomicOr per warp that saw a hit.
struct VerifySel { float x[8], y[8]; uint32_t n; float thr; };
__global__ void __launch_bounds__(256) verify_input_kernel(const float2* __restrict__ world, uint64_t n_points, const VerifySel S, uint32_t* __restrict__ flags_out) {
  uint32_t flags = 0;
  const uint64_t n2 = n_points >> 1, stride = (uint64_t)gridDim.x * blockDim.x;
  const float4* w4 = reinterpret_cast<const float4*>(world);
  for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n2; q += stride) {
    const float4 v = __ldg(w4 + q);
#pragma unroll
    for (uint32_t i = 0; i < 8; ++i) {
      if (i < S.n) {
        const float ax = fsub(v.x, S.x[i]), ay = fsub(v.y, S.y[i]), bx = fsub(v.z, S.x[i]), by = fsub(v.w, S.y[i]);
        const float da = __fsqrt_rn(fadd(fmul(ax, ax), fmul(ay, ay))), db = __fsqrt_rn(fadd(fmul(bx, bx), fmul(by, by)));
        if (da < S.thr || db < S.thr) flags |= 1u << i;
      }
    }
  }
  if ((n_points & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    const float2 v = world[n_points - 1];
    for (uint32_t i = 0; i < S.n; ++i) {
      const float ax = fsub(v.x, S.x[i]), ay = fsub(v.y, S.y[i]);
      if (__fsqrt_rn(fadd(fmul(ax, ax), fmul(ay, ay))) < S.thr) flags |= 1u << i;
    }
  }
  flags = __reduce_or_sync(0xffffffffu, flags);
  if ((threadIdx.x & 31) == 0 && flags) atomicOr(flags_out, flags);
}

}  // namespace hitl

using namespace hitl;

extern "C" int hitl_verify_input(hitl_ctx* ctx, uint32_t n_selected, const float* sel_xy, float threshold, uint32_t* points_verified, uint32_t* seen_mask) {
  if (!ctx) return HITL_ERR_ARG;
  HITL_DEVICE(ctx);
  if (!points_verified || (n_selected && !sel_xy)) return fail(ctx, HITL_ERR_ARG, "hitl_verify_input: null argument");
  if (n_selected > 8) return fail(ctx, HITL_ERR_ARG, "hitl_verify_input: at most 8 selected points");
  if (!ctx->have_world) return fail(ctx, HITL_ERR_STATE, "hitl_verify_input: world clouds not set (hitl_world_transform / hitl_set_world_clouds)");
  VerifySel S; memset(&S, 0, sizeof(S));
  S.n = n_selected; S.thr = threshold;
  for (uint32_t i = 0; i < n_selected; ++i) { S.x[i] = sel_xy[2 * i]; S.y[i] = sel_xy[2 * i + 1]; }
  HITL_CUDA(ctx->d_ticket.ensure(4));
  HITL_CUDA(cudaMemsetAsync(ctx->d_ticket.p, 0, sizeof(uint32_t), ctx->stream));
  if (ctx->n_points && n_selected) {
    const uint64_t pairs = (ctx->n_points + 1) / 2;
    const uint32_t grid = (uint32_t)std::min<uint64_t>((pairs + 255) / 256, (uint64_t)ctx->sm_count * 8);
    verify_input_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->d_world.p, ctx->n_points, S, ctx->d_ticket.p);
    HITL_LAUNCH_CHECK("verify_input_kernel");
  }
  uint32_t* h = reinterpret_cast<uint32_t*>(ctx->h_pinned);
  HITL_CUDA(cudaMemcpyAsync(h, ctx->d_ticket.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  HITL_CUDA(cudaStreamSynchronize(ctx->stream));
  const uint32_t mask = h[0];
  uint32_t v = (uint32_t)__builtin_popcount(mask);
  // degenerate strokes void the whole input (HitLSLAM.cpp:238-241; the reference reads selected points 0..3 there)
  if (n_selected >= 4 && ((sel_xy[0] == sel_xy[2] && sel_xy[1] == sel_xy[3]) || (sel_xy[4] == sel_xy[6] && sel_xy[5] == sel_xy[7]))) v = 0;
  *points_verified = v;
  if (seen_mask) *seen_mask = mask;
  return HITL_OK;
}

extern "C" int hitl_world_transform(hitl_ctx* ctx, const float* poses_xyt, float* world_xy_out) {
  if (!ctx) return HITL_ERR_ARG;
  HITL_DEVICE(ctx);
  if (ctx->h_off.empty()) return fail(ctx, HITL_ERR_STATE, "hitl_world_transform: scans not set");
  if (!poses_xyt && ctx->n_poses) return fail(ctx, HITL_ERR_ARG, "hitl_world_transform: null poses");
  HITL_CUDA(ctx->d_world.ensure(ctx->n_points)); HITL_CUDA(ctx->d_poses_f.ensure(3 * (size_t)ctx->n_poses));
  ctx->em_boxes_valid = false;
  if (ctx->n_poses) {
    HITL_CUDA(cudaMemcpyAsync(ctx->d_poses_f.p, poses_xyt, 12 * (size_t)ctx->n_poses, cudaMemcpyHostToDevice, ctx->stream));
    const int threads = 256;
    HITL_KERNEL_BEGIN(HITL_K_WORLD_TRANSFORM);
    world_transform_kernel<<<((size_t)ctx->n_poses * 32 + threads - 1) / threads, threads, 0, ctx->stream>>>(ctx->d_pts.p, ctx->d_off.p, ctx->d_poses_f.p,
                                                                                                          ctx->n_poses, ctx->d_world.p);
    HITL_KERNEL_END(HITL_K_WORLD_TRANSFORM);
    HITL_LAUNCH_CHECK("world_transform_kernel");
  }
  if (world_xy_out && ctx->n_points) HITL_CUDA(cudaMemcpyAsync(world_xy_out, ctx->d_world.p, 8 * ctx->n_points, cudaMemcpyDeviceToHost, ctx->stream));
  HITL_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->have_world = true;
  return HITL_OK;
}

extern "C" int hitl_set_world_clouds(hitl_ctx* ctx, const float* world_xy) {
  if (!ctx) return HITL_ERR_ARG;
  HITL_DEVICE(ctx);
  if (ctx->h_off.empty()) return fail(ctx, HITL_ERR_STATE, "hitl_set_world_clouds: scans not set");
  if (!world_xy && ctx->n_points) return fail(ctx, HITL_ERR_ARG, "hitl_set_world_clouds: null clouds");
  HITL_CUDA(ctx->d_world.ensure(ctx->n_points));
  if (ctx->n_points) HITL_CUDA(cudaMemcpyAsync(ctx->d_world.p, world_xy, 8 * ctx->n_points, cudaMemcpyHostToDevice, ctx->stream));
  HITL_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->have_world = true; ctx->em_boxes_valid = false;
  return HITL_OK;
}

// host mirror of Vector2f::normalized() in float (same IEEE ops as the device would do)
static void make_seg(const float s[4], Seg* o) {
  o->p0x = s[0]; o->p0y = s[1]; o->p1x = s[2]; o->p1y = s[3];
  const float dx = s[2] - s[0], dy = s[3] - s[1];
  const float z = dx * dx + dy * dy;
  if (z > 0.0f) { const float n = sqrtf(z); o->dirx = dx / n; o->diry = dy / n; }
  else { o->dirx = dx; o->diry = dy; }
}
static void make_seg2(const float s[4], Seg2* o) {
  o->p1x = s[0]; o->p1y = s[1]; o->p2x = s[2]; o->p2y = s[3];
  o->dx = s[2] - s[0]; o->dy = s[3] - s[1];
  o->dd = o->dx * o->dx + o->dy * o->dy;
}

// One E-step launch on the context's stream.  `state`: n_chunks look-back words followed by the inlier total, `ticket`: one word;
// both zero when the kernel starts.  seg_dev != NULL: the stroke is read from the device (a chained round), `s` is ignored.
// The first E-step after the world clouds changed also records the chunks' bounding boxes; the later ones skip the chunks
// out of the stroke's reach.
static int launch_em_inliers(hitl_ctx* ctx, const Seg& s, const float* seg_dev, double thr, unsigned long long* state, uint32_t* ticket, uint64_t dcap,
                             uint32_t* out_pose, uint32_t* out_idx, float2* out_xy) {
  const uint32_t n_chunks = (uint32_t)((ctx->n_points + kEmChunk - 1) / kEmChunk);
  int mode = 0;
  if (ctx->em_cull) { HITL_CUDA(ctx->d_em_box.ensure(n_chunks)); mode = ctx->em_boxes_valid ? 2 : 1; }
  const uint32_t grid = std::min<uint32_t>(n_chunks, (uint32_t)ctx->sm_count * 8);
  auto kernel = mode == 2 ? em_inliers_kernel<2> : mode == 1 ? em_inliers_kernel<1> : em_inliers_kernel<0>;
  HITL_KERNEL_BEGIN(HITL_K_EM_INLIERS);
  kernel<<<grid, kEmThreads, 0, ctx->stream>>>(ctx->d_world.p, ctx->d_off.p, ctx->n_poses, ctx->n_points, s, seg_dev, thr, state, ticket, dcap, out_pose, out_idx,
                                               out_xy, state + n_chunks, ctx->d_em_box.p);
  HITL_KERNEL_END(HITL_K_EM_INLIERS);
  HITL_LAUNCH_CHECK("em_inliers_kernel");
  if (mode == 1) ctx->em_boxes_valid = true;
  return HITL_OK;
}

extern "C" int hitl_em_inliers(hitl_ctx* ctx, const float seg[4], double threshold, uint64_t cap, uint32_t* out_pose, uint32_t* out_idx,
                               float* out_xy, uint64_t* n_out) {
  if (!ctx) return HITL_ERR_ARG;
  HITL_DEVICE(ctx);
  if (!ctx->have_world) return fail(ctx, HITL_ERR_STATE, "hitl_em_inliers: world clouds not set");
  if (!seg || !n_out) return fail(ctx, HITL_ERR_ARG, "hitl_em_inliers: null argument");
  *n_out = 0;
  if (ctx->n_points == 0) return HITL_OK;
  const bool want = out_pose && out_idx && cap;
  const uint64_t dcap = want ? std::min<uint64_t>(cap, ctx->n_points) : 0;
  const uint32_t n_chunks = (uint32_t)((ctx->n_points + kEmChunk - 1) / kEmChunk);
  HITL_CUDA(ctx->d_scan_state.ensure(n_chunks + 1)); HITL_CUDA(ctx->d_ticket.ensure(4));
  if (want) { HITL_CUDA(ctx->d_em_pose.ensure(dcap)); HITL_CUDA(ctx->d_em_idx.ensure(dcap)); if (out_xy) HITL_CUDA(ctx->d_em_xy.ensure(dcap)); }
  HITL_CUDA(cudaMemsetAsync(ctx->d_scan_state.p, 0, 8 * (size_t)(n_chunks + 1), ctx->stream));
  HITL_CUDA(cudaMemsetAsync(ctx->d_ticket.p, 0, 4, ctx->stream));
  Seg s; make_seg(seg, &s);
  if (int rc = launch_em_inliers(ctx, s, nullptr, threshold, (unsigned long long*)ctx->d_scan_state.p, ctx->d_ticket.p, dcap, want ? ctx->d_em_pose.p : nullptr,
                                 want ? ctx->d_em_idx.p : nullptr, (want && out_xy) ? ctx->d_em_xy.p : nullptr))
    return rc;
  HITL_CUDA(cudaMemcpyAsync(ctx->h_pinned, ctx->d_scan_state.p + n_chunks, 8, cudaMemcpyDeviceToHost, ctx->stream));
  HITL_CUDA(cudaStreamSynchronize(ctx->stream));
  const uint64_t n = ctx->h_pinned[0];
  *n_out = n;
  if (want && n) {
    const uint64_t m = std::min<uint64_t>(n, dcap);
    HITL_CUDA(cudaMemcpyAsync(out_pose, ctx->d_em_pose.p, 4 * m, cudaMemcpyDeviceToHost, ctx->stream));
    HITL_CUDA(cudaMemcpyAsync(out_idx, ctx->d_em_idx.p, 4 * m, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_xy) HITL_CUDA(cudaMemcpyAsync(out_xy, ctx->d_em_xy.p, 8 * m, cudaMemcpyDeviceToHost, ctx->stream));
    HITL_CUDA(cudaStreamSynchronize(ctx->stream));
    if (n > cap) return fail(ctx, HITL_ERR_OVERFLOW, "hitl_em_inliers: more inliers than cap");
  }
  return HITL_OK;
}

// EM rounds chained on the device.  For every stroke, round r's E-step reads the stroke that round r-1's fit left in device memory
// (round 0 reads the caller's), and its M-step refits it; all n_strokes x rounds (E-step, M-step) pairs are enqueued back to back
// and the host waits ONCE, for all the results.  Each round is exactly the round hitl_em_refit would run on the previous round's
// output, so a caller that applies its convergence rule to the returned sequence and ignores the rounds past convergence gets
// the bits of the one-call-per-round loop (EMinput.cpp:107-147) without a host round trip per round.
constexpr uint32_t kChainMaxStrokes = 2, kChainMaxRounds = 4;
static_assert(sizeof(FitResult) == 64, "FitResult slots are copied back as 64-byte records");
extern "C" int hitl_em_refit_chain(hitl_ctx* ctx, uint32_t n_strokes, const float* segs_in, double inlier_threshold, int32_t max_iterations, uint32_t rounds,
                                   float* segs_out, hitl_em_fit_info* info) {
  if (!ctx) return HITL_ERR_ARG;
  HITL_DEVICE(ctx);
  if (!ctx->have_world) return fail(ctx, HITL_ERR_STATE, "hitl_em_refit: world clouds not set");
  if (!segs_in || !segs_out || max_iterations < 0 || n_strokes < 1 || n_strokes > kChainMaxStrokes || rounds < 1 || rounds > kChainMaxRounds)
    return fail(ctx, HITL_ERR_ARG, "hitl_em_refit: bad argument");
  const uint32_t L = n_strokes * rounds;                       // launches pairs; slot of (stroke s, round r) = r * n_strokes + s
  for (uint32_t r = 0; r < rounds; ++r)
    for (uint32_t q = 0; q < 4 * n_strokes; ++q) segs_out[4 * n_strokes * r + q] = segs_in[q];
  if (info) memset(info, 0, sizeof(hitl_em_fit_info) * L);
  if (ctx->n_points == 0) return HITL_OK;
  // the inliers' coordinates stay resident, in (pose, index) order (every point may be an inlier: the list cannot overflow)
  const uint64_t dcap = ctx->n_points;
  const uint32_t n_chunks = (uint32_t)((ctx->n_points + kEmChunk - 1) / kEmChunk);
  const size_t state_words = (size_t)n_chunks + 1;
  HITL_CUDA(ctx->d_scan_state.ensure(state_words * L)); HITL_CUDA(ctx->d_ticket.ensure(4 * (size_t)L));
  HITL_CUDA(ctx->d_em_pose.ensure(dcap)); HITL_CUDA(ctx->d_em_idx.ensure(dcap)); HITL_CUDA(ctx->d_em_xy.ensure(dcap));
  const int fit_blocks = std::min(kFitMaxBlocks, ctx->sm_count);
  HITL_CUDA(ctx->d_fit_partial.ensure(2 * (size_t)kFitMaxBlocks * 3)); HITL_CUDA(ctx->d_fit_out.ensure(sizeof(FitResult) * (size_t)L));
  HITL_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
  HITL_CUDA(cudaMemsetAsync(ctx->d_scan_state.p, 0, 8 * state_words * L, ctx->stream));
  HITL_CUDA(cudaMemsetAsync(ctx->d_ticket.p, 0, 16 * (size_t)L, ctx->stream));      // per launch pair: word 0 E-step ticket, words 2-3 the fit's grid barrier
  FitResult* d_res = reinterpret_cast<FitResult*>(ctx->d_fit_out.p);
  for (uint32_t sk = 0; sk < n_strokes; ++sk)
    for (uint32_t r = 0; r < rounds; ++r) {
      const uint32_t slot = r * n_strokes + sk;
      unsigned long long* state = (unsigned long long*)ctx->d_scan_state.p + state_words * slot;
      uint32_t* ticket = ctx->d_ticket.p + 4 * (size_t)slot;
      const float* seg_dev = r ? d_res[slot - n_strokes].seg : nullptr;
      Seg s; make_seg(segs_in + 4 * sk, &s);
      if (int rc = launch_em_inliers(ctx, s, seg_dev, inlier_threshold, state, ticket, dcap, ctx->d_em_pose.p, ctx->d_em_idx.p, ctx->d_em_xy.p)) return rc;
      // M-step: the whole LM loop in one cooperative launch (co-resident CTAs, software grid barrier)
      const float2* d_xy = ctx->d_em_xy.p;
      const unsigned long long* d_n = state + n_chunks;
      double p1x = segs_in[4 * sk], p1y = segs_in[4 * sk + 1], p2x = segs_in[4 * sk + 2], p2y = segs_in[4 * sk + 3];
      int iters = max_iterations;
      double* d_partial = ctx->d_fit_partial.p;
      unsigned int* d_barrier = ticket + 2;
      FitResult* d_out = d_res + slot;
      void* args[] = {(void*)&d_xy, (void*)&d_n, (void*)&p1x, (void*)&p1y, (void*)&p2x, (void*)&p2y, (void*)&seg_dev, (void*)&iters, (void*)&d_partial,
                      (void*)&d_barrier, (void*)&d_out};
      HITL_KERNEL_BEGIN(HITL_K_EM_FIT);
      HITL_CUDA(cudaLaunchCooperativeKernel((const void*)em_fit_kernel, dim3(fit_blocks), dim3(kFitThreads), args, 0, ctx->stream));
      HITL_KERNEL_END(HITL_K_EM_FIT);
      HITL_LAUNCH_CHECK("em_fit_kernel");
    }
  HITL_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
  static_assert(sizeof(FitResult) * kChainMaxStrokes * kChainMaxRounds <= 64 * sizeof(uint64_t), "h_pinned holds every FitResult of a chain");
  FitResult* h = reinterpret_cast<FitResult*>(ctx->h_pinned);
  HITL_CUDA(cudaMemcpyAsync(h, d_res, sizeof(FitResult) * (size_t)L, cudaMemcpyDeviceToHost, ctx->stream));
  HITL_CUDA(cudaStreamSynchronize(ctx->stream));
  float ms = 0.f;
  HITL_CUDA(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
  for (uint32_t slot = 0; slot < L; ++slot) {
    // no inliers: the fit kernel skips the solve and rebuilds the stroke from theta_0, as the reference does (EMinput.cpp:178-190)
    for (int q = 0; q < 4; ++q) segs_out[4 * slot + q] = h[slot].seg[q];
    if (info) {
      hitl_em_fit_info& inf = info[slot];
      inf.theta = h[slot].theta; inf.initial_cost = h[slot].cost0; inf.final_cost = h[slot].cost; inf.n_inliers = h[slot].n; inf.iterations = h[slot].iterations;
      inf.evaluations = h[slot].evaluations; inf.termination = h[slot].termination;
      inf.ms = ms;                                            // device time of the whole chain, the same in every record
    }
  }
  return HITL_OK;
}

extern "C" int hitl_em_refit(hitl_ctx* ctx, const float seg_in[4], double inlier_threshold, int32_t max_iterations, float seg_out[4], hitl_em_fit_info* info) {
  return hitl_em_refit_chain(ctx, 1, seg_in, inlier_threshold, max_iterations, 1, seg_out, info);
}

extern "C" int hitl_debug_set_em_cull(hitl_ctx* ctx, int on) {
  if (!ctx) return HITL_ERR_ARG;
  ctx->em_cull = on != 0;
  ctx->em_boxes_valid = false;
  return HITL_OK;
}

extern "C" int hitl_em_assign(hitl_ctx* ctx, const float segs[8], double threshold, uint32_t min_obs, uint32_t n_sets[2], uint32_t* set_pose0,
                              uint64_t* set_off0, uint32_t* obs0, uint32_t* set_pose1, uint64_t* set_off1, uint32_t* obs1) {
  if (!ctx) return HITL_ERR_ARG;
  HITL_DEVICE(ctx);
  if (!ctx->have_world) return fail(ctx, HITL_ERR_STATE, "hitl_em_assign: world clouds not set");
  if (!segs || !n_sets) return fail(ctx, HITL_ERR_ARG, "hitl_em_assign: null argument");
  n_sets[0] = n_sets[1] = 0;
  if (set_off0) set_off0[0] = 0;
  if (set_off1) set_off1[0] = 0;
  const uint32_t n = ctx->n_poses;
  if (n == 0) return HITL_OK;
  DevBuf<uint32_t>* slots = ctx->d_em_slots; DevBuf<uint32_t>& slot_of_pose = ctx->d_em_slotof;
  for (int f = 0; f < 2; ++f) {
    HITL_CUDA(slots[f].ensure(ctx->n_points)); HITL_CUDA(ctx->d_em_obs[f].ensure(ctx->n_points)); HITL_CUDA(ctx->d_em_cnt[f].ensure(n));
    HITL_CUDA(ctx->d_em_setpose[f].ensure(n)); HITL_CUDA(ctx->d_em_setoff[f].ensure(n + 1));
  }
  HITL_CUDA(slot_of_pose.ensure(n)); HITL_CUDA(ctx->d_counters.ensure(8));
  Seg2 sa, sb; make_seg2(segs, &sa); make_seg2(segs + 4, &sb);
  const int threads = 256;
  const uint32_t grid = (uint32_t)(((size_t)n * 32 + threads - 1) / threads);
  HITL_KERNEL_BEGIN(HITL_K_EM_ASSIGN);
  em_assign_kernel<<<grid, threads, 0, ctx->stream>>>(ctx->d_world.p, ctx->d_off.p, n, sa, sb, threshold, slots[0].p, slots[1].p, ctx->d_em_cnt[0].p,
                                                      ctx->d_em_cnt[1].p);
  HITL_KERNEL_END(HITL_K_EM_ASSIGN);
  HITL_LAUNCH_CHECK("em_assign_kernel");
  uint32_t* set_pose[2] = {set_pose0, set_pose1}; uint64_t* set_off[2] = {set_off0, set_off1}; uint32_t* obs[2] = {obs0, obs1};
  for (int f = 0; f < 2; ++f) {
    em_sets_scan_kernel<<<1, 1024, 0, ctx->stream>>>(ctx->d_em_cnt[f].p, n, min_obs, ctx->d_em_setpose[f].p, (unsigned long long*)ctx->d_em_setoff[f].p,
                                                     slot_of_pose.p, (unsigned long long*)ctx->d_counters.p + 2 * f);
    HITL_LAUNCH_CHECK("em_sets_scan_kernel");
    if (!obs[f]) continue;                 // the caller reads the observing poses only: the index lists are not assembled
    em_sets_gather_kernel<<<grid, threads, 0, ctx->stream>>>(slots[f].p, ctx->d_off.p, ctx->d_em_cnt[f].p, slot_of_pose.p,
                                                            (unsigned long long*)ctx->d_em_setoff[f].p, n, ctx->d_em_obs[f].p);
    HITL_LAUNCH_CHECK("em_sets_gather_kernel");
  }
  HITL_CUDA(cudaMemcpyAsync(ctx->h_pinned, ctx->d_counters.p, 4 * 8, cudaMemcpyDeviceToHost, ctx->stream));
  HITL_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int f = 0; f < 2; ++f) {
    const uint64_t ns = ctx->h_pinned[2 * f], no = ctx->h_pinned[2 * f + 1];
    n_sets[f] = (uint32_t)ns;
    if (set_pose[f] && ns) HITL_CUDA(cudaMemcpyAsync(set_pose[f], ctx->d_em_setpose[f].p, 4 * ns, cudaMemcpyDeviceToHost, ctx->stream));
    if (set_off[f]) HITL_CUDA(cudaMemcpyAsync(set_off[f], ctx->d_em_setoff[f].p, 8 * (ns + 1), cudaMemcpyDeviceToHost, ctx->stream));
    if (obs[f] && no) HITL_CUDA(cudaMemcpyAsync(obs[f], ctx->d_em_obs[f].p, 4 * no, cudaMemcpyDeviceToHost, ctx->stream));
  }
  HITL_CUDA(cudaStreamSynchronize(ctx->stream));
  return HITL_OK;
}
