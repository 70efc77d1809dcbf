// eval.cu — batched residual / Jacobian evaluation and normal-equation assembly on sm_100a.
//
// Replaces the per-block AutoDiffCostFunction<Functor,...>::Evaluate calls Ceres makes for the
// blocks the reference adds (paths relative to HitL-SLAM/src/human_in_the_loop_slam/):
//   PoseConstraint                      residual_functors.h:1054-1133 (AddOdometryConstraints, JointOptimization.cpp:736-825)
//   Colocation/Colinear/Perpendicular/ParallelHumanImposedConstraint
//                                       residual_functors.h:1299-1415 (AddHumanConstraints, JointOptimization.cpp:969-1054)
//   PointToPointGlobConstraint          residual_functors.h:768-848   (AddSTFConstraints, JointOptimization.cpp:539-559)
//   PointToLineGlobConstraint / PointToLineConstraint   residual_functors.h:314-385, :557-622 (no call site in the reference)
//
// Derivatives are analytic (hand-derived from the functor text) instead of Jets: same values
// to rounding (<= 1e-9 relative in FP64 is the contract; FP32 mode <= 1e-5), ~85 flop per
// correspondence instead of ~7x that, so the STF kernel stays bound by its 32 B/correspondence
// gather.  Jacobian layout = Ceres': row-major [residual][param] per parameter block.
#include <cub/device/device_radix_sort.cuh>
#include "hitl_internal.h"

namespace hitl {

struct NeqOut {            // optional normal-equation accumulation (all device pointers, may be null)
  double* H_diag;          // n_poses x 9
  double* g;               // n_poses x 3
  double* H_off;           // n_binary_blocks x 9
  double* cost;            // 1
};

template <int R>
__device__ __forceinline__ void accumulate_unary(const NeqOut& q, uint32_t pose, const double* r, const double* J /* R x 3 */) {
  double* H = q.H_diag + 9 * (size_t)pose; double* g = q.g + 3 * (size_t)pose;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    double ga = 0;
#pragma unroll
    for (int i = 0; i < R; ++i) ga += J[3 * i + a] * r[i];
    atomicAdd(g + a, ga);
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      double h = 0;
#pragma unroll
      for (int i = 0; i < R; ++i) h += J[3 * i + a] * J[3 * i + b];
      atomicAdd(H + 3 * a + b, h);
    }
  }
}
template <int R>
__device__ __forceinline__ void accumulate_binary(const NeqOut& q, size_t off_slot, uint32_t pa, uint32_t pb, const double* r, const double* Ja,
                                                  const double* Jb) {
  accumulate_unary<R>(q, pa, r, Ja);
  accumulate_unary<R>(q, pb, r, Jb);
  if (q.H_off) {
    double* H = q.H_off + 9 * off_slot;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        double h = 0;
#pragma unroll
        for (int i = 0; i < R; ++i) h += Ja[3 * i + a] * Jb[3 * i + b];
        H[3 * a + b] = h;
      }
  }
}
__device__ __forceinline__ void accumulate_cost(const NeqOut& q, double c) {
  for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c != 0.0) atomicAdd(q.cost, c);
}

// ---- odometry blocks: one thread per block ---------------------------------------------------
template <typename T>
__global__ void eval_odometry_kernel(const float* __restrict__ consts, const double* __restrict__ pose, uint32_t n_blocks, double* __restrict__ r_out,
                                     double* __restrict__ J_out, NeqOut neq, int want_neq) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  double cost = 0;
  if (b < n_blocks) {
    const float* c = consts + 9 * (size_t)b;
    const T a00 = c[0], a01 = c[1], a10 = c[2], a11 = c[3];
    const T isr = T(1) / T(c[4]), ist = T(1) / T(c[5]), isa = T(1) / T(c[6]);
    const T x1 = (T)pose[3 * b], y1 = (T)pose[3 * b + 1], t1 = (T)pose[3 * b + 2];
    const T x2 = (T)pose[3 * b + 3], y2 = (T)pose[3 * b + 4], t2 = (T)pose[3 * b + 5];
    const T tx = x2 - x1, ty = y2 - y1;
    const T cs = cos(-t1), sn = sin(-t1);          // Rotation2D(-theta1)
    const T rx = cs * tx - sn * ty, ry = sn * tx + cs * ty;
    const T ax = a00 * rx + a01 * ry, ay = a10 * rx + a11 * ry;
    const T d = t2 - t1 - T(c[8]);
    double r[3], J1[9], J2[9];
    r[0] = (double)((ax - T(c[7])) * isr);
    r[1] = (double)(ay * ist);
    r[2] = (double)(atan2(sin(d), cos(d)) * isa);
    // d(rx,ry)/d(x2,y2) = Rot(-t1); d/d(x1,y1) = -Rot(-t1); d(rx,ry)/dt1 = (ry, -rx)
    const T m00 = a00 * cs + a01 * sn, m01 = a00 * (-sn) + a01 * cs;   // A * Rot(-t1)
    const T m10 = a10 * cs + a11 * sn, m11 = a10 * (-sn) + a11 * cs;
    const T dax = a00 * ry - a01 * rx, day = a10 * ry - a11 * rx;
    J2[0] = (double)(m00 * isr); J2[1] = (double)(m01 * isr); J2[2] = 0;
    J2[3] = (double)(m10 * ist); J2[4] = (double)(m11 * ist); J2[5] = 0;
    J2[6] = 0; J2[7] = 0; J2[8] = (double)isa;
    J1[0] = -J2[0]; J1[1] = -J2[1]; J1[2] = (double)(dax * isr);
    J1[3] = -J2[3]; J1[4] = -J2[4]; J1[5] = (double)(day * ist);
    J1[6] = 0; J1[7] = 0; J1[8] = -(double)isa;
    if (r_out) { r_out[3 * (size_t)b] = r[0]; r_out[3 * (size_t)b + 1] = r[1]; r_out[3 * (size_t)b + 2] = r[2]; }
    if (J_out) {
      double* J = J_out + 18 * (size_t)b;
#pragma unroll
      for (int i = 0; i < 9; ++i) { J[i] = J1[i]; J[9 + i] = J2[i]; }
    }
    if (want_neq) {
      accumulate_binary<3>(neq, b, b, b + 1, r, J1, J2);
      cost = 0.5 * (r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    }
  }
  if (want_neq) accumulate_cost(neq, cost);
}

// ---- human blocks: one thread per block ------------------------------------------------------
template <typename T>
__global__ void eval_human_kernel(const int32_t* __restrict__ type_pose, const double* __restrict__ tg, const double* __restrict__ pose,
                                  uint32_t n_blocks, double* __restrict__ r_out, double* __restrict__ J_out, NeqOut neq, int want_neq) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  double cost = 0;
  if (b < n_blocks) {
    const int type = type_pose[2 * b]; const uint32_t p = (uint32_t)type_pose[2 * b + 1];
    const T x = (T)pose[3 * p], y = (T)pose[3 * p + 1], th = (T)pose[3 * p + 2];
    const T xt = (T)tg[4 * b], yt = (T)tg[4 * b + 1], tt = (T)tg[4 * b + 2];
    double r[3] = {0, 0, 0}, J[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (type == 2) {                 // colocation (kLineSegmentCorrection): 3 residuals
      r[0] = (double)(xt - x); r[1] = (double)(yt - y); r[2] = (double)(tt - th);
      J[0] = -1; J[4] = -1; J[8] = -1;
    } else if (type == 4) {          // colinear: 2 residuals
      const double pd = tg[4 * b + 3];
      const T cx = (T)cos(pd), cy = (T)sin(pd);
      r[0] = (double)(cx * (xt - x) + cy * (yt - y)); r[1] = (double)(tt - th);
      J[0] = -(double)cx; J[1] = -(double)cy; J[5] = -1;
    } else {                         // perpendicular / parallel: 1 residual
      r[0] = (double)(tt - th);
      J[2] = -1;
    }
    if (r_out) { r_out[3 * (size_t)b] = r[0]; r_out[3 * (size_t)b + 1] = r[1]; r_out[3 * (size_t)b + 2] = r[2]; }
    if (J_out) { double* o = J_out + 9 * (size_t)b;
#pragma unroll
      for (int i = 0; i < 9; ++i) o[i] = J[i]; }
    if (want_neq) { accumulate_unary<3>(neq, p, r, J); cost = 0.5 * (r[0] * r[0] + r[1] * r[1] + r[2] * r[2]); }
  }
  if (want_neq) accumulate_cost(neq, cost);
}

// ---- per-pose trigonometry, once per evaluation point ---------------------------------------------
// (c, s, x, y) per pose: every STF block needs both of its poses' sin/cos; computing them per block
// costs more FP64 instructions than the block's own matches (~90 on average).
__global__ void pose_trig_kernel(const double* __restrict__ pose, uint32_t n_poses, double2* __restrict__ trig) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_poses) return;
  double s, c;
  sincos(pose[3 * i + 2], &s, &c);
  trig[2 * i] = make_double2(c, s);
  trig[2 * i + 1] = make_double2(pose[3 * i], pose[3 * i + 1]);
}

// ---- STF blocks: one warp per block ----------------------------------------------------------
// r0 = sqrt(S0 / M), S0 = sum a_m^2, a_m = (n0g . (p1g - p0g)) * cf / sd      (likewise r1 with n1g)
// dr0/dx = (sum a_m da_m/dx) / (M r0); r = 0 with zero Jacobian when the sum is exactly 0.
template <typename T>
struct StfAccum {
  T S0, S1, G0[6], G1[6];
};

template <typename T>
__device__ __forceinline__ void stf_accumulate(StfAccum<T>& A, const float2 P0, const float2 N0, const float2 P1, const float2 N1, const double c0d, const double s0d,
                                               const double c1d, const double s1d, const double tdx, const double tdy, const T w) {
  // Pose trigonometry and the world-frame difference p1g - p0g are always formed in FP64: the
  // difference cancels ~10 m coordinates down to centimetres, which FP32 cannot carry to 1e-5.
  // Everything downstream (normals, projections, derivative sums) runs in T.
  const T c0 = (T)c0d, s0 = (T)s0d, c1 = (T)c1d, s1 = (T)s1d;
  const double r0xd = c0d * P0.x - s0d * P0.y, r0yd = s0d * P0.x + c0d * P0.y;   // rotated (not translated) points
  const double r1xd = c1d * P1.x - s1d * P1.y, r1yd = s1d * P1.x + c1d * P1.y;
  const T r0x = (T)r0xd, r0y = (T)r0yd, r1x = (T)r1xd, r1y = (T)r1yd;
  const T n0x = c0 * T(N0.x) - s0 * T(N0.y), n0y = s0 * T(N0.x) + c0 * T(N0.y);
  const T n1x = c1 * T(N1.x) - s1 * T(N1.y), n1y = s1 * T(N1.x) + c1 * T(N1.y);
  const T dx = (T)((r1xd - r0xd) + tdx), dy = (T)((r1yd - r0yd) + tdy);
  const T u = n0x * dx + n0y * dy, v = n1x * dx + n1y * dy;
  const T a = u * w, bb = v * w;
  A.S0 += a * a; A.S1 += bb * bb;
  // perp(w) = (-w.y, w.x) = d/dtheta of a rotated vector
  // du/dth0 = perp(n0g).d - n0g.perp(r0) ; du/dth1 = n0g.perp(r1)
  const T du_t0 = (-n0y * dx + n0x * dy) - (n0x * (-r0y) + n0y * r0x);
  const T du_t1 = n0x * (-r1y) + n0y * r1x;
  // dv/dth0 = -n1g.perp(r0) ; dv/dth1 = perp(n1g).d + n1g.perp(r1)
  const T dv_t0 = -(n1x * (-r0y) + n1y * r0x);
  const T dv_t1 = (-n1y * dx + n1x * dy) + (n1x * (-r1y) + n1y * r1x);
  const T aw = a * w, bw = bb * w;
  A.G0[0] += aw * (-n0x); A.G0[1] += aw * (-n0y); A.G0[2] += aw * du_t0; A.G0[3] += aw * n0x; A.G0[4] += aw * n0y; A.G0[5] += aw * du_t1;
  A.G1[0] += bw * (-n1x); A.G1[1] += bw * (-n1y); A.G1[2] += bw * dv_t0; A.G1[3] += bw * n1x; A.G1[4] += bw * n1y; A.G1[5] += bw * dv_t1;
}

constexpr int kStfThreads = 256;
#ifndef HITL_STF_MINBLOCKS
#define HITL_STF_MINBLOCKS 4   // 64 registers, 32 warps / SM: measured 0.53 / 0.42 / 0.41 ms at 2 / 3 / 4 on B200 (profiles/eval_variants.py)
#endif

template <typename T>
__global__ void __launch_bounds__(kStfThreads, HITL_STF_MINBLOCKS) eval_stf_kernel(const float2* __restrict__ pts, const float2* __restrict__ nrm, const uint32_t* __restrict__ off,
                                                               const uint32_t* __restrict__ pair_i, const uint32_t* __restrict__ pair_j,
                                                               const unsigned long long* __restrict__ pair_off, const uint32_t* __restrict__ kk,
                                                               const uint32_t* __restrict__ idx, const double2* __restrict__ trig, uint64_t n_blocks,
                                                               float std_dev, float corr, double* __restrict__ r_out, double* __restrict__ J_out, NeqOut neq,
                                                               int want_neq, size_t off_slot0) {
  const uint64_t b = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  if (b >= n_blocks) return;
  const uint32_t i = pair_i[b], j = pair_j[b];
  const unsigned long long m0 = pair_off[b], m1 = pair_off[b + 1];
  const double2 ti = __ldg(trig + 2 * i), xi = __ldg(trig + 2 * i + 1), tj = __ldg(trig + 2 * j), xj = __ldg(trig + 2 * j + 1);   // (cos, sin), (x, y)
  const double tdx = xj.x - xi.x, tdy = xj.y - xi.y;
  const T w = T(corr) / T(std_dev);
  const float2* __restrict__ pi = pts + off[i]; const float2* __restrict__ ni = nrm + off[i];
  const float2* __restrict__ pj = pts + off[j]; const float2* __restrict__ nj = nrm + off[j];
  StfAccum<T> A;
  A.S0 = 0; A.S1 = 0;
#pragma unroll
  for (int q = 0; q < 6; ++q) { A.G0[q] = 0; A.G1[q] = 0; }
  unsigned long long m = m0 + lane;
  // two matches per lane and iteration: all eight gathers are issued before the first use
  for (; m + 32 < m1; m += 64) {
    const uint32_t ka = __ldg(kk + m), qa = __ldg(idx + m), kb = __ldg(kk + m + 32), qb = __ldg(idx + m + 32);
    const float2 P0a = __ldg(pi + ka), N0a = __ldg(ni + ka), P1a = __ldg(pj + qa), N1a = __ldg(nj + qa);
    const float2 P0b = __ldg(pi + kb), N0b = __ldg(ni + kb), P1b = __ldg(pj + qb), N1b = __ldg(nj + qb);
    stf_accumulate<T>(A, P0a, N0a, P1a, N1a, ti.x, ti.y, tj.x, tj.y, tdx, tdy, w);
    stf_accumulate<T>(A, P0b, N0b, P1b, N1b, ti.x, ti.y, tj.x, tj.y, tdx, tdy, w);
  }
  if (m < m1) {
    const uint32_t ka = __ldg(kk + m), qa = __ldg(idx + m);
    stf_accumulate<T>(A, __ldg(pi + ka), __ldg(ni + ka), __ldg(pj + qa), __ldg(nj + qa), ti.x, ti.y, tj.x, tj.y, tdx, tdy, w);
  }
  // fixed-order butterfly: every lane ends with the same sums
  for (int o = 16; o; o >>= 1) {
    A.S0 += __shfl_xor_sync(0xffffffffu, A.S0, o); A.S1 += __shfl_xor_sync(0xffffffffu, A.S1, o);
#pragma unroll
    for (int q = 0; q < 6; ++q) { A.G0[q] += __shfl_xor_sync(0xffffffffu, A.G0[q], o); A.G1[q] += __shfl_xor_sync(0xffffffffu, A.G1[q], o); }
  }
  // every lane holds the block's sums: lanes share the epilogue (lane q < 12 owns Jacobian entry q)
  const T M = (T)(double)(m1 - m0);
  T r0 = 0, r1 = 0, k0 = 0, k1 = 0;
  if (A.S0 != T(0)) { r0 = sqrt(A.S0 / M); k0 = T(1) / (M * r0); }
  if (A.S1 != T(0)) { r1 = sqrt(A.S1 / M); k1 = T(1) / (M * r1); }
  double r[2], Ji[6], Jj[6];
  r[0] = (double)r0; r[1] = (double)r1;
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    Ji[q] = (double)(A.G0[q] * k0); Ji[3 + q] = (double)(A.G1[q] * k1);
    Jj[q] = (double)(A.G0[3 + q] * k0); Jj[3 + q] = (double)(A.G1[3 + q] * k1);
  }
  if (lane == 0 && r_out) { r_out[2 * b] = r[0]; r_out[2 * b + 1] = r[1]; }
  if (J_out && lane < 12) {
    double v = 0;
#pragma unroll
    for (int q = 0; q < 6; ++q) { if (lane == (uint32_t)q) v = Ji[q]; if (lane == (uint32_t)(6 + q)) v = Jj[q]; }
    J_out[12 * b + lane] = v;                                       // one coalesced 96 B store per block
  }
  if (want_neq) {
    // lanes 0..8: H_ii entry, 9..17: H_jj entry, 18..20: g_i, 21..23: g_j, 24: cost; H_off by lanes 0..8 as plain stores
    const uint32_t e = lane < 9 ? lane : lane < 18 ? lane - 9 : 0, a3 = e / 3, b3 = e % 3;
    if (lane < 18) {
      const double* Jx = lane < 9 ? Ji : Jj;
      double h = 0;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        double ja = 0, jb = 0;
#pragma unroll
        for (int c = 0; c < 3; ++c) { if (a3 == (uint32_t)c) ja = Jx[3 * q + c]; if (b3 == (uint32_t)c) jb = Jx[3 * q + c]; }
        h += ja * jb;
      }
      atomicAdd(neq.H_diag + 9 * (size_t)(lane < 9 ? i : j) + e, h);
      if (lane < 9 && neq.H_off) {
        double ho = 0;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          double ja = 0, jb = 0;
#pragma unroll
          for (int c = 0; c < 3; ++c) { if (a3 == (uint32_t)c) ja = Ji[3 * q + c]; if (b3 == (uint32_t)c) jb = Jj[3 * q + c]; }
          ho += ja * jb;
        }
        neq.H_off[9 * (off_slot0 + b) + e] = ho;
      }
    } else if (lane < 24) {
      const uint32_t c3 = (lane - 18) % 3;
      const double* Jx = lane < 21 ? Ji : Jj;
      double gq = 0;
#pragma unroll
      for (int c = 0; c < 3; ++c) if (c3 == (uint32_t)c) gq = Jx[c] * r[0] + Jx[3 + c] * r[1];
      atomicAdd(neq.g + 3 * (size_t)(lane < 21 ? i : j) + c3, gq);
    } else if (lane == 24) {
      const double c = 0.5 * (r[0] * r[0] + r[1] * r[1]);
      if (c != 0.0) atomicAdd(neq.cost, c);
    }
  }
}

// ---- point-to-line-glob blocks: one warp per block; residual = sum (d cf/sd)^2 (no sqrt) ------
template <typename T>
__global__ void eval_p2l_glob_kernel(const uint32_t* __restrict__ blk_pose, const unsigned long long* __restrict__ blk_off, const float2* __restrict__ pts,
                                     const float2* __restrict__ ln, const float* __restrict__ lo, const uint8_t* __restrict__ valid,
                                     const double* __restrict__ pose, uint32_t n_blocks, float std_dev, float corr, double* __restrict__ r_out,
                                     double* __restrict__ J_out, NeqOut neq, int want_neq) {
  const uint32_t b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (b >= n_blocks) return;
  const uint32_t p = blk_pose[b];
  const T tx = (T)pose[3 * p], ty = (T)pose[3 * p + 1], th = (T)pose[3 * p + 2];
  const T c = cos(th), s = sin(th), w = T(corr) / T(std_dev);
  T S = 0, G[3] = {0, 0, 0};
  for (unsigned long long m = blk_off[b] + lane; m < blk_off[b + 1]; m += 32) {
    if (!valid[m]) continue;
    const float2 P = pts[m], N = ln[m];
    const T rx = c * T(P.x) - s * T(P.y), ry = s * T(P.x) + c * T(P.y);
    const T e = ((rx + tx) * T(N.x) + (ry + ty) * T(N.y)) + T(lo[m]);
    const T a = e * T(corr) / T(std_dev);
    S += a * a;
    const T de_t = (-ry) * T(N.x) + rx * T(N.y);
    G[0] += T(2) * a * w * T(N.x); G[1] += T(2) * a * w * T(N.y); G[2] += T(2) * a * w * de_t;
  }
  for (int o = 16; o; o >>= 1) {
    S += __shfl_xor_sync(0xffffffffu, S, o);
#pragma unroll
    for (int q = 0; q < 3; ++q) G[q] += __shfl_xor_sync(0xffffffffu, G[q], o);
  }
  if (lane == 0) {
    double r[1] = {(double)S}, J[3] = {(double)G[0], (double)G[1], (double)G[2]};
    if (r_out) r_out[b] = r[0];
    if (J_out) { J_out[3 * (size_t)b] = J[0]; J_out[3 * (size_t)b + 1] = J[1]; J_out[3 * (size_t)b + 2] = J[2]; }
    if (want_neq) { accumulate_unary<1>(neq, p, r, J); atomicAdd(neq.cost, 0.5 * r[0] * r[0]); }
  }
}

// ---- single point-to-line blocks: one thread per block ----------------------------------------
template <typename T>
__global__ void eval_p2l_kernel(const uint32_t* __restrict__ pose_idx, const float2* __restrict__ pts, const float2* __restrict__ ln,
                                const float* __restrict__ lo, const uint8_t* __restrict__ valid, const double* __restrict__ pose, uint64_t n_blocks,
                                float std_dev, float corr, double* __restrict__ r_out, double* __restrict__ J_out, NeqOut neq, int want_neq) {
  const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double cost = 0;
  if (b < n_blocks) {
    double r[1] = {0}, J[3] = {0, 0, 0};
    const uint32_t p = pose_idx[b];
    if (valid[b]) {
      const T tx = (T)pose[3 * p], ty = (T)pose[3 * p + 1], th = (T)pose[3 * p + 2];
      const T c = cos(th), s = sin(th);
      const float2 P = pts[b], N = ln[b];
      const T rx = c * T(P.x) - s * T(P.y), ry = s * T(P.x) + c * T(P.y);
      const T e = ((rx + tx) * T(N.x) + (ry + ty) * T(N.y)) + T(lo[b]);
      const T w = T(corr) / T(std_dev);
      r[0] = (double)(e * T(corr) / T(std_dev));
      J[0] = (double)(w * T(N.x)); J[1] = (double)(w * T(N.y)); J[2] = (double)(w * ((-ry) * T(N.x) + rx * T(N.y)));
    }
    if (r_out) r_out[b] = r[0];
    if (J_out) { J_out[3 * b] = J[0]; J_out[3 * b + 1] = J[1]; J_out[3 * b + 2] = J[2]; }
    if (want_neq) { accumulate_unary<1>(neq, p, r, J); cost = 0.5 * r[0] * r[0]; }
  }
  if (want_neq) accumulate_cost(neq, cost);
}

}  // namespace hitl

using namespace hitl;

// ---- block registration ------------------------------------------------------------------------
extern "C" int hitl_set_stf_blocks_from_search(hitl_ctx* ctx, float std_dev, float corr) {
  if (!ctx) return HITL_ERR_ARG;
  HITL_DEVICE(ctx);
  if (!ctx->have_stf) return fail(ctx, HITL_ERR_STATE, "hitl_set_stf_blocks_from_search: no search result");
  ctx->eval_valid = ctx->neq_valid = false; ctx->inc_valid = false;
  ctx->stf_from_search = true; ctx->nb_stf = ctx->n_pairs; ctx->stf_std = std_dev; ctx->stf_corr = corr;
  return HITL_OK;
}
extern "C" int hitl_set_stf_blocks(hitl_ctx* ctx, uint64_t n_pairs, const uint32_t* pair_i, const uint32_t* pair_j, const uint64_t* pair_off,
                                   const uint32_t* k, const uint32_t* idx, float std_dev, float corr) {
  if (!ctx) return HITL_ERR_ARG;
  HITL_DEVICE(ctx);
  if (ctx->h_off.empty()) return fail(ctx, HITL_ERR_STATE, "hitl_set_stf_blocks: scans not set");
  if (n_pairs && (!pair_i || !pair_j || !pair_off || !k || !idx)) return fail(ctx, HITL_ERR_ARG, "hitl_set_stf_blocks: null argument");
  const uint64_t nm = n_pairs ? pair_off[n_pairs] : 0;
  for (uint64_t b = 0; b < n_pairs; ++b) {
    if (pair_i[b] >= ctx->n_poses || pair_j[b] >= ctx->n_poses || pair_off[b + 1] <= pair_off[b])
      return fail(ctx, HITL_ERR_ARG, "hitl_set_stf_blocks: bad pair (pose out of range or empty block)");
    const uint32_t ni = ctx->h_off[pair_i[b] + 1] - ctx->h_off[pair_i[b]], nj = ctx->h_off[pair_j[b] + 1] - ctx->h_off[pair_j[b]];
    for (uint64_t m = pair_off[b]; m < pair_off[b + 1]; ++m)
      if (k[m] >= ni || idx[m] >= nj) return fail(ctx, HITL_ERR_ARG, "hitl_set_stf_blocks: point index out of range");
  }
  HITL_CUDA(ctx->d_blk_i.ensure(n_pairs)); HITL_CUDA(ctx->d_blk_j.ensure(n_pairs)); HITL_CUDA(ctx->d_blk_off.ensure(n_pairs + 1));
  HITL_CUDA(ctx->d_blk_k.ensure(nm)); HITL_CUDA(ctx->d_blk_idx.ensure(nm));
  if (n_pairs) {
    HITL_CUDA(cudaMemcpyAsync(ctx->d_blk_i.p, pair_i, 4 * n_pairs, cudaMemcpyHostToDevice, ctx->stream));
    HITL_CUDA(cudaMemcpyAsync(ctx->d_blk_j.p, pair_j, 4 * n_pairs, cudaMemcpyHostToDevice, ctx->stream));
    HITL_CUDA(cudaMemcpyAsync(ctx->d_blk_off.p, pair_off, 8 * (n_pairs + 1), cudaMemcpyHostToDevice, ctx->stream));
    HITL_CUDA(cudaMemcpyAsync(ctx->d_blk_k.p, k, 4 * nm, cudaMemcpyHostToDevice, ctx->stream));
    HITL_CUDA(cudaMemcpyAsync(ctx->d_blk_idx.p, idx, 4 * nm, cudaMemcpyHostToDevice, ctx->stream));
    HITL_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  ctx->eval_valid = ctx->neq_valid = false; ctx->inc_valid = false;
  ctx->stf_from_search = false; ctx->nb_stf = n_pairs; ctx->stf_std = std_dev; ctx->stf_corr = corr;
  return HITL_OK;
}
extern "C" int hitl_set_odometry_blocks(hitl_ctx* ctx, uint32_t n_blocks, const float* consts9) {
  if (!ctx) return HITL_ERR_ARG;
  HITL_DEVICE(ctx);
  if (n_blocks && !consts9) return fail(ctx, HITL_ERR_ARG, "hitl_set_odometry_blocks: null argument");
  if (n_blocks && n_blocks + 1 > ctx->n_poses) return fail(ctx, HITL_ERR_ARG, "hitl_set_odometry_blocks: more blocks than pose pairs");
  for (uint32_t b = 0; b < n_blocks; ++b)
    for (int q = 4; q < 7; ++q) if (!(consts9[9 * b + q] > 0.0f)) return fail(ctx, HITL_ERR_ARG, "hitl_set_odometry_blocks: std-dev must be > 0");
  HITL_CUDA(ctx->d_odo.ensure(9 * (size_t)n_blocks));
  if (n_blocks) HITL_CUDA(cudaMemcpy(ctx->d_odo.p, consts9, 36 * (size_t)n_blocks, cudaMemcpyHostToDevice));
  ctx->nb_odo = n_blocks; ctx->eval_valid = ctx->neq_valid = false; ctx->inc_valid = false;
  return HITL_OK;
}
extern "C" int hitl_set_human_blocks(hitl_ctx* ctx, uint32_t n_blocks, const int32_t* type_pose, const double* targets4) {
  if (!ctx) return HITL_ERR_ARG;
  HITL_DEVICE(ctx);
  if (n_blocks && (!type_pose || !targets4)) return fail(ctx, HITL_ERR_ARG, "hitl_set_human_blocks: null argument");
  for (uint32_t b = 0; b < n_blocks; ++b) {
    const int t = type_pose[2 * b];
    if ((t != 2 && t != 4 && t != 5 && t != 6) || type_pose[2 * b + 1] < 0 || (uint32_t)type_pose[2 * b + 1] >= ctx->n_poses)
      return fail(ctx, HITL_ERR_ARG, "hitl_set_human_blocks: unsupported type or pose out of range");
  }
  HITL_CUDA(ctx->d_hum_i.ensure(2 * (size_t)n_blocks)); HITL_CUDA(ctx->d_hum_d.ensure(4 * (size_t)n_blocks));
  if (n_blocks) {
    HITL_CUDA(cudaMemcpy(ctx->d_hum_i.p, type_pose, 8 * (size_t)n_blocks, cudaMemcpyHostToDevice));
    HITL_CUDA(cudaMemcpy(ctx->d_hum_d.p, targets4, 32 * (size_t)n_blocks, cudaMemcpyHostToDevice));
  }
  ctx->nb_human = n_blocks; ctx->eval_valid = ctx->neq_valid = false; ctx->inc_valid = false;
  return HITL_OK;
}
extern "C" int hitl_set_p2l_glob_blocks(hitl_ctx* ctx, uint32_t n_blocks, const uint32_t* blk_pose, const uint64_t* blk_off, const float* pts_xy,
                                        const float* ln_xy, const float* lo, const uint8_t* valid, float std_dev, float corr) {
  if (!ctx) return HITL_ERR_ARG;
  HITL_DEVICE(ctx);
  if (n_blocks && (!blk_pose || !blk_off || !pts_xy || !ln_xy || !lo || !valid)) return fail(ctx, HITL_ERR_ARG, "hitl_set_p2l_glob_blocks: null argument");
  for (uint32_t b = 0; b < n_blocks; ++b) if (blk_pose[b] >= ctx->n_poses) return fail(ctx, HITL_ERR_ARG, "hitl_set_p2l_glob_blocks: pose out of range");
  if (n_blocks && blk_off[0] != 0) return fail(ctx, HITL_ERR_ARG, "hitl_set_p2l_glob_blocks: blk_off must start at 0");
  for (uint32_t b = 0; b < n_blocks; ++b) if (blk_off[b + 1] < blk_off[b]) return fail(ctx, HITL_ERR_ARG, "hitl_set_p2l_glob_blocks: blk_off must be non-decreasing");
  const uint64_t m = n_blocks ? blk_off[n_blocks] : 0;
  HITL_CUDA(ctx->d_p2lg_pose.ensure(n_blocks)); HITL_CUDA(ctx->d_p2lg_off.ensure(n_blocks + 1)); HITL_CUDA(ctx->d_p2lg_pts.ensure(m));
  HITL_CUDA(ctx->d_p2lg_n.ensure(m)); HITL_CUDA(ctx->d_p2lg_o.ensure(m)); HITL_CUDA(ctx->d_p2lg_v.ensure(m));
  if (n_blocks) {
    HITL_CUDA(cudaMemcpy(ctx->d_p2lg_pose.p, blk_pose, 4 * (size_t)n_blocks, cudaMemcpyHostToDevice));
    HITL_CUDA(cudaMemcpy(ctx->d_p2lg_off.p, blk_off, 8 * (size_t)(n_blocks + 1), cudaMemcpyHostToDevice));
    if (m) {
      HITL_CUDA(cudaMemcpy(ctx->d_p2lg_pts.p, pts_xy, 8 * m, cudaMemcpyHostToDevice)); HITL_CUDA(cudaMemcpy(ctx->d_p2lg_n.p, ln_xy, 8 * m, cudaMemcpyHostToDevice));
      HITL_CUDA(cudaMemcpy(ctx->d_p2lg_o.p, lo, 4 * m, cudaMemcpyHostToDevice)); HITL_CUDA(cudaMemcpy(ctx->d_p2lg_v.p, valid, m, cudaMemcpyHostToDevice));
    }
  }
  ctx->eval_valid = ctx->neq_valid = false; ctx->inc_valid = false;
  ctx->nb_p2lg = n_blocks; ctx->p2lg_std = std_dev; ctx->p2lg_corr = corr;
  return HITL_OK;
}
extern "C" int hitl_set_p2l_blocks(hitl_ctx* ctx, uint64_t n, const uint32_t* pose_idx, const float* pts_xy, const float* ln_xy, const float* lo,
                                   const uint8_t* valid, float std_dev, float corr) {
  if (!ctx) return HITL_ERR_ARG;
  HITL_DEVICE(ctx);
  if (n && (!pose_idx || !pts_xy || !ln_xy || !lo || !valid)) return fail(ctx, HITL_ERR_ARG, "hitl_set_p2l_blocks: null argument");
  for (uint64_t b = 0; b < n; ++b) if (pose_idx[b] >= ctx->n_poses) return fail(ctx, HITL_ERR_ARG, "hitl_set_p2l_blocks: pose out of range");
  HITL_CUDA(ctx->d_p2l_pose.ensure(n)); HITL_CUDA(ctx->d_p2l_pts.ensure(n)); HITL_CUDA(ctx->d_p2l_n.ensure(n)); HITL_CUDA(ctx->d_p2l_o.ensure(n));
  HITL_CUDA(ctx->d_p2l_v.ensure(n));
  if (n) {
    HITL_CUDA(cudaMemcpy(ctx->d_p2l_pose.p, pose_idx, 4 * n, cudaMemcpyHostToDevice)); HITL_CUDA(cudaMemcpy(ctx->d_p2l_pts.p, pts_xy, 8 * n, cudaMemcpyHostToDevice));
    HITL_CUDA(cudaMemcpy(ctx->d_p2l_n.p, ln_xy, 8 * n, cudaMemcpyHostToDevice)); HITL_CUDA(cudaMemcpy(ctx->d_p2l_o.p, lo, 4 * n, cudaMemcpyHostToDevice));
    HITL_CUDA(cudaMemcpy(ctx->d_p2l_v.p, valid, n, cudaMemcpyHostToDevice));
  }
  ctx->eval_valid = ctx->neq_valid = false; ctx->inc_valid = false;
  ctx->nb_p2l = n; ctx->p2l_std = std_dev; ctx->p2l_corr = corr;
  return HITL_OK;
}

extern "C" int hitl_eval_layout_get(hitl_ctx* ctx, hitl_eval_layout* L) {
  if (!ctx || !L) return HITL_ERR_ARG;
  L->n_odometry = ctx->nb_odo; L->n_human = ctx->nb_human; L->n_stf = ctx->nb_stf; L->n_p2l_glob = ctx->nb_p2lg; L->n_p2l = ctx->nb_p2l;
  L->n_residuals = 3 * ctx->nb_odo + 3 * ctx->nb_human + 2 * ctx->nb_stf + ctx->nb_p2lg + ctx->nb_p2l;
  L->n_jacobian = 18 * ctx->nb_odo + 9 * ctx->nb_human + 12 * ctx->nb_stf + 3 * ctx->nb_p2lg + 3 * ctx->nb_p2l;
  return HITL_OK;
}

template <typename T>
static int launch_all(hitl_ctx* ctx, double* d_r, double* d_J, const NeqOut& neq, int want_neq) {
  const double* pose = ctx->d_pose.p;
  size_t ro = 0, jo = 0;
  if (ctx->nb_odo) {
    eval_odometry_kernel<T><<<(uint32_t)((ctx->nb_odo + 127) / 128), 128, 0, ctx->stream>>>(ctx->d_odo.p, pose, (uint32_t)ctx->nb_odo, d_r ? d_r + ro : nullptr,
                                                                                          d_J ? d_J + jo : nullptr, neq, want_neq);
    HITL_LAUNCH_CHECK("eval_odometry_kernel");
  }
  ro += 3 * ctx->nb_odo; jo += 18 * ctx->nb_odo;
  if (ctx->nb_human) {
    eval_human_kernel<T><<<(uint32_t)((ctx->nb_human + 127) / 128), 128, 0, ctx->stream>>>(ctx->d_hum_i.p, ctx->d_hum_d.p, pose, (uint32_t)ctx->nb_human,
                                                                                         d_r ? d_r + ro : nullptr, d_J ? d_J + jo : nullptr, neq, want_neq);
    HITL_LAUNCH_CHECK("eval_human_kernel");
  }
  ro += 3 * ctx->nb_human; jo += 9 * ctx->nb_human;
  if (ctx->nb_stf) {
    const bool fs = ctx->stf_from_search;
    const uint64_t nb = ctx->nb_stf;
    HITL_CUDA(ctx->d_trig.ensure(2 * (size_t)ctx->n_poses));
    pose_trig_kernel<<<(ctx->n_poses + 255) / 256, 256, 0, ctx->stream>>>(pose, ctx->n_poses, ctx->d_trig.p);
    HITL_LAUNCH_CHECK("pose_trig_kernel");
    HITL_KERNEL_BEGIN(HITL_K_EVAL_STF);
    eval_stf_kernel<T><<<(uint32_t)((nb * 32 + kStfThreads - 1) / kStfThreads), kStfThreads, 0, ctx->stream>>>(
        ctx->d_pts.p, ctx->d_nrm.p, ctx->d_off.p, fs ? ctx->d_pair_i.p : ctx->d_blk_i.p, fs ? ctx->d_pair_j.p : ctx->d_blk_j.p,
        (const unsigned long long*)(fs ? ctx->d_pair_off.p : ctx->d_blk_off.p), fs ? ctx->d_k.p : ctx->d_blk_k.p, fs ? ctx->d_idx.p : ctx->d_blk_idx.p, ctx->d_trig.p, nb,
        ctx->stf_std, ctx->stf_corr, d_r ? d_r + ro : nullptr, d_J ? d_J + jo : nullptr, neq, want_neq, (size_t)ctx->nb_odo);
    HITL_KERNEL_END(HITL_K_EVAL_STF);
    HITL_LAUNCH_CHECK("eval_stf_kernel");
  }
  ro += 2 * ctx->nb_stf; jo += 12 * ctx->nb_stf;
  if (ctx->nb_p2lg) {
    eval_p2l_glob_kernel<T><<<(uint32_t)((ctx->nb_p2lg * 32 + 127) / 128), 128, 0, ctx->stream>>>(
        ctx->d_p2lg_pose.p, (const unsigned long long*)ctx->d_p2lg_off.p, ctx->d_p2lg_pts.p, ctx->d_p2lg_n.p, ctx->d_p2lg_o.p, ctx->d_p2lg_v.p, pose,
        (uint32_t)ctx->nb_p2lg, ctx->p2lg_std, ctx->p2lg_corr, d_r ? d_r + ro : nullptr, d_J ? d_J + jo : nullptr, neq, want_neq);
    HITL_LAUNCH_CHECK("eval_p2l_glob_kernel");
  }
  ro += ctx->nb_p2lg; jo += 3 * ctx->nb_p2lg;
  if (ctx->nb_p2l) {
    eval_p2l_kernel<T><<<(uint32_t)((ctx->nb_p2l + 127) / 128), 128, 0, ctx->stream>>>(ctx->d_p2l_pose.p, ctx->d_p2l_pts.p, ctx->d_p2l_n.p, ctx->d_p2l_o.p,
                                                                                     ctx->d_p2l_v.p, pose, ctx->nb_p2l, ctx->p2l_std, ctx->p2l_corr,
                                                                                     d_r ? d_r + ro : nullptr, d_J ? d_J + jo : nullptr, neq, want_neq);
    HITL_LAUNCH_CHECK("eval_p2l_kernel");
  }
  return HITL_OK;
}

extern "C" int hitl_eval(hitl_ctx* ctx, const double* pose_array, int precision, double* r_out, double* J_out, float* ms_out) {
  if (!ctx) return HITL_ERR_ARG;
  HITL_DEVICE(ctx);
  if (!pose_array) return fail(ctx, HITL_ERR_ARG, "hitl_eval: null poses");
  if (ctx->nb_stf && ctx->stf_from_search && !ctx->have_stf) return fail(ctx, HITL_ERR_STATE, "hitl_eval: search result was invalidated");
  hitl_eval_layout L; hitl_eval_layout_get(ctx, &L);
  HITL_CUDA(ctx->d_pose.ensure(3 * (size_t)ctx->n_poses));
  HITL_CUDA(ctx->d_r.ensure(L.n_residuals)); if (J_out) HITL_CUDA(ctx->d_J.ensure(L.n_jacobian));
  HITL_CUDA(cudaMemcpyAsync(ctx->d_pose.p, pose_array, 24 * (size_t)ctx->n_poses, cudaMemcpyHostToDevice, ctx->stream));
  HITL_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
  NeqOut none; none.H_diag = none.g = none.H_off = none.cost = nullptr;
  int rc = precision == 1 ? launch_all<float>(ctx, ctx->d_r.p, J_out ? ctx->d_J.p : nullptr, none, 0)
                          : launch_all<double>(ctx, ctx->d_r.p, J_out ? ctx->d_J.p : nullptr, none, 0);
  if (rc) return rc;
  ctx->eval_valid = J_out != nullptr;
  HITL_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
  if (r_out && L.n_residuals) HITL_CUDA(cudaMemcpyAsync(r_out, ctx->d_r.p, 8 * L.n_residuals, cudaMemcpyDeviceToHost, ctx->stream));
  if (J_out && L.n_jacobian) HITL_CUDA(cudaMemcpyAsync(J_out, ctx->d_J.p, 8 * L.n_jacobian, cudaMemcpyDeviceToHost, ctx->stream));
  HITL_CUDA(cudaStreamSynchronize(ctx->stream));
  if (ms_out) HITL_CUDA(cudaEventElapsedTime(ms_out, ctx->ev[0], ctx->ev[1]));
  return HITL_OK;
}

namespace hitl {
// ---- deterministic normal equations (hitl_set_deterministic) ------------------------------------------------------------------
// The default path adds every block's J^T J / J^T r into the per-pose buffers with FP64 atomics: correct to rounding, but the order
// of the additions — hence the last bits — varies from run to run.  The deterministic path evaluates r and J per block (each block is
// a fixed-order reduction already), then GATHERS: an incidence list per pose, built once per block registration by a stable radix
// sort of (pose, block reference) pairs generated in the canonical block order [odometry | human | stf | p2l_glob | p2l], and one
// warp per pose in which lane e < 9 owns H entry e and lanes 9..11 own g, each walking the pose's list in order.  The cost is a
// two-level sum of fixed shape.  Same inputs, same bits, on every run and for every launch geometry of the other kernels.
constexpr uint64_t kRefKindShift = 60, kRefSideBit = 1ull << 59;
__global__ void incidence_fill_kernel(uint64_t n_odo, uint64_t n_hum, uint64_t n_stf, uint64_t n_p2lg, uint64_t n_p2l, const int32_t* __restrict__ hum_i,
                                      const uint32_t* __restrict__ stf_i, const uint32_t* __restrict__ stf_j, const uint32_t* __restrict__ p2lg_pose,
                                      const uint32_t* __restrict__ p2l_pose, uint32_t* __restrict__ keys, uint64_t* __restrict__ refs) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t a = 2 * n_odo, b = a + n_hum, c = b + 2 * n_stf, d = c + n_p2lg, e = d + n_p2l;
  if (t >= e) return;
  if (t < a) { const uint64_t blk = t >> 1, side = t & 1; keys[t] = (uint32_t)(blk + side); refs[t] = (0ull << kRefKindShift) | (side ? kRefSideBit : 0) | blk; }
  else if (t < b) { const uint64_t blk = t - a; keys[t] = (uint32_t)hum_i[2 * blk + 1]; refs[t] = (1ull << kRefKindShift) | blk; }
  else if (t < c) { const uint64_t u = t - b, blk = u >> 1, side = u & 1; keys[t] = side ? stf_j[blk] : stf_i[blk]; refs[t] = (2ull << kRefKindShift) | (side ? kRefSideBit : 0) | blk; }
  else if (t < d) { const uint64_t blk = t - c; keys[t] = p2lg_pose[blk]; refs[t] = (3ull << kRefKindShift) | blk; }
  else { const uint64_t blk = t - d; keys[t] = p2l_pose[blk]; refs[t] = (4ull << kRefKindShift) | blk; }
}
__global__ void incidence_offsets_kernel(const uint32_t* __restrict__ sorted_keys, uint64_t n_inc, uint32_t n_poses, uint64_t* __restrict__ inc_off) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p > n_poses) return;
  uint64_t lo = 0, hi = n_inc;                                      // first incidence with key >= p
  while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (sorted_keys[mid] < p) lo = mid + 1; else hi = mid; }
  inc_off[p] = lo;
}
struct GatherLayout { uint64_t r_off[5], j_off[5]; };
__global__ void neq_gather_kernel(const uint64_t* __restrict__ inc_off, const uint64_t* __restrict__ refs, uint32_t n_poses, const double* __restrict__ r,
                                  const double* __restrict__ J, GatherLayout L, double* __restrict__ H_diag, double* __restrict__ g) {
  const uint32_t p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, e = threadIdx.x & 31;
  if (p >= n_poses || e >= 12) return;
  const int kRes[5] = {3, 3, 2, 1, 1}, kJac[5] = {18, 9, 12, 3, 3};
  double acc = 0.0;
  for (uint64_t q = inc_off[p]; q < inc_off[p + 1]; ++q) {
    const uint64_t ref = refs[q];
    const int kind = (int)(ref >> kRefKindShift), side = (ref & kRefSideBit) ? 1 : 0;
    const uint64_t blk = ref & (kRefSideBit - 1);
    const int nr = kRes[kind];
    const double* Js = J + L.j_off[kind] + blk * kJac[kind] + side * 3 * nr;      // [rows x 3] wrt this pose
    const double* rs = r + L.r_off[kind] + blk * nr;
    double v = 0.0;
    if (e < 9) { const int a = e / 3, b = e % 3; for (int row = 0; row < nr; ++row) v += Js[3 * row + a] * Js[3 * row + b]; }
    else { const int c = e - 9; for (int row = 0; row < nr; ++row) v += Js[3 * row + c] * rs[row]; }
    acc += v;
  }
  if (e < 9) H_diag[9 * (size_t)p + e] = acc; else g[3 * (size_t)p + (e - 9)] = acc;
}
// H_off of the binary blocks (odometry, then stf): J_a^T J_b, 9 entries per block, from the per-block Jacobians
__global__ void neq_hoff_kernel(uint64_t n_odo, uint64_t n_stf, const double* __restrict__ J, GatherLayout L, double* __restrict__ H_off) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t b = t / 9; const int e = (int)(t % 9);
  if (b >= n_odo + n_stf) return;
  const bool odo = b < n_odo;
  const int nr = odo ? 3 : 2;
  const double* Jb = odo ? J + L.j_off[0] + b * 18 : J + L.j_off[2] + (b - n_odo) * 12;
  const int a3 = e / 3, b3 = e % 3;
  double h = 0.0;
  for (int row = 0; row < nr; ++row) h += Jb[3 * row + a3] * Jb[3 * nr + 3 * row + b3];
  H_off[9 * b + e] = h;
}
constexpr int kCostBlocks = 256, kCostThreads = 256;
__global__ void __launch_bounds__(kCostThreads) neq_cost_partial_kernel(const double* __restrict__ r, uint64_t n, double* __restrict__ partial) {
  __shared__ double sm[kCostThreads];
  double acc = 0.0;
  for (uint64_t i = (uint64_t)blockIdx.x * kCostThreads + threadIdx.x; i < n; i += (uint64_t)kCostBlocks * kCostThreads) acc += r[i] * r[i];
  sm[threadIdx.x] = acc;
  __syncthreads();
  for (int o = kCostThreads / 2; o; o >>= 1) { if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) partial[blockIdx.x] = sm[0];
}
__global__ void neq_cost_final_kernel(const double* __restrict__ partial, double* __restrict__ cost) {
  double acc = 0.0;
  for (int q = 0; q < kCostBlocks; ++q) acc += partial[q];
  *cost = 0.5 * acc;
}

static int ensure_incidence(hitl_ctx* ctx) {
  if (ctx->inc_valid) return HITL_OK;
  const uint64_t n_inc = 2 * ctx->nb_odo + ctx->nb_human + 2 * ctx->nb_stf + ctx->nb_p2lg + ctx->nb_p2l;
  if (n_inc >= 0x7FFFFFFFull) return fail(ctx, HITL_ERR_ARG, "deterministic normal equations: too many block incidences");
  ctx->n_inc = n_inc;
  HITL_CUDA(ctx->d_inc_key.ensure(2 * n_inc)); HITL_CUDA(ctx->d_inc_ref.ensure(2 * n_inc)); HITL_CUDA(ctx->d_inc_off.ensure((size_t)ctx->n_poses + 2));
  if (n_inc) {
    const bool fs = ctx->stf_from_search;
    incidence_fill_kernel<<<(uint32_t)((n_inc + 255) / 256), 256, 0, ctx->stream>>>(ctx->nb_odo, ctx->nb_human, ctx->nb_stf, ctx->nb_p2lg, ctx->nb_p2l, ctx->d_hum_i.p,
                                                                                    fs ? ctx->d_pair_i.p : ctx->d_blk_i.p, fs ? ctx->d_pair_j.p : ctx->d_blk_j.p,
                                                                                    ctx->d_p2lg_pose.p, ctx->d_p2l_pose.p, ctx->d_inc_key.p, ctx->d_inc_ref.p);
    HITL_LAUNCH_CHECK("incidence_fill_kernel");
    size_t tmp = 0;
    HITL_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp, ctx->d_inc_key.p, ctx->d_inc_key.p + n_inc, ctx->d_inc_ref.p, ctx->d_inc_ref.p + n_inc, (int)n_inc, 0, 32, ctx->stream));
    HITL_CUDA(ctx->d_sort_tmp.ensure(tmp));
    HITL_CUDA(cub::DeviceRadixSort::SortPairs(ctx->d_sort_tmp.p, tmp, ctx->d_inc_key.p, ctx->d_inc_key.p + n_inc, ctx->d_inc_ref.p, ctx->d_inc_ref.p + n_inc, (int)n_inc, 0, 32, ctx->stream));
    ctx->launches += 1;
  }
  incidence_offsets_kernel<<<(ctx->n_poses + 1 + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_inc_key.p + n_inc, n_inc, ctx->n_poses, ctx->d_inc_off.p);
  HITL_LAUNCH_CHECK("incidence_offsets_kernel");
  ctx->inc_valid = true;
  return HITL_OK;
}

static int normal_eq_launch_deterministic(hitl_ctx* ctx) {
  const size_t n = ctx->n_poses;
  hitl_eval_layout L; hitl_eval_layout_get(ctx, &L);
  HITL_CUDA(ctx->d_r.ensure(L.n_residuals)); HITL_CUDA(ctx->d_J.ensure(L.n_jacobian)); HITL_CUDA(ctx->d_cost_partial.ensure(kCostBlocks));
  int rc = ensure_incidence(ctx);
  if (rc) return rc;
  NeqOut none; none.H_diag = none.g = none.H_off = none.cost = nullptr;
  rc = launch_all<double>(ctx, ctx->d_r.p, ctx->d_J.p, none, 0);
  if (rc) return rc;
  ctx->eval_valid = true;                            // r and J of every block are resident as after hitl_eval
  GatherLayout G;
  const uint64_t counts[5] = {ctx->nb_odo, ctx->nb_human, ctx->nb_stf, ctx->nb_p2lg, ctx->nb_p2l};
  const int kRes[5] = {3, 3, 2, 1, 1}, kJac[5] = {18, 9, 12, 3, 3};
  uint64_t ro = 0, jo = 0;
  for (int k = 0; k < 5; ++k) { G.r_off[k] = ro; G.j_off[k] = jo; ro += counts[k] * kRes[k]; jo += counts[k] * kJac[k]; }
  if (n) {
    neq_gather_kernel<<<(uint32_t)((n * 32 + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_inc_off.p, ctx->d_inc_ref.p + ctx->n_inc, (uint32_t)n, ctx->d_r.p, ctx->d_J.p, G, ctx->d_neq.p,
                                                                                ctx->d_neq.p + 9 * n);
    HITL_LAUNCH_CHECK("neq_gather_kernel");
  }
  const uint64_t nbin = ctx->nb_odo + ctx->nb_stf;
  if (nbin) {
    neq_hoff_kernel<<<(uint32_t)((nbin * 9 + 255) / 256), 256, 0, ctx->stream>>>(ctx->nb_odo, ctx->nb_stf, ctx->d_J.p, G, ctx->d_hoff.p);
    HITL_LAUNCH_CHECK("neq_hoff_kernel");
  }
  neq_cost_partial_kernel<<<kCostBlocks, kCostThreads, 0, ctx->stream>>>(ctx->d_r.p, L.n_residuals, ctx->d_cost_partial.p);
  HITL_LAUNCH_CHECK("neq_cost_partial_kernel");
  neq_cost_final_kernel<<<1, 1, 0, ctx->stream>>>(ctx->d_cost_partial.p, ctx->d_neq.p + 12 * n);
  HITL_LAUNCH_CHECK("neq_cost_final_kernel");
  return HITL_OK;
}

// Asynchronous part of hitl_normal_eq (also the first half of hitl_normal_eq_allreduce): pose upload, zero fill, the evaluation kernels.
int normal_eq_launch(hitl_ctx* ctx, const double* pose_array) {
  if (ctx->nb_stf && ctx->stf_from_search && !ctx->have_stf) return fail(ctx, HITL_ERR_STATE, "hitl_normal_eq: search result was invalidated");
  const size_t n = ctx->n_poses, nbin = ctx->nb_odo + ctx->nb_stf;
  ctx->neq_valid = false;
  HITL_CUDA(ctx->d_pose.ensure(3 * n));
  HITL_CUDA(ctx->d_neq.ensure(12 * n + 1)); HITL_CUDA(ctx->d_hoff.ensure(9 * nbin));
  HITL_CUDA(cudaMemcpyAsync(ctx->d_pose.p, pose_array, 24 * n, cudaMemcpyHostToDevice, ctx->stream));
  if (ctx->deterministic) {
    const int rcd = normal_eq_launch_deterministic(ctx);
    if (rcd) return rcd;
    ctx->neq_valid = true;
    return HITL_OK;
  }
  HITL_CUDA(cudaMemsetAsync(ctx->d_neq.p, 0, 8 * (12 * n + 1), ctx->stream));
  NeqOut q; q.H_diag = ctx->d_neq.p; q.g = ctx->d_neq.p + 9 * n; q.cost = ctx->d_neq.p + 12 * n; q.H_off = ctx->d_hoff.p;
  const int rc = launch_all<double>(ctx, nullptr, nullptr, q, 1);
  if (rc) return rc;
  ctx->neq_valid = true;
  return HITL_OK;
}
}  // namespace hitl

extern "C" int hitl_normal_eq(hitl_ctx* ctx, const double* pose_array, double* H_diag, double* g, double* H_off, double* cost, float* ms_out) {
  if (!ctx) return HITL_ERR_ARG;
  HITL_DEVICE(ctx);
  if (!pose_array) return fail(ctx, HITL_ERR_ARG, "hitl_normal_eq: null poses");
  const size_t n = ctx->n_poses, nbin = ctx->nb_odo + ctx->nb_stf;
  HITL_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
  const int rc = hitl::normal_eq_launch(ctx, pose_array);
  if (rc) return rc;
  HITL_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
  if (H_diag && n) HITL_CUDA(cudaMemcpyAsync(H_diag, ctx->d_neq.p, 72 * n, cudaMemcpyDeviceToHost, ctx->stream));
  if (g && n) HITL_CUDA(cudaMemcpyAsync(g, ctx->d_neq.p + 9 * n, 24 * n, cudaMemcpyDeviceToHost, ctx->stream));
  if (H_off && nbin) HITL_CUDA(cudaMemcpyAsync(H_off, ctx->d_hoff.p, 72 * nbin, cudaMemcpyDeviceToHost, ctx->stream));
  if (cost) HITL_CUDA(cudaMemcpyAsync(cost, ctx->d_neq.p + 12 * n, 8, cudaMemcpyDeviceToHost, ctx->stream));
  HITL_CUDA(cudaStreamSynchronize(ctx->stream));
  if (ms_out) HITL_CUDA(cudaEventElapsedTime(ms_out, ctx->ev[0], ctx->ev[1]));
  return HITL_OK;
}

extern "C" int hitl_set_deterministic(hitl_ctx* ctx, int on) {
  if (!ctx) return HITL_ERR_ARG;
  ctx->deterministic = on ? 1 : 0;
  ctx->neq_valid = false;
  return HITL_OK;
}

extern "C" int hitl_normal_eq_device(hitl_ctx* ctx, void** dev_ptr, uint64_t* n_doubles) {
  if (!ctx || !dev_ptr || !n_doubles) return HITL_ERR_ARG;
  if (!ctx->d_neq.p) return fail(ctx, HITL_ERR_STATE, "hitl_normal_eq_device: call hitl_normal_eq first");
  *dev_ptr = ctx->d_neq.p; *n_doubles = 12 * (uint64_t)ctx->n_poses + 1;
  return HITL_OK;
}
