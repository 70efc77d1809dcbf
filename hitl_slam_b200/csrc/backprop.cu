// backprop.cu — the pose update of COP-SLAM back-propagation on the device, bit for bit as the host loops.
//
// Replaces the two O(L^2) loops of Backprop::BackPropagateError (human_in_the_loop_slam/Backprop.cpp:170-199; L = poses
// between the two features of a correction, up to all of them):
//   rotation     for i in [lo, hi):  d_i = rot_weights[i - lo] * theta;  post_i = T(t_i) * R(d_i) * T(-t_i);  angle_i += d_i;
//                                    for k in (i, hi]:  angle_k += d_i;  t_k = post_i * t_k
//   translation  trans = destination - t_hi;  for i in [lo, hi):  for k in (i, hi]:  t_k += trans_weights[i - lo] * trans
// The reference applies these one after the other on one thread, 0.15 s at L = 10 k.  Pose k receives the SAME sequence of
// float operations here: one CTA walks i = lo .. hi-1; the thread that owns pose i has, by then, applied every earlier step to
// it, so it forms post_i from the current t_i and broadcasts it through shared memory (double-buffered: one barrier per
// step); every thread then applies post_i to the poses k > i it owns.  sin / cos of d_i do not depend on the evolving state
// and are computed for all i up front with the library's glibc-identical sinf_rn / cosf_rn.  The translation pass has no
// dependency between poses at all.  Compiled with --fmad=false; Eigen's operation order:
//   T(t) * R * T(-t):  linear = R,  translation = t + R * (-t);     post * v = (R * v) + translation.
#include "hitl_internal.h"
#include "hitl_math.h"

namespace hitl {
namespace {

constexpr int kBpThreads = 1024;

struct BpStep { float c, s, tx, ty, d; float pad[3]; };

__global__ void __launch_bounds__(kBpThreads) backprop_poses_kernel(float* __restrict__ poses, uint32_t lo, uint32_t hi, const float* __restrict__ rot_w,
                                                                    const float* __restrict__ trans_w, float theta, float dest_x, float dest_y,
                                                                    float2* __restrict__ cs) {
  __shared__ BpStep step[2];
  __shared__ float s_trans[2];
  const uint32_t L = hi - lo + 1, tid = threadIdx.x;
  // d_i, cos d_i, sin d_i for every step (state-independent)
  for (uint32_t a = tid; a + 1 < L; a += kBpThreads) {
    const float d = fmul(rot_w[a], theta);
    cs[a] = make_float2(cosf_rn(d), sinf_rn(d));
  }
  __syncthreads();
  // ---- rotation ----
  for (uint32_t a = 0; a + 1 < L; ++a) {
    BpStep& S = step[a & 1];
    if ((a % kBpThreads) == tid) {                      // owner of pose lo + a
      float* p = poses + 3 * (size_t)(lo + a);
      const float2 r = cs[a];
      const float d = fmul(rot_w[a], theta);
      const float tx = p[0], ty = p[1];
      float rx, ry;
      rot_apply(r.x, r.y, -tx, -ty, &rx, &ry);          // R * (-t)
      S.c = r.x; S.s = r.y; S.tx = fadd(tx, rx); S.ty = fadd(ty, ry); S.d = d;
      p[2] = fadd(p[2], d);
    }
    __syncthreads();
    const float c = S.c, s = S.s, ptx = S.tx, pty = S.ty, d = S.d;
    // poses k = lo + b, b > a, owned by this thread: b = tid (mod kBpThreads)
    uint32_t b = (a + 1) - ((a + 1) % kBpThreads) + tid;
    if (b < a + 1) b += kBpThreads;
    for (; b < L; b += kBpThreads) {
      float* p = poses + 3 * (size_t)(lo + b);
      float rx, ry;
      rot_apply(c, s, p[0], p[1], &rx, &ry);
      p[0] = fadd(rx, ptx); p[1] = fadd(ry, pty);
      p[2] = fadd(p[2], d);
    }
    // the other buffer is rewritten only after the next barrier: no second barrier needed
  }
  __syncthreads();
  // ---- translation ----
  if (tid == 0) { s_trans[0] = fsub(dest_x, poses[3 * (size_t)hi]); s_trans[1] = fsub(dest_y, poses[3 * (size_t)hi + 1]); }
  __syncthreads();
  const float trx = s_trans[0], try_ = s_trans[1];
  for (uint32_t b = tid; b < L; b += kBpThreads) {
    float* p = poses + 3 * (size_t)(lo + b);
    float x = p[0], y = p[1];
    for (uint32_t a = 0; a < b && a + 1 < L; ++a) {
      const float w = trans_w[a];
      x = fadd(x, fmul(w, trx)); y = fadd(y, fmul(w, try_));
    }
    p[0] = x; p[1] = y;
  }
}

}  // namespace
}  // namespace hitl

using namespace hitl;

extern "C" int hitl_backprop_poses(hitl_ctx* ctx, uint32_t n_poses, float* poses_xyt, uint32_t lo, uint32_t hi, const float* rot_weights,
                                   const float* trans_weights, float theta, const float* destination_xy, float* ms_out) {
  if (!ctx) return HITL_ERR_ARG;
  HITL_DEVICE(ctx);
  if (!poses_xyt || !rot_weights || !trans_weights || !destination_xy) return fail(ctx, HITL_ERR_ARG, "hitl_backprop_poses: null argument");
  if (hi >= n_poses || lo >= hi) return fail(ctx, HITL_ERR_ARG, "hitl_backprop_poses: need lo < hi < n_poses");
  const uint32_t L = hi - lo + 1;
  // context-owned scratch, grown on demand: a correction must not pay four cudaMalloc / cudaFree round trips
  DevBuf<float>& d_poses = ctx->d_bp_poses; DevBuf<float>& d_rw = ctx->d_bp_rw; DevBuf<float>& d_tw = ctx->d_bp_tw; DevBuf<float2>& d_cs = ctx->d_bp_cs;
  HITL_CUDA(d_poses.ensure(3 * (size_t)n_poses)); HITL_CUDA(d_rw.ensure(L)); HITL_CUDA(d_tw.ensure(L)); HITL_CUDA(d_cs.ensure(L));
  HITL_CUDA(cudaMemcpyAsync(d_poses.p, poses_xyt, 12 * (size_t)n_poses, cudaMemcpyHostToDevice, ctx->stream));
  HITL_CUDA(cudaMemcpyAsync(d_rw.p, rot_weights, 4 * (size_t)(L - 1), cudaMemcpyHostToDevice, ctx->stream));
  HITL_CUDA(cudaMemcpyAsync(d_tw.p, trans_weights, 4 * (size_t)(L - 1), cudaMemcpyHostToDevice, ctx->stream));
  HITL_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
  backprop_poses_kernel<<<1, kBpThreads, 0, ctx->stream>>>(d_poses.p, lo, hi, d_rw.p, d_tw.p, theta, destination_xy[0], destination_xy[1], d_cs.p);
  HITL_LAUNCH_CHECK("backprop_poses_kernel");
  HITL_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
  HITL_CUDA(cudaMemcpyAsync(poses_xyt, d_poses.p, 12 * (size_t)n_poses, cudaMemcpyDeviceToHost, ctx->stream));
  HITL_CUDA(cudaStreamSynchronize(ctx->stream));
  if (ms_out) HITL_CUDA(cudaEventElapsedTime(ms_out, ctx->ev[0], ctx->ev[1]));
  return HITL_OK;
}
