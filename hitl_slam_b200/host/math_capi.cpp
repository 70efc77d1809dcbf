// math_capi.cpp — host-side entry points over csrc/hitl_math.h (the host/device-shared scalar
// code), so the libm-parity claim of DESIGN.md can be checked without a GPU.
#include <math.h>
#include <string.h>
#include <atomic>
#include <thread>
#include <vector>
#include "hitl_host.h"
#include <algorithm>
#include "hitl_math.h"
#include "stdsort_exact.h"

extern "C" {
float hitl_host_sinf(float x) { return hitl::sinf_rn(x); }
float hitl_host_cosf(float x) { return hitl::cosf_rn(x); }

// Compares hitl::sinf_rn / cosf_rn with the platform libm over float bit patterns
// first, first+stride, ... (count of them); returns the number of finite inputs whose sin or
// cos bits differ.  stride 1 over [0, 2^32) is the exhaustive check.
uint64_t hitl_host_sincos_mismatches(uint64_t first, uint64_t count, uint64_t stride) {
  unsigned nt = std::thread::hardware_concurrency();
  if (nt == 0) nt = 1;
  if (nt > 64) nt = 64;
  std::atomic<uint64_t> bad(0);
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; ++t)
    th.emplace_back([=, &bad]() {
      uint64_t local = 0;
      for (uint64_t i = t; i < count; i += nt) {
        const uint32_t b = (uint32_t)(first + i * stride);
        float f; memcpy(&f, &b, 4);
        if (!(fabsf(f) <= 3.4028235e38f)) continue;
        const float a = sinf(f), c = cosf(f), a2 = hitl::sinf_rn(f), c2 = hitl::cosf_rn(f);
        if (memcmp(&a, &a2, 4) || memcmp(&c, &c2, 4)) ++local;
      }
      bad += local;
    });
  for (auto& x : th) x.join();
  return bad.load();
}

// Sorts ids 0..n-1 by keys[id] twice: with this image's std::sort and with the restatement of libstdc++'s introsort the GPU
// tree builder uses for segments that contain equal keys (csrc/stdsort_exact.h).  Returns the number of positions that differ.
uint64_t hitl_host_stdsort_mismatches(const float* keys, uint32_t n, uint32_t* out_std, uint32_t* out_restated, uint32_t* heap_sorts) {
  std::vector<uint32_t> a(n), b(n);
  for (uint32_t i = 0; i < n; ++i) a[i] = b[i] = i;
  std::sort(a.begin(), a.end(), [keys](uint32_t x, uint32_t y) { return keys[x] < keys[y]; });
  const uint32_t hs = hitl::stdsort::sort_ids(b.data(), (int64_t)n, [keys](uint32_t id) { return keys[id]; });
  if (heap_sorts) *heap_sorts = hs;
  uint64_t bad = 0;
  for (uint32_t i = 0; i < n; ++i) bad += a[i] != b[i];
  if (out_std) memcpy(out_std, a.data(), 4 * (size_t)n);
  if (out_restated) memcpy(out_restated, b.data(), 4 * (size_t)n);
  return bad;
}

void hitl_host_relative_pose(const double* pose_array, uint32_t src, uint32_t dst, float* out6) {
  const hitl::Aff2 a = hitl::pose_affine(pose_array[3 * src], pose_array[3 * src + 1], pose_array[3 * src + 2]);
  const hitl::Aff2 b = hitl::pose_affine(pose_array[3 * dst], pose_array[3 * dst + 1], pose_array[3 * dst + 2]);
  const hitl::Aff2 T = hitl::affine_mul(hitl::affine_inverse(b), a);
  out6[0] = T.m00; out6[1] = T.m01; out6[2] = T.m10; out6[3] = T.m11; out6[4] = T.tx; out6[5] = T.ty;
}
}
