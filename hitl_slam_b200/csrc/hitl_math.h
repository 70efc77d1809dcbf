// hitl_math.h — scalar building blocks shared by the host mirror and the sm_100a kernels.
//
// Everything in here is written so that the host compiler (g++ -ffp-contract=off) and
// nvcc (--fmad=false) produce the SAME bits:
//   * float geometry is spelled out one IEEE operation at a time, in the order Eigen 3
//     evaluates the expressions the reference uses;
//   * sinf/cosf are a restatement of glibc 2.39's x86-64 FMA ifunc variant
//     (__sinf_fma/__cosf_fma: double-precision reduction + polynomial, explicit fused
//     multiply-adds exactly where that build contracts them), because the reference's
//     Rotation2Df(theta) calls the platform libm and CUDA's sinf/cosf differ from it.
//     tests/test_sincos.py checks this file against the host libm over every float in
//     the working range (|x| < 120 exhaustive stride, large path sampled).
//
// Reference call sites that consume these:
//   Rotation2Df(angle)                JointOptimization.cpp:296-305, :604-606
//   Affine2f inverse / product        JointOptimization.cpp:304
//   Affine2f * Vector2f               JointOptimization.cpp:602
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define HITL_HD __host__ __device__ __forceinline__
#else
#define HITL_HD inline
#endif

namespace hitl {

// ---- exact-order double helpers ------------------------------------------------------
HITL_HD double dmul(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b);
#else
  return a * b;
#endif
}
HITL_HD double dfma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
  return __fma_rn(a, b, c);
#else
  return __builtin_fma(a, b, c);
#endif
}
HITL_HD uint32_t f2u(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}

// ---- exact-order float helpers (never contracted) -------------------------------------
HITL_HD float fmul(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fmul_rn(a, b);
#else
  return a * b;
#endif
}
HITL_HD float fadd(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  return a + b;
#endif
}
HITL_HD float fsub(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fsub_rn(a, b);
#else
  return a - b;
#endif
}
// a*b + c*d, two products rounded, then the sum rounded (Eigen's 2-term dot product).
HITL_HD float dot2(float a, float b, float c, float d) { return fadd(fmul(a, b), fmul(c, d)); }

// ---- glibc 2.39 sinf/cosf (sysdeps/ieee754/flt-32/s_sincosf.h, x86-64 FMA build) --------
// Coefficients read out of the installed libm's __sincosf_table ("libm parity").
// The library keeps a second table with the cosine coefficients negated (selected when
// quadrant bit 1 is set); negating every coefficient negates every intermediate exactly,
// so that table is expressed here as a final sign flip of the cosine polynomial.
constexpr double kHpiInv = 0x1.45f306dc9c883p+23;   // 2/pi * 2^24
constexpr double kHpi = 0x1.921fb54442d18p+0;       // pi/2
constexpr double kC0 = 0x1p0, kC1 = -0x1.ffffffd0c621cp-2, kC2 = 0x1.55553e1068f19p-5,
                 kC3 = -0x1.6c087e89a359dp-10, kC4 = 0x1.99343027bf8c3p-16;
constexpr double kS1 = -0x1.555545995a603p-3, kS2 = 0x1.1107605230bc4p-7,
                 kS3 = -0x1.994eb3774cf24p-13;
constexpr double kPi63 = 0x1.921fb54442d18p-62;     // pi * 2^-63 ... (pi/4 * 2^-61 scale)

// 2/pi bits for the |x| >= 120 reduction (libm's __inv_pio4).
#define HITL_INV_PIO4_INIT { \
    0xa2u, 0xa2f9u, 0xa2f983u, 0xa2f9836eu, 0xf9836e4eu, 0x836e4e44u, 0x6e4e4415u, 0x4e441529u, \
    0x441529fcu, 0x1529fc27u, 0x29fc2757u, 0xfc2757d1u, 0x2757d1f5u, 0x57d1f534u, 0xd1f534ddu, \
    0xf534ddc0u, 0x34ddc0dbu, 0xddc0db62u, 0xc0db6295u, 0xdb629599u, 0x6295993cu, 0x95993c43u, \
    0x993c4390u, 0x3c439041u }
static const uint32_t kInvPio4Host[24] = HITL_INV_PIO4_INIT;
#if defined(__CUDACC__)
static __device__ const uint32_t kInvPio4Dev[24] = HITL_INV_PIO4_INIT;
#endif
HITL_HD uint32_t inv_pio4(int i) {
#if defined(__CUDA_ARCH__)
  return kInvPio4Dev[i];
#else
  return kInvPio4Host[i];
#endif
}

// sin polynomial if n even, cos polynomial if n odd (negated when `neg`); fused exactly as
// the FMA build of the library does.
HITL_HD float sincos_poly(double x, double x2, int neg, int n) {
  if ((n & 1) == 0) {
    const double x3 = dmul(x, x2);
    const double s1 = dfma(x2, kS3, kS2);
    const double x7 = dmul(x3, x2);
    const double s = dfma(x3, kS1, x);
    return (float)dfma(x7, s1, s);
  } else {
    const double x4 = dmul(x2, x2);
    const double c2 = dfma(x2, kC4, kC3);
    const double c1 = dfma(x2, kC1, kC0);
    const double x6 = dmul(x4, x2);
    const double c = dfma(x4, kC2, c1);
    const double r = dfma(x6, c2, c);
    return (float)(neg ? -r : r);
  }
}

HITL_HD double reduce_fast(double x, int* np) {
  const double r = dmul(x, kHpiInv);
  const int n = ((int32_t)r + 0x800000) >> 24;
  *np = n;
  return dfma(-(double)n, kHpi, x);   // vfnmadd: x - n*hpi with one rounding
}

HITL_HD double reduce_large(uint32_t xi, int* np) {
  const int base = (xi >> 26) & 15;
  const int shift = (xi >> 23) & 7;
  xi = (xi & 0xffffff) | 0x800000;
  xi <<= shift;
  uint64_t res0 = (uint32_t)(xi * inv_pio4(base));
  const uint64_t res1 = (uint64_t)xi * inv_pio4(base + 4);
  const uint64_t res2 = (uint64_t)xi * inv_pio4(base + 8);
  res0 = (res2 >> 32) | (res0 << 32);
  res0 += res1;
  const uint64_t n = (res0 + (1ULL << 61)) >> 62;
  res0 -= n << 62;
  const double x = (double)(int64_t)res0;
  *np = (int)n;
  return dmul(x, kPi63);
}

// sign[] of the library table: {+1,-1,-1,+1}[q & 3]
HITL_HD double quad_sign(double x, int q) { q &= 3; return (q == 1 || q == 2) ? -x : x; }

// is_cos: 0 -> sinf, 1 -> cosf (the library evaluates cosf as the polynomial of n^1).
HITL_HD float sincosf_core(float y, int is_cos) {
  const uint32_t yi = f2u(y);
  const uint32_t top = (yi >> 20) & 0x7ff;
  double x = (double)y;
  int n;
  if (top <= 0x3f3) {                         // |y| < pi/4
    if (top <= 0x397) {                       // |y| < 2^-12
      return is_cos ? 1.0f : y;
    }
    return sincos_poly(x, dmul(x, x), 0, is_cos);
  } else if (top <= 0x42e) {                  // |y| < 120
    x = reduce_fast(x, &n);
    return sincos_poly(quad_sign(x, n), dmul(x, x), n & 2, n ^ is_cos);
  } else if (top <= 0x7f7) {                  // finite
    const int sign = (int)(yi >> 31);
    x = reduce_large(yi, &n);
    const int ns = n + sign;
    return sincos_poly(quad_sign(x, ns), dmul(x, x), ns & 2, n ^ is_cos);
  }
  return y - y;                               // inf/nan -> nan (errno side effects dropped)
}
HITL_HD float sinf_rn(float y) { return sincosf_core(y, 0); }
HITL_HD float cosf_rn(float y) { return sincosf_core(y, 1); }

// ---- 2-D float affine algebra in Eigen's evaluation order --------
struct Aff2 {   // [ m00 m01 | tx ]
  float m00, m01, m10, m11, tx, ty;   // [ m10 m11 | ty ]
};

// Translation2Df(x,y) * Rotation2Df(theta): doubles are narrowed to float first.
HITL_HD Aff2 pose_affine(double x, double y, double theta) {
  const float th = (float)theta;
  const float s = sinf_rn(th), c = cosf_rn(th);
  Aff2 a; a.m00 = c; a.m01 = -s; a.m10 = s; a.m11 = c; a.tx = (float)x; a.ty = (float)y;
  return a;
}
// Transform::inverse(Eigen::Affine): 2x2 adjugate * (1/det), translation = -(Linv * t).
HITL_HD Aff2 affine_inverse(const Aff2& a) {
  const float det = fsub(fmul(a.m00, a.m11), fmul(a.m10, a.m01));
  const float inv = 1.0f / det;
  Aff2 r;
  r.m00 = fmul(a.m11, inv);
  r.m10 = fmul(-a.m10, inv);
  r.m01 = fmul(-a.m01, inv);
  r.m11 = fmul(a.m00, inv);
  r.tx = dot2(-r.m00, a.tx, -r.m01, a.ty);
  r.ty = dot2(-r.m10, a.tx, -r.m11, a.ty);
  return r;
}
// lhs * rhs: linear = L1*L2, translation = L1*t2 + t1.
HITL_HD Aff2 affine_mul(const Aff2& l, const Aff2& r) {
  Aff2 o;
  o.m00 = dot2(l.m00, r.m00, l.m01, r.m10);
  o.m01 = dot2(l.m00, r.m01, l.m01, r.m11);
  o.m10 = dot2(l.m10, r.m00, l.m11, r.m10);
  o.m11 = dot2(l.m10, r.m01, l.m11, r.m11);
  o.tx = fadd(dot2(l.m00, r.tx, l.m01, r.ty), l.tx);
  o.ty = fadd(dot2(l.m10, r.tx, l.m11, r.ty), l.ty);
  return o;
}
HITL_HD void affine_apply(const Aff2& a, float x, float y, float* ox, float* oy) {
  *ox = fadd(dot2(a.m00, x, a.m01, y), a.tx);
  *oy = fadd(dot2(a.m10, x, a.m11, y), a.ty);
}
// Rotation2Df(angle) * v
HITL_HD void rot_apply(float c, float s, float x, float y, float* ox, float* oy) {
  *ox = dot2(c, x, -s, y);
  *oy = dot2(s, x, c, y);
}

}  // namespace hitl
