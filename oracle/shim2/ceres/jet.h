// Minimal stand-in for <ceres/jet.h> (Ceres Solver is absent from this image): forward-mode dual
// numbers with the arithmetic of ceres::Jet as published, ONLY so that the
// reference's residual_functors.h can be compiled where it lies.  Test infrastructure.
#pragma once
#include <cmath>
namespace ceres {
template <typename T, int N>
struct Jet {
  T a;
  T v[N];
  Jet() : a() { for (int i = 0; i < N; ++i) v[i] = T(); }
  explicit Jet(const T& s) : a(s) { for (int i = 0; i < N; ++i) v[i] = T(); }
  Jet(const T& s, int k) : a(s) { for (int i = 0; i < N; ++i) v[i] = T(); v[k] = T(1); }
};
#define HITL_JET template <typename T, int N> inline
HITL_JET Jet<T, N> operator+(const Jet<T, N>& f, const Jet<T, N>& g) { Jet<T, N> h; h.a = f.a + g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] + g.v[i]; return h; }
HITL_JET Jet<T, N> operator-(const Jet<T, N>& f, const Jet<T, N>& g) { Jet<T, N> h; h.a = f.a - g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] - g.v[i]; return h; }
HITL_JET Jet<T, N> operator-(const Jet<T, N>& f) { Jet<T, N> h; h.a = -f.a; for (int i = 0; i < N; ++i) h.v[i] = -f.v[i]; return h; }
HITL_JET Jet<T, N> operator*(const Jet<T, N>& f, const Jet<T, N>& g) { Jet<T, N> h; h.a = f.a * g.a; for (int i = 0; i < N; ++i) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h; }
HITL_JET Jet<T, N> operator/(const Jet<T, N>& f, const Jet<T, N>& g) {
  Jet<T, N> h; const T gi = T(1.0) / g.a; h.a = f.a * gi;
  for (int i = 0; i < N; ++i) h.v[i] = (f.v[i] - h.a * g.v[i]) * gi;
  return h;
}
HITL_JET Jet<T, N> operator+(const Jet<T, N>& f, T s) { Jet<T, N> h = f; h.a = f.a + s; return h; }
HITL_JET Jet<T, N> operator+(T s, const Jet<T, N>& f) { Jet<T, N> h = f; h.a = s + f.a; return h; }
HITL_JET Jet<T, N> operator-(const Jet<T, N>& f, T s) { Jet<T, N> h = f; h.a = f.a - s; return h; }
HITL_JET Jet<T, N> operator-(T s, const Jet<T, N>& f) { Jet<T, N> h; h.a = s - f.a; for (int i = 0; i < N; ++i) h.v[i] = -f.v[i]; return h; }
HITL_JET Jet<T, N> operator*(const Jet<T, N>& f, T s) { Jet<T, N> h; h.a = f.a * s; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * s; return h; }
HITL_JET Jet<T, N> operator*(T s, const Jet<T, N>& f) { Jet<T, N> h; h.a = f.a * s; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * s; return h; }
HITL_JET Jet<T, N> operator/(const Jet<T, N>& f, T s) { const T si = T(1.0) / s; Jet<T, N> h; h.a = f.a * si; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * si; return h; }
HITL_JET Jet<T, N>& operator+=(Jet<T, N>& f, const Jet<T, N>& g) { f = f + g; return f; }
HITL_JET bool operator<(const Jet<T, N>& f, const Jet<T, N>& g) { return f.a < g.a; }
HITL_JET bool operator>(const Jet<T, N>& f, const Jet<T, N>& g) { return f.a > g.a; }
HITL_JET bool operator<=(const Jet<T, N>& f, const Jet<T, N>& g) { return f.a <= g.a; }
HITL_JET bool operator>=(const Jet<T, N>& f, const Jet<T, N>& g) { return f.a >= g.a; }
HITL_JET bool operator==(const Jet<T, N>& f, const Jet<T, N>& g) { return f.a == g.a; }
HITL_JET bool operator!=(const Jet<T, N>& f, const Jet<T, N>& g) { return f.a != g.a; }
HITL_JET bool operator<(const Jet<T, N>& f, T s) { return f.a < s; }
HITL_JET bool operator>(const Jet<T, N>& f, T s) { return f.a > s; }
HITL_JET bool operator<=(const Jet<T, N>& f, T s) { return f.a <= s; }
HITL_JET bool operator>=(const Jet<T, N>& f, T s) { return f.a >= s; }
HITL_JET bool operator<(T s, const Jet<T, N>& f) { return s < f.a; }
HITL_JET bool operator>(T s, const Jet<T, N>& f) { return s > f.a; }
HITL_JET Jet<T, N> sqrt(const Jet<T, N>& f) { Jet<T, N> h; h.a = std::sqrt(f.a); const T t = T(1.0) / (T(2.0) * h.a); for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * t; return h; }
HITL_JET Jet<T, N> sin(const Jet<T, N>& f) { Jet<T, N> h; h.a = std::sin(f.a); const T c = std::cos(f.a); for (int i = 0; i < N; ++i) h.v[i] = c * f.v[i]; return h; }
HITL_JET Jet<T, N> cos(const Jet<T, N>& f) { Jet<T, N> h; h.a = std::cos(f.a); const T s = -std::sin(f.a); for (int i = 0; i < N; ++i) h.v[i] = s * f.v[i]; return h; }
HITL_JET Jet<T, N> atan2(const Jet<T, N>& g, const Jet<T, N>& f) {
  Jet<T, N> h; h.a = std::atan2(g.a, f.a); const T t = T(1.0) / (f.a * f.a + g.a * g.a);
  for (int i = 0; i < N; ++i) h.v[i] = t * (f.a * g.v[i] - g.a * f.v[i]);
  return h;
}
HITL_JET Jet<T, N> pow(const Jet<T, N>& f, double p) { Jet<T, N> h; h.a = std::pow(f.a, p); const T t = p * std::pow(f.a, p - 1.0); for (int i = 0; i < N; ++i) h.v[i] = t * f.v[i]; return h; }
HITL_JET bool IsFinite(const Jet<T, N>& f) { return std::isfinite(f.a); }
#undef HITL_JET
inline double sqrt(double x) { return std::sqrt(x); }
inline double sin(double x) { return std::sin(x); }
inline double cos(double x) { return std::cos(x); }
inline double atan2(double y, double x) { return std::atan2(y, x); }
inline double pow(double x, double p) { return std::pow(x, p); }
inline bool IsFinite(double x) { return std::isfinite(x); }
}  // namespace ceres
