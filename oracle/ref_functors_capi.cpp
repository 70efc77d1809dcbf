// ref_functors_capi.cpp — C entry points over the REFERENCE's own residual functors
// (human_in_the_loop_slam/residual_functors.h) and DistanceToLineSegment (shared/math/eigen_helper.h),
// compiled from /root/reference where they lie against oracle/shim2/ (Eigen / ceres::Jet / glog
// stand-ins).  Output: oracle/_ref/libfunctors_ref.so.  TEST INFRASTRUCTURE ONLY: pins the oracle's
// restated functors (values and auto-diff Jacobians) to the reference's code.
#include <stdint.h>
#include <vector>
#include "residual_functors.h"
#include "eigen_helper.h"

typedef ceres::Jet<double, 6> J6;
typedef ceres::Jet<double, 3> J3;

template <typename F, int R>
static void autodiff2(const F& f, const double* x0, const double* x1, double* res, double* j0, double* j1) {
  J6 a[3], b[3], r[R];
  for (int i = 0; i < 3; ++i) { a[i] = J6(x0[i], i); b[i] = J6(x1[i], 3 + i); }
  f(a, b, r);
  for (int q = 0; q < R; ++q) { res[q] = r[q].a; for (int c = 0; c < 3; ++c) { j0[3 * q + c] = r[q].v[c]; j1[3 * q + c] = r[q].v[3 + c]; } }
}
template <typename F, int R>
static void autodiff1(const F& f, const double* x0, double* res, double* j0) {
  J3 a[3], r[R];
  for (int i = 0; i < 3; ++i) a[i] = J3(x0[i], i);
  f(a, r);
  for (int q = 0; q < R; ++q) { res[q] = r[q].a; for (int c = 0; c < 3; ++c) j0[3 * q + c] = r[q].v[c]; }
}
static std::vector<Eigen::Vector2f> vecs(const float* p, uint32_t n) {
  std::vector<Eigen::Vector2f> v(n);
  for (uint32_t i = 0; i < n; ++i) v[i] = Eigen::Vector2f(p[2 * i], p[2 * i + 1]);
  return v;
}

extern "C" {
// PointToPointGlobConstraint (residual_functors.h:768-848): res[2], j0/j1 row-major 2x3; value-only evaluation in res_plain.
void ref_p2p_glob(uint32_t m, const float* p0, const float* p1, const float* n0, const float* n1, float std_dev, float corr, const double* x0, const double* x1,
                  double* res, double* j0, double* j1, double* res_plain) {
  PointToPointGlobConstraint f(0, 1, vecs(p0, m), vecs(p1, m), vecs(n0, m), vecs(n1, m), std_dev, corr);
  autodiff2<PointToPointGlobConstraint, 2>(f, x0, x1, res, j0, j1);
  f(x0, x1, res_plain);
}
// PoseConstraint (:1054-1133): consts9 = axis_transform row-major, 3 std-devs, radial_translation, rotation.
void ref_pose_constraint(const float* c, const double* x0, const double* x1, double* res, double* j0, double* j1) {
  Eigen::Matrix2f A; A(0, 0) = c[0]; A(0, 1) = c[1]; A(1, 0) = c[2]; A(1, 1) = c[3];
  PoseConstraint f(A, c[4], c[5], c[6], c[7], c[8]);
  autodiff2<PoseConstraint, 3>(f, x0, x1, res, j0, j1);
}
// Human-imposed constraints (:1299-1415): type 2 colocation (3), 4 colinear (2), 5 perpendicular (1), 6 parallel (1).
int ref_human(int type, const double* tg, const double* x, double* res, double* j0) {
  if (type == 2) { ColocationHumanImposedConstraint f(tg[0], tg[1], tg[2]); autodiff1<ColocationHumanImposedConstraint, 3>(f, x, res, j0); return 3; }
  if (type == 4) { ColinearHumanImposedConstraint f(tg[0], tg[1], tg[2], tg[3]); autodiff1<ColinearHumanImposedConstraint, 2>(f, x, res, j0); return 2; }
  if (type == 5) { PerpendicularHumanImposedConstraint f(tg[2]); autodiff1<PerpendicularHumanImposedConstraint, 1>(f, x, res, j0); return 1; }
  if (type == 6) { ParallelHumanImposedConstraint f(tg[2]); autodiff1<ParallelHumanImposedConstraint, 1>(f, x, res, j0); return 1; }
  return 0;
}
// PointToLineGlobConstraint (:314-385) and PointToLineConstraint (:557-622).
void ref_p2l_glob(uint32_t m, const float* pts, const float* ln, const float* lo, const uint8_t* valid, float std_dev, float corr, const double* x, double* res, double* j0) {
  std::vector<float> off(lo, lo + m);
  std::vector<bool> v(m);
  for (uint32_t i = 0; i < m; ++i) v[i] = valid[i] != 0;
  const std::vector<Eigen::Vector2f> points = vecs(pts, m);   // the functor keeps a REFERENCE to this vector (residual_functors.h:368)
  PointToLineGlobConstraint f(0, points, vecs(ln, m), off, v, std_dev, corr);
  autodiff1<PointToLineGlobConstraint, 1>(f, x, res, j0);
}
void ref_p2l(const float* pt, const float* ln, float lo, int valid, float std_dev, float corr, const double* x, double* res, double* j0) {
  PointToLineConstraint f(0, 0, Eigen::Vector2f(pt[0], pt[1]), Eigen::Vector2f(ln[0], ln[1]), lo, valid != 0, std_dev, corr);
  autodiff1<PointToLineConstraint, 1>(f, x, res, j0);
}
// Eigen::DistanceToLineSegment (shared/math/eigen_helper.h:66-81), float.
void ref_distance_to_line_segment(uint32_t n, const float* p0, const float* p1, const float* pts, float* out) {
  const Eigen::Vector2f a(p0[0], p0[1]), b(p1[0], p1[1]);
  for (uint32_t i = 0; i < n; ++i) out[i] = Eigen::DistanceToLineSegment(a, b, Eigen::Vector2f(pts[2 * i], pts[2 * i + 1]));
}
}
