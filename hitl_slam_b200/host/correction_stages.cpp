// correction_stages.cpp — see correction_stages.h.  Built with -ffp-contract=off; rotations use hitl::sinf_rn / cosf_rn
// (bit-identical to glibc's sinf / cosf), acosf is the platform's as in the reference (acos(float) resolves to the float
// overload because <math.h> is in scope there, SURVEY.md §8a note).
#include "correction_stages.h"
#include <math.h>
#include <stdexcept>
#include "hitl_math.h"
#include "../../include/hitl_gpu.h"

namespace hitl {
namespace {

inline Vector2f unit(Vector2f v) {                           // Eigen normalized(): v / sqrt(squaredNorm) when non-zero
  const float z = v.x * v.x + v.y * v.y;
  if (z > 0.0f) { const float n = sqrtf(z); return Vector2f(v.x / n, v.y / n); }
  return v;
}
inline Vector2f rotate(float angle, Vector2f v) {            // Rotation2Df(angle) * v
  float ox, oy;
  rot_apply(cosf_rn(angle), sinf_rn(angle), v.x, v.y, &ox, &oy);
  return Vector2f(ox, oy);
}
inline Vector2f mid(Vector2f a, Vector2f b) { return Vector2f((a.x + b.x) * 0.5f, (a.y + b.y) * 0.5f); }   // (a + b) / 2 == 0.5 * (a + b) exactly

}  // namespace

// Every mode is "rotate the corrected poses by theta about cmA, then carry cmA to `target`":  p1 = target + R (p0 - cmA).
//   line segment   target = cmB,                         theta = angle from A to B           (:150-181)
//   colinear       target = cmB + ((cmA - cmB) . B) B,   theta = angle from A to B           (:229-257)
//   perpendicular  target = cmA,                         theta = angle from A to B -/+ pi/2  (:259-293, double arithmetic)
//   parallel       target = cmA,                         theta = angle from A to B           (:295-316)
void AppExpCorrect::CalculateExplicitCorrections(std::vector<CorrectionPair>* corrections) const {
  if (selected_points_.size() != 4) throw std::runtime_error("AppExpCorrect: four selected points expected");
  const Vector2f* sp = selected_points_.data();
  const Vector2f cmA = mid(sp[1], sp[0]), cmB = mid(sp[3], sp[2]);
  const Vector2f A = unit(sp[1] - sp[0]), B = unit(sp[3] - sp[2]);
  const float cross = A.x * B.y - A.y * B.x;
  const float ang = acosf(dot(A, B));
  float theta;
  Vector2f target = cmA;
  switch (correction_type_) {
    case CorrectionType::kLineSegmentCorrection: theta = cross < 0.0f ? -ang : ang; target = cmB; break;
    case CorrectionType::kColinearCorrection: {
      theta = cross >= 0.0f ? ang : -ang;
      const float alpha = dot(cmA - cmB, B);
      target = cmB + alpha * B;
    } break;
    case CorrectionType::kPerpendicularCorrection: {
      double t = cross < 0.0f ? -(double)ang : (double)ang;
      if (t == M_PI / 2.0 || t == -M_PI / 2.0) t = 0.0;
      else if (t > 0.0) t = -(-t + M_PI / 2.0);
      else t = -(-t - M_PI / 2.0);
      theta = (float)t;
    } break;
    case CorrectionType::kParallelCorrection: theta = cross >= 0.0f ? ang : -ang; break;
    default: return;                                         // point / corner corrections: unsupported in the reference as well
  }
  for (size_t i = 0; i < corrected_poses_.size(); ++i) {
    const int id = corrected_poses_[i];
    const Vector2f p0 = poses_.at((size_t)id).translation;
    const Vector2f p1 = target + rotate(theta, p0 - cmA);
    corrections->push_back(CorrectionPair(id, Vector3f{{p1.x - p0.x, p1.y - p0.y, theta}}));
  }
}

// Corrections are grouped into runs of consecutive pose ids (FindContiguousGroups, :360-385); only the first run is applied
// (:417-445): its poses move by their own correction, every later pose follows rigidly the last pose of the run (:387-415).
Vector3f AppExpCorrect::AppExpCorrections() {
  std::vector<CorrectionPair> corrections;
  CalculateExplicitCorrections(&corrections);
  std::vector<int> slot(poses_.size(), -1);                  // last correction naming each pose
  for (size_t j = 0; j < corrections.size(); ++j) slot[(size_t)corrections[j].first] = (int)j;
  size_t first = 0;
  while (first < poses_.size() && slot[first] < 0) ++first;
  applied_ = first < poses_.size();
  if (!applied_) return correction_;
  size_t last = first;
  while (last + 1 < poses_.size() && slot[last + 1] >= 0) ++last;
  for (size_t i = first; i <= last; ++i) {
    const Vector3f& c = corrections[(size_t)slot[i]].second;
    poses_[i].translation.x += c[0]; poses_[i].translation.y += c[1]; poses_[i].angle += c[2];
  }
  const Vector3f lc = corrections[(size_t)slot[last]].second;
  const Vector2f pivot = poses_[last].translation;
  for (size_t k = last + 1; k < poses_.size(); ++k) {
    poses_[k].angle += lc[2];
    const Vector2f moved = pivot + rotate(lc[2], poses_[k].translation - pivot);
    poses_[k].translation = moved + Vector2f(lc[0], lc[1]);
  }
  return corrections[(size_t)slot[first]].second;
}

void AppExpCorrect::Run() {
  correction_ = AppExpCorrections();
  new_human_constraints_ = CalculateConstraintTargets(poses_, selected_points_, correction_type_, anchor_poses_, corrected_poses_);
}

// Backprop.cpp:98-200.  Variances: rotation = cov(2,2), translation = mean of cov(0,0), cov(1,1); weights over [lo, hi] with the
// destination's variance fused in; covariances of [lo, hi) shrink by beta (the reference scales (0,2) twice and (1,2) never —
// kept); then the O(L^2) pose update, on the device.
void Backprop::BackPropagateError() {
  const int lo = backprop_bounds_.first, hi = backprop_bounds_.second;
  if (lo < 0 || hi >= (int)poses_.size() || covariances_.size() < poses_.size()) throw std::runtime_error("Backprop: bounds outside the pose graph");
  const float dest_rot_var = 0.0001f, dest_trans_var = 0.001f;
  const float destination[2] = {poses_[hi].translation.x + correction_[0], poses_[hi].translation.y + correction_[1]};
  auto rot_sigma = [&](int i) { return covariances_[(size_t)i][8]; };
  auto trans_sigma = [&](int i) { return (float)((covariances_[(size_t)i][0] + covariances_[(size_t)i][4]) / 2.0); };
  float sum_rot = 0.0f, sum_trans = 0.0f;
  for (int i = lo; i <= hi; ++i) { sum_rot += rot_sigma(i); sum_trans += trans_sigma(i); }
  sum_rot += dest_rot_var; sum_trans += dest_trans_var;
  std::vector<float> rot_w, trans_w;
  for (int i = lo; i <= hi; ++i) { rot_w.push_back(rot_sigma(i) / sum_rot); trans_w.push_back(trans_sigma(i) / sum_trans); }
  const float rot_beta = 1 / (1 + (rot_sigma(hi - 1) / dest_rot_var));
  const float trans_beta = 1 / (1 + (trans_sigma(hi - 1) / dest_trans_var));
  for (int i = lo; i < hi; ++i) {
    Matrix3f& c = covariances_[(size_t)i];
    c[0] *= trans_beta; c[1] *= trans_beta; c[3] *= trans_beta; c[4] *= trans_beta;
    c[2] *= rot_beta; c[2] *= rot_beta;
    c[6] *= rot_beta; c[7] *= rot_beta; c[8] *= rot_beta;
  }
  if (!ctx_) throw std::runtime_error("Backprop needs a GPU context (no CPU fallback)");
  std::vector<float> xyt(3 * poses_.size());
  for (size_t i = 0; i < poses_.size(); ++i) { xyt[3 * i] = poses_[i].translation.x; xyt[3 * i + 1] = poses_[i].translation.y; xyt[3 * i + 2] = poses_[i].angle; }
  if (hitl_backprop_poses(ctx_, (uint32_t)poses_.size(), xyt.data(), (uint32_t)lo, (uint32_t)hi, rot_w.data(), trans_w.data(), correction_[2], destination, &last_device_ms_) != HITL_OK)
    throw std::runtime_error(std::string("hitl_backprop_poses: ") + hitl_last_error(ctx_));
  for (size_t i = 0; i < poses_.size(); ++i) { poses_[i].translation = Vector2f(xyt[3 * i], xyt[3 * i + 1]); poses_[i].angle = xyt[3 * i + 2]; }
}

void Backprop::Run() {
  if (backprop_bounds_.first < backprop_bounds_.second) BackPropagateError();
}

}  // namespace hitl
