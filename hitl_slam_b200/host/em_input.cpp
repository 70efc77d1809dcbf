// em_input.cpp — see em_input.h.
#include "em_input.h"

#include <math.h>
#include <algorithm>
#include <stdexcept>

namespace hitl {

namespace {

// Distance of one inlier to a segment of fixed midpoint (cmx, cmy) and half-length len whose only
// free parameter is its direction angle theta — the residual SegFitEM minimises (EMinput.cpp:107-149).
// T is double or a Jet; pow(x, 2) keeps the reference's derivative form.
struct SegmentAngleResidual {
  SegmentAngleResidual(double px, double py, double cmx, double cmy, double len) : px_(px), py_(py), cmx_(cmx), cmy_(cmy), len_(len) {}
  template <typename T>
  bool operator()(const T* const theta, T* residual) const {
    using ceres::cos; using ceres::sin; using ceres::sqrt; using ceres::pow;
    T ax = cos(theta[0]), ay = sin(theta[0]);
    const T inv = sqrt(ax * ax + ay * ay);                 // alpha.normalize()
    ax = ax / inv; ay = ay / inv;
    const T e1x = T(cmx_) + T(len_) * ax, e1y = T(cmy_) + T(len_) * ay;
    const T e2x = T(cmx_) - T(len_) * ax, e2y = T(cmy_) - T(len_) * ay;
    const T dx = e2x - e1x, dy = e2y - e1y;
    const T t = ((T(px_) - e1x) * dx + (T(py_) - e1y) * dy) / (pow(dx, 2) + pow(dy, 2));
    if (t < 0.0) {
      residual[0] = sqrt(pow(T(px_) - e1x, 2) + pow(T(py_) - e1y, 2));
    } else if (t > 1.0) {
      residual[0] = sqrt(pow(T(px_) - e2x, 2) + pow(T(py_) - e2y, 2));
    } else {
      const T qx = e1x + t * dx, qy = e1y + t * dy;
      residual[0] = sqrt(pow(T(px_) - qx, 2) + pow(T(py_) - qy, 2));
    }
    return true;
  }
  const double px_, py_, cmx_, cmy_, len_;
};

}  // namespace

EMInput::EMInput(hitl_ctx* ctx) : ctx_(ctx) {
  if (!ctx) throw std::invalid_argument("EMInput needs a hitl_ctx: the hot path has no CPU implementation");
  backprop_bounds_ = std::make_pair(0, 0);
}
EMInput::~EMInput() {}

void EMInput::check(int rc, const char* where) {
  if (rc == HITL_OK) return;
  last_error_ = std::string(where) + ": " + hitl_last_error(ctx_);
  throw std::runtime_error(last_error_);
}

void EMInput::UploadWorldClouds() {
  size_t total = 0;
  for (const auto& c : local_version_point_clouds_) total += c.size();
  std::vector<float> w(2 * std::max<size_t>(total, 1));
  size_t o = 0;
  for (const auto& c : local_version_point_clouds_)
    for (const Vector2f& p : c) { w[2 * o] = p.x; w[2 * o + 1] = p.y; ++o; }
  check(hitl_set_world_clouds(ctx_, w.data()), "hitl_set_world_clouds");
  world_clouds_resident_ = true;
}

std::vector<Vector2f> EMInput::SegFitEM(double* p1, double* p2, double* cm, double* data, int size) {
  (void)cm;   // unused by the reference as well (EMinput.cpp:152-191)
  return FitSegmentAngle(p1, p2, data, size);
}

// The M-step proper: pure host code, no device state.
std::vector<Vector2f> FitSegmentAngle(const double* p1, const double* p2, const double* data, int size, double* theta_out, int* iterations_out) {
  const double icm[2] = {(p1[0] + p2[0]) / 2.0, (p1[1] + p2[1]) / 2.0};
  const double hy = sqrt(pow(p1[0] - p2[0], 2) + pow(p1[1] - p2[1], 2));
  const double ad = fabs(p1[0] - p2[0]);
  double theta[1] = {acos(ad / hy)};   // in [0, pi/2]: the sign of the slope is dropped, as in the reference
  ceres::Problem problem;
  for (int i = 0; i < size; ++i)
    problem.AddResidualBlock(new ceres::AutoDiffCostFunction<SegmentAngleResidual, 1, 1>(new SegmentAngleResidual(data[2 * i], data[2 * i + 1], icm[0], icm[1], hy / 2.0)),
                             NULL, theta);
  ceres::Solver::Options options;
  options.max_num_iterations = 25;
  options.linear_solver_type = ceres::DENSE_QR;
  options.minimizer_progress_to_stdout = false;
  ceres::Solver::Summary summary;
  if (size > 0) ceres::Solve(options, &problem, &summary);
  if (theta_out) *theta_out = theta[0];
  if (iterations_out) *iterations_out = summary.num_successful_steps + summary.num_unsuccessful_steps;
  double ax = cos(theta[0]), ay = sin(theta[0]);
  const double len = sqrt(ax * ax + ay * ay);
  ax /= len; ay /= len;
  std::vector<Vector2f> fit(2);
  fit[0] = Vector2f((float)(icm[0] + (hy / 2.0) * ax), (float)(icm[1] + (hy / 2.0) * ay));
  fit[1] = Vector2f((float)(icm[0] - (hy / 2.0) * ax), (float)(icm[1] - (hy / 2.0) * ay));
  return fit;
}

void EMInput::AutomaticEndpointAdjustment() {
  if (!world_clouds_resident_) UploadWorldClouds();
  if (device_m_step_) {
    // E-step AND M-step on the device (SegFitEM's LM runs on the resident inliers), the rounds of both strokes chained: the strokes
    // are independent (each round reads its own stroke and the fixed world clouds), so their loops run side by side and the host
    // waits once per device_chain_rounds_ rounds.  Each stroke's rounds are consumed in order under the reference's loop condition.
    const double thresh = 0.05;
    const size_t ns = std::min<size_t>(selected_points_.size() / 2, 2);
    bool active[2] = {false, false};
    for (size_t k = 0; k < ns; ++k) { em_rounds_[k] = 0; active[k] = max_em_rounds_ > 0; }
    for (;;) {
      uint32_t idx[2], na = 0;
      int rounds = std::max(1, std::min(device_chain_rounds_, 4));
      for (size_t k = 0; k < ns; ++k)
        if (active[k]) { idx[na++] = (uint32_t)k; rounds = std::min(rounds, max_em_rounds_ - em_rounds_[k]); }
      if (na == 0) break;
      float in[8], out[4 * 2 * 4];
      hitl_em_fit_info fi[2 * 4];
      for (uint32_t a = 0; a < na; ++a) {
        in[4 * a] = selected_points_[2 * idx[a]].x; in[4 * a + 1] = selected_points_[2 * idx[a]].y;
        in[4 * a + 2] = selected_points_[2 * idx[a] + 1].x; in[4 * a + 3] = selected_points_[2 * idx[a] + 1].y;
      }
      check(hitl_em_refit_chain(ctx_, na, in, 0.03, 25, (uint32_t)rounds, out, fi), "hitl_em_refit_chain");
      for (uint32_t a = 0; a < na; ++a) {
        const size_t k = idx[a];
        for (int r = 0; r < rounds && active[k]; ++r) {
          const size_t slot = (size_t)r * na + a;
          const Vector2f fit0(out[4 * slot], out[4 * slot + 1]), fit1(out[4 * slot + 2], out[4 * slot + 3]);
          const double adjustment1 = norm(selected_points_[2 * k] - fit0), adjustment2 = norm(selected_points_[2 * k + 1] - fit1);
          selected_points_[2 * k] = fit0;
          selected_points_[2 * k + 1] = fit1;
          em_inliers_[k] = fi[slot].n_inliers;
          last_theta_[k] = fi[slot].theta;
          ++em_rounds_[k];
          active[k] = (adjustment1 > thresh || adjustment2 > thresh) && em_rounds_[k] < max_em_rounds_;
        }
      }
    }
    return;
  }
  std::vector<float> xy;
  std::vector<uint32_t> in_pose, in_idx;
  std::vector<double> data;
  for (size_t k = 0; k < selected_points_.size() / 2 && k < 2; ++k) {
    const double thresh = 0.05;
    double adjustment1 = 2 * thresh, adjustment2 = 2 * thresh;
    em_rounds_[k] = 0;
    while ((adjustment1 > thresh || adjustment2 > thresh) && em_rounds_[k] < max_em_rounds_) {
      // E-step on the device: every world point within 3 cm of the stroke, in (pose, index) order.
      const float seg[4] = {selected_points_[2 * k].x, selected_points_[2 * k].y, selected_points_[2 * k + 1].x, selected_points_[2 * k + 1].y};
      uint64_t n = 0;
      if (in_pose.empty()) { in_pose.resize(1 << 16); in_idx.resize(1 << 16); xy.resize(2 << 16); }
      int rc = hitl_em_inliers(ctx_, seg, 0.03, in_pose.size(), in_pose.data(), in_idx.data(), xy.data(), &n);
      if (rc == HITL_ERR_OVERFLOW) {   // one pass is enough unless the stroke covers more points than ever before
        in_pose.resize(n); in_idx.resize(n); xy.resize(2 * n);
        rc = hitl_em_inliers(ctx_, seg, 0.03, n, in_pose.data(), in_idx.data(), xy.data(), &n);
      }
      check(rc, "hitl_em_inliers");
      em_inliers_[k] = n;
      // M-step on the host: refit the stroke's direction to the inliers.
      double cmx = 0, cmy = 0;
      data.resize(2 * n);
      for (uint64_t j = 0; j < n; ++j) { data[2 * j] = xy[2 * j]; data[2 * j + 1] = xy[2 * j + 1]; cmx += xy[2 * j]; cmy += xy[2 * j + 1]; }
      double CM[2] = {n ? cmx / (double)n : 0.0, n ? cmy / (double)n : 0.0};
      double P1[2] = {selected_points_[2 * k].x, selected_points_[2 * k].y};
      double P2[2] = {selected_points_[2 * k + 1].x, selected_points_[2 * k + 1].y};
      const std::vector<Vector2f> fit = SegFitEM(P1, P2, CM, data.data(), (int)n);
      adjustment1 = norm(selected_points_[2 * k] - fit[0]);
      adjustment2 = norm(selected_points_[2 * k + 1] - fit[1]);
      selected_points_[2 * k] = fit[0];
      selected_points_[2 * k + 1] = fit[1];
      ++em_rounds_[k];
    }
  }
}

std::pair<PoseObservations, PoseObservations> EMInput::EstablishObservationSets() { return ObservationSets(true); }

// with_indices = false returns the observing POSES only (empty index vectors): all that OrderAndFilterUserInput and
// SetCorrectionRelations read (EMinput.cpp:253-267, 325-455 use `.first` of every entry), without copying the index lists back.
std::pair<PoseObservations, PoseObservations> EMInput::ObservationSets(bool with_indices) {
  if (!world_clouds_resident_) UploadWorldClouds();
  const size_t n = local_version_point_clouds_.size();
  size_t total = 0;
  for (const auto& c : local_version_point_clouds_) total += c.size();
  const float segs[8] = {selected_points_[0].x, selected_points_[0].y, selected_points_[1].x, selected_points_[1].y,
                         selected_points_[2].x, selected_points_[2].y, selected_points_[3].x, selected_points_[3].y};
  uint32_t n_sets[2] = {0, 0};
  std::vector<uint32_t> set_pose[2], obs[2];
  std::vector<uint64_t> set_off[2];
  for (int f = 0; f < 2; ++f) { set_pose[f].resize(std::max<size_t>(n, 1)); if (with_indices) { set_off[f].resize(n + 1); obs[f].resize(std::max<size_t>(total, 1)); } }
  check(hitl_em_assign(ctx_, segs, 0.03, 5, n_sets, set_pose[0].data(), with_indices ? set_off[0].data() : nullptr, with_indices ? obs[0].data() : nullptr,
                       set_pose[1].data(), with_indices ? set_off[1].data() : nullptr, with_indices ? obs[1].data() : nullptr),
        "hitl_em_assign");
  std::pair<PoseObservations, PoseObservations> out;
  PoseObservations* dst[2] = {&out.first, &out.second};
  for (int f = 0; f < 2; ++f)
    for (uint32_t s = 0; s < n_sets[f]; ++s)
      dst[f]->push_back(std::make_pair((int)set_pose[f][s], with_indices ? std::vector<int>(obs[f].begin() + set_off[f][s], obs[f].begin() + set_off[f][s + 1]) : std::vector<int>()));
  return out;
}

void EMInput::SetCorrectionRelations(const PoseObservations& first_poses_obs, const PoseObservations& second_poses_obs) {
  corrected_poses_.clear(); anchor_poses_.clear();
  for (const auto& p : first_poses_obs) corrected_poses_.push_back(p.first);
  for (const auto& p : second_poses_obs) anchor_poses_.push_back(p.first);
}

// Integer bookkeeping of EMinput.cpp:325-455: drop poses that saw both strokes (four overlap
// cases), decide which stroke is older, possibly swap the strokes, emit the pose lists and the
// back-propagation bounds.  (-1, -1) signals an unusable selection, as in the reference.
void EMInput::OrderAndFilterUserInput() {
  if (selected_points_.size() != 4) throw std::invalid_argument("OrderAndFilterUserInput: 4 selected points expected");
  const std::pair<PoseObservations, PoseObservations> sets = ObservationSets(fetch_observation_indices_);
  std::vector<int> first, second;
  for (const auto& p : sets.first) first.push_back(p.first);
  for (const auto& p : sets.second) second.push_back(p.first);
  std::vector<int> overlaps;
  for (int s : second)
    for (int f : first)
      if (s == f) overlaps.push_back(f);
  auto drop = [&](std::vector<int>* v) {
    for (int o : overlaps) v->erase(std::remove(v->begin(), v->end(), o), v->end());
  };
  int backprop_start = 0, backprop_end = 0;
  if (overlaps.size() == first.size() && overlaps.size() == second.size()) {
    backprop_bounds_ = std::make_pair(-1, -1);   // complete overlap; the reference then reads [0] of empty lists (UB) — stop here
    corrected_poses_.clear(); anchor_poses_.clear();
    return;
  } else if (overlaps.size() == first.size()) {
    drop(&second);
  } else if (overlaps.size() == second.size()) {
    drop(&first);
  } else if (!overlaps.empty()) {
    drop(&first); drop(&second);
  }
  if (first.empty() || second.empty()) {          // nothing observed by one stroke: undefined in the reference
    backprop_bounds_ = std::make_pair(-1, -1);
    corrected_poses_.clear(); anchor_poses_.clear();
    return;
  }
  const int first_min = first.front(), first_max = first.back(), second_min = second.front(), second_max = second.back();
  auto keep = [](const PoseObservations& all, const std::vector<int>& ids) {
    PoseObservations out;
    for (int id : ids)
      for (const auto& p : all)
        if (p.first == id) out.push_back(p);
    return out;
  };
  const PoseObservations new_first = keep(sets.first, first), new_second = keep(sets.second, second);
  if (first_min > second_max) {                   // first stroke on the later visit: as drawn
    SetCorrectionRelations(new_first, new_second);
    backprop_start = second_max + 1;
    backprop_end = first_min - 1;
  } else if (first_max < second_min) {            // drawn in the other order: swap the strokes
    std::vector<Vector2f> reordered(selected_points_.begin() + 2, selected_points_.end());
    reordered.insert(reordered.end(), selected_points_.begin(), selected_points_.begin() + 2);
    selected_points_ = reordered;
    SetCorrectionRelations(new_second, new_first);
    backprop_start = first_max + 1;
    backprop_end = second_min - 1;
  } else {                                        // interleaved visits
    backprop_start = -1;
    backprop_end = -1;
  }
  backprop_bounds_ = std::make_pair(backprop_start, backprop_end);
}

void EMInput::Run() {
  AutomaticEndpointAdjustment();
  if (correction_type_ != CorrectionType::kPointCorrection && correction_type_ != CorrectionType::kCornerCorrection) OrderAndFilterUserInput();
}

}  // namespace hitl
