// host_capi.cpp — C entry points over the C++ host mirror (JointOpt, EMInput) for the Python
// harness (tests/, bench.py).  One session = one JointOpt + one EMInput on one hitl_ctx, wired the
// way HitLSLAM::Run wires the reference stages (human_in_the_loop_slam/HitLSLAM.cpp:379-484):
// world clouds -> EMInput::Run -> constraint targets -> JointOpt::Run.
//
// Also here, because a replayed correction needs them (host-side, O(#poses), outside the GPU path):
//   * CalculateConstraintTargets — AppExpCorrect::calculateConstraintTargets (ApplyExplicitCorrection.cpp:447-487)
//   * the session-log reader / writer — LoadLogFile / LogActivity (HitLSLAM_main.cpp:676-764, :776-822)
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <time.h>
#include <exception>
#include <stdexcept>
#include <string>
#include <vector>
#include "correction_stages.h"
#include "em_input.h"
#include "hitl_host.h"
#include "joint_optimization.h"
#include "hitl_math.h"

namespace hitl {

// One HumanConstraint per (anchor, corrected) pair, deltas measured in the anchor's frame from the
// CURRENT poses; relative_penalty_dir from the direction of the second stroke (selected_points[2..3]).
std::vector<HumanConstraint> CalculateConstraintTargets(const std::vector<Pose2Df>& poses, const std::vector<Vector2f>& selected_points, CorrectionType type,
                                                        const std::vector<int>& anchor_poses, const std::vector<int>& corrected_poses) {
  std::vector<HumanConstraint> out;
  const Vector2f dir = selected_points[3] - selected_points[2];
  const float correction_angle = atan2f(dir.y, dir.x);
  for (int a : anchor_poses) {
    const float anchor_angle = poses[a].angle;
    const float rel_pen_dir = (float)(atan2f(sinf(correction_angle - anchor_angle), cosf(correction_angle - anchor_angle)) + M_PI / 2.0);
    const Vector2f anchor_loc = poses[a].translation;
    const Vector2f p(cosf(anchor_angle), sinf(anchor_angle)), n(-p.y, p.x);
    for (int c : corrected_poses) {
      const Vector2f rel = poses[c].translation - anchor_loc;
      const float pose_angle = poses[c].angle;
      HumanConstraint h;
      h.constraint_type = type; h.anchor_pose_id = a; h.constrained_pose_id = c;
      h.delta_parallel = dot(p, rel); h.delta_perpendicular = dot(n, rel);
      h.delta_angle = atan2f(sinf(pose_angle - anchor_angle), cosf(pose_angle - anchor_angle));
      h.relative_penalty_dir = rel_pen_dir;
      out.push_back(h);
    }
  }
  return out;
}

struct Session {
  explicit Session(hitl_ctx* c) : ctx(c), jopt(c), em(c) {}
  hitl_ctx* ctx;
  JointOpt jopt;
  EMInput em;
  std::string error;
};

}  // namespace hitl

using namespace hitl;

namespace {
// Powell's singular function, the standard smoke test of a non-linear least-squares stack:
// four residual blocks over four scalar parameter blocks, minimum 0 at the origin.
struct PowellF1 { template <typename T> bool operator()(const T* x1, const T* x2, T* r) const { r[0] = x1[0] + 10.0 * x2[0]; return true; } };
struct PowellF2 { template <typename T> bool operator()(const T* x3, const T* x4, T* r) const { r[0] = ::sqrt(5.0) * (x3[0] - x4[0]); return true; } };
struct PowellF3 { template <typename T> bool operator()(const T* x2, const T* x3, T* r) const { const T d = x2[0] - 2.0 * x3[0]; r[0] = d * d; return true; } };
struct ChainRel { double dx, dy; template <typename T> bool operator()(const T* a, const T* b, T* r) const { r[0] = b[0] - a[0] - dx + 0.1 * sin(a[1]); r[1] = b[1] - a[1] - dy; return true; } };
struct ChainPin { double tx, ty; template <typename T> bool operator()(const T* a, T* r) const { r[0] = 3.0 * (a[0] - tx); r[1] = 3.0 * (a[1] - ty) + 0.2 * (a[0] - tx); return true; } };
struct PowellF4 { template <typename T> bool operator()(const T* x1, const T* x4, T* r) const { const T d = x1[0] - x4[0]; r[0] = ::sqrt(10.0) * d * d; return true; } };
}  // namespace

#define HOST_TRY(s, ...)                                                    \
  try { __VA_ARGS__; return 0; }                                            \
  catch (const std::exception& e) { (s)->error = e.what(); return -1; }     \
  catch (...) { (s)->error = "unknown exception"; return -1; }

extern "C" {
static void fill_summary(const JointOpt& J, double out[6]);

void* hitl_host_session_create(void* ctx) {
  if (!ctx) return nullptr;
  try { return new Session(static_cast<hitl_ctx*>(ctx)); } catch (...) { return nullptr; }
}
void hitl_host_session_destroy(void* s) { delete static_cast<Session*>(s); }
const char* hitl_host_session_error(void* s) { return s ? static_cast<Session*>(s)->error.c_str() : "null session"; }

int hitl_host_session_set_map(void* sp, uint32_t n_poses, const float* poses_xyt, const uint32_t* off, const float* pts_xy, const float* nrm_xy) {
  Session* s = static_cast<Session*>(sp);
  HOST_TRY(s, {
    JointOpt& J = s->jopt;
    J.ClearPoses();
    J.poses_.resize(n_poses); J.robot_frame_point_clouds_.resize(n_poses); J.robot_frame_normal_clouds_.resize(n_poses);
    for (uint32_t i = 0; i < n_poses; ++i) {
      J.poses_[i] = Pose2Df(poses_xyt[3 * i + 2], poses_xyt[3 * i], poses_xyt[3 * i + 1]);
      const uint32_t n = off[i + 1] - off[i];
      J.robot_frame_point_clouds_[i].resize(n); J.robot_frame_normal_clouds_[i].resize(n);
      for (uint32_t k = 0; k < n; ++k) {
        J.robot_frame_point_clouds_[i][k] = Vector2f(pts_xy[2 * (size_t)(off[i] + k)], pts_xy[2 * (size_t)(off[i] + k) + 1]);
        J.robot_frame_normal_clouds_[i][k] = Vector2f(nrm_xy[2 * (size_t)(off[i] + k)], nrm_xy[2 * (size_t)(off[i] + k) + 1]);
      }
    }
    J.human_constraints_.clear();
    J.BuildKDTrees();
    J.SetParams();
    s->em.world_clouds_resident_ = false;
  })
}

int hitl_host_session_set_poses(void* sp, const float* poses_xyt) {
  Session* s = static_cast<Session*>(sp);
  HOST_TRY(s, {
    for (size_t i = 0; i < s->jopt.poses_.size(); ++i) s->jopt.poses_[i] = Pose2Df(poses_xyt[3 * i + 2], poses_xyt[3 * i], poses_xyt[3 * i + 1]);
    s->jopt.SetParams();
    s->em.world_clouds_resident_ = false;
  })
}
int hitl_host_session_get_poses(void* sp, float* poses_xyt, double* pose_array) {
  Session* s = static_cast<Session*>(sp);
  for (size_t i = 0; i < s->jopt.poses_.size(); ++i) {
    if (poses_xyt) { poses_xyt[3 * i] = s->jopt.poses_[i].translation.x; poses_xyt[3 * i + 1] = s->jopt.poses_[i].translation.y; poses_xyt[3 * i + 2] = s->jopt.poses_[i].angle; }
  }
  if (pose_array && !s->jopt.pose_array_.empty()) memcpy(pose_array, s->jopt.pose_array_.data(), sizeof(double) * s->jopt.pose_array_.size());
  return 0;
}

// World-frame clouds from the session's poses, resident on the device for the EM calls
// (HitLSLAM::transformPointCloudsToWorldFrame, HitLSLAM.cpp:245-254 + the copy at :400).
int hitl_host_session_world_transform(void* sp, int keep_host_copy) {
  Session* s = static_cast<Session*>(sp);
  HOST_TRY(s, {
    if (keep_host_copy) {
      s->jopt.copy_world_frame_clouds_to_host_ = true;
      s->jopt.CopyTempLaserScans();
      s->jopt.copy_world_frame_clouds_to_host_ = false;
      s->em.local_version_point_clouds_ = s->jopt.world_frame_point_clouds_;
    } else {
      std::vector<float> p(3 * s->jopt.poses_.size());
      for (size_t i = 0; i < s->jopt.poses_.size(); ++i) { p[3 * i] = s->jopt.poses_[i].translation.x; p[3 * i + 1] = s->jopt.poses_[i].translation.y; p[3 * i + 2] = s->jopt.poses_[i].angle; }
      if (hitl_world_transform(s->ctx, p.data(), nullptr) != HITL_OK) throw std::runtime_error(hitl_last_error(s->ctx));
      // EstablishObservationSets sizes its outputs from the cloud shapes: keep the shapes, not the data
      s->em.local_version_point_clouds_.resize(s->jopt.robot_frame_point_clouds_.size());
      for (size_t i = 0; i < s->em.local_version_point_clouds_.size(); ++i) s->em.local_version_point_clouds_[i].resize(s->jopt.robot_frame_point_clouds_[i].size());
    }
    s->em.world_clouds_resident_ = true;
  })
}

// EMInput::Run on the 4 selected points (in/out).  info = {n_corrected, n_anchor, backprop_start,
// backprop_end, rounds_stroke0, rounds_stroke1}.
int hitl_host_session_em_run(void* sp, int correction_type, float sel_xy[8], int32_t info[6]) {
  Session* s = static_cast<Session*>(sp);
  HOST_TRY(s, {
    EMInput& E = s->em;
    E.correction_type_ = (CorrectionType)correction_type;
    E.selected_points_.resize(4);
    for (int i = 0; i < 4; ++i) E.selected_points_[i] = Vector2f(sel_xy[2 * i], sel_xy[2 * i + 1]);
    E.Run();
    for (int i = 0; i < 4; ++i) { sel_xy[2 * i] = E.selected_points_[i].x; sel_xy[2 * i + 1] = E.selected_points_[i].y; }
    info[0] = (int32_t)E.corrected_poses_.size(); info[1] = (int32_t)E.anchor_poses_.size();
    info[2] = E.backprop_bounds_.first; info[3] = E.backprop_bounds_.second; info[4] = E.em_rounds_[0]; info[5] = E.em_rounds_[1];
  })
}
int hitl_host_session_em_poses(void* sp, int32_t* corrected, int32_t* anchor) {
  Session* s = static_cast<Session*>(sp);
  for (size_t i = 0; i < s->em.corrected_poses_.size(); ++i) corrected[i] = s->em.corrected_poses_[i];
  for (size_t i = 0; i < s->em.anchor_poses_.size(); ++i) anchor[i] = s->em.anchor_poses_[i];
  return 0;
}

// Appends one correction's constraints (anchor x corrected from the last EM run, targets from the
// current poses) to JointOpt::human_constraints_.  Returns the number of constraints added via n_out.
int hitl_host_session_add_constraints_from_em(void* sp, uint32_t* n_out) {
  Session* s = static_cast<Session*>(sp);
  HOST_TRY(s, {
    std::vector<HumanConstraint> hc = CalculateConstraintTargets(s->jopt.poses_, s->em.selected_points_, s->em.correction_type_, s->em.anchor_poses_, s->em.corrected_poses_);
    if (n_out) *n_out = (uint32_t)hc.size();
    s->jopt.human_constraints_.push_back(hc);
  })
}
// Explicit list: ids = {type, constrained, anchor} per constraint, deltas = {parallel, perpendicular, angle, penalty_dir}.
int hitl_host_session_add_constraints(void* sp, uint32_t n, const int32_t* ids3, const float* deltas4) {
  Session* s = static_cast<Session*>(sp);
  HOST_TRY(s, {
    std::vector<HumanConstraint> hc(n);
    for (uint32_t i = 0; i < n; ++i) {
      hc[i].constraint_type = (CorrectionType)ids3[3 * i]; hc[i].constrained_pose_id = ids3[3 * i + 1]; hc[i].anchor_pose_id = ids3[3 * i + 2];
      hc[i].delta_parallel = deltas4[4 * i]; hc[i].delta_perpendicular = deltas4[4 * i + 1]; hc[i].delta_angle = deltas4[4 * i + 2]; hc[i].relative_penalty_dir = deltas4[4 * i + 3];
    }
    s->jopt.human_constraints_.push_back(hc);
  })
}
// ---- the two stages between EM and JointOpt (f3) -------------------------------------------------------------------
// AppExpCorrect::Run on explicit inputs.  poses_xyt in/out; C3 = correction handed to Backprop; returns 1 if a group of
// corrected poses was applied, 0 if none, -1 on error.
int hitl_host_app_exp_correct(int correction_type, const float sel_xy[8], uint32_t n_poses, float* poses_xyt, uint32_t n_corrected, const int32_t* corrected,
                              float C3[3]) {
  try {
    AppExpCorrect A;
    A.correction_type_ = (CorrectionType)correction_type;
    for (int i = 0; i < 4; ++i) A.selected_points_.push_back(Vector2f(sel_xy[2 * i], sel_xy[2 * i + 1]));
    A.corrected_poses_.assign(corrected, corrected + n_corrected);
    A.poses_.resize(n_poses);
    for (uint32_t i = 0; i < n_poses; ++i) A.poses_[i] = Pose2Df(poses_xyt[3 * i + 2], poses_xyt[3 * i], poses_xyt[3 * i + 1]);
    A.Run();
    for (uint32_t i = 0; i < n_poses; ++i) { poses_xyt[3 * i] = A.poses_[i].translation.x; poses_xyt[3 * i + 1] = A.poses_[i].translation.y; poses_xyt[3 * i + 2] = A.poses_[i].angle; }
    for (int i = 0; i < 3; ++i) C3[i] = A.correction_[i];
    return A.applied_ ? 1 : 0;
  } catch (...) { return -1; }
}
// AppExpCorrect::calculateConstraintTargets (ApplyExplicitCorrection.cpp:447-487) on explicit inputs: one HumanConstraint per
// (anchor, corrected) pair, anchor-major.  ids3 = (type, constrained, anchor), deltas4 = (parallel, perpendicular, angle, penalty dir).
// Buffers hold n_anchor * n_corrected entries; returns that count, or -1 on error.
int hitl_host_constraint_targets(int correction_type, const float sel_xy[8], uint32_t n_poses, const float* poses_xyt, uint32_t n_corrected, const int32_t* corrected,
                                 uint32_t n_anchor, const int32_t* anchor, int32_t* ids3, float* deltas4) {
  try {
    std::vector<Pose2Df> poses(n_poses);
    for (uint32_t i = 0; i < n_poses; ++i) poses[i] = Pose2Df(poses_xyt[3 * i + 2], poses_xyt[3 * i], poses_xyt[3 * i + 1]);
    std::vector<Vector2f> sel;
    for (int i = 0; i < 4; ++i) sel.push_back(Vector2f(sel_xy[2 * i], sel_xy[2 * i + 1]));
    for (uint32_t i = 0; i < n_corrected; ++i) if (corrected[i] < 0 || (uint32_t)corrected[i] >= n_poses) return -1;
    for (uint32_t i = 0; i < n_anchor; ++i) if (anchor[i] < 0 || (uint32_t)anchor[i] >= n_poses) return -1;
    const std::vector<HumanConstraint> hc = CalculateConstraintTargets(poses, sel, (CorrectionType)correction_type, std::vector<int>(anchor, anchor + n_anchor),
                                                                       std::vector<int>(corrected, corrected + n_corrected));
    for (size_t b = 0; b < hc.size(); ++b) {
      ids3[3 * b] = (int32_t)hc[b].constraint_type; ids3[3 * b + 1] = hc[b].constrained_pose_id; ids3[3 * b + 2] = hc[b].anchor_pose_id;
      deltas4[4 * b] = hc[b].delta_parallel; deltas4[4 * b + 1] = hc[b].delta_perpendicular; deltas4[4 * b + 2] = hc[b].delta_angle; deltas4[4 * b + 3] = hc[b].relative_penalty_dir;
    }
    return (int)hc.size();
  } catch (...) { return -1; }
}
// Backprop::Run on explicit inputs (pose update on the GPU of `ctx`).  poses_xyt, cov9 in/out.
int hitl_host_backprop(void* ctx, uint32_t n_poses, float* poses_xyt, float* cov9, int32_t lo, int32_t hi, const float C3[3], float* device_ms) {
  try {
    Backprop B(static_cast<hitl_ctx*>(ctx));
    B.backprop_bounds_ = std::make_pair((int)lo, (int)hi);
    B.correction_ = Vector3f{{C3[0], C3[1], C3[2]}};
    B.poses_.resize(n_poses); B.covariances_.resize(n_poses);
    for (uint32_t i = 0; i < n_poses; ++i) {
      B.poses_[i] = Pose2Df(poses_xyt[3 * i + 2], poses_xyt[3 * i], poses_xyt[3 * i + 1]);
      for (int k = 0; k < 9; ++k) B.covariances_[i][k] = cov9[9 * (size_t)i + k];
    }
    B.Run();
    for (uint32_t i = 0; i < n_poses; ++i) {
      poses_xyt[3 * i] = B.poses_[i].translation.x; poses_xyt[3 * i + 1] = B.poses_[i].translation.y; poses_xyt[3 * i + 2] = B.poses_[i].angle;
      for (int k = 0; k < 9; ++k) cov9[9 * (size_t)i + k] = B.covariances_[i][k];
    }
    if (device_ms) *device_ms = B.last_device_ms_;
    return 0;
  } catch (...) { return -1; }
}

// One full human correction on the session's map, wired as HitLSLAM::Run (HitLSLAM.cpp:379-484): EMInput::Run (strokes refit,
// corrected / anchor poses, bounds) -> AppExpCorrect::Run -> Backprop::Run -> angle wrap -> constraints appended ->
// JointOpt::Run (unless solve = 0).  cov9: per-pose covariances, in/out (may be NULL: 1e-4 / 1e-5 diagonals).
// info = {n_corrected, n_anchor, bound_lo, bound_hi, em_rounds_a, em_rounds_b, n_new_constraints, applied};
// ms = {em, explicit correction, backprop (host + device), backprop device only, joint optimisation}.
int hitl_host_session_correct(void* sp, int correction_type, float sel_xy[8], float* cov9, int solve, int32_t info[8], double ms[5], double summary[6]) {
  Session* s = static_cast<Session*>(sp);
  HOST_TRY(s, {
    auto now = []() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; };
    for (int i = 0; i < 8; ++i) info[i] = 0;
    for (int i = 0; i < 5; ++i) ms[i] = 0;
    const double t0 = now();
    EMInput& E = s->em;
    E.correction_type_ = (CorrectionType)correction_type;
    E.selected_points_.resize(4);
    for (int i = 0; i < 4; ++i) E.selected_points_[i] = Vector2f(sel_xy[2 * i], sel_xy[2 * i + 1]);
    E.Run();
    for (int i = 0; i < 4; ++i) { sel_xy[2 * i] = E.selected_points_[i].x; sel_xy[2 * i + 1] = E.selected_points_[i].y; }
    info[0] = (int32_t)E.corrected_poses_.size(); info[1] = (int32_t)E.anchor_poses_.size();
    info[2] = E.backprop_bounds_.first; info[3] = E.backprop_bounds_.second; info[4] = E.em_rounds_[0]; info[5] = E.em_rounds_[1];
    const double t1 = now();
    ms[0] = t1 - t0;
    if (!(E.backprop_bounds_.first >= 0 && E.backprop_bounds_.second >= 1)) return 0;      // HitLSLAM.cpp:413
    AppExpCorrect A;
    A.correction_type_ = E.correction_type_; A.selected_points_ = E.selected_points_;
    A.corrected_poses_ = E.corrected_poses_; A.anchor_poses_ = E.anchor_poses_; A.poses_ = s->jopt.poses_;
    A.Run();
    info[6] = (int32_t)A.new_human_constraints_.size(); info[7] = A.applied_ ? 1 : 0;
    s->jopt.human_constraints_.push_back(A.new_human_constraints_);
    const double t2 = now();
    ms[1] = t2 - t1;
    Backprop B(s->ctx);
    B.poses_ = A.poses_; B.correction_ = A.correction_; B.backprop_bounds_ = E.backprop_bounds_;
    B.covariances_.resize(A.poses_.size());
    for (size_t i = 0; i < B.covariances_.size(); ++i)
      for (int k = 0; k < 9; ++k) B.covariances_[i][k] = cov9 ? cov9[9 * i + k] : ((k == 0 || k == 4) ? 1e-4f : (k == 8 ? 1e-5f : 0.f));
    if (A.applied_) B.Run();
    if (cov9) for (size_t i = 0; i < B.covariances_.size(); ++i) for (int k = 0; k < 9; ++k) cov9[9 * i + k] = B.covariances_[i][k];
    for (size_t i = 0; i < B.poses_.size(); ++i) { const float a = B.poses_[i].angle; B.poses_[i].angle = atan2f(sinf_rn(a), cosf_rn(a)); }   // :443-447
    s->jopt.poses_ = B.poses_;
    s->em.world_clouds_resident_ = false;
    const double t3 = now();
    ms[2] = t3 - t2; ms[3] = B.last_device_ms_;
    if (solve) {
      s->jopt.Run();
      if (summary) fill_summary(s->jopt, summary);
    }
    ms[4] = now() - t3;
  })
}

// HitLSLAM::verifyUserInput (HitLSLAM.cpp:218-243) over the world clouds resident on the session's GPU context: replayLog / Run go on
// only when every selected point was verified (:315, :383).
int hitl_host_session_verify_input(void* sp, const float sel_xy[8], uint32_t* points_verified) {
  Session* s = static_cast<Session*>(sp);
  HOST_TRY(s, {
    if (hitl_verify_input(s->ctx, 4, sel_xy, 0.05f, points_verified, nullptr) != HITL_OK) throw std::runtime_error(hitl_last_error(s->ctx));
  })
}

int hitl_host_session_clear_constraints(void* sp) { static_cast<Session*>(sp)->jopt.human_constraints_.clear(); return 0; }

// Solver knobs: which = 0 (SolveHumanConstraints) or 1 (PostHumanOptimization); negative values keep the default.
int hitl_host_session_solver_options(void* sp, int which, int max_iterations, double function_tolerance, double gradient_tolerance, double parameter_tolerance,
                                     int precision, int verbose) {
  Session* s = static_cast<Session*>(sp);
  ceres::Solver::Options& o = which ? s->jopt.post_solver_options_ : s->jopt.human_solver_options_;
  if (max_iterations >= 0) o.max_num_iterations = max_iterations;
  if (function_tolerance >= 0) o.function_tolerance = function_tolerance;
  if (gradient_tolerance >= 0) o.gradient_tolerance = gradient_tolerance;
  if (parameter_tolerance >= 0) o.parameter_tolerance = parameter_tolerance;
  if (precision >= 0) s->jopt.precision_ = precision;
  if (verbose >= 0) s->jopt.verbose_ = verbose != 0;
  return 0;
}

static void fill_summary(const JointOpt& J, double out[6]) {
  out[0] = J.last_summary_.initial_cost; out[1] = J.last_summary_.final_cost;
  out[2] = J.last_summary_.num_successful_steps; out[3] = J.last_summary_.num_unsuccessful_steps;
  out[4] = (double)J.last_summary_.termination_type; out[5] = J.num_hc_residuals_;
}

// JointOpt::Run: SolveHumanConstraints (+ PostHumanOptimization when post != 0).  summary = {initial
// cost, final cost, successful steps, unsuccessful steps, termination type, num_hc_residuals}.
int hitl_host_session_joint_opt_run(void* sp, int post, double summary[6]) {
  Session* s = static_cast<Session*>(sp);
  HOST_TRY(s, {
    s->jopt.enable_post_human_optimization_ = post != 0;
    s->jopt.Run();
    s->em.world_clouds_resident_ = false;   // poses moved
    if (summary) fill_summary(s->jopt, summary);
  })
}
// The two solves separately, on pose_array_ (no CopyParams): mode 0 = human, 1 = post-HitL (search + STF blocks).
int hitl_host_session_solve(void* sp, int mode, double summary[6]) {
  Session* s = static_cast<Session*>(sp);
  HOST_TRY(s, {
    const ceres::TerminationType t = mode ? s->jopt.PostHumanOptimization(0, (int)s->jopt.poses_.size() - 1) : s->jopt.SolveHumanConstraints();
    if (summary) fill_summary(s->jopt, summary);
    if (t == ceres::FAILURE || t == ceres::USER_FAILURE) throw std::runtime_error("solver failure: " + s->jopt.last_error_);
  })
}
int hitl_host_session_copy_params(void* sp) { static_cast<Session*>(sp)->jopt.CopyParams(); return 0; }

// Multi-GPU in one process: extra contexts (other devices) that share the search and the STF blocks with the session's context
// (JointOpt::UseShardContexts).  n_extra = 0 goes back to one context.  The map must be set again afterwards.
int hitl_host_session_use_shard_contexts(void* sp, uint32_t n_extra, void** extra_ctx) {
  Session* s = static_cast<Session*>(sp);
  HOST_TRY(s, {
    std::vector<hitl_ctx*> v;
    for (uint32_t q = 0; q < n_extra; ++q) v.push_back(static_cast<hitl_ctx*>(extra_ctx[q]));
    s->jopt.UseShardContexts(v);
  })
}
// Source-pose range and STF block count of every context in the last FindSTFCorrespondences: out = {lo, hi, blocks} per context.
int hitl_host_session_shard_info(void* sp, uint32_t cap, uint64_t* out3, uint32_t* n_contexts) {
  Session* s = static_cast<Session*>(sp);
  const JointOpt& J = s->jopt;
  if (n_contexts) *n_contexts = (uint32_t)J.NumContexts();
  for (size_t r = 0; r < J.shard_ranges_.size() && r < cap; ++r) { out3[3 * r] = J.shard_ranges_[r].first; out3[3 * r + 1] = J.shard_ranges_[r].second; out3[3 * r + 2] = r < J.shard_blocks_.size() ? J.shard_blocks_[r] : 0; }
  return 0;
}
// FindSTFCorrespondences through the mirror; counts = {n_pairs, n_matches, n_queries}.
int hitl_host_session_find_stf(void* sp, uint64_t min_pose, uint64_t max_pose, uint64_t counts[3]) {
  Session* s = static_cast<Session*>(sp);
  HOST_TRY(s, {
    s->jopt.FindSTFCorrespondences(min_pose, max_pose);
    const StfCorrespondenceSet& S = s->jopt.point_point_glob_correspondences_;
    counts[0] = S.pair_i.size(); counts[1] = S.k.size(); counts[2] = S.n_queries;
  })
}
int hitl_host_session_get_stf(void* sp, uint32_t* pair_i, uint32_t* pair_j, uint64_t* pair_off, uint32_t* k, uint32_t* idx) {
  const StfCorrespondenceSet& S = static_cast<Session*>(sp)->jopt.point_point_glob_correspondences_;
  if (!S.pair_i.empty()) { memcpy(pair_i, S.pair_i.data(), 4 * S.pair_i.size()); memcpy(pair_j, S.pair_j.data(), 4 * S.pair_j.size()); }
  if (!S.pair_off.empty()) memcpy(pair_off, S.pair_off.data(), 8 * S.pair_off.size());
  if (!S.k.empty()) { memcpy(k, S.k.data(), 4 * S.k.size()); memcpy(idx, S.idx.data(), 4 * S.idx.size()); }
  return 0;
}
// Problem::Evaluate of the last post-HitL problem: gradient (3 per non-constant pose) and CRS Jacobian sizes.
int hitl_host_session_gradient(void* sp, uint64_t cap, double* gradient, uint64_t* n_out, uint64_t jac_dims[3]) {
  Session* s = static_cast<Session*>(sp);
  const std::vector<double>& g = s->jopt.gradients_;
  if (n_out) *n_out = g.size();
  for (size_t i = 0; i < g.size() && i < cap; ++i) gradient[i] = g[i];
  if (jac_dims) { jac_dims[0] = s->jopt.ceres_jacobian_.num_rows; jac_dims[1] = s->jopt.ceres_jacobian_.num_cols; jac_dims[2] = s->jopt.ceres_jacobian_.values.size(); }
  return 0;
}

// ---- one cost function at a time, the way Ceres would call it (parity of the drop-in blocks) ----
// Builds the odometry + human + (optionally) STF problem at the session's pose_array_ and evaluates residual
// block `block` (AddResidualBlock order: odometry, human, stf) through CostFunction::Evaluate.
// Returns the number of residuals via nres; jac0/jac1 are row-major [nres x 3] (jac1 only for binary blocks).
int hitl_host_session_evaluate_block(void* sp, int with_stf, uint64_t block, const double* pose_array_override, int32_t* nres, int32_t* nblocks,
                                     double* residuals, double* jac0, double* jac1, uint64_t* n_total_blocks) {
  Session* s = static_cast<Session*>(sp);
  HOST_TRY(s, {
    JointOpt& J = s->jopt;
    if (pose_array_override) J.pose_array_.assign(pose_array_override, pose_array_override + 3 * J.poses_.size());
    ceres::Problem problem(J.BeginProblem());
    J.AddOdometryConstraints(&problem);
    J.AddHumanConstraints(&problem);
    if (with_stf) J.AddSTFConstraints(&problem);
    if (n_total_blocks) *n_total_blocks = (uint64_t)problem.NumResidualBlocks();
    if (block >= (uint64_t)problem.NumResidualBlocks()) throw std::out_of_range("residual block index");
    const ceres::Problem::ResidualBlock& rb = problem.residual_blocks()[block];
    const double* params[2] = {nullptr, nullptr};
    for (size_t q = 0; q < rb.blocks.size(); ++q) params[q] = problem.parameter_blocks()[rb.blocks[q]].values;
    double* jac[2] = {jac0, jac1};
    *nres = rb.cost->num_residuals(); *nblocks = (int32_t)rb.blocks.size();
    if (!rb.cost->Evaluate(params, residuals, jac)) throw std::runtime_error("CostFunction::Evaluate returned false");
  })
}

// ---- device-free pieces (CPU tests) ---------------------------------------------------------------------
int hitl_host_seg_fit_em_theta(const double p1[2], const double p2[2], const double* data, int size, float out4[4], double* theta, int* iterations) {
  try {
    const std::vector<Vector2f> fit = FitSegmentAngle(p1, p2, data, size, theta, iterations);
    out4[0] = fit[0].x; out4[1] = fit[0].y; out4[2] = fit[1].x; out4[3] = fit[1].y;
    return 0;
  } catch (...) { return -1; }
}
// Selects where EMInput's M-step runs in this session: 1 = device (hitl_em_refit, default), 0 = host LM (the checker).
int hitl_host_session_set_device_m_step(void* sp, int on) { static_cast<Session*>(sp)->em.device_m_step_ = on != 0; return 0; }
// Device M-step: EM rounds enqueued per host wait (1..4; hitl_em_refit_chain).  1 = one wait per round and stroke pair.
int hitl_host_session_set_em_chain_rounds(void* sp, int rounds) { static_cast<Session*>(sp)->em.device_chain_rounds_ = rounds < 1 ? 1 : (rounds > 4 ? 4 : rounds); return 0; }
int hitl_host_seg_fit_em(const double p1[2], const double p2[2], const double* data, int size, float out4[4]) {
  try {
    const std::vector<Vector2f> fit = FitSegmentAngle(p1, p2, data, size);
    out4[0] = fit[0].x; out4[1] = fit[0].y; out4[2] = fit[1].x; out4[3] = fit[1].y;
    return 0;
  } catch (...) { return -1; }
}
void hitl_host_odometry_consts(const float* poses_xyt, uint32_t n_poses, float* consts9) {
  for (uint32_t i = 1; i < n_poses; ++i)
    OdometryBlockConstants(Pose2Df(poses_xyt[3 * i - 1], poses_xyt[3 * i - 3], poses_xyt[3 * i - 2]), Pose2Df(poses_xyt[3 * i + 2], poses_xyt[3 * i], poses_xyt[3 * i + 1]),
                           consts9 + 9 * (size_t)(i - 1));
}
void hitl_host_human_targets(const float* poses_xyt, uint32_t n_poses, uint32_t n, const int32_t* ids3, const float* deltas4, double* targets4) {
  std::vector<Pose2Df> poses(n_poses);
  for (uint32_t i = 0; i < n_poses; ++i) poses[i] = Pose2Df(poses_xyt[3 * i + 2], poses_xyt[3 * i], poses_xyt[3 * i + 1]);
  for (uint32_t i = 0; i < n; ++i) {
    HumanConstraint c;
    c.constraint_type = (CorrectionType)ids3[3 * i]; c.constrained_pose_id = ids3[3 * i + 1]; c.anchor_pose_id = ids3[3 * i + 2];
    c.delta_parallel = deltas4[4 * i]; c.delta_perpendicular = deltas4[4 * i + 1]; c.delta_angle = deltas4[4 * i + 2]; c.relative_penalty_dir = deltas4[4 * i + 3];
    HumanBlockTargets(poses, c, targets4 + 4 * (size_t)i);
  }
}


// Runs ceres::Solve on Powell's function from x; out = {initial cost, final cost, iterations, termination, #constant blocks honoured}.
// hold_x1 != 0 keeps the first parameter constant (SetParameterBlockConstant).
int hitl_host_solver_selftest(double x[4], int max_iterations, int hold_x1, int force_cg, double out[4]) {
  ceres::Problem problem;
  problem.AddResidualBlock(new ceres::AutoDiffCostFunction<PowellF1, 1, 1, 1>(new PowellF1), NULL, &x[0], &x[1]);
  problem.AddResidualBlock(new ceres::AutoDiffCostFunction<PowellF2, 1, 1, 1>(new PowellF2), NULL, &x[2], &x[3]);
  problem.AddResidualBlock(new ceres::AutoDiffCostFunction<PowellF3, 1, 1, 1>(new PowellF3), NULL, &x[1], &x[2]);
  problem.AddResidualBlock(new ceres::AutoDiffCostFunction<PowellF4, 1, 1, 1>(new PowellF4), NULL, &x[0], &x[3]);
  if (hold_x1) problem.SetParameterBlockConstant(&x[0]);
  ceres::Solver::Options o;
  o.max_num_iterations = max_iterations;
  o.function_tolerance = 1e-20; o.gradient_tolerance = 1e-14; o.parameter_tolerance = 1e-14;
  if (force_cg) o.dense_limit = 0;
  ceres::Solver::Summary s;
  ceres::Solve(o, &problem, &s);
  out[0] = s.initial_cost; out[1] = s.final_cost; out[2] = s.num_successful_steps + s.num_unsuccessful_steps; out[3] = (double)s.termination_type;
  return 0;
}

// Problem::Evaluate on Powell's function with x1 optionally constant: gradient (as many entries as the problem has parameters — constant
// blocks keep zero entries, as in Ceres) and the CRS Jacobian's dimensions {rows, cols, non-zeros}.  Returns the gradient length.
int hitl_host_evaluate_selftest(const double x_in[4], int hold_x1, double grad_out[4], uint64_t dims[3], double* cost) {
  double x[4] = {x_in[0], x_in[1], x_in[2], x_in[3]};
  ceres::Problem problem;
  problem.AddResidualBlock(new ceres::AutoDiffCostFunction<PowellF1, 1, 1, 1>(new PowellF1), NULL, &x[0], &x[1]);
  problem.AddResidualBlock(new ceres::AutoDiffCostFunction<PowellF2, 1, 1, 1>(new PowellF2), NULL, &x[2], &x[3]);
  problem.AddResidualBlock(new ceres::AutoDiffCostFunction<PowellF3, 1, 1, 1>(new PowellF3), NULL, &x[1], &x[2]);
  problem.AddResidualBlock(new ceres::AutoDiffCostFunction<PowellF4, 1, 1, 1>(new PowellF4), NULL, &x[0], &x[3]);
  if (hold_x1) problem.SetParameterBlockConstant(&x[0]);
  std::vector<double> residuals, gradient;
  ceres::CRSMatrix jac;
  ceres::Problem::EvaluateOptions eo;
  if (!problem.Evaluate(eo, cost, &residuals, &gradient, &jac)) return -1;
  for (size_t i = 0; i < gradient.size() && i < 4; ++i) grad_out[i] = gradient[i];
  dims[0] = jac.num_rows; dims[1] = jac.num_cols; dims[2] = jac.values.size();
  return (int)gradient.size();
}

// Problem bookkeeping that the pose-chain hot path never exercises: parameter blocks added in DESCENDING address order (the pointer
// index leaves its append-only form), one cost function shared by two residual blocks (deleted once), a residual block over three
// parameter blocks (block list beyond its inline storage), SetParameterBlockConstant on a block found through each index form.
// out = {cost, #parameter blocks, #residuals, gradient entry of the constant block, live cost functions after the problem died}.
namespace {
int g_live_sum3 = 0;
struct Sum3 : ceres::CostFunction {        // r = a + 2 b + 3 c - 1 over three scalar blocks
  Sum3() { ++g_live_sum3; set_num_residuals(1); for (int q = 0; q < 3; ++q) mutable_parameter_block_sizes()->push_back(1); }
  ~Sum3() override { --g_live_sum3; }
  bool Evaluate(double const* const* p, double* r, double** J) const override {
    r[0] = p[0][0] + 2.0 * p[1][0] + 3.0 * p[2][0] - 1.0;
    if (J) for (int q = 0; q < 3; ++q) if (J[q]) J[q][0] = q + 1.0;
    return true;
  }
};
}  // namespace
int hitl_host_problem_selftest(double out[5]) {
  double x[6] = {0.5, -1.0, 2.0, 0.25, 4.0, -3.0};
  g_live_sum3 = 0;
  {
    ceres::Problem problem;
    Sum3* shared = new Sum3;
    problem.AddResidualBlock(shared, NULL, std::vector<double*>{&x[1], &x[2], &x[3]});      // ascending so far
    problem.SetParameterBlockConstant(&x[2]);                                                   // found by bisection
    problem.AddResidualBlock(shared, NULL, std::vector<double*>{&x[5], &x[4], &x[0]});      // x4 and x0 arrive out of order
    problem.AddResidualBlock(new Sum3, NULL, std::vector<double*>{&x[0], &x[2], &x[5]});
    problem.SetParameterBlockConstant(&x[0]);                                                   // found through the hash map
    problem.SetParameterBlockVariable(&x[2]);
    std::vector<double> residuals, gradient;
    double cost = 0;
    ceres::Problem::EvaluateOptions eo;
    if (!problem.Evaluate(eo, &cost, &residuals, &gradient, NULL)) return -1;
    out[0] = cost; out[1] = problem.NumParameterBlocks(); out[2] = problem.NumResiduals();
    // blocks in insertion order: x1 x2 x3 x5 x4 x0 -> the constant x0 is the last gradient entry
    out[3] = gradient.size() == 6 ? gradient[5] : -1.0;
    if (g_live_sum3 != 2) return -2;
  }
  out[4] = g_live_sum3;
  return 0;
}

// A pose-chain problem of the human-constraint shape — n blocks of 2 parameters, relative factors between neighbours,
// unary factors on every 7th block, non-linear through a sine — solved with the dense path (mode 0), the chain-direct
// path (mode 1: block-tridiagonal elimination) or PCG (mode 2).  x: 2 n values in/out.
int hitl_host_solver_chain_selftest(double* x, int n, int mode, double out[4]) {
  ceres::Problem problem;
  for (int i = 0; i + 1 < n; ++i) problem.AddResidualBlock(new ceres::AutoDiffCostFunction<ChainRel, 2, 2, 2>(new ChainRel{0.3 + 0.01 * (i % 5), -0.1 + 0.02 * (i % 3)}), NULL, &x[2 * i], &x[2 * i + 2]);
  for (int i = 3; i < n; i += 7) problem.AddResidualBlock(new ceres::AutoDiffCostFunction<ChainPin, 2, 2>(new ChainPin{0.31 * i, -0.07 * i}), NULL, &x[2 * i]);
  problem.SetParameterBlockConstant(&x[0]);
  ceres::Solver::Options o;
  o.max_num_iterations = 200;
  o.function_tolerance = 1e-18; o.gradient_tolerance = 1e-13; o.parameter_tolerance = 1e-14;
  o.dense_limit = mode == 0 ? (1 << 30) : 0;
  if (mode == 2) o.chain_direct = false;
  ceres::Solver::Summary s;
  ceres::Solve(o, &problem, &s);
  out[0] = s.initial_cost; out[1] = s.final_cost; out[2] = s.num_successful_steps + s.num_unsuccessful_steps; out[3] = (double)s.termination_type;
  return 0;
}

// ---- session log (HitLSLAM_main.cpp:676-764 reader, :806-820 writer) -------------------------------------
// File: "<count> \n", then per entry "<type>, <undone>\n" followed by one "x, y\n" line per selected point:
// 2 points for type 1, 8 for type 3, 4 otherwise.  Flat outputs: types[n], undone[n], npts[n], pts (x,y pairs).
int hitl_host_load_log(const char* path, uint32_t cap_entries, uint32_t cap_points, int32_t* types, int32_t* undone, int32_t* npts, float* pts_xy,
                       uint32_t* n_entries, uint32_t* n_points) {
  FILE* f = fopen(path, "r");
  if (!f) return -1;
  int count = 0;
  if (fscanf(f, "%d", &count) != 1 || count < 0) { fclose(f); return -2; }
  uint32_t ne = 0, np = 0;
  for (int e = 0; e < count; ++e) {
    int type = 0, und = 0;
    if (fscanf(f, "%d, %d", &type, &und) != 2) { fclose(f); return -2; }
    const int m = type == 1 ? 2 : type == 3 ? 8 : (type == 2 || type == 4 || type == 5 || type == 6) ? 4 : 0;   // unknown types carry no points
    if (ne < cap_entries) { types[ne] = type; undone[ne] = und; npts[ne] = m; }
    for (int q = 0; q < m; ++q) {
      float x, y;
      if (fscanf(f, "%f, %f", &x, &y) != 2) { fclose(f); return -2; }
      if (np < cap_points) { pts_xy[2 * np] = x; pts_xy[2 * np + 1] = y; }
      ++np;
    }
    ++ne;
  }
  fclose(f);
  if (n_entries) *n_entries = ne;
  if (n_points) *n_points = np;
  return (ne > cap_entries || np > cap_points) ? -3 : 0;
}
int hitl_host_save_log(const char* path, uint32_t n_entries, const int32_t* types, const int32_t* undone, const int32_t* npts, const float* pts_xy) {
  FILE* f = fopen(path, "w");
  if (!f) return -1;
  fprintf(f, "%d \n", (int)n_entries);
  size_t p = 0;
  for (uint32_t e = 0; e < n_entries; ++e) {
    fprintf(f, "%d, %d\n", types[e], undone[e]);
    for (int q = 0; q < npts[e]; ++q, ++p) fprintf(f, "%.4f, %.4f\n", pts_xy[2 * p], pts_xy[2 * p + 1]);
  }
  fclose(f);
  return 0;
}

}  // extern "C"
