// joint_optimization.cpp — see joint_optimization.h.  Host-side problem building in the reference's
// order and arithmetic (float constants frozen from the current poses), every O(points) step on the GPU.
#include "joint_optimization.h"

#include <math.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <stdexcept>
#include <thread>

namespace hitl {

namespace {
// shared/math/util.h:433-439 instantiated for float: the arithmetic runs in double (M_2PI is a
// double constant) and the result narrows on assignment.
inline float angle_mod_f(float angle) {
  angle -= (2.0 * M_PI) * rint(angle / (2.0 * M_PI));
  return angle;
}
inline double angle_mod_d(double angle) {
  angle -= (2.0 * M_PI) * rint(angle / (2.0 * M_PI));
  return angle;
}
struct Mat2 { float m00, m01, m10, m11; };
inline Mat2 rotation2d(float a) { const float s = sinf(a), c = cosf(a); return Mat2{c, -s, s, c}; }
inline Vector2f mul(const Mat2& m, Vector2f v) { return Vector2f(m.m00 * v.x + m.m01 * v.y, m.m10 * v.x + m.m11 * v.y); }
}  // namespace

void OdometryBlockConstants(const Pose2Df& prev, const Pose2Df& cur, float out9[9]) {
  const float kEpsilon = 1e-6f;
  const Vector2f translation = cur.translation - prev.translation;
  Vector2f radial;
  float radial_translation = 0.0f;
  if (fabsf(translation.x) < kEpsilon && fabsf(translation.y) < kEpsilon) {
    radial = Vector2f(cosf(cur.angle), sinf(cur.angle));           // standing still: axes from the heading
  } else {
    const Vector2f local = mul(rotation2d(-prev.angle), translation);
    const float len2 = local.x * local.x + local.y * local.y;
    radial = local;
    if (len2 > 0.0f) { const float len = sqrtf(len2); radial = Vector2f(local.x / len, local.y / len); }
    radial_translation = norm(translation);
  }
  const Vector2f tangential = mul(rotation2d((float)M_PI_2), radial);
  out9[0] = radial.x; out9[1] = radial.y; out9[2] = tangential.x; out9[3] = tangential.y;
  out9[4] = 0.03f; out9[5] = 0.03f; out9[6] = 0.01f;                 // hard-coded std-devs (:770, :776, :783)
  out9[7] = radial_translation;
  out9[8] = angle_mod_f(cur.angle - prev.angle);
}

void HumanBlockTargets(const std::vector<Pose2Df>& poses, const HumanConstraint& c, double out4[4]) {
  const Pose2Df& anchor = poses[c.anchor_pose_id];
  const float t_angle = anchor.angle + c.delta_angle;
  const float target_angle = atan2f(sinf(t_angle), cosf(t_angle));
  out4[0] = out4[1] = out4[3] = 0.0;
  out4[2] = (double)target_angle;
  if (c.constraint_type == CorrectionType::kLineSegmentCorrection || c.constraint_type == CorrectionType::kColinearCorrection) {
    const Vector2f para(cosf(anchor.angle), sinf(anchor.angle));
    const Vector2f perp(-para.y, para.x);
    const Vector2f target = (anchor.translation + c.delta_parallel * para) + c.delta_perpendicular * perp;
    out4[0] = (double)target.x; out4[1] = (double)target.y;
    if (c.constraint_type == CorrectionType::kColinearCorrection) out4[3] = (double)(anchor.angle + c.relative_penalty_dir);
  }
}

JointOpt::JointOpt(hitl_ctx* ctx) : ctx_(ctx) {
  if (!ctx) throw std::invalid_argument("JointOpt needs a hitl_ctx: the hot path has no CPU implementation");
  memset(&last_search_info_, 0, sizeof(last_search_info_));
  human_solver_options_.max_num_iterations = 100;                      // :1062
  post_solver_options_.linear_solver_type = ceres::SPARSE_SCHUR;         // :148-161
  post_solver_options_.trust_region_strategy_type = ceres::LEVENBERG_MARQUARDT;
  post_solver_options_.minimizer_type = ceres::TRUST_REGION;
  post_solver_options_.function_tolerance = 0.000001;
  post_solver_options_.update_state_every_iteration = true;
  post_solver_options_.max_num_iterations = 100;                        // :1168
}
JointOpt::~JointOpt() {}

void JointOpt::check(int rc, const char* where) {
  if (rc == HITL_OK) return;
  last_error_ = std::string(where) + ": " + hitl_last_error(ctx_);
  throw std::runtime_error(last_error_);
}

void JointOpt::UseShardContexts(const std::vector<hitl_ctx*>& extra) {
  for (hitl_ctx* c : extra) if (!c || c == ctx_) throw std::invalid_argument("UseShardContexts: null or duplicate context");
  shard_ctx_ = extra;
  kdtrees_built_ = false;
  if (evaluator_) evaluator_->SetShardContexts(shard_ctx_);
}

// Runs fn(rank, ctx) for every context, the extra ones on their own host threads; rethrows the first failure.
template <typename F>
static void for_each_context(hitl_ctx* first, const std::vector<hitl_ctx*>& extra, F fn) {
  std::vector<std::string> errors(extra.size());
  std::vector<std::thread> workers;
  for (size_t q = 0; q < extra.size(); ++q)
    workers.emplace_back([&, q]() { try { fn(q + 1, extra[q]); } catch (const std::exception& e) { errors[q] = e.what(); if (errors[q].empty()) errors[q] = "error"; } });
  std::string mine;
  try { fn(0, first); } catch (const std::exception& e) { mine = e.what(); if (mine.empty()) mine = "error"; }
  for (std::thread& w : workers) w.join();
  if (!mine.empty()) throw std::runtime_error(mine);
  for (const std::string& e : errors) if (!e.empty()) throw std::runtime_error("shard context: " + e);
}
static void check_ctx(hitl_ctx* c, int rc, const char* where) {
  if (rc != HITL_OK) throw std::runtime_error(std::string(where) + ": " + hitl_last_error(c));
}

void JointOpt::ClearPoses() {
  poses_.clear(); robot_frame_point_clouds_.clear(); robot_frame_normal_clouds_.clear(); covariances_.clear();
  pose_array_.clear(); kdtrees_built_ = false;
}

void JointOpt::SetParams() {
  pose_array_.assign(poses_.size() * 3, 0.0);
  for (size_t i = 0; i < poses_.size(); ++i) {
    pose_array_[3 * i + 0] = poses_[i].translation.x;
    pose_array_[3 * i + 1] = poses_[i].translation.y;
    pose_array_[3 * i + 2] = poses_[i].angle;
  }
}

void JointOpt::CopyParams() {
  for (size_t i = 0; i < poses_.size(); ++i) {
    poses_[i].translation.x = (float)pose_array_[3 * i + 0];
    poses_[i].translation.y = (float)pose_array_[3 * i + 1];
    poses_[i].angle = (float)angle_mod_d(pose_array_[3 * i + 2]);
  }
}

// World-frame copy of the scans "as they are displayed in the gui" (:404-419), on the GPU.
void JointOpt::CopyTempLaserScans() {
  if (!kdtrees_built_) BuildKDTrees();
  std::vector<float> p(3 * poses_.size());
  for (size_t i = 0; i < poses_.size(); ++i) { p[3 * i] = poses_[i].translation.x; p[3 * i + 1] = poses_[i].translation.y; p[3 * i + 2] = poses_[i].angle; }
  size_t total = 0;
  for (const PointCloudf& c : robot_frame_point_clouds_) total += c.size();
  if (!copy_world_frame_clouds_to_host_) {   // clouds stay on the device (the EM stage reads them there)
    check(hitl_world_transform(ctx_, p.data(), nullptr), "hitl_world_transform");
    return;
  }
  std::vector<float> w(2 * std::max<size_t>(total, 1));
  check(hitl_world_transform(ctx_, p.data(), w.data()), "hitl_world_transform");
  world_frame_point_clouds_.resize(robot_frame_point_clouds_.size());
  size_t o = 0;
  for (size_t i = 0; i < robot_frame_point_clouds_.size(); ++i) {
    world_frame_point_clouds_[i].resize(robot_frame_point_clouds_[i].size());
    for (size_t k = 0; k < world_frame_point_clouds_[i].size(); ++k, ++o) world_frame_point_clouds_[i][k] = Vector2f(w[2 * o], w[2 * o + 1]);
  }
}

// Scans and their KD-trees become resident once per session (:1307 builds only when kdtrees_ is empty).
void JointOpt::BuildKDTrees() {
  const size_t n = robot_frame_point_clouds_.size();
  if (robot_frame_normal_clouds_.size() != n) throw std::invalid_argument("BuildKDTrees: point and normal clouds differ in size");
  std::vector<uint32_t> off(n + 1, 0);
  for (size_t i = 0; i < n; ++i) {
    if (robot_frame_normal_clouds_[i].size() != robot_frame_point_clouds_[i].size()) throw std::invalid_argument("BuildKDTrees: scan with mismatched normals");
    off[i + 1] = off[i] + (uint32_t)robot_frame_point_clouds_[i].size();
  }
  std::vector<float> pts(2 * std::max<size_t>(off[n], 1)), nrm(2 * std::max<size_t>(off[n], 1));
  for (size_t i = 0; i < n; ++i)
    for (size_t k = 0; k < robot_frame_point_clouds_[i].size(); ++k) {
      const size_t o = off[i] + k;
      pts[2 * o] = robot_frame_point_clouds_[i][k].x; pts[2 * o + 1] = robot_frame_point_clouds_[i][k].y;
      nrm[2 * o] = robot_frame_normal_clouds_[i][k].x; nrm[2 * o + 1] = robot_frame_normal_clouds_[i][k].y;
    }
  // scans + trees are replicated on every context (the search shards by SOURCE pose; every shard needs all targets)
  try {
    for_each_context(ctx_, shard_ctx_, [&](size_t, hitl_ctx* c) {
      check_ctx(c, hitl_set_scans(c, (uint32_t)n, off.data(), pts.data(), nrm.data()), "hitl_set_scans");
      check_ctx(c, hitl_build_kdtrees(c), "hitl_build_kdtrees");
    });
  } catch (const std::exception& e) { last_error_ = e.what(); throw; }
  kdtrees_built_ = true;
}

hitl_stf_opts JointOpt::search_options() const {
  hitl_stf_opts o;
  o.point_match_threshold = localization_options_.kPointMatchThreshold;
  o.min_cosine_angle = (float)cos((double)localization_options_.kMaxStfAngleError);   // :564 (float <- cos of a float option)
  o.max_correspondences_per_point = localization_options_.kMaxCorrespondencesPerPoint;
  o.num_skip_readings = localization_options_.num_skip_readings;
  o.min_inter_pose_correspondence = localization_options_.kMinInterPoseCorrespondence;
  o.disable_culling = 0;
  return o;
}

void JointOpt::FindSTFCorrespondences(size_t min_poses, size_t max_poses) {
  if (!kdtrees_built_) BuildKDTrees();
  if (pose_array_.size() != 3 * poses_.size()) SetParams();
  StfCorrespondenceSet& S = point_point_glob_correspondences_;
  S = StfCorrespondenceSet();
  const hitl_stf_opts o = search_options();
  const uint32_t lo = (uint32_t)std::min<size_t>(min_poses, 0xFFFFFFFFu), hi = (uint32_t)std::min<size_t>(max_poses, 0xFFFFFFFEu);
  uint32_t dummy = 0;
  if (shard_ctx_.empty()) {
    check(hitl_find_stf(ctx_, pose_array_.data(), lo, hi, 0, 0xFFFFFFFFu, &o, &last_search_info_), "hitl_find_stf");
    const uint64_t np = last_search_info_.n_pairs, nm = last_search_info_.n_matches;
    S.pair_i.resize(np); S.pair_j.resize(np); S.pair_off.assign(np + 1, 0); S.k.resize(nm); S.idx.resize(nm);
    S.n_queries = last_search_info_.n_queries;
    check(hitl_get_stf(ctx_, np ? S.pair_i.data() : &dummy, np ? S.pair_j.data() : &dummy, S.pair_off.data(), nm ? S.k.data() : &dummy, nm ? S.idx.data() : &dummy),
          "hitl_get_stf");
    shard_ranges_.assign(1, std::make_pair(0u, (uint32_t)poses_.size())); shard_blocks_.assign(1, np);
    return;
  }
  // One contiguous source-pose range per context, balanced by point count; every range is searched against ALL targets on its own GPU
  // and the per-range lists concatenate to the reference's (i, j, k) order.
  const size_t R = 1 + shard_ctx_.size(), n = poses_.size();
  std::vector<uint64_t> cum(n + 1, 0);
  for (size_t i = 0; i < n; ++i) cum[i + 1] = cum[i] + robot_frame_point_clouds_[i].size();
  shard_ranges_.assign(R, std::make_pair(0u, 0u));
  uint32_t prev = 0;
  for (size_t r = 0; r < R; ++r) {
    uint32_t end = (uint32_t)n;
    if (r + 1 < R) { const uint64_t target = cum[n] * (r + 1) / R; end = (uint32_t)(std::lower_bound(cum.begin(), cum.end(), target) - cum.begin()); end = std::max(prev, std::min<uint32_t>(end, (uint32_t)n)); }
    shard_ranges_[r] = std::make_pair(prev, end);
    prev = end;
  }
  std::vector<StfCorrespondenceSet> part(R);
  std::vector<hitl_stf_info> infos(R);
  try {
    for_each_context(ctx_, shard_ctx_, [&](size_t r, hitl_ctx* c) {
      memset(&infos[r], 0, sizeof(hitl_stf_info));
      check_ctx(c, hitl_find_stf(c, pose_array_.data(), lo, hi, shard_ranges_[r].first, shard_ranges_[r].second, &o, &infos[r]), "hitl_find_stf");
      StfCorrespondenceSet& P = part[r];
      const uint64_t np = infos[r].n_pairs, nm = infos[r].n_matches;
      uint32_t dmy = 0;
      P.pair_i.resize(np); P.pair_j.resize(np); P.pair_off.assign(np + 1, 0); P.k.resize(nm); P.idx.resize(nm);
      check_ctx(c, hitl_get_stf(c, np ? P.pair_i.data() : &dmy, np ? P.pair_j.data() : &dmy, P.pair_off.data(), nm ? P.k.data() : &dmy, nm ? P.idx.data() : &dmy), "hitl_get_stf");
    });
  } catch (const std::exception& e) { last_error_ = e.what(); throw; }
  memset(&last_search_info_, 0, sizeof(last_search_info_));
  shard_blocks_.assign(R, 0);
  S.pair_off.assign(1, 0);
  for (size_t r = 0; r < R; ++r) {
    const StfCorrespondenceSet& P = part[r];
    const uint64_t base = S.k.size();
    S.pair_i.insert(S.pair_i.end(), P.pair_i.begin(), P.pair_i.end()); S.pair_j.insert(S.pair_j.end(), P.pair_j.begin(), P.pair_j.end());
    for (size_t b = 1; b < P.pair_off.size(); ++b) S.pair_off.push_back(base + P.pair_off[b]);
    S.k.insert(S.k.end(), P.k.begin(), P.k.end()); S.idx.insert(S.idx.end(), P.idx.begin(), P.idx.end());
    shard_blocks_[r] = infos[r].n_pairs;
    last_search_info_.n_pairs += infos[r].n_pairs; last_search_info_.n_matches += infos[r].n_matches; last_search_info_.n_raw_matches += infos[r].n_raw_matches;
    last_search_info_.n_queries += infos[r].n_queries; last_search_info_.n_traversals += infos[r].n_traversals; last_search_info_.n_tile_pairs += infos[r].n_tile_pairs;
    last_search_info_.ms_search = std::max(last_search_info_.ms_search, infos[r].ms_search); last_search_info_.ms_total = std::max(last_search_info_.ms_total, infos[r].ms_total);
  }
  S.n_queries = last_search_info_.n_queries;
  (void)dummy;
}

PointToPointGlobCorrespondence JointOpt::GlobCorrespondence(size_t b) const {
  const StfCorrespondenceSet& S = point_point_glob_correspondences_;
  PointToPointGlobCorrespondence c;
  c.pose_index0 = S.pair_i[b]; c.pose_index1 = S.pair_j[b];
  for (uint64_t m = S.pair_off[b]; m < S.pair_off[b + 1]; ++m) {
    c.points0_indices.push_back(S.k[m]); c.points1_indices.push_back(S.idx[m]);
    c.points0.push_back(robot_frame_point_clouds_[c.pose_index0][S.k[m]]); c.points1.push_back(robot_frame_point_clouds_[c.pose_index1][S.idx[m]]);
    c.normals0.push_back(robot_frame_normal_clouds_[c.pose_index0][S.k[m]]); c.normals1.push_back(robot_frame_normal_clouds_[c.pose_index1][S.idx[m]]);
  }
  return c;
}

void JointOpt::FindVisualOdometryCorrespondences(int min_poses, int max_poses) {
  if (!kdtrees_built_) BuildKDTrees();
  if (pose_array_.size() != 3 * poses_.size()) SetParams();
  const hitl_stf_opts o = search_options();
  uint64_t n = 0;
  check(hitl_find_vo(ctx_, pose_array_.data(), min_poses, max_poses, &o, &n), "hitl_find_vo");
  std::vector<uint32_t> sp(n ? n : 1), sk(n ? n : 1), tk(n ? n : 1);
  check(hitl_get_vo(ctx_, sp.data(), sk.data(), tk.data()), "hitl_get_vo");
  for (uint64_t m = 0; m < n; ++m) {   // the reference appends (never clears) this list
    PointToPointCorrespondence c; c.source_pose = sp[m]; c.target_pose = sp[m] + 1; c.source_point = sk[m]; c.target_point = tk[m];
    point_point_correspondences_.push_back(c);
  }
}

ceres::Problem::Options JointOpt::BeginProblem() {
  if (!kdtrees_built_) BuildKDTrees();
  if (pose_array_.size() != 3 * poses_.size()) SetParams();
  const float z9[9] = {1, 0, 0, 1, 1, 1, 1, 0, 0};
  const int32_t zi[2] = {2, 0}; const double zd[4] = {0, 0, 0, 0};
  check(hitl_set_odometry_blocks(ctx_, 0, z9), "hitl_set_odometry_blocks(0)");
  check(hitl_set_human_blocks(ctx_, 0, zi, zd), "hitl_set_human_blocks(0)");
  check(hitl_set_stf_blocks(ctx_, 0, nullptr, nullptr, nullptr, nullptr, nullptr, 0.05f, 0.025f), "hitl_set_stf_blocks(0)");
  check(hitl_set_p2l_glob_blocks(ctx_, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 1.f, 1.f), "hitl_set_p2l_glob_blocks(0)");
  check(hitl_set_p2l_blocks(ctx_, 0, nullptr, nullptr, nullptr, nullptr, nullptr, 1.f, 1.f), "hitl_set_p2l_blocks(0)");
  for (hitl_ctx* c : shard_ctx_) {                  // the extra contexts only ever hold STF blocks
    check_ctx(c, hitl_set_odometry_blocks(c, 0, z9), "hitl_set_odometry_blocks(0)"); check_ctx(c, hitl_set_human_blocks(c, 0, zi, zd), "hitl_set_human_blocks(0)");
    check_ctx(c, hitl_set_stf_blocks(c, 0, nullptr, nullptr, nullptr, nullptr, nullptr, 0.05f, 0.025f), "hitl_set_stf_blocks(0)");
    check_ctx(c, hitl_set_p2l_glob_blocks(c, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 1.f, 1.f), "hitl_set_p2l_glob_blocks(0)");
    check_ctx(c, hitl_set_p2l_blocks(c, 0, nullptr, nullptr, nullptr, nullptr, nullptr, 1.f, 1.f), "hitl_set_p2l_blocks(0)");
  }
  // one evaluator per JointOpt, re-bound per problem: its page-locked staging buffers survive from one problem to the next
  if (!evaluator_) evaluator_.reset(new GpuBlockEvaluator(ctx_, pose_array_.data(), poses_.size(), precision_));
  else evaluator_->Rebind(pose_array_.data(), poses_.size(), precision_);
  evaluator_->SetShardContexts(shard_ctx_);
  ceres::Problem::Options po;
  po.evaluation_callback = evaluator_.get();
  return po;
}

void JointOpt::AddSTFConstraints(ceres::Problem* problem) {
  const StfCorrespondenceSet& S = point_point_glob_correspondences_;
  // The blocks are the kept pairs of the search that is still resident on the device; nothing is re-uploaded.
  check(hitl_set_stf_blocks_from_search(ctx_, localization_options_.kLaserStdDev, localization_options_.kPointPointCorrelationFactor), "hitl_set_stf_blocks_from_search");
  for (hitl_ctx* c : shard_ctx_)                    // every context registers the pairs ITS search found; block b of the concatenated list lives on the context that found it
    check_ctx(c, hitl_set_stf_blocks_from_search(c, localization_options_.kLaserStdDev, localization_options_.kPointPointCorrelationFactor), "hitl_set_stf_blocks_from_search");
  evaluator_->Refresh();
  for (size_t b = 0; b < S.size(); ++b)
    problem->AddResidualBlock(new GpuPointToPointGlobConstraint(evaluator_.get(), b, (int)S.pair_i[b], (int)S.pair_j[b]), NULL, &pose_array_[3 * (size_t)S.pair_i[b]],
                              &pose_array_[3 * (size_t)S.pair_j[b]]);
}

void JointOpt::AddOdometryConstraints(ceres::Problem* problem) {
  const size_t nb = poses_.size() > 0 ? poses_.size() - 1 : 0;
  std::vector<float>& consts = odometry_consts_;      // scratch kept between problems: no 180 KB allocation per build
  consts.resize(9 * std::max<size_t>(nb, 1));
  for (size_t i = 1; i < poses_.size(); ++i) OdometryBlockConstants(poses_[i - 1], poses_[i], &consts[9 * (i - 1)]);
  check(hitl_set_odometry_blocks(ctx_, (uint32_t)nb, consts.data()), "hitl_set_odometry_blocks");
  evaluator_->Refresh();
  for (size_t i = 1; i < poses_.size(); ++i)
    problem->AddResidualBlock(new GpuPoseConstraint(evaluator_.get(), i - 1), NULL, &pose_array_[3 * i - 3], &pose_array_[3 * i]);
  if (!poses_.empty()) {
    problem->AddParameterBlock(&pose_array_[0], 3);
    problem->SetParameterBlockConstant(&pose_array_[0]);                // :824
  }
}

void JointOpt::AddHumanConstraints(ceres::Problem* problem) {
  num_hc_residuals_ = 0;
  std::vector<int32_t> type_pose;
  std::vector<double> targets;
  for (size_t i = 0; i < human_constraints_.size(); ++i)
    for (size_t j = 0; j < human_constraints_[i].size(); ++j) {
      const HumanConstraint& c = human_constraints_[i][j];
      const CorrectionType t = c.constraint_type;
      if (t != CorrectionType::kLineSegmentCorrection && t != CorrectionType::kColinearCorrection && t != CorrectionType::kPerpendicularCorrection &&
          t != CorrectionType::kParallelCorrection)
        continue;                                                      // the reference adds nothing for other types
      double tg[4];
      HumanBlockTargets(poses_, c, tg);
      type_pose.push_back((int32_t)t); type_pose.push_back(c.constrained_pose_id);
      targets.insert(targets.end(), tg, tg + 4);
    }
  const uint32_t nb = (uint32_t)(type_pose.size() / 2);
  const int32_t zi[2] = {2, 0}; const double zd[4] = {0, 0, 0, 0};
  check(hitl_set_human_blocks(ctx_, nb, nb ? type_pose.data() : zi, nb ? targets.data() : zd), "hitl_set_human_blocks");
  evaluator_->Refresh();
  for (uint32_t b = 0; b < nb; ++b) {
    const int pose = type_pose[2 * b + 1];
    double* x = &pose_array_[3 * (size_t)pose];
    switch ((CorrectionType)type_pose[2 * b]) {
      case CorrectionType::kLineSegmentCorrection: problem->AddResidualBlock(new GpuColocationHumanImposedConstraint(evaluator_.get(), b, pose), NULL, x); num_hc_residuals_ += 3; break;
      case CorrectionType::kColinearCorrection: problem->AddResidualBlock(new GpuColinearHumanImposedConstraint(evaluator_.get(), b, pose), NULL, x); num_hc_residuals_ += 2; break;
      case CorrectionType::kPerpendicularCorrection: problem->AddResidualBlock(new GpuPerpendicularHumanImposedConstraint(evaluator_.get(), b, pose), NULL, x); num_hc_residuals_ += 1; break;
      default: problem->AddResidualBlock(new GpuParallelHumanImposedConstraint(evaluator_.get(), b, pose), NULL, x); num_hc_residuals_ += 1; break;
    }
  }
}

ceres::TerminationType JointOpt::SolveHumanConstraints() {
  ceres::Solver::Options solver_options = human_solver_options_;
  solver_options.minimizer_progress_to_stdout = verbose_;
  ceres::Problem problem(BeginProblem());
  AddOdometryConstraints(&problem);
  AddHumanConstraints(&problem);
  ceres::Solve(solver_options, &problem, &last_summary_);
  ceres_cost_.push_back((float)last_summary_.final_cost);
  if (verbose_) printf("%s\n", last_summary_.FullReport().c_str());
  if (!evaluator_->ok()) { last_error_ = evaluator_->error(); return ceres::FAILURE; }
  return last_summary_.termination_type;
}

ceres::TerminationType JointOpt::PostHumanOptimization(int min_pose, int max_pose) {
  (void)min_pose; (void)max_pose;   // the reference ignores them too and searches the whole graph (:1188-1189)
  ceres::Solver::Options solver_options = post_solver_options_;
  solver_options.minimizer_progress_to_stdout = verbose_;
  ceres::Problem problem(BeginProblem());
  const int last = (int)(pose_array_.size() / 3) - 1;
  FindVisualOdometryCorrespondences(0, last);
  FindSTFCorrespondences(0, (size_t)std::max(last, 0));
  AddSTFConstraints(&problem);
  if (!pose_array_.empty()) {
    problem.AddParameterBlock(&pose_array_[0], 3);
    problem.SetParameterBlockConstant(&pose_array_[0]);               // :1197
  }
  ceres::Solve(solver_options, &problem, &last_summary_);
  if (verbose_) printf("%s\n", last_summary_.FullReport().c_str());
  if (!evaluator_->ok()) { last_error_ = evaluator_->error(); return ceres::FAILURE; }
  // gradient + CRS Jacobian of the final problem (:1241-1252)
  std::vector<double> residuals;
  gradients_.clear();
  ceres::Problem::EvaluateOptions eo;
  problem.Evaluate(eo, NULL, &residuals, &gradients_, &ceres_jacobian_);
  return last_summary_.termination_type;
}

void JointOpt::Run() {
  if (!kdtrees_built_) BuildKDTrees();
  CopyTempLaserScans();
  SetParams();
  ceres::TerminationType t = SolveHumanConstraints();
  if (t == ceres::FAILURE || t == ceres::USER_FAILURE) throw std::runtime_error("JointOpt::Run: solver failure: " + last_error_);   // the reference exit(1)s
  CopyParams();
  if (enable_post_human_optimization_) {
    SetParams();
    t = PostHumanOptimization(0, (int)poses_.size() - 1);
    if (t == ceres::FAILURE || t == ceres::USER_FAILURE) throw std::runtime_error("JointOpt::Run: post-HitL solver failure: " + last_error_);
    CopyParams();
  }
}

}  // namespace hitl
