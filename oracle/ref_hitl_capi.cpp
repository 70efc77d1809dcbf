// C API over the REFERENCE's own back-end translation units, compiled where they lie under /root/reference by
// oracle/Makefile (-> oracle/_ref/libhitl_ref.so):
//   human_in_the_loop_slam/{JointOptimization,EMinput,ApplyExplicitCorrection,Backprop,HitLSLAM}.cpp,
//   perception_tools/kdtree.cpp, shared/util/helpers.cpp
// against the stand-in headers of oracle/shim3 (Eigen 2-D subset, Ceres API slice with a small dense LM, glog, CImg;
// none of those libraries exists in this image).  TEST INFRASTRUCTURE: it pins the oracle restatement — and through it
// the CUDA path — to the reference's own loops: FindSTFCorrespondences / FindVisualOdometryCorrespondences /
// RelativePoseTransform / BuildKDTrees (JointOptimization.cpp:296-305, 432-468, 514-537, 561-642), the residual blocks
// that AddOdometryConstraints / AddHumanConstraints / AddSTFConstraints build (:539-559, 736-825, 969-1054), EMInput::Run
// and EstablishObservationSets (EMinput.cpp:195-455), AppExpCorrect::Run (ApplyExplicitCorrection.cpp), Backprop::Run
// (Backprop.cpp:98-210) and the whole correction chain HitLSLAM::replayLog (HitLSLAM.cpp:311-398).
// No reference source is copied: this file only CALLS the reference classes (private members are reached by compiling
// this one translation unit with `private` spelled `public`, which changes no symbol name or layout).
#include <stdint.h>
#include <stdio.h>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <queue>
#include <sstream>
#include <string>
#include <utility>
#include <vector>
#include <pthread.h>
#include <semaphore.h>
#include <eigen3/Eigen/Dense>
#include "ceres/ceres.h"

#define private public
#define protected public
#include "HitLSLAM.h"
#undef private
#undef protected

using Eigen::Vector2f;
using perception_2d::Pose2Df;

namespace {

// std::cout of the reference code is silenced while a call runs (it prints per call / per overlap pose).
struct Quiet {
  std::streambuf* old;
  std::ostringstream sink;
  Quiet() : old(std::cout.rdbuf(sink.rdbuf())) {}
  ~Quiet() { std::cout.rdbuf(old); }
};

void fill_clouds(uint32_t n, const uint32_t* off, const float* xy, std::vector<std::vector<Vector2f> >* out) {
  out->assign(n, std::vector<Vector2f>());
  for (uint32_t i = 0; i < n; ++i) {
    (*out)[i].resize(off[i + 1] - off[i]);
    for (uint32_t k = off[i]; k < off[i + 1]; ++k) (*out)[i][k - off[i]] = Vector2f(xy[2 * k], xy[2 * k + 1]);
  }
}
void fill_poses(uint32_t n, const float* xyt, std::vector<Pose2Df>* out) {
  out->resize(n);
  for (uint32_t i = 0; i < n; ++i) { (*out)[i].translation = Vector2f(xyt[3 * i], xyt[3 * i + 1]); (*out)[i].angle = xyt[3 * i + 2]; }
}

struct RefJointOpt {
  JointOpt jo;
  cimg_library::CImg<float> info;
  std::vector<ceres::Problem::Block> recorded;
  std::vector<std::vector<vector2f> > parked_sources;   // point_clouds_g_ entries parked by ref_jo_restrict_sources
};

}  // namespace

extern "C" {

// ---- JointOpt ----------------------------------------------------------------------------------
void* ref_jo_create(uint32_t n, const uint32_t* off, const float* pts_xy, const float* nrm_xy, const float* poses_xyt) {
  Quiet q;
  RefJointOpt* h = new RefJointOpt();
  fill_poses(n, poses_xyt, &h->jo.poses_);
  fill_clouds(n, off, pts_xy, &h->jo.robot_frame_point_clouds_);
  fill_clouds(n, off, nrm_xy, &h->jo.robot_frame_normal_clouds_);
  h->jo.covariances_.assign(n, Eigen::Matrix3f::Zero());
  h->info = cimg_library::CImg<float>(n, n, 1, 1, 0);
  h->jo.info_mat_ = &h->info;
  // what JointOpt::Run does before the solves (JointOptimization.cpp:1305-1309)
  h->jo.ConvertPointClouds();
  h->jo.CopyTempLaserScans();
  h->jo.BuildKDTrees();
  h->jo.SetParams();
  return h;
}
void ref_jo_destroy(void* hp) {
  RefJointOpt* h = (RefJointOpt*)hp;
  if (!h) return;
  for (size_t i = 0; i < h->jo.kdtrees_.size(); ++i) delete h->jo.kdtrees_[i];
  delete h;
}
void ref_jo_set_options(void* hp, float thr, float max_angle, int cap, uint32_t skip, float laser_std, float corr) {
  JointOpt& jo = ((RefJointOpt*)hp)->jo;
  jo.localization_options_.kPointMatchThreshold = thr;
  jo.localization_options_.kMaxStfAngleError = max_angle;
  jo.localization_options_.kMaxCorrespondencesPerPoint = cap;
  jo.localization_options_.num_skip_readings = skip;
  jo.localization_options_.kLaserStdDev = laser_std;
  jo.localization_options_.kPointPointCorrelationFactor = corr;
}
// The cosine gate exactly as FindSTFCorrespondences forms it (:564): unqualified cos() of the float option, stored to a float.
float ref_min_cos(float max_angle) {
  using namespace std;
  const float min_cosine_angle = cos(max_angle);
  return min_cosine_angle;
}
void ref_jo_set_pose_array(void* hp, const double* pose_array) {
  JointOpt& jo = ((RefJointOpt*)hp)->jo;
  std::copy(pose_array, pose_array + jo.pose_array_.size(), jo.pose_array_.begin());
}
void ref_jo_get_pose_array(void* hp, double* pose_array) {
  JointOpt& jo = ((RefJointOpt*)hp)->jo;
  std::copy(jo.pose_array_.begin(), jo.pose_array_.end(), pose_array);
}
void ref_jo_set_poses(void* hp, const float* poses_xyt) {   // poses_ (float) and pose_array_ = SetParams()
  JointOpt& jo = ((RefJointOpt*)hp)->jo;
  fill_poses((uint32_t)jo.poses_.size(), poses_xyt, &jo.poses_);
  jo.SetParams();
}
void ref_jo_world_clouds(void* hp, float* world_xy) {       // CopyTempLaserScans (:404-419)
  JointOpt& jo = ((RefJointOpt*)hp)->jo;
  jo.CopyTempLaserScans();
  size_t o = 0;
  for (size_t i = 0; i < jo.world_frame_point_clouds_.size(); ++i)
    for (size_t k = 0; k < jo.world_frame_point_clouds_[i].size(); ++k, ++o) { world_xy[2 * o] = jo.world_frame_point_clouds_[i][k].x(); world_xy[2 * o + 1] = jo.world_frame_point_clouds_[i][k].y(); }
}
void ref_jo_relative_pose(void* hp, uint32_t n_pairs, const uint32_t* src, const uint32_t* dst, float* out6) {
  JointOpt& jo = ((RefJointOpt*)hp)->jo;
  for (uint32_t p = 0; p < n_pairs; ++p) {
    const Eigen::Affine2f T = jo.RelativePoseTransform(src[p], dst[p]);
    out6[6 * p + 0] = T.linear()(0, 0); out6[6 * p + 1] = T.linear()(0, 1); out6[6 * p + 2] = T.linear()(1, 0); out6[6 * p + 3] = T.linear()(1, 1);
    out6[6 * p + 4] = T.translation()(0); out6[6 * p + 5] = T.translation()(1);
  }
}
// Bounded SAMPLE of the full-map search for timing: JointOpt::FindSTFCorrespondences takes the SOURCE point count of pose i from
// point_clouds_g_[i].size() (:577, :593) and everything about the TARGETS from kdtrees_[j] / robot_frame_point_clouds_[j], so
// parking the point_clouds_g_ entry of every pose outside `keep` makes the reference's own, unmodified loop search exactly the
// kept source poses against ALL target poses of the full map (each source pose is independent: private cap counters, :577).
// keep == NULL or n_keep == 0 restores every pose.
void ref_jo_restrict_sources(void* hp, uint32_t n_keep, const uint32_t* keep) {
  RefJointOpt* h = (RefJointOpt*)hp;
  JointOpt& jo = h->jo;
  const size_t n = jo.point_clouds_g_.size();
  if (h->parked_sources.size() != n) h->parked_sources.assign(n, std::vector<vector2f>());
  for (size_t i = 0; i < n; ++i)
    if (jo.point_clouds_g_[i].empty() && !h->parked_sources[i].empty()) jo.point_clouds_g_[i].swap(h->parked_sources[i]);   // restore
  if (!keep || n_keep == 0) return;
  std::vector<char> kept(n, 0);
  for (uint32_t q = 0; q < n_keep; ++q) if (keep[q] < n) kept[keep[q]] = 1;
  for (size_t i = 0; i < n; ++i)
    if (!kept[i]) jo.point_clouds_g_[i].swap(h->parked_sources[i]);
}
// counts[0] = kept pose pairs, counts[1] = matches in them
void ref_jo_find_stf(void* hp, uint64_t min_pose, uint64_t max_pose, uint64_t* counts) {
  JointOpt& jo = ((RefJointOpt*)hp)->jo;
  jo.FindSTFCorrespondences(min_pose, max_pose);
  counts[0] = jo.point_point_glob_correspondences_.size();
  uint64_t m = 0;
  for (size_t b = 0; b < jo.point_point_glob_correspondences_.size(); ++b) m += jo.point_point_glob_correspondences_[b].points0_indices.size();
  counts[1] = m;
}
// CSR copy of point_point_glob_correspondences_, plus the point / normal copies each entry carries (xy4 = p0, p1, n0, n1 per match; may be null)
void ref_jo_get_stf(void* hp, uint32_t* pair_i, uint32_t* pair_j, uint64_t* pair_off, uint32_t* k, uint32_t* idx, float* xy8) {
  JointOpt& jo = ((RefJointOpt*)hp)->jo;
  uint64_t m = 0;
  for (size_t b = 0; b < jo.point_point_glob_correspondences_.size(); ++b) {
    const vector_localization::VectorMapping::PointToPointGlobCorrespondence& c = jo.point_point_glob_correspondences_[b];
    pair_i[b] = (uint32_t)c.pose_index0; pair_j[b] = (uint32_t)c.pose_index1; pair_off[b] = m;
    for (size_t q = 0; q < c.points0_indices.size(); ++q, ++m) {
      k[m] = (uint32_t)c.points0_indices[q]; idx[m] = (uint32_t)c.points1_indices[q];
      if (xy8) {
        xy8[8 * m + 0] = c.points0[q].x(); xy8[8 * m + 1] = c.points0[q].y(); xy8[8 * m + 2] = c.points1[q].x(); xy8[8 * m + 3] = c.points1[q].y();
        xy8[8 * m + 4] = c.normals0[q].x(); xy8[8 * m + 5] = c.normals0[q].y(); xy8[8 * m + 6] = c.normals1[q].x(); xy8[8 * m + 7] = c.normals1[q].y();
      }
    }
  }
  pair_off[jo.point_point_glob_correspondences_.size()] = m;
}
uint64_t ref_jo_find_vo(void* hp, int min_pose, int max_pose) {
  JointOpt& jo = ((RefJointOpt*)hp)->jo;
  jo.point_point_correspondences_.clear();
  jo.FindVisualOdometryCorrespondences(min_pose, max_pose);
  return jo.point_point_correspondences_.size();
}
void ref_jo_get_vo(void* hp, uint32_t* source_pose, uint32_t* source_point, uint32_t* target_point) {
  JointOpt& jo = ((RefJointOpt*)hp)->jo;
  for (size_t m = 0; m < jo.point_point_correspondences_.size(); ++m) {
    source_pose[m] = (uint32_t)jo.point_point_correspondences_[m].source_pose;
    source_point[m] = (uint32_t)jo.point_point_correspondences_[m].source_point;
    target_point[m] = (uint32_t)jo.point_point_correspondences_[m].target_point;
  }
}
// KD trees as JointOpt::BuildKDTrees made them: nearest-point-normal / nearest-point queries on scan `scan`
void ref_jo_kd_query(void* hp, uint32_t scan, uint32_t nq, const float* q_xy, float thr, int mode, float* dist, int32_t* index) {
  JointOpt& jo = ((RefJointOpt*)hp)->jo;
  for (uint32_t i = 0; i < nq; ++i) {
    KDNodeValue<float, 2> v; v.index = -1;
    const Vector2f p(q_xy[2 * i], q_xy[2 * i + 1]);
    dist[i] = mode == 0 ? jo.kdtrees_[scan]->FindNearestPointNormal(p, thr, &v) : jo.kdtrees_[scan]->FindNearestPoint(p, thr, &v);
    index[i] = v.index;
  }
}
void ref_jo_set_human_constraints(void* hp, uint32_t n_groups, const uint32_t* group_off, const int32_t* hc_i, const float* hc_f) {
  JointOpt& jo = ((RefJointOpt*)hp)->jo;
  jo.human_constraints_.assign(n_groups, std::vector<HumanConstraint>());
  for (uint32_t g = 0; g < n_groups; ++g)
    for (uint32_t b = group_off[g]; b < group_off[g + 1]; ++b) {
      HumanConstraint c;
      c.constraint_type = static_cast<CorrectionType>(hc_i[3 * b]); c.constrained_pose_id = hc_i[3 * b + 1]; c.anchor_pose_id = hc_i[3 * b + 2];
      c.delta_parallel = hc_f[4 * b]; c.delta_perpendicular = hc_f[4 * b + 1]; c.delta_angle = hc_f[4 * b + 2]; c.relative_penalty_dir = hc_f[4 * b + 3];
      jo.human_constraints_[g].push_back(c);
    }
}
// Builds the residual blocks with the reference's own Add*Constraints (which = 0 odometry, 1 human, 2 STF from the last
// FindSTFCorrespondences) on a ceres::Problem, then evaluates every block at `pose_array` (which replaces pose_array_;
// the float poses_ the constants are frozen from are untouched).  Per block: r padded to `r_stride` doubles, J as
// [parameter block][residual][3] padded to `j_stride` doubles.  Returns the number of blocks (or -1 when out_cap is too small);
// n_res[b] = residual count of block b.
int64_t ref_jo_eval_blocks(void* hp, int which, const double* pose_array, uint64_t out_cap, int r_stride, int j_stride, double* r_out, double* J_out, int32_t* n_res) {
  Quiet q;
  JointOpt& jo = ((RefJointOpt*)hp)->jo;
  ceres::Problem problem;
  if (which == 0) jo.AddOdometryConstraints(&problem);
  else if (which == 1) jo.AddHumanConstraints(&problem);
  else jo.AddSTFConstraints(&problem);
  const std::vector<ceres::Problem::Block>& B = problem.blocks();
  if (B.size() > out_cap) return -1;
  std::vector<double> saved = jo.pose_array_;
  std::copy(pose_array, pose_array + jo.pose_array_.size(), jo.pose_array_.begin());
  for (size_t b = 0; b < B.size(); ++b) {
    double r[8], Jb[2][64]; double* Jp[2] = {Jb[0], Jb[1]};
    B[b].cost->Evaluate(B[b].params.data(), r, Jp);
    const int nr = B[b].cost->num_residuals();
    n_res[b] = nr;
    for (int k = 0; k < r_stride; ++k) r_out[b * r_stride + k] = k < nr ? r[k] : 0.0;
    for (int k = 0; k < j_stride; ++k) J_out[b * j_stride + k] = 0.0;
    size_t o = 0;
    for (size_t p = 0; p < B[b].params.size(); ++p) for (int k = 0; k < nr * 3; ++k) J_out[b * j_stride + o++] = Jb[p][k];
  }
  jo.pose_array_ = saved;
  return (int64_t)B.size();
}
// JointOpt::Run (:1295-1385): odometry + human solve through the stand-in LM, CopyParams; returns the float poses.
void ref_jo_run(void* hp, float* poses_xyt_out, double* pose_array_out) {
  Quiet q;
  RefJointOpt* h = (RefJointOpt*)hp;
  h->jo.Run();
  h->jo.info_mat_ = &h->info;   // Run points info_mat_ at a local image
  for (size_t i = 0; i < h->jo.poses_.size(); ++i) { poses_xyt_out[3 * i] = h->jo.poses_[i].translation.x(); poses_xyt_out[3 * i + 1] = h->jo.poses_[i].translation.y(); poses_xyt_out[3 * i + 2] = h->jo.poses_[i].angle; }
  if (pose_array_out) std::copy(h->jo.pose_array_.begin(), h->jo.pose_array_.end(), pose_array_out);
}

// JointOpt::PostHumanOptimization (:1156-1256) on the reference's own CPU code: FindVisualOdometryCorrespondences, FindSTFCorrespondences,
// AddSTFConstraints, Solve (stand-in LM), Problem::Evaluate.  counts = {STF blocks, STF matches, consecutive-pose correspondences, gradient entries}.
int ref_jo_post_human_optimization(void* hp, double* pose_array_out, uint64_t counts[4]) {
  Quiet q;
  JointOpt& jo = ((RefJointOpt*)hp)->jo;
  jo.point_point_correspondences_.clear();
  const int t = (int)jo.PostHumanOptimization(0, (int)jo.pose_array_.size() / 3 - 1);
  std::copy(jo.pose_array_.begin(), jo.pose_array_.end(), pose_array_out);
  uint64_t m = 0;
  for (size_t b = 0; b < jo.point_point_glob_correspondences_.size(); ++b) m += jo.point_point_glob_correspondences_[b].points0_indices.size();
  counts[0] = jo.point_point_glob_correspondences_.size(); counts[1] = m; counts[2] = jo.point_point_correspondences_.size(); counts[3] = jo.gradients_.size();
  return t;
}
// The CPU counterpart of dropin_evaluate_stf_problem: FindSTFCorrespondences + AddSTFConstraints + Problem::Evaluate, all reference code.
int64_t ref_jo_evaluate_stf_problem(void* hp, const double* pose_array, double* cost, double* residuals, uint64_t res_cap, double* gradient) {
  Quiet q;
  JointOpt& jo = ((RefJointOpt*)hp)->jo;
  std::copy(pose_array, pose_array + jo.pose_array_.size(), jo.pose_array_.begin());
  jo.FindSTFCorrespondences(0, jo.pose_array_.size() / 3 - 1);
  ceres::Problem problem;
  jo.AddSTFConstraints(&problem);
  problem.SetParameterBlockConstant(&jo.pose_array_[0]);
  std::vector<double> res, grad;
  ceres::CRSMatrix jac;
  problem.Evaluate(ceres::Problem::EvaluateOptions(), cost, &res, &grad, &jac);
  if (res.size() > res_cap) return -1;
  std::copy(res.begin(), res.end(), residuals);
  std::fill(gradient, gradient + jo.pose_array_.size(), 0.0);
  const std::vector<double*>& order = problem.parameter_blocks();
  for (size_t i = 0; i < order.size(); ++i) {
    const size_t pose = (size_t)(order[i] - &jo.pose_array_[0]) / 3;
    for (int e = 0; e < 3; ++e) gradient[3 * pose + e] = grad[3 * i + e];
  }
  return (int64_t)problem.NumResidualBlocks();
}
void ref_jo_get_gradient(void* hp, double* out) {
  JointOpt& jo = ((RefJointOpt*)hp)->jo;
  std::copy(jo.gradients_.begin(), jo.gradients_.end(), out);
}

// ---- EMInput -----------------------------------------------------------------------------------
// EMInput::Run on world-frame clouds. ret = {n_corrected, n_anchor, backprop first, backprop second}
void ref_em_run(uint32_t n, const uint32_t* off, const float* world_xy, float segs[8], int type, int32_t* ret, int32_t* corrected, int32_t* anchor) {
  Quiet q;
  EMInput em;
  fill_clouds(n, off, world_xy, &em.local_version_point_clouds_);
  for (int k = 0; k < 4; ++k) em.selected_points_.push_back(Vector2f(segs[2 * k], segs[2 * k + 1]));
  em.correction_type_ = static_cast<CorrectionType>(type);
  em.backprop_bounds_ = std::make_pair(0, 0);
  em.Run();
  for (int k = 0; k < 4; ++k) { segs[2 * k] = em.selected_points_[k].x(); segs[2 * k + 1] = em.selected_points_[k].y(); }
  ret[0] = (int32_t)em.corrected_poses_.size(); ret[1] = (int32_t)em.anchor_poses_.size(); ret[2] = em.backprop_bounds_.first; ret[3] = em.backprop_bounds_.second;
  std::copy(em.corrected_poses_.begin(), em.corrected_poses_.end(), corrected);
  std::copy(em.anchor_poses_.begin(), em.anchor_poses_.end(), anchor);
}
// EMInput::EstablishObservationSets (private): the two lists of (pose, point indices) as CSR
void ref_em_observation_sets(uint32_t n, const uint32_t* off, const float* world_xy, const float segs[8], uint32_t n_sets[2], uint32_t* pose0, uint64_t* off0,
                             uint32_t* idx0, uint32_t* pose1, uint64_t* off1, uint32_t* idx1) {
  EMInput em;
  fill_clouds(n, off, world_xy, &em.local_version_point_clouds_);
  for (int k = 0; k < 4; ++k) em.selected_points_.push_back(Vector2f(segs[2 * k], segs[2 * k + 1]));
  const std::pair<std::vector<std::pair<int, std::vector<int> > >, std::vector<std::pair<int, std::vector<int> > > > s = em.EstablishObservationSets();
  const std::vector<std::pair<int, std::vector<int> > >* L[2] = {&s.first, &s.second};
  uint32_t* pose[2] = {pose0, pose1}; uint64_t* offs[2] = {off0, off1}; uint32_t* idx[2] = {idx0, idx1};
  for (int f = 0; f < 2; ++f) {
    uint64_t m = 0;
    n_sets[f] = (uint32_t)L[f]->size();
    for (size_t i = 0; i < L[f]->size(); ++i) {
      pose[f][i] = (uint32_t)(*L[f])[i].first; offs[f][i] = m;
      for (size_t k = 0; k < (*L[f])[i].second.size(); ++k) idx[f][m++] = (uint32_t)(*L[f])[i].second[k];
    }
    offs[f][L[f]->size()] = m;
  }
}
double ref_em_dist_to_line_seg(const float p1[2], const float p2[2], const float p[2]) {
  EMInput em;
  return em.distToLineSeg(Vector2f(p1[0], p1[1]), Vector2f(p2[0], p2[1]), Vector2f(p[0], p[1]));
}
void ref_em_seg_fit(const double p1[2], const double p2[2], const double* data, int n, float out4[4]) {
  EMInput em;
  double a[2] = {p1[0], p1[1]}, b[2] = {p2[0], p2[1]}, cm[2] = {0, 0};
  std::vector<double> d(data, data + 2 * n);
  const std::vector<Vector2f> fit = em.SegFitEM(a, b, cm, d.data(), n);
  out4[0] = fit[0].x(); out4[1] = fit[0].y(); out4[2] = fit[1].x(); out4[3] = fit[1].y();
}

// ---- AppExpCorrect / Backprop ------------------------------------------------------------------
// AppExpCorrect::Run: poses in/out, correction_ out, new_human_constraints_ out (hc_i = type, constrained, anchor; hc_f = dpar, dperp, dangle, rel_pen_dir).
// Returns the number of human constraints.
uint32_t ref_app_exp_run(int type, const float sel8[8], float* poses_xyt, uint32_t n_poses, const int32_t* corrected, uint32_t n_corrected, const int32_t* anchor,
                         uint32_t n_anchor, float C3[3], int32_t* hc_i, float* hc_f) {
  Quiet q;
  AppExpCorrect a;
  a.correction_type_ = static_cast<CorrectionType>(type);
  for (int k = 0; k < 4; ++k) a.selected_points_.push_back(Vector2f(sel8[2 * k], sel8[2 * k + 1]));
  a.corrected_poses_.assign(corrected, corrected + n_corrected);
  a.anchor_poses_.assign(anchor, anchor + n_anchor);
  fill_poses(n_poses, poses_xyt, &a.poses_);
  a.Run();
  for (uint32_t i = 0; i < n_poses; ++i) { poses_xyt[3 * i] = a.poses_[i].translation.x(); poses_xyt[3 * i + 1] = a.poses_[i].translation.y(); poses_xyt[3 * i + 2] = a.poses_[i].angle; }
  C3[0] = a.correction_(0); C3[1] = a.correction_(1); C3[2] = a.correction_(2);
  for (size_t b = 0; b < a.new_human_constraints_.size(); ++b) {
    const HumanConstraint& c = a.new_human_constraints_[b];
    hc_i[3 * b] = (int32_t)c.constraint_type; hc_i[3 * b + 1] = c.constrained_pose_id; hc_i[3 * b + 2] = c.anchor_pose_id;
    hc_f[4 * b] = c.delta_parallel; hc_f[4 * b + 1] = c.delta_perpendicular; hc_f[4 * b + 2] = c.delta_angle; hc_f[4 * b + 3] = c.relative_penalty_dir;
  }
  return (uint32_t)a.new_human_constraints_.size();
}
void ref_backprop_run(float* poses_xyt, float* cov9, uint32_t n_poses, int lo, int hi, const float C3[3]) {
  Quiet q;
  Backprop b;
  fill_poses(n_poses, poses_xyt, &b.poses_);
  b.covariances_.resize(n_poses);
  for (uint32_t i = 0; i < n_poses; ++i) for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) b.covariances_[i](r, c) = cov9[9 * i + 3 * r + c];
  b.correction_ = Eigen::Vector3f(C3[0], C3[1], C3[2]);
  b.backprop_bounds_ = std::make_pair(lo, hi);
  b.Run();
  for (uint32_t i = 0; i < n_poses; ++i) {
    poses_xyt[3 * i] = b.poses_[i].translation.x(); poses_xyt[3 * i + 1] = b.poses_[i].translation.y(); poses_xyt[3 * i + 2] = b.poses_[i].angle;
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) cov9[9 * i + 3 * r + c] = b.covariances_[i](r, c);
  }
}

// ---- HitLSLAM: the whole correction chain ------------------------------------------------------
void* ref_session_create(uint32_t n, const uint32_t* off, const float* pts_xy, const float* nrm_xy, const float* poses_xyt, const float* cov9) {
  Quiet q;
  HitLSLAM* s = new HitLSLAM();
  std::vector<Pose2Df> poses; fill_poses(n, poses_xyt, &poses);
  std::vector<std::vector<Vector2f> > pts, nrm; fill_clouds(n, off, pts_xy, &pts); fill_clouds(n, off, nrm_xy, &nrm);
  std::vector<Eigen::Matrix3f> cov(n);
  for (uint32_t i = 0; i < n; ++i) for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) cov[i](r, c) = cov9 ? cov9[9 * i + 3 * r + c] : 0.0f;
  s->init(poses, pts, nrm, cov, poses);
  return s;
}
void ref_session_destroy(void* sp) { delete (HitLSLAM*)sp; }
// HitLSLAM::replayLog of one logged input; returns the number of human-constraint groups held afterwards
// (it grows by one when the correction was applied).
uint32_t ref_session_replay(void* sp, int type, const float sel8[8]) {
  Quiet q;
  HitLSLAM* s = (HitLSLAM*)sp;
  SingleInput in;
  in.type_of_constraint = static_cast<CorrectionType>(type); in.undone = 0;
  for (int k = 0; k < 4; ++k) in.input_points.push_back(Vector2f(sel8[2 * k], sel8[2 * k + 1]));
  s->replayLog(in);
  return (uint32_t)s->human_constraints_.size();
}
void ref_session_get(void* sp, float* poses_xyt, float* cov9, float* world_xy) {
  HitLSLAM* s = (HitLSLAM*)sp;
  for (size_t i = 0; i < s->poses_.size(); ++i) {
    poses_xyt[3 * i] = s->poses_[i].translation.x(); poses_xyt[3 * i + 1] = s->poses_[i].translation.y(); poses_xyt[3 * i + 2] = s->poses_[i].angle;
    if (cov9) for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) cov9[9 * i + 3 * r + c] = s->covariances_[i](r, c);
  }
  if (world_xy) {
    size_t o = 0;
    for (size_t i = 0; i < s->WORLD_FRAME_point_clouds_.size(); ++i)
      for (size_t k = 0; k < s->WORLD_FRAME_point_clouds_[i].size(); ++k, ++o) { world_xy[2 * o] = s->WORLD_FRAME_point_clouds_[i][k].x(); world_xy[2 * o + 1] = s->WORLD_FRAME_point_clouds_[i][k].y(); }
  }
}
// the human constraints of group g (as AppExpCorrect produced them) — count, then fill
uint32_t ref_session_constraints(void* sp, uint32_t g, int32_t* hc_i, float* hc_f) {
  HitLSLAM* s = (HitLSLAM*)sp;
  if (g >= s->human_constraints_.size()) return 0;
  const std::vector<HumanConstraint>& v = s->human_constraints_[g];
  if (hc_i && hc_f)
    for (size_t b = 0; b < v.size(); ++b) {
      hc_i[3 * b] = (int32_t)v[b].constraint_type; hc_i[3 * b + 1] = v[b].constrained_pose_id; hc_i[3 * b + 2] = v[b].anchor_pose_id;
      hc_f[4 * b] = v[b].delta_parallel; hc_f[4 * b + 1] = v[b].delta_perpendicular; hc_f[4 * b + 2] = v[b].delta_angle; hc_f[4 * b + 3] = v[b].relative_penalty_dir;
    }
  return (uint32_t)v.size();
}
size_t ref_session_verify(void* sp, int type, const float sel8[8]) {   // HitLSLAM::verifyUserInput (:218-243)
  HitLSLAM* s = (HitLSLAM*)sp;
  s->selected_points_.clear();
  for (int k = 0; k < 4; ++k) s->selected_points_.push_back(Vector2f(sel8[2 * k], sel8[2 * k + 1]));
  const size_t v = s->verifyUserInput();
  s->selected_points_.clear();
  (void)type;
  return v;
}

}  // extern "C"
