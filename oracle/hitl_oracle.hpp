// hitl_oracle.hpp — CPU restatement of the reference's hot path.  TEST INFRASTRUCTURE ONLY.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// build, load or call anything in oracle/.  The product (hitl_slam_b200/) never does.
//
// PARITY STATUS: the reference (ut-amrl/hitl-slam) ships no tests, golden vectors or fixtures
// for this path, and its translation units cannot be built here as a whole (Eigen, Ceres, glog,
// ROS absent).  What CAN be compiled from the reference's own sources, where they lie, is
// (oracle/Makefile `ref`, outputs in oracle/_ref/, never copied into this repo):
//   * perception_tools/kdtree.cpp            against oracle/shim   -> libkdtree_ref.so
//       pins the tree build and all three queries of this file, node for node / query for query
//       (tests/test_cpu_oracle.py::test_oracle_tree_matches_reference_kdtree);
//   * human_in_the_loop_slam/residual_functors.h + shared/math/eigen_helper.h (header-only code)
//     against oracle/shim2 (Eigen 2-vector / 2x2 / Rotation2D, ceres::Jet, glog stand-ins) -> libfunctors_ref.so
//       pins PointToPointGlob, PoseConstraint, the four human-imposed functors, both point-to-line
//       functors (values and auto-diff Jacobians, <= 1e-12) and DistanceToLineSegment (identical
//       inlier sets) (tests/test_oracle_functors_ref.py).
//   * human_in_the_loop_slam/{JointOptimization,EMinput,ApplyExplicitCorrection,Backprop,HitLSLAM}.cpp + kdtree.cpp + helpers.cpp
//     against oracle/shim3 (2-D Eigen subset incl. Affine2f / Translation2f, a Ceres API slice with a small dense LM, glog, CImg)
//     behind oracle/ref_hitl_capi.cpp -> libhitl_ref.so
//       pins, on the reference's OWN code: FindSTFCorrespondences / FindVisualOdometryCorrespondences / RelativePoseTransform /
//       BuildKDTrees (index lists, transforms and query answers bit for bit), the residual blocks AddOdometryConstraints /
//       AddHumanConstraints / AddSTFConstraints build (<= 1e-12), CopyTempLaserScans / transformPointCloudsToWorldFrame,
//       verifyUserInput, EMInput::Run / EstablishObservationSets / distToLineSeg / SegFitEM, AppExpCorrect::Run (incl.
//       calculateConstraintTargets), Backprop::Run and the whole HitLSLAM::replayLog chain
//       (tests/test_oracle_ref_backend.py; the CUDA path against the same library: tests/test_gpu_vs_reference.py).
// What stays restated-only is THIRD-PARTY arithmetic that is absent from /root/reference and from this image: Eigen 3's
// evaluation order for the 2-D expressions used (the shims restate it a second time, so the pin is on the
// reference's loops and constants, not on Eigen's internals), Ceres' Jet rules and trust-region loop (1.x documentation), and
// libstdc++'s std::sort for equal keys (pinned against this image's libstdc++).
//
// Every function cites the reference lines it follows (paths relative to
// /root/reference/HitL-SLAM/src/).  Float expressions follow Eigen 3's evaluation order as
// listed in SURVEY.md Appendix C; the file must be compiled with -ffp-contract=off for the
// parity build.  libm calls (sinf, cosf, ...) are the platform's, as in the reference.
#pragma once
#include <float.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <algorithm>
#include <utility>
#include <vector>

namespace orc {

// ------------------------------------------------------------------------------------------
// Minimal stand-ins for Eigen::Vector2f / Rotation2Df / Affine2f (operation order: App. C)
// ------------------------------------------------------------------------------------------
struct V2 {
  float x, y;
  V2() : x(0), y(0) {}
  V2(float a, float b) : x(a), y(b) {}
};
inline V2 operator+(V2 a, V2 b) { return V2(a.x + b.x, a.y + b.y); }
inline V2 operator-(V2 a, V2 b) { return V2(a.x - b.x, a.y - b.y); }
inline V2 operator-(V2 a) { return V2(-a.x, -a.y); }
inline V2 operator*(float s, V2 a) { return V2(s * a.x, s * a.y); }
inline V2 operator/(V2 a, float s) { return V2(a.x / s, a.y / s); }
inline bool operator!=(V2 a, V2 b) { return a.x != b.x || a.y != b.y; }
inline float dot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
inline float sqnorm(V2 a) { return a.x * a.x + a.y * a.y; }
inline float norm(V2 a) { return sqrtf(sqnorm(a)); }
inline V2 normalized(V2 a) {
  const float z = sqnorm(a);
  if (z > 0.0f) return a / sqrtf(z);
  return a;
}
struct M2 { float m00, m01, m10, m11; };
inline V2 operator*(const M2& m, V2 v) { return V2(m.m00 * v.x + m.m01 * v.y, m.m10 * v.x + m.m11 * v.y); }
inline M2 operator*(const M2& a, const M2& b) {
  M2 r;
  r.m00 = a.m00 * b.m00 + a.m01 * b.m10;
  r.m01 = a.m00 * b.m01 + a.m01 * b.m11;
  r.m10 = a.m10 * b.m00 + a.m11 * b.m10;
  r.m11 = a.m10 * b.m01 + a.m11 * b.m11;
  return r;
}
// Rotation2Df(angle).toRotationMatrix()
inline M2 rot2(float angle) {
  const float s = sinf(angle), c = cosf(angle);
  M2 m; m.m00 = c; m.m01 = -s; m.m10 = s; m.m11 = c;
  return m;
}
struct Aff { M2 L; V2 t; };
inline Aff aff_inverse(const Aff& a) {
  const float det = a.L.m00 * a.L.m11 - a.L.m10 * a.L.m01;
  const float inv = 1.0f / det;
  Aff r;
  r.L.m00 = a.L.m11 * inv;
  r.L.m10 = -a.L.m10 * inv;
  r.L.m01 = -a.L.m01 * inv;
  r.L.m11 = a.L.m00 * inv;
  M2 neg; neg.m00 = -r.L.m00; neg.m01 = -r.L.m01; neg.m10 = -r.L.m10; neg.m11 = -r.L.m11;
  r.t = neg * a.t;
  return r;
}
inline Aff operator*(const Aff& l, const Aff& r) {
  Aff o; o.L = l.L * r.L; o.t = (l.L * r.t) + l.t; return o;
}
inline V2 operator*(const Aff& a, V2 v) { return (a.L * v) + a.t; }

// ------------------------------------------------------------------------------------------
// KD-tree — perception_tools/kdtree.h:26-96, kdtree.cpp:37-273
// ------------------------------------------------------------------------------------------
struct KDValue { V2 point, normal; int index; KDValue() : index(0) {} };

struct KDTree {
  int dim;
  KDValue value;
  KDTree *left, *right;
  KDTree() : dim(0), left(NULL), right(NULL) {}
  ~KDTree() { delete left; delete right; }

  static float coord(const V2& v, int d) { return d == 0 ? v.x : v.y; }

  // kdtree.cpp:37-69 — sequential float mean, sequential float sum of squared deviations,
  // dimension 0 wins ties and the all-zero case.
  static int splitting_plane(const std::vector<KDValue>& v) {
    V2 mean(0, 0), dev(0, 0);
    for (unsigned i = 0; i < v.size(); ++i) mean = mean + v[i].point;
    mean = mean / static_cast<float>(v.size());
    for (unsigned i = 0; i < v.size(); ++i) {
      dev.x = dev.x + (v[i].point.x - mean.x) * (v[i].point.x - mean.x);
      dev.y = dev.y + (v[i].point.y - mean.y) * (v[i].point.y - mean.y);
    }
    int plane = 0;
    float best = 0.0f;
    if (dev.x > best) { plane = 0; best = dev.x; }
    if (dev.y > best) { plane = 1; best = dev.y; }
    return plane;
  }

  // kdtree.cpp:106-139 — by-value copy, std::sort on one coordinate, median n/2, recurse on
  // fresh sub-vectors (children see the parent's sorted order).
  void build(std::vector<KDValue> v) {
    dim = splitting_plane(v);
    const int d = dim;
    std::sort(v.begin(), v.end(),
              [d](const KDValue& a, const KDValue& b) { return coord(a.point, d) < coord(b.point, d); });
    const unsigned ind = v.size() / 2;
    value = v[ind];
    left = NULL;
    if (ind > 0) {
      left = new KDTree();
      left->build(std::vector<KDValue>(v.begin(), v.begin() + ind));
    }
    right = NULL;
    if (ind < v.size() - 1) {
      right = new KDTree();
      right->build(std::vector<KDValue>(v.begin() + ind + 1, v.end()));
    }
  }

  // kdtree.cpp:141-197 — point-to-plane ranked, Euclidean gated, lossy-pruned search.
  float nearest_point_normal(const V2& q, const float& thr, KDValue* out) const {
    float best = FLT_MAX;
    if (sqnorm(value.point - q) < thr * thr) {
      *out = value;
      best = fabsf(dot(value.normal, q - value.point));
      if (best < FLT_MIN) return 0.0f;
    }
    const float s = coord(q, dim) - coord(value.point, dim);
    const KDTree* other = NULL;
    if (s <= 0.0 && left != NULL) {
      KDValue cand;
      const float d = left->nearest_point_normal(q, thr, &cand);
      if (d < best) { best = d; *out = cand; }
      other = right;
    }
    if (s >= 0.0 && right != NULL) {
      KDValue cand;
      const float d = right->nearest_point_normal(q, thr, &cand);
      if (d < best) { best = d; *out = cand; }
      other = left;
    }
    if (other != NULL && s != 0.0 && fabsf(s) < (best < thr ? best : thr)) {
      KDValue cand;
      const float d = other->nearest_point_normal(q, thr, &cand);
      if (d < best) { best = d; *out = cand; }
    }
    return best;
  }

  // kdtree.cpp:220-273 — Euclidean NN, bound min(best, thr) handed down.
  float nearest_point(const V2& q, const float& thr, KDValue* out) const {
    float best = norm(value.point - q);
    *out = value;
    if (best < FLT_MIN) return 0.0f;
    const float s = coord(q, dim) - coord(value.point, dim);
    const KDTree* other = NULL;
    if (s <= 0.0 && left != NULL) {
      KDValue cand;
      const float d = left->nearest_point(q, (best < thr ? best : thr), &cand);
      if (d < best) { best = d; *out = cand; }
      other = right;
    }
    if (s >= 0.0 && right != NULL) {
      KDValue cand;
      const float d = right->nearest_point(q, (best < thr ? best : thr), &cand);
      if (d < best) { best = d; *out = cand; }
      other = left;
    }
    if (other != NULL && s != 0.0 && fabsf(s) < (best < thr ? best : thr)) {
      KDValue cand;
      const float d = other->nearest_point(q, (best < thr ? best : thr), &cand);
      if (d < best) { best = d; *out = cand; }
    }
    return best;
  }

  // kdtree.cpp:199-218 — radius query (no caller in the reference).
  void neighbor_points(const V2& q, const float& thr, std::vector<KDValue>* out) const {
    if (norm(value.point - q) < thr) out->push_back(value);
    const float s = coord(q, dim) - coord(value.point, dim);
    if (s < thr && left != NULL) left->neighbor_points(q, thr, out);
    if (s > -thr && right != NULL) right->neighbor_points(q, thr, out);
  }

  // Preorder flattening used to compare tree shapes with the product's flat builder.
  void flatten(std::vector<KDValue>* vals, std::vector<int>* dims) const {
    vals->push_back(value);
    dims->push_back(dim);
    if (left) left->flatten(vals, dims);
    if (right) right->flatten(vals, dims);
  }
};

// ------------------------------------------------------------------------------------------
// Scan set + correspondence search — JointOptimization.cpp:296-305, 432-468, 514-537, 561-642
// ------------------------------------------------------------------------------------------
struct StfOptions {                 // vector_mapping.h:121-186 (fields the path reads)
  float kPointMatchThreshold;       // config point_match_threshold = 0.15
  float min_cosine_angle;           // cos(kMaxStfAngleError), JointOptimization.cpp:564
  int kMaxCorrespondencesPerPoint;  // 6
  unsigned num_skip_readings;       // 1
  size_t kMinInterPoseCorrespondence;  // 10, JointOptimization.cpp:563
};

struct GlobCorrespondence {         // vector_mapping.h:102-119 (index part)
  size_t pose_index0, pose_index1;
  std::vector<size_t> points0_indices, points1_indices;
};
struct PointCorrespondence {        // vector_mapping.h:89-98
  size_t source_pose, source_point, target_pose, target_point;
};

struct ScanSet {
  std::vector<std::vector<V2> > points, normals;   // robot frame
  std::vector<KDTree*> trees;
  ~ScanSet() { for (size_t i = 0; i < trees.size(); ++i) delete trees[i]; }

  // JointOptimization.cpp:514-537. Empty scans: the reference leaves an uninitialised root;
  // here an empty tree is NULL and never matches.
  void build_trees() {
    for (size_t i = 0; i < trees.size(); ++i) delete trees[i];
    trees.assign(points.size(), NULL);
    for (size_t i = 0; i < points.size(); ++i) {
      std::vector<KDValue> values(points[i].size());
      for (size_t j = 0; j < points[i].size(); ++j) {
        values[j].index = j;
        values[j].point = points[i][j];
        values[j].normal = normals[i][j];
      }
      if (!values.empty()) { trees[i] = new KDTree(); trees[i]->build(values); }
    }
  }
};

// JointOptimization.cpp:296-305
inline Aff relative_pose_transform(const double* pose_array, unsigned source, unsigned target) {
  Aff s, t;
  s.L = rot2((float)pose_array[3 * source + 2]);
  s.t = V2((float)pose_array[3 * source], (float)pose_array[3 * source + 1]);
  t.L = rot2((float)pose_array[3 * target + 2]);
  t.t = V2((float)pose_array[3 * target], (float)pose_array[3 * target + 1]);
  return aff_inverse(t) * s;
}

// JointOptimization.cpp:561-642. Returns kept pairs in (i asc, j asc) order; *n_queries
// counts the KD queries the reference executes (not skipped by the per-point cap).
// src_lo/src_hi restrict the SOURCE pose range (for shard tests) and src_stride keeps every
// src_stride-th pose of it (timing samples spread over the trajectory); targets always span
// [min_poses, poses_end).
inline void find_stf(const ScanSet& S, const double* pose_array, size_t min_poses, size_t max_poses,
                     const StfOptions& o, std::vector<GlobCorrespondence>* out, uint64_t* n_queries,
                     size_t src_lo = 0, size_t src_hi = (size_t)-1, size_t src_stride = 1) {
  const size_t poses_end = std::min(max_poses + 1, S.points.size());
  if (src_stride == 0) src_stride = 1;
  out->clear();
  if (poses_end <= min_poses) { if (n_queries) *n_queries = 0; return; }
  const size_t lo = std::max(min_poses, src_lo), hi = std::min(poses_end, src_hi);
  const size_t n_src = hi > lo ? (hi - lo + src_stride - 1) / src_stride : 0;
  std::vector<std::vector<GlobCorrespondence> > per_pose(n_src);
  uint64_t queries = 0;
#if defined(_OPENMP)
#pragma omp parallel for schedule(static) reduction(+ : queries)
#endif
  for (size_t q = 0; q < n_src; ++q) {
    const size_t i = lo + q * src_stride;
    std::vector<int> count(S.points[i].size(), 0);
    for (size_t j = min_poses; j < poses_end; ++j) {
      if (i == j) continue;
      GlobCorrespondence c;
      c.pose_index0 = i;
      c.pose_index1 = j;
      const Aff T = relative_pose_transform(pose_array, i, j);
      for (size_t k = 0; k < S.points[i].size(); k += o.num_skip_readings) {
        if (count[k] >= o.kMaxCorrespondencesPerPoint) continue;
        const V2 q = T * S.points[i][k];
        const V2 n = rot2((float)(pose_array[3 * j + 2] - pose_array[3 * i + 2])) * S.normals[i][k];
        KDValue nb;
        float d = FLT_MAX;
        if (S.trees[j] != NULL) d = S.trees[j]->nearest_point_normal(q, o.kPointMatchThreshold, &nb);
        ++queries;
        if (d < o.kPointMatchThreshold && dot(nb.normal, n) > o.min_cosine_angle) {
          c.points0_indices.push_back(k);
          c.points1_indices.push_back(nb.index);
          ++count[k];
        }
      }
      if (c.points0_indices.size() > o.kMinInterPoseCorrespondence) per_pose[q].push_back(c);
    }
  }
  for (size_t i = 0; i < per_pose.size(); ++i) out->insert(out->end(), per_pose[i].begin(), per_pose[i].end());
  if (n_queries) *n_queries = queries;
}

// JointOptimization.cpp:432-468
inline void find_vo(const ScanSet& S, const double* pose_array, int min_poses, int max_poses,
                    const StfOptions& o, std::vector<PointCorrespondence>* out) {
  out->clear();
  const size_t poses_end = std::min(static_cast<size_t>(max_poses + 1), S.points.size());
  if (int(poses_end) < min_poses + 1) return;
  for (size_t i = min_poses; i + 1 < poses_end; ++i) {
    const Aff T = relative_pose_transform(pose_array, i, i + 1);
    for (size_t k = 0; k < S.points[i].size(); ++k) {
      const V2 q = T * S.points[i][k];
      const V2 n = rot2((float)(pose_array[3 * i + 2 + 3] - pose_array[3 * i + 2])) * S.normals[i][k];
      KDValue nb;
      if (S.trees[i + 1] == NULL) continue;
      const float d = S.trees[i + 1]->nearest_point(q, o.kPointMatchThreshold, &nb);
      if (d < o.kPointMatchThreshold && dot(nb.normal, n) > o.min_cosine_angle) {
        PointCorrespondence c;
        c.source_pose = i; c.source_point = k; c.target_pose = i + 1; c.target_point = nb.index;
        out->push_back(c);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// EM point-to-feature assignment — EMinput.cpp:195-323, shared/math/eigen_helper.h:66-81
// ------------------------------------------------------------------------------------------
// eigen_helper.h:66-81 (t is in metres and is compared with 1.0 — kept as is)
inline float distance_to_line_segment(V2 p0, V2 p1, V2 p) {
  const V2 delta = p1 - p0;
  const V2 dir = normalized(delta);
  const float t = dot(p - p0, dir);
  if (t < 0.0f) return norm(p - p0);
  else if (t > 1.0f) return norm(p - p1);
  return fabsf(dot(V2(-dir.y, dir.x), p - p0));
}
// EMinput.cpp:269-279 (float sqrt, then widened; SURVEY.md §8a note)
inline double dist_to_line_seg(V2 p1, V2 p2, V2 p) {
  const float t = dot(p - p1, p2 - p1) / dot(p2 - p1, p2 - p1);
  if (t < 0.0) return double(sqrtf(dot(p - p1, p - p1)));
  else if (t > 1.0) return double(sqrtf(dot(p - p2, p - p2)));
  const V2 proj = p1 + t * (p2 - p1);
  return double(sqrtf(dot(p - proj, p - proj)));
}

struct Inlier { uint32_t pose, index; V2 p; };
// E-step of EMinput.cpp:207-218: inliers in (pose, index) order; threshold is a double.
inline void em_inliers(const std::vector<std::vector<V2> >& world, V2 a, V2 b, double threshold,
                       std::vector<Inlier>* out) {
  out->clear();
  for (size_t i = 0; i < world.size(); ++i)
    for (size_t j = 0; j < world[i].size(); ++j) {
      const double dst = distance_to_line_segment(a, b, world[i][j]);
      if (dst < threshold) { Inlier in; in.pose = i; in.index = j; in.p = world[i][j]; out->push_back(in); }
    }
}

typedef std::vector<std::pair<int, std::vector<int> > > ObsSets;
// EMinput.cpp:281-323
inline void establish_observation_sets(const std::vector<std::vector<V2> >& world, const V2 sel[4],
                                       double threshold, size_t min_obs /* 5: keep if size > 5 */,
                                       ObsSets* first, ObsSets* second) {
  first->clear(); second->clear();
  for (size_t i = 0; i < world.size(); ++i) {
    std::vector<int> a, b;
    for (size_t j = 0; j < world[i].size(); ++j) {
      if (dist_to_line_seg(sel[0], sel[1], world[i][j]) < threshold) a.push_back(j);
      if (dist_to_line_seg(sel[2], sel[3], world[i][j]) < threshold) b.push_back(j);
    }
    if (a.size() > min_obs) first->push_back(std::make_pair((int)i, a));
    if (b.size() > min_obs) second->push_back(std::make_pair((int)i, b));
  }
}

struct OrderResult {
  std::vector<int> corrected_poses, anchor_poses;
  int backprop_start, backprop_end;
  bool swapped;       // strokes were re-ordered (EMinput.cpp:421-432)
  bool valid;
};
// EMinput.cpp:325-455 + 253-267. The reference indexes [0] of possibly-empty lists
// (undefined); here that case returns valid=false with bounds -1.
inline OrderResult order_and_filter(const ObsSets& first_in, const ObsSets& second_in, V2 sel[4]) {
  OrderResult R; R.backprop_start = R.backprop_end = 0; R.swapped = false; R.valid = true;
  std::vector<int> fp, sp;
  for (size_t i = 0; i < first_in.size(); ++i) fp.push_back(first_in[i].first);
  for (size_t i = 0; i < second_in.size(); ++i) sp.push_back(second_in[i].first);
  std::vector<int> overlaps;
  for (size_t i = 0; i < sp.size(); ++i)
    for (size_t j = 0; j < fp.size(); ++j)
      if (sp[i] == fp[j]) overlaps.push_back(fp[j]);
  auto erase_all = [](std::vector<int>& v, const std::vector<int>& w) {
    for (size_t i = 0; i < w.size(); ++i) v.erase(std::remove(v.begin(), v.end(), w[i]), v.end());
  };
  if (overlaps.size() == fp.size() && overlaps.size() == sp.size()) {
    R.backprop_start = R.backprop_end = -1;
  } else if (overlaps.size() == fp.size()) {
    erase_all(sp, overlaps);
  } else if (overlaps.size() == sp.size()) {
    erase_all(fp, overlaps);
  } else if (overlaps.size() > 0) {
    erase_all(fp, overlaps);
    erase_all(sp, overlaps);
  }
  if (fp.empty() || sp.empty()) { R.valid = false; R.backprop_start = R.backprop_end = -1; return R; }
  const int fmin = fp[0], fmax = fp.back(), smin = sp[0], smax = sp.back();
  if (fmin > smax) {
    R.corrected_poses = fp; R.anchor_poses = sp;
    R.backprop_start = smax + 1; R.backprop_end = fmin - 1;
  } else if (fmax < smin) {
    V2 tmp[4] = {sel[2], sel[3], sel[0], sel[1]};
    for (int i = 0; i < 4; ++i) sel[i] = tmp[i];
    R.swapped = true;
    R.corrected_poses = sp; R.anchor_poses = fp;
    R.backprop_start = fmax + 1; R.backprop_end = smin - 1;
  } else {
    R.backprop_start = R.backprop_end = -1;
  }
  return R;
}

// ------------------------------------------------------------------------------------------
// ceres::Jet stand-in and the residual functors
// ------------------------------------------------------------------------------------------
template <int N> struct Jet {
  double a; double v[N];
  Jet() : a(0) { for (int i = 0; i < N; ++i) v[i] = 0; }
  Jet(double s) : a(s) { for (int i = 0; i < N; ++i) v[i] = 0; }
  Jet(double s, int k) : a(s) { for (int i = 0; i < N; ++i) v[i] = 0; v[k] = 1.0; }
};
template <int N> inline Jet<N> operator+(const Jet<N>& f, const Jet<N>& g) { Jet<N> h; h.a = f.a + g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] + g.v[i]; return h; }
template <int N> inline Jet<N> operator-(const Jet<N>& f, const Jet<N>& g) { Jet<N> h; h.a = f.a - g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] - g.v[i]; return h; }
template <int N> inline Jet<N> operator-(const Jet<N>& f) { Jet<N> h; h.a = -f.a; for (int i = 0; i < N; ++i) h.v[i] = -f.v[i]; return h; }
template <int N> inline Jet<N> operator*(const Jet<N>& f, const Jet<N>& g) { Jet<N> h; h.a = f.a * g.a; for (int i = 0; i < N; ++i) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h; }
template <int N> inline Jet<N> operator/(const Jet<N>& f, const Jet<N>& g) {
  Jet<N> h; const double gi = 1.0 / g.a; const double q = f.a * gi; h.a = q;
  for (int i = 0; i < N; ++i) h.v[i] = (f.v[i] - q * g.v[i]) * gi;
  return h;
}
template <int N> inline Jet<N> sqrt(const Jet<N>& f) { Jet<N> h; h.a = ::sqrt(f.a); const double t = 1.0 / (2.0 * h.a); for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * t; return h; }
template <int N> inline Jet<N> sin(const Jet<N>& f) { Jet<N> h; h.a = ::sin(f.a); const double c = ::cos(f.a); for (int i = 0; i < N; ++i) h.v[i] = c * f.v[i]; return h; }
template <int N> inline Jet<N> cos(const Jet<N>& f) { Jet<N> h; h.a = ::cos(f.a); const double s = -::sin(f.a); for (int i = 0; i < N; ++i) h.v[i] = s * f.v[i]; return h; }
template <int N> inline Jet<N> atan2(const Jet<N>& g, const Jet<N>& f) {
  Jet<N> h; h.a = ::atan2(g.a, f.a); const double t = 1.0 / (f.a * f.a + g.a * g.a);
  for (int i = 0; i < N; ++i) h.v[i] = t * (f.a * g.v[i] - g.a * f.v[i]);
  return h;
}
template <int N> inline Jet<N> pow2(const Jet<N>& f) { Jet<N> h; h.a = ::pow(f.a, 2); const double t = 2.0 * ::pow(f.a, 1); for (int i = 0; i < N; ++i) h.v[i] = t * f.v[i]; return h; }
template <int N> inline bool operator<(const Jet<N>& f, double s) { return f.a < s; }
template <int N> inline bool operator>(const Jet<N>& f, double s) { return f.a > s; }
template <int N> inline bool is_nonzero(const Jet<N>& f) { return f.a != 0.0; }   // Jet != T(0.0) compares .a
inline bool is_nonzero(double f) { return f != 0.0; }
inline double sqrt(double x) { return ::sqrt(x); }
inline double sin(double x) { return ::sin(x); }
inline double cos(double x) { return ::cos(x); }
inline double atan2(double y, double x) { return ::atan2(y, x); }
inline double pow2(double x) { return ::pow(x, 2); }
template <typename T> inline T sq(const T& x) { return x * x; }

// residual_functors.h:768-848
struct PointToPointGlob {
  std::vector<V2> points0, points1, normals0, normals1;
  float std_dev, correlation_factor;
  template <typename T> bool operator()(const T* pose0, const T* pose1, T* residuals) const {
    const T t0x = pose0[0], t0y = pose0[1], t1x = pose1[0], t1y = pose1[1];
    const T c0 = cos(pose0[2]), s0 = sin(pose0[2]), c1 = cos(pose1[2]), s1 = sin(pose1[2]);
    T r0(0.0), r1(0.0);
    for (size_t i = 0; i < points0.size(); ++i) {
      const T p0x = (c0 * T(points0[i].x) + (-s0) * T(points0[i].y)) + t0x;
      const T p0y = (s0 * T(points0[i].x) + c0 * T(points0[i].y)) + t0y;
      const T p1x = (c1 * T(points1[i].x) + (-s1) * T(points1[i].y)) + t1x;
      const T p1y = (s1 * T(points1[i].x) + c1 * T(points1[i].y)) + t1y;
      const T n0x = c0 * T(normals0[i].x) + (-s0) * T(normals0[i].y);
      const T n0y = s0 * T(normals0[i].x) + c0 * T(normals0[i].y);
      const T n1x = c1 * T(normals1[i].x) + (-s1) * T(normals1[i].y);
      const T n1y = s1 * T(normals1[i].x) + c1 * T(normals1[i].y);
      const T dx = p1x - p0x, dy = p1y - p0y;
      r0 = r0 + sq((n0x * dx + n0y * dy) * T(correlation_factor) / T(std_dev));
      r1 = r1 + sq((n1x * dx + n1y * dy) * T(correlation_factor) / T(std_dev));
    }
    if (is_nonzero(r0)) r0 = sqrt(r0 / T(static_cast<double>(points0.size())));
    if (is_nonzero(r1)) r1 = sqrt(r1 / T(static_cast<double>(points0.size())));
    residuals[0] = r0; residuals[1] = r1;
    return true;
  }
};

// residual_functors.h:314-385 (sum of squares, no sqrt — :360-364 is commented out there)
struct PointToLineGlob {
  std::vector<V2> points, line_normals; std::vector<float> line_offsets; std::vector<uint8_t> valid;
  float std_dev, correlation_factor;
  template <typename T> bool operator()(const T* pose, T* residuals) const {
    const T c = cos(pose[2]), s = sin(pose[2]);
    T r(0.0);
    for (size_t i = 0; i < points.size(); ++i) {
      if (!valid[i]) continue;
      const T gx = (c * T(points[i].x) + (-s) * T(points[i].y)) + pose[0];
      const T gy = (s * T(points[i].x) + c * T(points[i].y)) + pose[1];
      const T err = (gx * T(line_normals[i].x) + gy * T(line_normals[i].y)) + T(line_offsets[i]);
      r = r + sq(err * T(correlation_factor) / T(std_dev));
    }
    residuals[0] = r;
    return true;
  }
};
// residual_functors.h:557-622
struct PointToLine {
  V2 point, line_normal; float line_offset; bool valid; float std_dev, correlation_factor;
  template <typename T> bool operator()(const T* pose, T* residuals) const {
    if (!valid) { residuals[0] = T(0.0); return true; }
    const T c = cos(pose[2]), s = sin(pose[2]);
    const T gx = (c * T(point.x) + (-s) * T(point.y)) + pose[0];
    const T gy = (s * T(point.x) + c * T(point.y)) + pose[1];
    const T err = (gx * T(line_normal.x) + gy * T(line_normal.y)) + T(line_offset);
    residuals[0] = err * T(correlation_factor) / T(std_dev);
    return true;
  }
};

// residual_functors.h:1054-1133
struct PoseConstraint {
  float a00, a01, a10, a11;   // axis_transform (Matrix2f)
  float radial_std_dev, tangential_std_dev, angular_std_dev, radial_translation, rotation;
  template <typename T> bool operator()(const T* pose1, const T* pose2, T* residuals) const {
    T tx = pose2[0] - pose1[0], ty = pose2[1] - pose1[1];
    const T c = cos(-pose1[2]), s = sin(-pose1[2]);
    const T rx = c * tx + (-s) * ty, ry = s * tx + c * ty;
    const T ax = T(a00) * rx + T(a01) * ry, ay = T(a10) * rx + T(a11) * ry;
    residuals[0] = (ax - T(radial_translation)) / T(radial_std_dev);
    residuals[1] = ay / T(tangential_std_dev);
    const T err = atan2(sin(pose2[2] - pose1[2] - T(rotation)), cos(pose2[2] - pose1[2] - T(rotation)));
    residuals[2] = err / T(angular_std_dev);
    return true;
  }
};

// shared/math/util.h:433-439
inline float angle_mod_f(float angle) { angle -= (2.0 * M_PI) * rint(angle / (2.0 * M_PI)); return angle; }
inline double angle_mod_d(double angle) { angle -= (2.0 * M_PI) * rint(angle / (2.0 * M_PI)); return angle; }

// JointOptimization.cpp:736-825 — constants of block i (between poses i-1 and i) from float poses.
inline PoseConstraint make_odometry_block(const float* poses_xyt /* 3 floats per pose */, size_t i) {
  const float kEpsilon = 1e-6;
  const V2 ti(poses_xyt[3 * i], poses_xyt[3 * i + 1]), tp(poses_xyt[3 * i - 3], poses_xyt[3 * i - 2]);
  const float ai = poses_xyt[3 * i + 2], ap = poses_xyt[3 * i - 1];
  const V2 translation = ti - tp;
  V2 radial, tangential;
  float rotation, radial_translation = 0.0;
  if (fabsf(translation.x) < kEpsilon && fabsf(translation.y) < kEpsilon) {
    radial = V2(cosf(ai), sinf(ai));
    tangential = rot2((float)M_PI_2) * radial;
    rotation = angle_mod_f(ai - ap);
    radial_translation = 0.0;
  } else {
    radial = normalized(rot2(-ap) * translation);
    tangential = rot2((float)M_PI_2) * radial;
    rotation = angle_mod_f(ai - ap);
    radial_translation = norm(translation);
  }
  PoseConstraint pc;
  pc.a00 = radial.x; pc.a01 = radial.y; pc.a10 = tangential.x; pc.a11 = tangential.y;
  pc.radial_std_dev = 0.03; pc.tangential_std_dev = 0.03; pc.angular_std_dev = 0.01;
  pc.radial_translation = radial_translation; pc.rotation = rotation;
  return pc;
}

// human_constraints.h:8-47
enum { kLineSegmentCorrection = 2, kColinearCorrection = 4, kPerpendicularCorrection = 5, kParallelCorrection = 6 };
struct HumanConstraint {
  uint32_t constraint_type; int constrained_pose_id, anchor_pose_id;
  float delta_parallel, delta_perpendicular, delta_angle, relative_penalty_dir;
};
// Unary block produced by JointOptimization.cpp:969-1054 (targets frozen from float poses).
struct HumanBlock {
  uint32_t type; int pose; double x_target, y_target, t_target, penalty_dir;
  int num_residuals() const { return type == kLineSegmentCorrection ? 3 : type == kColinearCorrection ? 2 : 1; }
  // residual_functors.h:1299-1415
  template <typename T> bool operator()(const T* p, T* r) const {
    if (type == kLineSegmentCorrection) {
      r[0] = T(1.0) * (T(x_target) - p[0]); r[1] = T(1.0) * (T(y_target) - p[1]); r[2] = T(1.0) * (T(t_target) - p[2]);
    } else if (type == kColinearCorrection) {
      const T xd = T(::cos(penalty_dir)), yd = T(::sin(penalty_dir));
      r[0] = T(1.0) * (xd * (T(x_target) - p[0]) + yd * (T(y_target) - p[1]));
      r[1] = T(1.0) * (T(t_target) - p[2]);
    } else {
      r[0] = T(1.0) * (T(t_target) - p[2]);
    }
    return true;
  }
};
inline HumanBlock make_human_block(const float* poses_xyt, const HumanConstraint& c) {
  HumanBlock b; b.type = c.constraint_type; b.pose = c.constrained_pose_id;
  b.x_target = b.y_target = b.t_target = b.penalty_dir = 0.0;
  const V2 loc(poses_xyt[3 * c.anchor_pose_id], poses_xyt[3 * c.anchor_pose_id + 1]);
  const float ang = poses_xyt[3 * c.anchor_pose_id + 2];
  const float t_angle = ang + c.delta_angle;
  const float target_angle = atan2f(sinf(t_angle), cosf(t_angle));
  b.t_target = double(target_angle);
  if (c.constraint_type == kLineSegmentCorrection || c.constraint_type == kColinearCorrection) {
    const V2 para(cosf(ang), sinf(ang));
    const V2 perp(-para.y, para.x);
    const V2 target = (loc + c.delta_parallel * para) + c.delta_perpendicular * perp;
    b.x_target = double(target.x); b.y_target = double(target.y);
    if (c.constraint_type == kColinearCorrection) { const float pd = ang + c.relative_penalty_dir; b.penalty_dir = double(pd); }
  }
  return b;
}

// AutoDiffCostFunction<F, R, 3, 3>::Evaluate stand-in: residuals + row-major Jacobians per
// parameter block (either may be NULL).
template <typename F, int R> inline void autodiff2(const F& f, const double* x0, const double* x1, double* res, double* J0, double* J1) {
  Jet<6> a[3], b[3], r[R];
  for (int i = 0; i < 3; ++i) { a[i] = Jet<6>(x0[i], i); b[i] = Jet<6>(x1[i], 3 + i); }
  f(a, b, r);
  for (int i = 0; i < R; ++i) {
    res[i] = r[i].a;
    for (int c = 0; c < 3; ++c) { if (J0) J0[i * 3 + c] = r[i].v[c]; if (J1) J1[i * 3 + c] = r[i].v[3 + c]; }
  }
}
template <typename F> inline void autodiff1(const F& f, int R, const double* x0, double* res, double* J0) {
  Jet<3> a[3], r[3];
  for (int i = 0; i < 3; ++i) a[i] = Jet<3>(x0[i], i);
  f(a, r);
  for (int i = 0; i < R; ++i) { res[i] = r[i].a; if (J0) for (int c = 0; c < 3; ++c) J0[i * 3 + c] = r[i].v[c]; }
}

// ------------------------------------------------------------------------------------------
// SegFitEM — EMinput.cpp:107-191.  One-parameter fit; the solver is Ceres (absent), so the LM
// below follows Ceres' documented trust-region defaults on the
// auto-differentiated residual.  Pose-level agreement is tolerance-level only.
// ------------------------------------------------------------------------------------------
struct SegDistResidual {
  double px, py, cmx, cmy, len;
  template <typename T> bool operator()(const T* theta, T* residual) const {
    T ax = cos(theta[0]), ay = sin(theta[0]);
    const T nrm = sqrt(ax * ax + ay * ay);
    ax = ax / nrm; ay = ay / nrm;
    const T p1x = T(cmx) + T(len) * ax, p1y = T(cmy) + T(len) * ay;
    const T p2x = T(cmx) - T(len) * ax, p2y = T(cmy) - T(len) * ay;
    const T t = ((T(px) - p1x) * (p2x - p1x) + (T(py) - p1y) * (p2y - p1y)) / (pow2(p2x - p1x) + pow2(p2y - p1y));
    T res;
    if (t < 0.0) res = sqrt(pow2(T(px) - p1x) + pow2(T(py) - p1y));
    else if (t > 1.0) res = sqrt(pow2(T(px) - p2x) + pow2(T(py) - p2y));
    else {
      const T projx = p1x + t * (p2x - p1x), projy = p1y + t * (p2y - p1y);
      res = sqrt(pow2(T(px) - projx) + pow2(T(py) - projy));
    }
    residual[0] = res;
    return true;
  }
};

// Levenberg-Marquardt on one parameter with Ceres 1.x defaults: radius 1e4, decrease
// factor 2 (doubling on consecutive failures), radius *= 1/max(1/3, 1-(2rho-1)^3), Jacobi
// scaling 1/(1+sqrt(J^T J)), min_relative_decrease 1e-3, ftol 1e-6, gtol 1e-10, ptol 1e-8.
template <typename CostFn> inline double lm_scalar(CostFn&& cost_grad, double x, int max_iter) {
  double f, g, h;            // cost = 1/2 sum r^2, g = J^T r, h = J^T J
  cost_grad(x, &f, &g, &h);
  double radius = 1e4, decrease = 2.0;
  const double scale = 1.0 / (1.0 + ::sqrt(h));      // jacobi_scaling, fixed at the first point
  if (fabs(g) <= 1e-10) return x;
  for (int it = 0; it < max_iter; ++it) {
    const double hs = h * scale * scale, gs = g * scale;
    double diag = hs; if (diag < 1e-6) diag = 1e-6; if (diag > 1e32) diag = 1e32;
    const double step_s = -gs / (hs + diag / radius);
    const double step = step_s * scale;
    const double model_change = -(step_s * gs + 0.5 * step_s * hs * step_s);
    double f2, g2, h2;
    cost_grad(x + step, &f2, &g2, &h2);
    const double rho = model_change > 0 ? (f - f2) / model_change : -1.0;
    if (fabs(step) <= 1e-8 * (fabs(x) + 1e-8)) break;
    if (rho > 1e-3) {
      const double cost_change = f - f2;
      x += step; f = f2; g = g2; h = h2;
      const double t = 2.0 * rho - 1.0;
      radius = radius / std::max(1.0 / 3.0, 1.0 - t * t * t);
      if (radius > 1e16) radius = 1e16;
      decrease = 2.0;
      if (fabs(g) <= 1e-10) break;
      if (fabs(cost_change) <= 1e-6 * f) break;
    } else {
      radius = radius / decrease; decrease *= 2.0;
      if (radius < 1e-32) break;
    }
  }
  return x;
}

inline void seg_fit_em(const double p1[2], const double p2[2], const double* data, int size, V2* ep1, V2* ep2) {
  const double icm0 = (p1[0] + p2[0]) / 2.0, icm1 = (p1[1] + p2[1]) / 2.0;
  const double hy = ::sqrt(::pow(p1[0] - p2[0], 2) + ::pow(p1[1] - p2[1], 2));
  const double ad = fabs(p1[0] - p2[0]);
  double theta = acos(ad / hy);
  auto cg = [&](double th, double* f, double* g, double* h) {
    double F = 0, G = 0, H = 0;
    for (int i = 0; i < size; ++i) {
      SegDistResidual r; r.px = data[2 * i]; r.py = data[2 * i + 1]; r.cmx = icm0; r.cmy = icm1; r.len = hy / 2.0;
      Jet<1> t(th, 0), out;
      r(&t, &out);
      F += 0.5 * out.a * out.a; G += out.v[0] * out.a; H += out.v[0] * out.v[0];
    }
    *f = F; *g = G; *h = H;
  };
  if (size > 0) theta = lm_scalar(cg, theta, 25);
  double ax = ::cos(theta), ay = ::sin(theta);
  const double nrm = ::sqrt(ax * ax + ay * ay);
  ax /= nrm; ay /= nrm;
  ep1->x = icm0 + (hy / 2.0) * ax; ep1->y = icm1 + (hy / 2.0) * ay;
  ep2->x = icm0 - (hy / 2.0) * ax; ep2->y = icm1 - (hy / 2.0) * ay;
}

// EMinput.cpp:195-250 (max_rounds guards the reference's unbounded while loop)
inline int automatic_endpoint_adjustment(const std::vector<std::vector<V2> >& world, V2 sel[4], int max_rounds = 50) {
  int rounds = 0;
  for (size_t k = 0; k < 2; ++k) {
    const double thresh = 0.05;
    double adj1 = 2 * thresh, adj2 = 2 * thresh;
    int guard = 0;
    while ((adj1 > thresh || adj2 > thresh) && guard++ < max_rounds) {
      std::vector<Inlier> in;
      em_inliers(world, sel[2 * k], sel[2 * k + 1], 0.03, &in);
      const double P1[2] = {double(sel[2 * k].x), double(sel[2 * k].y)};
      const double P2[2] = {double(sel[2 * k + 1].x), double(sel[2 * k + 1].y)};
      std::vector<double> data(2 * in.size());
      for (size_t j = 0; j < in.size(); ++j) { data[2 * j] = double(in[j].p.x); data[2 * j + 1] = double(in[j].p.y); }
      V2 e1, e2;
      seg_fit_em(P1, P2, data.data(), (int)in.size(), &e1, &e2);
      adj1 = norm(sel[2 * k] - e1);
      adj2 = norm(sel[2 * k + 1] - e2);
      sel[2 * k] = e1; sel[2 * k + 1] = e2;
      ++rounds;
    }
  }
  return rounds;
}

// ------------------------------------------------------------------------------------------
// .stfs.covars reader — HitLSLAM_main.cpp:192-300 (normals translated like points: kept)
// ------------------------------------------------------------------------------------------
struct PoseGraph {
  std::vector<float> poses;                       // x, y, theta per pose
  std::vector<float> covariances;                 // 9 per pose
  std::vector<std::vector<V2> > points, normals;  // robot frame
};
inline bool load_pose_graph(const char* path, PoseGraph* g) {
  FILE* f = fopen(path, "r");
  if (!f) return false;
  char name[64]; double ts;
  if (fscanf(f, "%63s\n", name) != 1) { fclose(f); return false; }
  if (fscanf(f, "%lf\n", &ts) != 1) { fclose(f); return false; }
  float px, py, pa, ox, oy, nx, ny, c[9];
  std::vector<V2> pc, nc;
  auto flush = [&]() {
    const size_t n = g->poses.size() / 3 - 1;
    const M2 R = rot2(-g->poses[3 * n + 2]);
    const V2 loc(-g->poses[3 * n], -g->poses[3 * n + 1]);
    for (size_t i = 0; i < pc.size(); ++i) { pc[i] = R * (pc[i] + loc); nc[i] = R * (nc[i] + loc); }
    g->points.push_back(pc); g->normals.push_back(nc);
    pc.clear(); nc.clear();
  };
  while (fscanf(f, "%f,%f,%f,%f,%f, %f,%f,%f, %f, %f, %f, %f, %f, %f, %f, %f\n", &px, &py, &pa, &ox, &oy, &nx, &ny,
                &c[0], &c[1], &c[2], &c[3], &c[4], &c[5], &c[6], &c[7], &c[8]) == 16) {
    const size_t n = g->poses.size() / 3;
    const bool first = n == 0;
    const bool changed = !first && (px != g->poses[3 * n - 3] || py != g->poses[3 * n - 2] || pa != g->poses[3 * n - 1]);
    if (changed) flush();
    if (first || changed) {
      g->poses.push_back(px); g->poses.push_back(py); g->poses.push_back(pa);
      for (int k = 0; k < 9; ++k) g->covariances.push_back(c[k]);
    }
    pc.push_back(V2(ox, oy)); nc.push_back(V2(nx, ny));
  }
  if (!pc.empty()) flush();
  fclose(f);
  return true;
}

// HitLSLAM.cpp:245-254 / JointOptimization.cpp:404-419 — robot frame -> world frame.
inline void world_transform(const float* poses_xyt, const std::vector<std::vector<V2> >& robot,
                            std::vector<std::vector<V2> >* world) {
  world->resize(robot.size());
  for (size_t i = 0; i < robot.size(); ++i) {
    Aff T; T.L = rot2(poses_xyt[3 * i + 2]); T.t = V2(poses_xyt[3 * i], poses_xyt[3 * i + 1]);
    (*world)[i].resize(robot[i].size());
    for (size_t j = 0; j < robot[i].size(); ++j) (*world)[i][j] = T * robot[i][j];
  }
}

// HitLSLAM.cpp:218-243 — verifyUserInput: how many selected points have a world point closer than 0.05 (float), scanned per
// selected point with a break at the first hit; degenerate strokes (sel[0] == sel[1] or sel[2] == sel[3]) void the input.
// seen_mask (optional): bit i = selected point i was seen.
inline size_t verify_user_input(const std::vector<std::vector<V2> >& world, const V2* sel, size_t n_sel, uint32_t* seen_mask = nullptr,
                                float local_select_thresh = 0.05f) {
  size_t points_verified = 0;
  uint32_t mask = 0;
  for (size_t i = 0; i < n_sel; ++i) {
    bool seen = false;
    for (size_t j = 0; j < world.size() && !seen; ++j)
      for (size_t k = 0; k < world[j].size(); ++k)
        if (norm(world[j][k] - sel[i]) < local_select_thresh) { ++points_verified; seen = true; mask |= 1u << i; break; }
  }
  if (n_sel >= 4 && (!(sel[0] != sel[1]) || !(sel[2] != sel[3]))) points_verified = 0;
  if (seen_mask) *seen_mask = mask;
  return points_verified;
}

// ------------------------------------------------------------------------------------------
// Explicit correction + COP-SLAM back-propagation (the two host stages between EM and JointOpt) — "next" row f3.
// Restated from ApplyExplicitCorrection.cpp:150-181, 229-316, 318-445 and Backprop.cpp:98-210; pinned bit for bit against those
// translation units compiled where they lie (oracle/_ref/libhitl_ref.so, tests/test_oracle_ref_backend.py).
// ------------------------------------------------------------------------------------------
struct Pose2Df { V2 translation; float angle; };                 // perception_2d.h:32-34
struct CorrectionPair { int pose; float c[3]; };                 // ApplyExplicitCorrection.h: pair<int, Vector3f>
enum CorrectionKind { kPoint = 1, kLineSegment = 2, kCorner = 3, kColinear = 4, kPerpendicular = 5, kParallel = 6 };   // human_constraints.h:8-17

inline float scalar_cross(V2 a, V2 b) { return a.x * b.y - a.y * b.x; }   // eigen_helper.h:25-29

// CalculateExplicitCorrections (ApplyExplicitCorrection.cpp:318-356) with the four supported modes
// (:150-181 line-to-line, :229-257 colinear, :259-293 perpendicular, :295-316 parallel).  sel = 4 points.
inline void calculate_explicit_corrections(int type, const V2 sel[4], const std::vector<Pose2Df>& poses, const std::vector<int>& corrected,
                                           std::vector<CorrectionPair>* out) {
  V2 anchor_point, cmA;
  float theta_f = 0.0f;
  if (type == kLineSegment) {
    cmA = (sel[1] + sel[0]) / 2.0f;
    const V2 cmB = (sel[3] + sel[2]) / 2.0f;
    const V2 A = normalized(sel[1] - sel[0]), B = normalized(sel[3] - sel[2]);
    double theta = acosf(dot(A, B));                              // acos(float) -> acosf, widened
    if (A.x * B.y - A.y * B.x < 0.0) theta = -theta;              // z of the 3-D cross product
    theta_f = (float)theta;
    anchor_point = cmB;
  } else if (type == kColinear) {
    cmA = 0.5f * (sel[1] + sel[0]);
    const V2 cmB = 0.5f * (sel[3] + sel[2]);
    const V2 A = normalized(sel[1] - sel[0]), B = normalized(sel[3] - sel[2]);
    theta_f = (scalar_cross(A, B) >= 0.0) ? acosf(dot(A, B)) : -acosf(dot(A, B));
    const float alpha = dot(cmA - cmB, B);
    anchor_point = cmB + alpha * B;                               // new_cmA
  } else if (type == kPerpendicular) {
    cmA = (sel[1] + sel[0]) / 2.0f;
    const V2 A = normalized(sel[1] - sel[0]), B = normalized(sel[3] - sel[2]);
    double theta = 0.0;
    if (A.x * B.y - A.y * B.x < 0.0) theta = -acosf(dot(A, B)); else theta = acosf(dot(A, B));
    if (theta == M_PI / 2.0 || theta == -M_PI / 2.0) theta = 0.0;
    else if (theta > 0.0) theta = -(-theta + M_PI / 2.0);
    else theta = -(-theta - M_PI / 2.0);
    theta_f = (float)theta;
    anchor_point = cmA;
  } else if (type == kParallel) {
    cmA = 0.5f * (sel[1] + sel[0]);
    const V2 A = normalized(sel[1] - sel[0]), B = normalized(sel[3] - sel[2]);
    theta_f = (scalar_cross(A, B) >= 0.0) ? acosf(dot(A, B)) : -acosf(dot(A, B));
    anchor_point = cmA;
  } else {
    return;                                                       // point / corner: "not currently supported"
  }
  const M2 R = rot2(theta_f);
  for (size_t i = 0; i < corrected.size(); ++i) {
    const V2 p0 = poses[corrected[i]].translation;
    const V2 p1 = anchor_point + (R * (p0 - cmA));
    const V2 T = p1 - p0;
    CorrectionPair c; c.pose = corrected[i]; c.c[0] = T.x; c.c[1] = T.y; c.c[2] = theta_f;
    out->push_back(c);
  }
}

// AppExpCorrections (:417-445) = CalculateExplicitCorrections + FindContiguousGroups (:360-385) + ApplyExplicitCorrections
// (:387-415) for group 0 only.  Returns the correction C handed to Backprop (first correction of the first group);
// *applied = false when there was no group (C untouched).
inline void app_exp_corrections(int type, const V2 sel[4], std::vector<Pose2Df>* poses_io, const std::vector<int>& corrected, float C[3], bool* applied) {
  std::vector<Pose2Df>& poses = *poses_io;
  std::vector<CorrectionPair> corrections;
  calculate_explicit_corrections(type, sel, poses, corrected, &corrections);
  std::vector<std::vector<CorrectionPair> > groups;
  std::vector<CorrectionPair> one;
  for (size_t i = 0; i + 1 <= poses.size(); ++i) {
    bool in_group = false; size_t which = 0;
    for (size_t j = 0; j < corrections.size(); ++j) if (corrections[j].pose == (int)i) { in_group = true; which = j; }
    if (in_group) one.push_back(corrections[which]);
    else if (!one.empty()) { groups.push_back(one); one.clear(); }
  }
  if (!one.empty()) groups.push_back(one);
  *applied = !groups.empty();
  if (groups.empty()) return;
  const std::vector<CorrectionPair>& g0 = groups[0];
  C[0] = g0[0].c[0]; C[1] = g0[0].c[1]; C[2] = g0[0].c[2];
  for (size_t j = 0; j < g0.size(); ++j) {
    poses[g0[j].pose].translation.x += g0[j].c[0];
    poses[g0[j].pose].translation.y += g0[j].c[1];
    poses[g0[j].pose].angle += g0[j].c[2];
  }
  const int last_pose = g0.back().pose;
  const float* lc = g0.back().c;
  for (size_t k = (size_t)last_pose + 1; k < poses.size(); ++k) {
    poses[k].angle += lc[2];
    const V2 ab = poses[k].translation - poses[last_pose].translation;
    const V2 new_ab = rot2(lc[2]) * ab;
    poses[k].translation = (poses[last_pose].translation + new_ab) + V2(lc[0], lc[1]);
  }
}

// Backprop::BackPropagateError (Backprop.cpp:98-200) behind Run()'s bounds test (:202-210).  cov: 9 floats per pose, row-major.
inline void backprop(std::vector<Pose2Df>* poses_io, std::vector<float>* cov_io, int min_poses, int max_poses, const float correction[3]) {
  if (!(min_poses < max_poses)) return;
  std::vector<Pose2Df>& poses = *poses_io;
  std::vector<float>& cov = *cov_io;
  const size_t n = cov.size() / 9;
  const V2 destination = poses[max_poses].translation + V2(correction[0], correction[1]);
  const float destination_rot_variance = 0.0001f, destination_trans_variance = 0.001f;
  std::vector<float> rot_sigmas, trans_sigmas;
  for (size_t i = 0; i < n; ++i) {
    rot_sigmas.push_back(cov[9 * i + 8]);
    trans_sigmas.push_back((float)((cov[9 * i + 0] + cov[9 * i + 4]) / 2.0));
  }
  std::vector<float> rot_weights, trans_weights;
  float sum_of_rot_var = 0.0f, sum_of_trans_var = 0.0f;
  for (int i = min_poses; i <= max_poses; ++i) { sum_of_rot_var += rot_sigmas[i]; sum_of_trans_var += trans_sigmas[i]; }
  sum_of_rot_var += destination_rot_variance;
  sum_of_trans_var += destination_trans_variance;
  for (int i = min_poses; i <= max_poses; ++i) { rot_weights.push_back(rot_sigmas[i] / sum_of_rot_var); trans_weights.push_back(trans_sigmas[i] / sum_of_trans_var); }
  const float rot_beta = 1 / (1 + (rot_sigmas[max_poses - 1] / destination_rot_variance));
  const float trans_beta = 1 / (1 + (trans_sigmas[max_poses - 1] / destination_trans_variance));
  for (int i = min_poses; i < max_poses; ++i) {
    float* c = &cov[9 * (size_t)i];
    c[0] *= trans_beta; c[1] *= trans_beta; c[3] *= trans_beta; c[4] *= trans_beta;
    c[2] *= rot_beta; c[2] *= rot_beta;                           // (0,2) twice, (1,2) never: as in the reference (:166-169)
    c[6] *= rot_beta; c[7] *= rot_beta;
    c[8] *= rot_beta;
  }
  const float theta = correction[2];
  for (int i = min_poses; i < max_poses; ++i) {
    const float delta_theta = rot_weights[i - min_poses] * theta;
    // Translation2Df(t) * Rotation2Df(d) * Translation2Df(-t): linear = R, translation = t + R * (-t)
    Aff post; post.L = rot2(delta_theta);
    post.t = poses[i].translation + (post.L * (-poses[i].translation));
    poses[i].angle += delta_theta;
    for (int k = i + 1; k <= max_poses; ++k) {
      poses[k].angle += delta_theta;
      poses[k].translation = post * poses[k].translation;
    }
  }
  const V2 trans = destination - poses[max_poses].translation;
  for (int i = min_poses; i < max_poses; ++i) {
    const V2 delta_trans = trans_weights[i - min_poses] * trans;
    for (int k = i + 1; k <= max_poses; ++k) poses[k].translation = poses[k].translation + delta_trans;
  }
}

}  // namespace orc
