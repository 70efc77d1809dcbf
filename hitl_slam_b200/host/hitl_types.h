// hitl_types.h — plain host types the mirror classes use where the reference uses Eigen / its own PODs.
//
// Mirrors (paths relative to HitL-SLAM/src/ of ut-amrl/hitl-slam):
//   Eigen::Vector2f, perception_2d::Pose2Df        perception_tools/perception_2d.h:30-86
//   CorrectionType, HumanConstraint                human_in_the_loop_slam/human_constraints.h:8-47
//   PointToPointCorrespondence / ...GlobCorrespondence   episodic_non_markov_localization/vector_mapping.h:89-119
//   VectorMappingOptions (the fields the path reads)     vector_mapping.h:121-186, config/non_markov_localization.cfg:11-50
// Eigen is not a dependency of this library: the geometry that decides bits lives in csrc/hitl_math.h.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <math.h>
#include <utility>
#include <vector>

namespace hitl {

struct Vector2f {
  float x = 0.f, y = 0.f;
  Vector2f() {}
  Vector2f(float x_, float y_) : x(x_), y(y_) {}
  float& operator[](int i) { return i ? y : x; }
  float operator[](int i) const { return i ? y : x; }
};
inline Vector2f operator+(Vector2f a, Vector2f b) { return Vector2f(a.x + b.x, a.y + b.y); }
inline Vector2f operator-(Vector2f a, Vector2f b) { return Vector2f(a.x - b.x, a.y - b.y); }
inline Vector2f operator*(float s, Vector2f a) { return Vector2f(s * a.x, s * a.y); }
inline float dot(Vector2f a, Vector2f b) { return a.x * b.x + a.y * b.y; }
inline float norm(Vector2f a) { return sqrtf(a.x * a.x + a.y * a.y); }

typedef std::vector<Vector2f> PointCloudf;
typedef std::vector<Vector2f> NormalCloudf;

struct Pose2Df {
  Vector2f translation;
  float angle = 0.f;
  Pose2Df() {}
  Pose2Df(float a, float x, float y) : translation(x, y), angle(a) {}
};

enum class CorrectionType : uint32_t {
  kUnknownCorrection = 0,
  kPointCorrection = 1,          // not supported by the reference either
  kLineSegmentCorrection = 2,
  kCornerCorrection = 3,         // not supported by the reference either
  kColinearCorrection = 4,
  kPerpendicularCorrection = 5,
  kParallelCorrection = 6,
};

struct HumanConstraint {
  CorrectionType constraint_type = CorrectionType::kUnknownCorrection;
  int constrained_pose_id = 0;
  int anchor_pose_id = 0;
  float delta_parallel = 0.f;
  float delta_perpendicular = 0.f;
  float delta_angle = 0.f;
  float relative_penalty_dir = 0.f;
};

// One matched point pair between consecutive poses (FindVisualOdometryCorrespondences).
struct PointToPointCorrespondence {
  size_t source_pose = 0, target_pose = 0;
  size_t source_point = 0, target_point = 0;
};

// All matches of one ordered pose pair (FindSTFCorrespondences).  The reference stores copies of
// the matched points and normals per block; here the block is a view into the CSR arrays the
// GPU search produced, and the copies are materialised only on request.
struct PointToPointGlobCorrespondence {
  size_t pose_index0 = 0, pose_index1 = 0;
  std::vector<size_t> points0_indices, points1_indices;
  std::vector<Vector2f> points0, points1, normals0, normals1;
};

// Correspondences of the last search in CSR form (reference order: pose_index0, pose_index1, point index).
struct StfCorrespondenceSet {
  std::vector<uint32_t> pair_i, pair_j;
  std::vector<uint64_t> pair_off;   // n_pairs + 1
  std::vector<uint32_t> k, idx;     // point index in scan pair_i / pair_j
  uint64_t n_queries = 0;           // KD queries the reference semantics execute
  size_t size() const { return pair_i.size(); }
};

// The option fields the hot path reads, with the values of config/non_markov_localization.cfg.
struct VectorMappingOptions {
  float kPointMatchThreshold = 0.15f;            // :47
  float kMaxStfAngleError = 0.436332313f;        // RAD(25.0), :48
  int kMaxCorrespondencesPerPoint = 6;           // :49
  float kPointPointCorrelationFactor = 1.0f / 40.0f;   // :50
  float kLaserStdDev = 0.05f;                    // :11
  unsigned int num_skip_readings = 1;            // :16
  unsigned int kMinInterPoseCorrespondence = 10; // JointOptimization.cpp:563 (static const in the function)
};

}  // namespace hitl
