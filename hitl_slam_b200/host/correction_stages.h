// correction_stages.h — host mirrors of the two stages that sit between EMInput and JointOpt in one human correction
// (HitLSLAM::Run, human_in_the_loop_slam/HitLSLAM.cpp:379-484): the closed-form rigid correction and COP-SLAM
// back-propagation.  Same public members and Run() as the reference classes
//   AppExpCorrect   ApplyExplicitCorrection.h / .cpp:150-181, 229-316 (the four supported modes), :360-445
//   Backprop        Backprop.h:47-60, Backprop.cpp:98-210
// Both are O(#poses) scalar code except Backprop's pose update, which is O(L^2) sequential float work in the reference
// and runs on the GPU here with the same operation sequence per pose (csrc/backprop.cu, hitl_backprop_poses).
#pragma once
#include <array>
#include <utility>
#include <vector>
#include "hitl_types.h"

struct hitl_ctx;

namespace hitl {

typedef std::array<float, 3> Vector3f;
typedef std::array<float, 9> Matrix3f;                       // row-major
typedef std::pair<int, Vector3f> CorrectionPair;

std::vector<HumanConstraint> CalculateConstraintTargets(const std::vector<Pose2Df>& poses, const std::vector<Vector2f>& selected_points, CorrectionType type,
                                                        const std::vector<int>& anchor_poses, const std::vector<int>& corrected_poses);

class AppExpCorrect {
 public:
  void Run();                                                // correction_ = AppExpCorrections(); calculateConstraintTargets()

  CorrectionType correction_type_ = CorrectionType::kUnknownCorrection;
  std::vector<Vector2f> selected_points_;                    // 4 points: feature A (to be moved), feature B
  std::vector<int> corrected_poses_, anchor_poses_;
  std::vector<Pose2Df> poses_;                               // in / out
  Vector3f correction_ = {{0.f, 0.f, 0.f}};                  // out: the correction handed to Backprop
  std::vector<HumanConstraint> new_human_constraints_;       // out
  bool applied_ = false;                                     // a contiguous group of corrected poses existed

  void CalculateExplicitCorrections(std::vector<CorrectionPair>* corrections) const;
  Vector3f AppExpCorrections();
};

class Backprop {
 public:
  explicit Backprop(hitl_ctx* ctx) : ctx_(ctx) {}
  void Run();                                                // BackPropagateError() when bounds.first < bounds.second

  std::pair<int, int> backprop_bounds_ = {0, 0};
  std::vector<Matrix3f> covariances_;                        // in / out
  std::vector<Pose2Df> poses_;                               // in / out
  Vector3f correction_ = {{0.f, 0.f, 0.f}};
  float last_device_ms_ = 0.f;                               // device time of the pose update

 private:
  void BackPropagateError();
  hitl_ctx* ctx_;
};

}  // namespace hitl
