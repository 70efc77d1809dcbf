// em_input.h — host mirror of the reference's EMInput stage (human_in_the_loop_slam/EMinput.h:46-101)
// with its two O(N*P) scans on the B200.
//
//   EMInput::AutomaticEndpointAdjustment   EMinput.cpp:195-250   E-step -> hitl_em_inliers (ordered inlier list),
//                                                                M-step  = SegFitEM on the host (1-parameter fit, ceres::Solve)
//   EMInput::SegFitEM / segDistResidualEM  EMinput.cpp:107-191   host, AutoDiffCostFunction<.,1,1> per inlier, DENSE_QR, <= 25 iterations
//   EMInput::EstablishObservationSets      EMinput.cpp:281-323   -> hitl_em_assign (both strokes in one pass)
//   EMInput::OrderAndFilterUserInput       EMinput.cpp:325-455   host (integer lists)
//   EMInput::Run                           EMinput.cpp:457-472
//
// The world-frame clouds stay resident on the device between the E-steps and the assignment:
// upload them once with SetWorldClouds() (or let JointOpt/hitl_world_transform produce them in place).
#pragma once
#include <string>
#include <utility>
#include <vector>
#include "../../include/hitl_gpu.h"
#include "hitl_ceres.h"
#include "hitl_types.h"

namespace hitl {

typedef std::vector<std::pair<int, std::vector<int>>> PoseObservations;

// SegFitEM's computation (EMinput.cpp:152-191): refit the direction of the segment p1-p2 about its
// fixed midpoint and length to the inliers data[2*size]; returns the two new endpoints.
std::vector<Vector2f> FitSegmentAngle(const double* p1, const double* p2, const double* data, int size, double* theta_out = nullptr, int* iterations_out = nullptr);

class EMInput {
 public:
  explicit EMInput(hitl_ctx* ctx);
  virtual ~EMInput();

  void Run();

  // ---- public state of the reference class (EMinput.h:55-70) ----
  std::vector<Vector2f> selected_points_;                       // in: 4 stroke endpoints; out: refit (and possibly swapped)
  std::vector<std::vector<Vector2f>> local_version_point_clouds_;   // WORLD frame, one cloud per pose
  std::vector<int> corrected_poses_;
  std::vector<int> anchor_poses_;
  std::pair<int, int> backprop_bounds_;
  CorrectionType correction_type_ = CorrectionType::kUnknownCorrection;

  // ---- mirror-only ----
  // true: the context already holds the world clouds of the current scans (hitl_world_transform /
  // hitl_set_world_clouds) and local_version_point_clouds_ is not uploaded again.
  bool world_clouds_resident_ = false;
  // M-step placement: true (default) = hitl_em_refit, E-step + SegFitEM's LM in one device call per round (nothing but 64 bytes crosses
  // PCIe); false = inliers copied back and FitSegmentAngle on the host LM (the checker: tests compare the two to 1e-9 in theta).
  bool device_m_step_ = true;
  // Device M-step only: rounds enqueued per host wait (hitl_em_refit_chain, both strokes together).  The stopping rule is applied to
  // the returned sequence, so a stroke that converges early simply ignores the extra rounds; 1 = one wait per round.
  int device_chain_rounds_ = 2;
  // OrderAndFilterUserInput reads only the observing poses; true also copies every pose's index list back (EstablishObservationSets always does)
  bool fetch_observation_indices_ = false;
  double last_theta_[2] = {0, 0};    // fitted direction angle of each stroke's last M-step
  int max_em_rounds_ = 1000;         // guard for the reference's unbounded while loop (:200)
  int em_rounds_[2] = {0, 0};        // E/M rounds each stroke took
  uint64_t em_inliers_[2] = {0, 0};  // inliers of the last E-step of each stroke
  std::string last_error_;

  // ---- the stage's steps (private in the reference) ----
  void UploadWorldClouds();
  void AutomaticEndpointAdjustment();
  std::vector<Vector2f> SegFitEM(double* p1, double* p2, double* cm, double* data, int size);
  std::pair<PoseObservations, PoseObservations> EstablishObservationSets();
  std::pair<PoseObservations, PoseObservations> ObservationSets(bool with_indices);
  void OrderAndFilterUserInput();
  void SetCorrectionRelations(const PoseObservations& first_poses_obs, const PoseObservations& second_poses_obs);

 private:
  void check(int rc, const char* where);
  hitl_ctx* ctx_;
};

}  // namespace hitl
