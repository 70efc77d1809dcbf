// joint_optimization.h — host mirror of the reference's JointOpt stage with the hot path on the B200.
//
// Same public data members and (formerly private) method names as
// human_in_the_loop_slam/JointOptimization.h:55-250 of ut-amrl/hitl-slam, so a caller written
// against the reference (HitLSLAM::Run, HitLSLAM.cpp:449-470) switches by changing the type:
//
//   JointOpt::BuildKDTrees                       :514-537   -> hitl_set_scans + hitl_build_kdtrees
//   JointOpt::FindSTFCorrespondences             :561-642   -> hitl_find_stf
//   JointOpt::FindVisualOdometryCorrespondences  :432-468   -> hitl_find_vo
//   JointOpt::AddSTFConstraints                  :539-559   -> hitl_set_stf_blocks_from_search + GpuPointToPointGlobConstraint blocks
//   JointOpt::AddOdometryConstraints             :736-825   -> hitl_set_odometry_blocks + GpuPoseConstraint blocks
//   JointOpt::AddHumanConstraints                :969-1054  -> hitl_set_human_blocks + Gpu*HumanImposedConstraint blocks
//   JointOpt::SolveHumanConstraints / PostHumanOptimization  :1064-1138, :1156-1256  (ceres::Solve on the host)
//   JointOpt::SetParams / CopyParams / CopyTempLaserScans    :380-419
//
// What does not carry over: the N x N info_mat_ debug image (:1313-1324, :1381-1382; O(N^2) memory,
// SURVEY.md §8 f4), the GVector copy of the clouds (:421-430) and the commented-out experiments.
#pragma once
#include <array>
#include <memory>
#include <string>
#include <utility>
#include <vector>
#include "../../include/hitl_gpu.h"
#include "gpu_cost_functions.h"
#include "hitl_ceres.h"
#include "hitl_types.h"

namespace hitl {

class JointOpt {
 public:
  // ctx: the B200 context every batched call runs on (there is no CPU path).
  explicit JointOpt(hitl_ctx* ctx);
  ~JointOpt();

  void Run();
  void ClearPoses();
  // Multi-GPU, one process: contexts on OTHER devices (created with hitl_create(&c, device)) that share the search and the STF blocks
  // with the constructor's context.  BuildKDTrees replicates scans + trees on every context; FindSTFCorrespondences gives each context a
  // contiguous source-pose range balanced by point count (OMP-over-i of the reference, :575, becomes one GPU per range) and
  // concatenates the per-range lists, which is the reference's block order; every evaluation point costs one hitl_eval per context,
  // concurrently.  Results are identical to the single-context stage (tests/test_gpu_multi.py).  Call before BuildKDTrees.
  void UseShardContexts(const std::vector<hitl_ctx*>& extra);
  size_t NumContexts() const { return 1 + shard_ctx_.size(); }
  std::vector<std::pair<uint32_t, uint32_t>> shard_ranges_;   // source-pose range [lo, hi) of every context in the last search
  std::vector<uint64_t> shard_blocks_;                         // kept pairs (= STF blocks) each context found
  std::vector<float> GetCeresCost() const { return ceres_cost_; }

  // ---- public state of the reference class (JointOptimization.h:64-89) ----
  std::vector<Pose2Df> poses_;
  std::vector<PointCloudf> robot_frame_point_clouds_;
  std::vector<NormalCloudf> robot_frame_normal_clouds_;
  std::vector<std::array<float, 9>> covariances_;
  std::vector<std::vector<HumanConstraint>> human_constraints_;
  std::vector<double> gradients_;
  ceres::CRSMatrix ceres_jacobian_;
  int num_hc_residuals_ = 0;

  // ---- knobs the reference hard-codes or leaves commented out ----
  VectorMappingOptions localization_options_;
  bool enable_post_human_optimization_ = false;   // the PostHumanOptimization call is inside /* */ at :1353-1373
  int precision_ = 0;                              // 0 = FP64 blocks, 1 = FP32 mode (hitl_eval)
  bool copy_world_frame_clouds_to_host_ = false;   // CopyTempLaserScans: also fill world_frame_point_clouds_ (nothing on the path reads it)
  bool verbose_ = false;
  ceres::Solver::Options human_solver_options_;    // defaults of SolveHumanConstraints (:1059-1063)
  ceres::Solver::Options post_solver_options_;     // SetSolverOptions (:145-162) + max_num_iterations 100

  // ---- the stage's steps (private in the reference; public here so each can be parity-tested) ----
  void SetParams();
  void CopyParams();
  void CopyTempLaserScans();
  void BuildKDTrees();
  void FindSTFCorrespondences(size_t min_poses, size_t max_poses);
  void FindVisualOdometryCorrespondences(int min_poses, int max_poses);
  void AddSTFConstraints(ceres::Problem* problem);
  void AddOdometryConstraints(ceres::Problem* problem);
  void AddHumanConstraints(ceres::Problem* problem);
  ceres::TerminationType SolveHumanConstraints();
  ceres::TerminationType PostHumanOptimization(int min_pose, int max_pose);
  // Starts a new problem on the context: drops every registered block kind and returns Problem
  // options carrying the evaluation callback that batches this problem's blocks on the GPU.
  ceres::Problem::Options BeginProblem();

  // ---- results ----
  std::vector<double> pose_array_;
  std::vector<PointCloudf> world_frame_point_clouds_;
  StfCorrespondenceSet point_point_glob_correspondences_;            // CSR view of the reference's vector of blocks
  std::vector<PointToPointCorrespondence> point_point_correspondences_;
  PointToPointGlobCorrespondence GlobCorrespondence(size_t block) const;   // materialises one reference-shaped block
  hitl_stf_info last_search_info_;
  ceres::Solver::Summary last_summary_;
  bool kdtrees_built_ = false;
  std::string last_error_;

 private:
  void check(int rc, const char* where);
  hitl_stf_opts search_options() const;
  hitl_ctx* ctx_;
  std::vector<hitl_ctx*> shard_ctx_;                           // extra contexts (ranks 1 ..)
  std::unique_ptr<GpuBlockEvaluator> evaluator_;
  std::vector<float> odometry_consts_;                         // AddOdometryConstraints' upload staging, kept between problems
  std::vector<float> ceres_cost_;
};

// Constants of the PoseConstraint block between poses i-1 and i (JointOptimization.cpp:743-783):
// 9 floats = axis_transform (row-major), radial / tangential / angular std-dev, radial_translation, rotation.
void OdometryBlockConstants(const Pose2Df& prev, const Pose2Df& cur, float out9[9]);
// Targets of one human block (JointOptimization.cpp:980-1049): {x, y, theta, penalty_dir}.
void HumanBlockTargets(const std::vector<Pose2Df>& poses, const HumanConstraint& c, double out4[4]);

}  // namespace hitl
