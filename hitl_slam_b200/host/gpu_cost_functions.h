// gpu_cost_functions.h — GPU-backed drop-ins for the Ceres cost blocks of the reference's hot path.
//
// The reference registers one AutoDiffCostFunction per residual block and Ceres calls
// CostFunction::Evaluate on each of them, one at a time, every iteration
// (human_in_the_loop_slam/JointOptimization.cpp:553-557, 817-821, 994-1049).  Here every block
// of a problem is evaluated by ONE batched hitl_eval() on the B200 per evaluation point —
// triggered through ceres::EvaluationCallback::PrepareForEvaluation — and each cost function's
// Evaluate() only copies its own slice (residuals + row-major Jacobians) out of the host staging
// buffer.  Class names follow the functors they replace (residual_functors.h):
//
//   GpuPoseConstraint                         PoseConstraint                    :1054-1133   <3,3,3>
//   GpuPointToPointGlobConstraint             PointToPointGlobConstraint        :768-848     <2,3,3>
//   GpuColocationHumanImposedConstraint       ColocationHumanImposedConstraint  :1299-1330   <3,3>
//   GpuColinearHumanImposedConstraint         ColinearHumanImposedConstraint    :1332-1368   <2,3>
//   GpuPerpendicularHumanImposedConstraint    Perpendicular...                  :1370-1391   <1,3>
//   GpuParallelHumanImposedConstraint         Parallel...                       :1393-1415   <1,3>
//   GpuPointToLineGlobConstraint              PointToLineGlobConstraint         :314-385     <1,3>
//   GpuPointToLineConstraint                  PointToLineConstraint             :557-622     <1,3>
//
// Works unchanged against real Ceres (>= 1.14, which has EvaluationCallback) with -DHITL_USE_SYSTEM_CERES.
#pragma once
#include <mutex>
#include <string>
#include <vector>
#include "../../include/hitl_gpu.h"
#include "hitl_ceres.h"

namespace hitl {

// Owns the host staging buffers of one problem and runs the batch.  pose_array is the caller's
// contiguous parameter array (JointOpt::pose_array_): 3 doubles per pose.
class GpuBlockEvaluator : public ceres::EvaluationCallback {
 public:
  enum Kind { kOdometry = 0, kHuman = 1, kStf = 2, kPointToLineGlob = 3, kPointToLine = 4, kNumKinds = 5 };
  GpuBlockEvaluator(hitl_ctx* ctx, const double* pose_array, size_t n_poses, int precision = 0);
  ~GpuBlockEvaluator();
  GpuBlockEvaluator(const GpuBlockEvaluator&) = delete;
  GpuBlockEvaluator& operator=(const GpuBlockEvaluator&) = delete;
  // Call after every hitl_set_*_blocks registration: the block layout is re-read from the context before the next batch
  // (lazily — a problem is built from several registrations and evaluated once).
  bool Refresh() { dirty_ = true; valid_ = false; return true; }
  // Multi-GPU: contexts on other devices that hold the STF blocks of the other source shards (and nothing else).  Block order of the
  // staging buffers stays [odometry | human | stf of ctx | stf of shard 1 | stf of shard 2 ...] — the reference's AddResidualBlock order
  // when the shards are ascending source ranges.  One batch = one hitl_eval per context, each on its own host thread, each
  // reading its slice back into the shared page-locked staging.
  void SetShardContexts(const std::vector<hitl_ctx*>& extra) { shards_ = extra; dirty_ = true; valid_ = false; }
  void Rebind(const double* pose_array, size_t n_poses) { pose_array_ = pose_array; n_poses_ = n_poses; valid_ = false; dirty_ = true; }
  void Rebind(const double* pose_array, size_t n_poses, int precision) { Rebind(pose_array, n_poses); precision_ = precision; ok_ = true; error_.clear(); }
  void PrepareForEvaluation(bool evaluate_jacobians, bool new_evaluation_point) override;

  // Slice accessors used by the cost functions. Return false when the batch failed.
  bool Fetch(Kind kind, uint64_t block, int pose0, const double* x0, int pose1, const double* x1, int num_residuals, double* residuals,
             double* jac0, double* jac1);
  bool ok() const { return ok_; }
  const char* error() const { return error_.c_str(); }
  uint64_t batches() const { return batches_; }
  float last_ms() const { return last_ms_; }

 private:
  bool run_batch(bool want_jac);
  bool resolve_layout();
  bool matches(int pose, const double* x) const;
  hitl_ctx* ctx_;
  const double* pose_array_;
  size_t n_poses_;
  int precision_;
  std::vector<hitl_ctx*> shards_;                 // extra contexts (ranks 1 ..)
  std::vector<uint64_t> shard_r_off_, shard_j_off_;   // where each extra context's residuals / Jacobians start in the staging buffers
  hitl_eval_layout layout_;
  uint64_t r_off_[kNumKinds], j_off_[kNumKinds];
  // staging of one batch: page-locked (hitl_host_alloc), so hitl_eval's read-back runs at PCIe rate without a driver staging copy;
  // grown on demand, never zero-filled (every block's slice is written by the batch)
  double* r_ = nullptr; double* J_ = nullptr;
  size_t r_cap_ = 0, J_cap_ = 0;
  std::vector<double> snapshot_;
  bool valid_ = false, have_jac_ = false, ok_ = true, dirty_ = true;
  std::string error_;
  uint64_t batches_ = 0;
  float last_ms_ = 0.f;
  std::mutex mu_;
};

namespace detail {
template <int kRes, int... Ns>
class GpuCostBase : public ceres::SizedCostFunction<kRes, Ns...> {
 public:
  GpuCostBase(GpuBlockEvaluator* ev, GpuBlockEvaluator::Kind kind, uint64_t block, int pose0, int pose1) : ev_(ev), kind_(kind), block_(block), pose0_(pose0), pose1_(pose1) {}
  bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const override {
    const bool two = sizeof...(Ns) == 2;
    return ev_->Fetch(kind_, block_, pose0_, parameters[0], two ? pose1_ : -1, two ? parameters[1] : nullptr, kRes, residuals,
                      jacobians ? jacobians[0] : nullptr, (jacobians && two) ? jacobians[1] : nullptr);
  }
 private:
  GpuBlockEvaluator* ev_;
  GpuBlockEvaluator::Kind kind_;
  uint64_t block_;
  int pose0_, pose1_;
};
}  // namespace detail

// `block` = index of the block inside its kind, in registration order (= AddResidualBlock order).
struct GpuPoseConstraint : detail::GpuCostBase<3, 3, 3> {
  GpuPoseConstraint(GpuBlockEvaluator* ev, uint64_t block) : GpuCostBase(ev, GpuBlockEvaluator::kOdometry, block, (int)block, (int)block + 1) {}
};
struct GpuPointToPointGlobConstraint : detail::GpuCostBase<2, 3, 3> {
  GpuPointToPointGlobConstraint(GpuBlockEvaluator* ev, uint64_t block, int pose_index0, int pose_index1)
      : GpuCostBase(ev, GpuBlockEvaluator::kStf, block, pose_index0, pose_index1) {}
};
struct GpuColocationHumanImposedConstraint : detail::GpuCostBase<3, 3> {
  GpuColocationHumanImposedConstraint(GpuBlockEvaluator* ev, uint64_t block, int pose) : GpuCostBase(ev, GpuBlockEvaluator::kHuman, block, pose, -1) {}
};
struct GpuColinearHumanImposedConstraint : detail::GpuCostBase<2, 3> {
  GpuColinearHumanImposedConstraint(GpuBlockEvaluator* ev, uint64_t block, int pose) : GpuCostBase(ev, GpuBlockEvaluator::kHuman, block, pose, -1) {}
};
struct GpuPerpendicularHumanImposedConstraint : detail::GpuCostBase<1, 3> {
  GpuPerpendicularHumanImposedConstraint(GpuBlockEvaluator* ev, uint64_t block, int pose) : GpuCostBase(ev, GpuBlockEvaluator::kHuman, block, pose, -1) {}
};
struct GpuParallelHumanImposedConstraint : detail::GpuCostBase<1, 3> {
  GpuParallelHumanImposedConstraint(GpuBlockEvaluator* ev, uint64_t block, int pose) : GpuCostBase(ev, GpuBlockEvaluator::kHuman, block, pose, -1) {}
};
struct GpuPointToLineGlobConstraint : detail::GpuCostBase<1, 3> {
  GpuPointToLineGlobConstraint(GpuBlockEvaluator* ev, uint64_t block, int pose) : GpuCostBase(ev, GpuBlockEvaluator::kPointToLineGlob, block, pose, -1) {}
};
struct GpuPointToLineConstraint : detail::GpuCostBase<1, 3> {
  GpuPointToLineConstraint(GpuBlockEvaluator* ev, uint64_t block, int pose) : GpuCostBase(ev, GpuBlockEvaluator::kPointToLine, block, pose, -1) {}
};

}  // namespace hitl
