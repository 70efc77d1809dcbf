// gpu_cost_functions.cpp — the batch behind the GPU-backed cost blocks (see gpu_cost_functions.h).
#include "gpu_cost_functions.h"

#include <string.h>
#include <thread>

namespace hitl {

// doubles per block in hitl_eval's output (include/hitl_gpu.h "r_out / J_out layout")
static const int kResPerBlock[GpuBlockEvaluator::kNumKinds] = {3, 3, 2, 1, 1};
static const int kJacPerBlock[GpuBlockEvaluator::kNumKinds] = {18, 9, 12, 3, 3};

GpuBlockEvaluator::GpuBlockEvaluator(hitl_ctx* ctx, const double* pose_array, size_t n_poses, int precision)
    : ctx_(ctx), pose_array_(pose_array), n_poses_(n_poses), precision_(precision) {
  memset(&layout_, 0, sizeof(layout_));
  memset(r_off_, 0, sizeof(r_off_)); memset(j_off_, 0, sizeof(j_off_));
}

GpuBlockEvaluator::~GpuBlockEvaluator() {
  if (r_) hitl_host_free(r_);
  if (J_) hitl_host_free(J_);
}

bool GpuBlockEvaluator::resolve_layout() {
  if (hitl_eval_layout_get(ctx_, &layout_) != HITL_OK) { ok_ = false; error_ = "hitl_eval_layout_get failed"; return false; }
  const uint64_t counts[kNumKinds] = {layout_.n_odometry, layout_.n_human, layout_.n_stf, layout_.n_p2l_glob, layout_.n_p2l};
  uint64_t ro = 0, jo = 0;
  for (int k = 0; k < kNumKinds; ++k) { r_off_[k] = ro; j_off_[k] = jo; ro += counts[k] * kResPerBlock[k]; jo += counts[k] * kJacPerBlock[k]; }
  // extra contexts carry STF blocks only; their slices follow the first context's, shard after shard
  shard_r_off_.assign(shards_.size(), 0); shard_j_off_.assign(shards_.size(), 0);
  uint64_t r_end = layout_.n_residuals, j_end = layout_.n_jacobian;
  if (!shards_.empty() && (layout_.n_p2l_glob || layout_.n_p2l)) { ok_ = false; error_ = "sharded evaluation does not support point-to-line blocks"; return false; }
  for (size_t q = 0; q < shards_.size(); ++q) {
    hitl_eval_layout lq;
    if (hitl_eval_layout_get(shards_[q], &lq) != HITL_OK) { ok_ = false; error_ = "hitl_eval_layout_get failed on a shard"; return false; }
    if (lq.n_odometry || lq.n_human || lq.n_p2l_glob || lq.n_p2l) { ok_ = false; error_ = "a shard context holds blocks other than STF blocks"; return false; }
    shard_r_off_[q] = r_end; shard_j_off_[q] = j_end;
    r_end += lq.n_residuals; j_end += lq.n_jacobian;
    layout_.n_stf += lq.n_stf;
  }
  layout_.n_residuals = r_end; layout_.n_jacobian = j_end;
  auto grow = [](double** p, size_t* cap, size_t need) {
    if (need <= *cap) return true;
    if (*p) hitl_host_free(*p);
    const size_t want = need + need / 4 + 64;
    *p = static_cast<double*>(hitl_host_alloc(want * sizeof(double)));
    *cap = *p ? want : 0;
    return *p != nullptr;
  };
  if (!grow(&r_, &r_cap_, layout_.n_residuals + 1) || !grow(&J_, &J_cap_, layout_.n_jacobian + 1)) { ok_ = false; error_ = "hitl_host_alloc failed"; return false; }
  dirty_ = false; ok_ = true;
  return true;
}

bool GpuBlockEvaluator::run_batch(bool want_jac) {
  if (dirty_ && !resolve_layout()) { valid_ = false; return false; }
  // one hitl_eval per context, the extra contexts on their own host threads (every ABI call selects its context's device)
  std::vector<int> rcs(shards_.size(), HITL_OK);
  std::vector<std::thread> workers;
  for (size_t q = 0; q < shards_.size(); ++q)
    workers.emplace_back([&, q]() { float ms = 0.f; rcs[q] = hitl_eval(shards_[q], pose_array_, precision_, r_ + shard_r_off_[q], want_jac ? J_ + shard_j_off_[q] : nullptr, &ms); });
  const int rc = hitl_eval(ctx_, pose_array_, precision_, r_, want_jac ? J_ : nullptr, &last_ms_);
  for (std::thread& w : workers) w.join();
  ++batches_;
  if (rc != HITL_OK) { ok_ = false; valid_ = false; error_ = hitl_last_error(ctx_); return false; }
  for (size_t q = 0; q < shards_.size(); ++q)
    if (rcs[q] != HITL_OK) { ok_ = false; valid_ = false; error_ = std::string("shard: ") + hitl_last_error(shards_[q]); return false; }
  snapshot_.assign(pose_array_, pose_array_ + 3 * n_poses_);
  valid_ = true; have_jac_ = want_jac; ok_ = true;
  return true;
}

void GpuBlockEvaluator::PrepareForEvaluation(bool evaluate_jacobians, bool new_evaluation_point) {
  std::lock_guard<std::mutex> lock(mu_);
  if (!new_evaluation_point && valid_ && (have_jac_ || !evaluate_jacobians) &&
      memcmp(snapshot_.data(), pose_array_, sizeof(double) * 3 * n_poses_) == 0) return;   // same point, already staged
  run_batch(evaluate_jacobians);
}

bool GpuBlockEvaluator::matches(int pose, const double* x) const {
  return pose < 0 || !x || memcmp(&snapshot_[3 * (size_t)pose], x, 3 * sizeof(double)) == 0;
}

bool GpuBlockEvaluator::Fetch(Kind kind, uint64_t block, int pose0, const double* x0, int pose1, const double* x1, int nres, double* residuals,
                              double* jac0, double* jac1) {
  const bool want_jac = jac0 || jac1;
  if (!valid_ || (want_jac && !have_jac_) || !matches(pose0, x0) || !matches(pose1, x1)) {
    // No callback ran for this point (a solver without EvaluationCallback support, or a direct
    // Evaluate call): the caller's pose array is the point — batch it now, once, for all blocks.
    std::lock_guard<std::mutex> lock(mu_);
    if (!valid_ || (want_jac && !have_jac_) || !matches(pose0, x0) || !matches(pose1, x1)) {
      if (!run_batch(true)) return false;
      if (!matches(pose0, x0) || !matches(pose1, x1)) {
        ok_ = false; error_ = "GPU cost function evaluated at a point that is not in the bound pose array (solver without EvaluationCallback?)";
        return false;
      }
    }
  }
  if (!ok_) return false;
  const double* r = r_ + r_off_[kind] + block * kResPerBlock[kind];
  for (int q = 0; q < nres; ++q) residuals[q] = r[q];
  if (want_jac) {
    const double* J = J_ + j_off_[kind] + block * kJacPerBlock[kind];
    const int half = kJacPerBlock[kind] / 2;   // binary kinds: [rows x 3 wrt pose0 | rows x 3 wrt pose1]
    if (jac0) memcpy(jac0, J, sizeof(double) * 3 * nres);
    if (jac1) memcpy(jac1, J + half, sizeof(double) * 3 * nres);
  }
  return true;
}

}  // namespace hitl
