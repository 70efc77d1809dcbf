// Stand-in for <ceres/ceres.h>, ONLY so that the reference's own JointOptimization.cpp / EMinput.cpp can be compiled
// where they lie (oracle/Makefile -> oracle/_ref/libhitl_ref.so).  Ceres Solver is absent from this image and the
// reference pins no version, so this header restates the slice of the 1.x API those files use:
//   CostFunction / SizedCostFunction / AutoDiffCostFunction (Jet seeding per SURVEY.md Appendix C), Problem
//   (residual blocks are RECORDED so the parity tests can evaluate the blocks the reference's own
//   AddOdometryConstraints / AddHumanConstraints / AddSTFConstraints built), Problem::Evaluate, Solver::Options /
//   Summary and a dense Levenberg-Marquardt Solve following the documented trust-region loop and defaults.
// What this pins is the REFERENCE'S code around the library (loops, constants, block construction); the library
// arithmetic itself stays "restated from the published algorithm".  Test infrastructure; not Ceres.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <sstream>
#include <string>
#include <vector>
#include <glog/logging.h>
#include "jet.h"

namespace ceres {

enum LinearSolverType { DENSE_NORMAL_CHOLESKY, DENSE_QR, SPARSE_NORMAL_CHOLESKY, DENSE_SCHUR, SPARSE_SCHUR, ITERATIVE_SCHUR, CGNR };
enum TrustRegionStrategyType { LEVENBERG_MARQUARDT, DOGLEG };
enum MinimizerType { LINE_SEARCH, TRUST_REGION };
enum TerminationType { CONVERGENCE, NO_CONVERGENCE, FAILURE, USER_SUCCESS, USER_FAILURE };
enum CallbackReturnType { SOLVER_CONTINUE, SOLVER_ABORT, SOLVER_TERMINATE_SUCCESSFULLY };
enum Ownership { DO_NOT_TAKE_OWNERSHIP, TAKE_OWNERSHIP };

struct IterationSummary { int iteration; double cost; double cost_change; double gradient_max_norm; double step_norm; double trust_region_radius; };
class IterationCallback { public: virtual ~IterationCallback() {} virtual CallbackReturnType operator()(const IterationSummary&) = 0; };
class LossFunction { public: virtual ~LossFunction() {} };

struct CRSMatrix { int num_rows, num_cols; std::vector<int> cols, rows; std::vector<double> values; CRSMatrix() : num_rows(0), num_cols(0) {} };

class CostFunction {
 public:
  virtual ~CostFunction() {}
  virtual bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const = 0;
  const std::vector<int>& parameter_block_sizes() const { return sizes_; }
  int num_residuals() const { return num_residuals_; }
 protected:
  std::vector<int> sizes_;
  int num_residuals_ = 0;
};

template <int kRes, int N0, int N1 = 0>
class SizedCostFunction : public CostFunction {
 public:
  SizedCostFunction() { num_residuals_ = kRes; sizes_.push_back(N0); if (N1) sizes_.push_back(N1); }
};

// AutoDiffCostFunction<F, kRes, N0[, N1]>: doubles when no Jacobian is asked for, otherwise Jet<double, N0 + N1> seeded with unit
// infinitesimals per parameter; jacobians[b][r * Nb + c].
template <typename F, int kRes, int N0, int N1 = 0>
class AutoDiffCostFunction : public SizedCostFunction<kRes, N0, N1> {
 public:
  explicit AutoDiffCostFunction(F* f) : f_(f) {}
  ~AutoDiffCostFunction() override { delete f_; }
  const F& functor() const { return *f_; }
  bool Evaluate(double const* const* p, double* r, double** J) const override {
    if (!J) return Call(*f_, p, r);
    typedef Jet<double, N0 + N1> JetT;
    JetT x[N0 + N1 + 1], y[kRes];
    for (int i = 0; i < N0; ++i) x[i] = JetT(p[0][i], i);
    for (int i = 0; i < N1; ++i) x[N0 + i] = JetT(p[1][i], N0 + i);
    const JetT* xp[2] = {x, x + N0};
    if (!CallJet(*f_, xp, y)) return false;
    for (int k = 0; k < kRes; ++k) r[k] = y[k].a;
    if (J[0]) for (int k = 0; k < kRes; ++k) for (int c = 0; c < N0; ++c) J[0][k * N0 + c] = y[k].v[c];
    if (N1 && J[1]) for (int k = 0; k < kRes; ++k) for (int c = 0; c < N1; ++c) J[1][k * N1 + c] = y[k].v[N0 + c];
    return true;
  }
 private:
  template <int M = N1> static typename std::enable_if<M == 0, bool>::type Call(const F& f, double const* const* p, double* r) { return f(p[0], r); }
  template <int M = N1> static typename std::enable_if<M != 0, bool>::type Call(const F& f, double const* const* p, double* r) { return f(p[0], p[1], r); }
  template <typename JetT, int M = N1> static typename std::enable_if<M == 0, bool>::type CallJet(const F& f, const JetT* const* p, JetT* r) { return f(p[0], r); }
  template <typename JetT, int M = N1> static typename std::enable_if<M != 0, bool>::type CallJet(const F& f, const JetT* const* p, JetT* r) { return f(p[0], p[1], r); }
  F* f_;
};
template <typename F, int Stride = 4> class DynamicAutoDiffCostFunction;

class Problem {
 public:
  struct Block { CostFunction* cost; std::vector<double*> params; };
  struct EvaluateOptions { std::vector<double*> parameter_blocks; int num_threads; EvaluateOptions() : num_threads(1) {} };
  Problem() {}
  ~Problem() { for (size_t i = 0; i < blocks_.size(); ++i) delete blocks_[i].cost; }
  void AddResidualBlock(CostFunction* c, LossFunction*, double* x0) { Block b; b.cost = c; b.params.push_back(x0); Note(x0, c->parameter_block_sizes()[0]); blocks_.push_back(b); Record(b); }
  void AddResidualBlock(CostFunction* c, LossFunction*, double* x0, double* x1) {
    Block b; b.cost = c; b.params.push_back(x0); b.params.push_back(x1);
    Note(x0, c->parameter_block_sizes()[0]); Note(x1, c->parameter_block_sizes()[1]); blocks_.push_back(b); Record(b);
  }
  void SetParameterBlockConstant(double* x) { constant_[x] = true; }
  int NumResiduals() const { int n = 0; for (size_t i = 0; i < blocks_.size(); ++i) n += blocks_[i].cost->num_residuals(); return n; }
  int NumResidualBlocks() const { return (int)blocks_.size(); }
  const std::vector<Block>& blocks() const { return blocks_; }
  const std::vector<double*>& parameter_blocks() const { return order_; }
  int block_size(double* x) const { return size_.find(x)->second; }
  bool is_constant(double* x) const { return constant_.count(x) != 0; }
  // residuals / gradient / Jacobian over ALL parameter blocks in insertion order (constant blocks keep zero columns)
  bool Evaluate(const EvaluateOptions&, double* cost, std::vector<double>* residuals, std::vector<double>* gradient, CRSMatrix* jac) {
    std::map<double*, int> col; int ncol = 0;
    for (size_t i = 0; i < order_.size(); ++i) { col[order_[i]] = ncol; ncol += size_[order_[i]]; }
    if (residuals) residuals->clear();
    if (gradient) gradient->assign(ncol, 0.0);
    if (jac) { jac->num_rows = NumResiduals(); jac->num_cols = ncol; jac->cols.clear(); jac->values.clear(); jac->rows.assign(1, 0); }
    double c = 0.0;
    for (size_t b = 0; b < blocks_.size(); ++b) {
      const Block& B = blocks_[b];
      const int nr = B.cost->num_residuals();
      double r[8]; double Jb[2][8 * 8]; double* Jp[2] = {Jb[0], Jb[1]};
      if (!B.cost->Evaluate(B.params.data(), r, Jp)) return false;
      for (int k = 0; k < nr; ++k) {
        c += 0.5 * r[k] * r[k];
        if (residuals) residuals->push_back(r[k]);
        for (size_t q = 0; q < B.params.size(); ++q) {
          const int nb = size_[B.params[q]], c0 = col[B.params[q]];
          const bool fixed = is_constant(B.params[q]);
          for (int e = 0; e < nb; ++e) {
            const double v = fixed ? 0.0 : Jb[q][k * nb + e];
            if (gradient) (*gradient)[c0 + e] += v * r[k];
            if (jac) { jac->cols.push_back(c0 + e); jac->values.push_back(v); }
          }
        }
        if (jac) jac->rows.push_back((int)jac->cols.size());
      }
    }
    if (cost) *cost = c;
    return true;
  }
  // Every Problem the reference code builds is also mirrored here so that a test harness can reach the blocks of a
  // Problem that lives on the reference's stack (SolveHumanConstraints / PostHumanOptimization).
  static std::vector<Block>*& recorder() { static std::vector<Block>* r = nullptr; return r; }
 private:
  void Note(double* x, int n) { if (!size_.count(x)) { size_[x] = n; order_.push_back(x); } }
  void Record(const Block& b) { if (recorder()) recorder()->push_back(b); }
  std::vector<Block> blocks_;
  std::vector<double*> order_;
  std::map<double*, int> size_;
  std::map<double*, bool> constant_;
};

class Solver {
 public:
  struct Options {
    MinimizerType minimizer_type; TrustRegionStrategyType trust_region_strategy_type; LinearSolverType linear_solver_type;
    int max_num_iterations; bool minimizer_progress_to_stdout; double function_tolerance, gradient_tolerance, parameter_tolerance;
    double initial_trust_region_radius, max_trust_region_radius, min_trust_region_radius, min_relative_decrease, min_lm_diagonal, max_lm_diagonal;
    bool update_state_every_iteration, use_nonmonotonic_steps, jacobi_scaling; int num_threads, num_linear_solver_threads;
    std::vector<IterationCallback*> callbacks;
    Options() : minimizer_type(TRUST_REGION), trust_region_strategy_type(LEVENBERG_MARQUARDT), linear_solver_type(SPARSE_NORMAL_CHOLESKY), max_num_iterations(50),
                minimizer_progress_to_stdout(false), function_tolerance(1e-6), gradient_tolerance(1e-10), parameter_tolerance(1e-8), initial_trust_region_radius(1e4),
                max_trust_region_radius(1e16), min_trust_region_radius(1e-32), min_relative_decrease(1e-3), min_lm_diagonal(1e-6), max_lm_diagonal(1e32),
                update_state_every_iteration(false), use_nonmonotonic_steps(false), jacobi_scaling(true), num_threads(1), num_linear_solver_threads(1) {}
  };
  struct Summary {
    TerminationType termination_type; double initial_cost, final_cost; int num_successful_steps, num_unsuccessful_steps, iterations;
    Summary() : termination_type(NO_CONVERGENCE), initial_cost(0), final_cost(0), num_successful_steps(0), num_unsuccessful_steps(0), iterations(0) {}
    std::string BriefReport() const { std::ostringstream s; s << "stand-in LM: iterations " << iterations << ", initial cost " << initial_cost << ", final cost " << final_cost; return s.str(); }
    std::string FullReport() const { return BriefReport(); }
  };
};

namespace shim_detail {
// Dense symmetric positive definite solve (Cholesky, in place); false when not positive definite.
inline bool cholesky_solve(std::vector<double>& A, std::vector<double>& b, int n) {
  for (int j = 0; j < n; ++j) {
    double d = A[(size_t)j * n + j];
    for (int k = 0; k < j; ++k) d -= A[(size_t)j * n + k] * A[(size_t)j * n + k];
    if (!(d > 0.0)) return false;
    d = std::sqrt(d); A[(size_t)j * n + j] = d;
    for (int i = j + 1; i < n; ++i) {
      double s = A[(size_t)i * n + j];
      for (int k = 0; k < j; ++k) s -= A[(size_t)i * n + k] * A[(size_t)j * n + k];
      A[(size_t)i * n + j] = s / d;
    }
  }
  for (int i = 0; i < n; ++i) { double s = b[i]; for (int k = 0; k < i; ++k) s -= A[(size_t)i * n + k] * b[k]; b[i] = s / A[(size_t)i * n + i]; }
  for (int i = n - 1; i >= 0; --i) { double s = b[i]; for (int k = i + 1; k < n; ++k) s -= A[(size_t)k * n + i] * b[k]; b[i] = s / A[(size_t)i * n + i]; }
  return true;
}
}  // namespace shim_detail

// Levenberg-Marquardt trust-region loop as documented for Ceres 1.x (dense normal equations; SURVEY.md Appendix C defaults).
inline void Solve(const Solver::Options& o_in, Problem* problem, Solver::Summary* summary) {
  using std::vector;
  Solver::Options o = o_in;
  // Test switch: run every solve of the reference code to full convergence (the reference hard-codes its options), so that a
  // parity test can compare MINIMISERS instead of two early-stopped iterates.
  if (std::getenv("HITL_SHIM_LM_TIGHT")) { o.max_num_iterations = 2000; o.function_tolerance = 1e-16; o.gradient_tolerance = 1e-14; o.parameter_tolerance = 1e-14; }
  Solver::Summary S;
  vector<double*> free_blocks; std::map<double*, int> col; int n = 0;
  const vector<double*>& order = problem->parameter_blocks();
  for (size_t i = 0; i < order.size(); ++i) if (!problem->is_constant(order[i])) { col[order[i]] = n; n += problem->block_size(order[i]); free_blocks.push_back(order[i]); }
  const vector<Problem::Block>& blocks = problem->blocks();
  vector<double> H((size_t)n * n), g(n), x(n), x_new(n), dx(n), Hs, gs;
  auto gather = [&](vector<double>& v) { for (size_t i = 0; i < free_blocks.size(); ++i) std::memcpy(&v[col[free_blocks[i]]], free_blocks[i], sizeof(double) * problem->block_size(free_blocks[i])); };
  auto scatter = [&](const vector<double>& v) { for (size_t i = 0; i < free_blocks.size(); ++i) std::memcpy(free_blocks[i], &v[col[free_blocks[i]]], sizeof(double) * problem->block_size(free_blocks[i])); };
  auto evaluate = [&](bool want_normal, double* cost) -> bool {
    double c = 0.0;
    if (want_normal) { std::fill(H.begin(), H.end(), 0.0); std::fill(g.begin(), g.end(), 0.0); }
    for (size_t b = 0; b < blocks.size(); ++b) {
      const Problem::Block& B = blocks[b];
      const int nr = B.cost->num_residuals();
      double r[8]; double Jb[2][64]; double* Jp[2] = {Jb[0], Jb[1]};
      if (!B.cost->Evaluate(B.params.data(), r, want_normal ? Jp : nullptr)) return false;
      for (int k = 0; k < nr; ++k) c += 0.5 * r[k] * r[k];
      if (!want_normal) continue;
      for (size_t p = 0; p < B.params.size(); ++p) {
        if (problem->is_constant(B.params[p])) continue;
        const int np = problem->block_size(B.params[p]), cp = col[B.params[p]];
        for (int k = 0; k < nr; ++k) for (int a = 0; a < np; ++a) g[cp + a] += Jb[p][k * np + a] * r[k];
        for (size_t q = 0; q < B.params.size(); ++q) {
          if (problem->is_constant(B.params[q])) continue;
          const int nq = problem->block_size(B.params[q]), cq = col[B.params[q]];
          for (int k = 0; k < nr; ++k) for (int a = 0; a < np; ++a) for (int e = 0; e < nq; ++e) H[(size_t)(cp + a) * n + cq + e] += Jb[p][k * np + a] * Jb[q][k * nq + e];
        }
      }
    }
    *cost = c;
    return true;
  };
  double cost = 0.0;
  if (!evaluate(true, &cost)) { S.termination_type = FAILURE; if (summary) *summary = S; return; }
  S.initial_cost = S.final_cost = cost;
  gather(x);
  double radius = o.initial_trust_region_radius, decrease_factor = 2.0;
  S.termination_type = NO_CONVERGENCE;
  auto max_abs = [&](const vector<double>& v) { double m = 0.0; for (size_t i = 0; i < v.size(); ++i) m = std::max(m, std::fabs(v[i])); return m; };
  if (n == 0 || max_abs(g) <= o.gradient_tolerance) S.termination_type = CONVERGENCE;
  for (int it = 0; S.termination_type == NO_CONVERGENCE && it < o.max_num_iterations; ++it) {
    S.iterations = it + 1;
    Hs = H; gs = g;
    for (int i = 0; i < n; ++i) {
      const double d = std::min(std::max(H[(size_t)i * n + i], o.min_lm_diagonal), o.max_lm_diagonal);
      Hs[(size_t)i * n + i] += d / radius;
      gs[i] = -g[i];
    }
    bool ok = shim_detail::cholesky_solve(Hs, gs, n);
    double model_change = 0.0;
    if (ok) {
      dx = gs;
      // model cost change = -dx^T (g + 0.5 H dx)
      for (int i = 0; i < n; ++i) { double hd = 0.0; for (int j = 0; j < n; ++j) hd += H[(size_t)i * n + j] * dx[j]; model_change -= dx[i] * (g[i] + 0.5 * hd); }
      ok = model_change > 0.0;
    }
    double rho = -1.0, new_cost = cost, step_norm = 0.0, x_norm = 0.0;
    if (ok) {
      for (int i = 0; i < n; ++i) { x_new[i] = x[i] + dx[i]; step_norm += dx[i] * dx[i]; x_norm += x[i] * x[i]; }
      step_norm = std::sqrt(step_norm); x_norm = std::sqrt(x_norm);
      if (step_norm <= o.parameter_tolerance * (x_norm + o.parameter_tolerance)) { S.termination_type = CONVERGENCE; break; }
      scatter(x_new);
      if (evaluate(false, &new_cost)) rho = (cost - new_cost) / model_change;
    }
    if (rho > o.min_relative_decrease) {
      const double change = cost - new_cost;
      x = x_new; ++S.num_successful_steps;
      evaluate(true, &cost);
      radius = std::min(o.max_trust_region_radius, radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rho - 1.0, 3)));
      decrease_factor = 2.0;
      S.final_cost = cost;
      if (max_abs(g) <= o.gradient_tolerance) { S.termination_type = CONVERGENCE; break; }
      if (std::fabs(change) <= o.function_tolerance * cost) { S.termination_type = CONVERGENCE; break; }
    } else {
      scatter(x); ++S.num_unsuccessful_steps;
      radius /= decrease_factor; decrease_factor *= 2.0;
      if (radius < o.min_trust_region_radius) { S.termination_type = CONVERGENCE; break; }
    }
  }
  scatter(x);
  if (summary) *summary = S;
}

}  // namespace ceres
