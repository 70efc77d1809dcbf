// hitl_ceres.h — the slice of the Ceres Solver API the reference's hot path is written against.
//
// The reference builds its problems with ceres::Problem::AddResidualBlock(new AutoDiffCostFunction<...>,
// NULL, &pose_array_[3*i] [, &pose_array_[3*j]]) and solves them with ceres::Solve
// (human_in_the_loop_slam/JointOptimization.cpp:553-557, 817-824, 994-1049, 1093, 1208; EMinput.cpp:167-178).
// Ceres is an un-vendored, un-pinned system dependency of the reference and is absent from this
// image, so this header provides the same names, argument meaning and ownership rules:
//
//   * with -DHITL_USE_SYSTEM_CERES the namespace below is an alias of the real ::ceres and the
//     GPU-backed cost functions (gpu_cost_functions.h) plug into the real solver unchanged;
//   * otherwise it is a small self-contained implementation: CostFunction / SizedCostFunction /
//     AutoDiffCostFunction (Jets), Problem, EvaluationCallback, CRSMatrix, and a Levenberg-Marquardt
//     trust-region Solve that follows Ceres' documented loop and defaults.
//     It is the host-side stand-in for "the sparse linear solve stays in Ceres": dense Cholesky for
//     small problems, block-Jacobi preconditioned CG on the block-sparse normal equations otherwise.
#pragma once
#ifdef HITL_USE_SYSTEM_CERES
#include <ceres/ceres.h>
namespace hitl { namespace ceres = ::ceres; }
#else
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <memory>
#include <type_traits>
#include <string>
#include <unordered_map>
#include <vector>

namespace hitl {
namespace ceres {

enum TerminationType { CONVERGENCE, NO_CONVERGENCE, FAILURE, USER_SUCCESS, USER_FAILURE };
enum LinearSolverType { DENSE_NORMAL_CHOLESKY, DENSE_QR, SPARSE_NORMAL_CHOLESKY, DENSE_SCHUR, SPARSE_SCHUR, ITERATIVE_SCHUR, CGNR };
enum MinimizerType { LINE_SEARCH, TRUST_REGION };
enum TrustRegionStrategyType { LEVENBERG_MARQUARDT, DOGLEG };
enum Ownership { DO_NOT_TAKE_OWNERSHIP, TAKE_OWNERSHIP };

class LossFunction { public: virtual ~LossFunction() {} };   // the path passes NULL everywhere

// ---- cost functions --------------------------------------------------------------------------
class CostFunction {
 public:
  CostFunction() : sizes_(&parameter_block_sizes_), num_residuals_(0) {}
  virtual ~CostFunction() {}
  CostFunction(const CostFunction&) = delete;
  CostFunction& operator=(const CostFunction&) = delete;
  // jacobians[i] is row-major [num_residuals x parameter_block_sizes()[i]]; jacobians or any
  // jacobians[i] may be NULL.
  virtual bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const = 0;
  const std::vector<int32_t>& parameter_block_sizes() const { return *sizes_; }
  int num_residuals() const { return num_residuals_; }

 protected:
  std::vector<int32_t>* mutable_parameter_block_sizes() { sizes_ = &parameter_block_sizes_; return &parameter_block_sizes_; }
  void set_num_residuals(int n) { num_residuals_ = n; }
  // Sized cost functions share one size list per type: a problem of thousands of blocks (one object per residual block, as the
  // reference builds it) then costs one allocation per block, not two.
  void share_parameter_block_sizes(const std::vector<int32_t>* shared) { sizes_ = shared; }

 private:
  friend class Problem;
  std::vector<int32_t> parameter_block_sizes_;
  const std::vector<int32_t>* sizes_;
  int num_residuals_;
  int problem_uses_ = 0;   // residual blocks of the owning Problem that use this object: a shared cost function is deleted once
};

template <int kNumResiduals, int... Ns>
class SizedCostFunction : public CostFunction {
 public:
  SizedCostFunction() {
    static const std::vector<int32_t> kSizes{Ns...};
    set_num_residuals(kNumResiduals);
    share_parameter_block_sizes(&kSizes);
  }
};

// Forward-mode dual number, value + N partials (what ceres::Jet<double, N> is).
template <int N>
struct Jet {
  double a;
  double v[N];
  Jet() : a(0.0) { for (int i = 0; i < N; ++i) v[i] = 0.0; }
  Jet(double s) : a(s) { for (int i = 0; i < N; ++i) v[i] = 0.0; }   // NOLINT: implicit like ceres::Jet
  Jet(double s, int k) : a(s) { for (int i = 0; i < N; ++i) v[i] = 0.0; v[k] = 1.0; }
};
template <int N> inline Jet<N> operator+(const Jet<N>& f, const Jet<N>& g) { Jet<N> h; h.a = f.a + g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] + g.v[i]; return h; }
template <int N> inline Jet<N> operator-(const Jet<N>& f, const Jet<N>& g) { Jet<N> h; h.a = f.a - g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] - g.v[i]; return h; }
template <int N> inline Jet<N> operator-(const Jet<N>& f) { Jet<N> h; h.a = -f.a; for (int i = 0; i < N; ++i) h.v[i] = -f.v[i]; return h; }
template <int N> inline Jet<N> operator*(const Jet<N>& f, const Jet<N>& g) { Jet<N> h; h.a = f.a * g.a; for (int i = 0; i < N; ++i) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h; }
template <int N> inline Jet<N> operator/(const Jet<N>& f, const Jet<N>& g) {
  Jet<N> h; const double gi = 1.0 / g.a; h.a = f.a * gi;
  for (int i = 0; i < N; ++i) h.v[i] = (f.v[i] - h.a * g.v[i]) * gi;
  return h;
}
template <int N> inline Jet<N> operator+(const Jet<N>& f, double s) { Jet<N> h = f; h.a += s; return h; }
template <int N> inline Jet<N> operator+(double s, const Jet<N>& f) { Jet<N> h = f; h.a += s; return h; }
template <int N> inline Jet<N> operator-(const Jet<N>& f, double s) { Jet<N> h = f; h.a -= s; return h; }
template <int N> inline Jet<N> operator-(double s, const Jet<N>& f) { Jet<N> h = -f; h.a += s; return h; }
template <int N> inline Jet<N> operator*(const Jet<N>& f, double s) { Jet<N> h; h.a = f.a * s; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * s; return h; }
template <int N> inline Jet<N> operator*(double s, const Jet<N>& f) { return f * s; }
template <int N> inline Jet<N> operator/(const Jet<N>& f, double s) { return f * (1.0 / s); }
template <int N> inline bool operator<(const Jet<N>& f, const Jet<N>& g) { return f.a < g.a; }
template <int N> inline bool operator>(const Jet<N>& f, const Jet<N>& g) { return f.a > g.a; }
template <int N> inline bool operator<(const Jet<N>& f, double s) { return f.a < s; }
template <int N> inline bool operator>(const Jet<N>& f, double s) { return f.a > s; }
template <int N> inline Jet<N> sqrt(const Jet<N>& f) { Jet<N> h; h.a = ::sqrt(f.a); const double t = 1.0 / (2.0 * h.a); for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * t; return h; }
template <int N> inline Jet<N> sin(const Jet<N>& f) { Jet<N> h; h.a = ::sin(f.a); const double c = ::cos(f.a); for (int i = 0; i < N; ++i) h.v[i] = c * f.v[i]; return h; }
template <int N> inline Jet<N> cos(const Jet<N>& f) { Jet<N> h; h.a = ::cos(f.a); const double s = -::sin(f.a); for (int i = 0; i < N; ++i) h.v[i] = s * f.v[i]; return h; }
template <int N> inline Jet<N> pow(const Jet<N>& f, double p) { Jet<N> h; h.a = ::pow(f.a, p); const double t = p * ::pow(f.a, p - 1.0); for (int i = 0; i < N; ++i) h.v[i] = t * f.v[i]; return h; }
template <int N> inline Jet<N> atan2(const Jet<N>& g, const Jet<N>& f) {
  Jet<N> h; h.a = ::atan2(g.a, f.a); const double t = 1.0 / (f.a * f.a + g.a * g.a);
  for (int i = 0; i < N; ++i) h.v[i] = t * (f.a * g.v[i] - g.a * f.v[i]);
  return h;
}
using ::sqrt; using ::sin; using ::cos; using ::pow; using ::atan2;

// AutoDiffCostFunction<Functor, kNumResiduals, N0[, N1]>: takes ownership of the functor.
template <typename Functor, int kNumResiduals, int N0, int N1 = 0>
class AutoDiffCostFunction : public CostFunction {
 public:
  explicit AutoDiffCostFunction(Functor* f) : functor_(f) {
    set_num_residuals(kNumResiduals);
    mutable_parameter_block_sizes()->push_back(N0);
    if (N1) mutable_parameter_block_sizes()->push_back(N1);
  }
  bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const override {
    if (!jacobians) return call(parameters, residuals);
    typedef Jet<N0 + N1> J;
    J x0[N0], x1[N1 ? N1 : 1], r[kNumResiduals];
    for (int i = 0; i < N0; ++i) x0[i] = J(parameters[0][i], i);
    for (int i = 0; i < N1; ++i) x1[i] = J(parameters[1][i], N0 + i);
    if (!call_jet(x0, x1, r)) return false;
    for (int q = 0; q < kNumResiduals; ++q) {
      residuals[q] = r[q].a;
      if (jacobians[0]) for (int i = 0; i < N0; ++i) jacobians[0][q * N0 + i] = r[q].v[i];
      if (N1 && jacobians[1]) for (int i = 0; i < N1; ++i) jacobians[1][q * N1 + i] = r[q].v[N0 + i];
    }
    return true;
  }

 private:
  template <int M = N1> typename std::enable_if<M == 0, bool>::type call(double const* const* p, double* r) const { return (*functor_)(p[0], r); }
  template <int M = N1> typename std::enable_if<M != 0, bool>::type call(double const* const* p, double* r) const { return (*functor_)(p[0], p[1], r); }
  template <typename J, int M = N1> typename std::enable_if<M == 0, bool>::type call_jet(const J* x0, const J*, J* r) const { return (*functor_)(x0, r); }
  template <typename J, int M = N1> typename std::enable_if<M != 0, bool>::type call_jet(const J* x0, const J* x1, J* r) const { return (*functor_)(x0, x1, r); }
  std::unique_ptr<Functor> functor_;
};

// ---- evaluation callback -----------------------------------------------------------------------
// Called once before the residual blocks are evaluated at a point; the user's parameter blocks hold
// that point when it is called (Ceres >= 1.14 semantics).  This is where the GPU batch runs.
class EvaluationCallback {
 public:
  virtual ~EvaluationCallback() {}
  virtual void PrepareForEvaluation(bool evaluate_jacobians, bool new_evaluation_point) = 0;
};

struct CRSMatrix {
  CRSMatrix() : num_rows(0), num_cols(0) {}
  int num_rows, num_cols;
  std::vector<int> cols, rows;
  std::vector<double> values;
};

// ---- problem --------------------------------------------------------------------------------------
typedef int ResidualBlockId;

class Problem {
 public:
  struct Options {
    Options() : cost_function_ownership(TAKE_OWNERSHIP), evaluation_callback(nullptr) {}
    Ownership cost_function_ownership;
    EvaluationCallback* evaluation_callback;
  };
  struct EvaluateOptions {
    EvaluateOptions() : apply_loss_function(true), num_threads(1) {}
    std::vector<double*> parameter_blocks;   // empty = all blocks in the order they were added
    bool apply_loss_function;
    int num_threads;
  };
  Problem() {}
  explicit Problem(const Options& o) : options_(o) {}
  ~Problem();
  Problem(const Problem&) = delete;
  Problem& operator=(const Problem&) = delete;

  ResidualBlockId AddResidualBlock(CostFunction* cost, LossFunction* loss, double* x0);
  ResidualBlockId AddResidualBlock(CostFunction* cost, LossFunction* loss, double* x0, double* x1);
  ResidualBlockId AddResidualBlock(CostFunction* cost, LossFunction* loss, const std::vector<double*>& blocks);
  void AddParameterBlock(double* values, int size);
  void SetParameterBlockConstant(double* values);
  void SetParameterBlockVariable(double* values);
  int NumParameterBlocks() const { return (int)blocks_.size(); }
  int NumParameters() const;
  int NumResidualBlocks() const { return (int)residuals_.size(); }
  int NumResiduals() const { return num_residuals_; }
  // cost = 1/2 sum r^2; gradient and jacobian over the non-constant requested blocks.
  bool Evaluate(const EvaluateOptions& options, double* cost, std::vector<double>* residuals, std::vector<double>* gradient, CRSMatrix* jacobian);

  // -- used by Solve --
  struct ParameterBlock { double* values; int size; bool constant; };
  // parameter blocks of one residual block: the hot path has one or two (inline storage); a longer list keeps its tail in storage
  // the Problem owns, so that a ResidualBlock stays trivially copyable and the block vector grows by plain copies
  class BlockList {
   public:
    BlockList() : n_(0), more_(nullptr) {}
    size_t size() const { return n_; }
    int operator[](size_t i) const { return i < kInline ? inl_[i] : more_[i - kInline]; }
   private:
    friend class Problem;
    static const size_t kInline = 2;
    uint32_t n_;
    int inl_[kInline];
    int* more_;
  };
  struct ResidualBlock { CostFunction* cost; BlockList blocks; int residual_offset; };
  const std::vector<ParameterBlock>& parameter_blocks() const { return blocks_; }
  const std::vector<ResidualBlock>& residual_blocks() const { return residuals_; }
  const Options& options() const { return options_; }

 private:
  int block_index(double* values, int size);
  int find_block(double* values) const;   // -1 when the pointer is not a parameter block of this problem
  ResidualBlockId add_block(CostFunction* cost, double* const* blocks, size_t n);
  Options options_;
  std::vector<ParameterBlock> blocks_;
  // Pointer -> block index.  While the blocks arrive in ascending address order (the hot path walks one contiguous pose array:
  // JointOptimization.cpp:817-821, 994-1049) blocks_ itself is the index (append / bisection, no allocation per block); the first
  // out-of-order pointer switches to the hash map for the rest of the problem's life.
  bool ascending_ = true;
  std::unordered_map<double*, int> index_;
  std::vector<ResidualBlock> residuals_;
  std::vector<std::unique_ptr<int[]>> spill_;   // tails of the block lists longer than BlockList::kInline
  int num_residuals_ = 0;
};

// ---- solver ------------------------------------------------------------------------------------------
struct Solver {
  struct Options {
    Options()
        : minimizer_type(TRUST_REGION), trust_region_strategy_type(LEVENBERG_MARQUARDT), linear_solver_type(SPARSE_NORMAL_CHOLESKY),
          max_num_iterations(50), function_tolerance(1e-6), gradient_tolerance(1e-10), parameter_tolerance(1e-8),
          initial_trust_region_radius(1e4), max_trust_region_radius(1e16), min_trust_region_radius(1e-32), min_relative_decrease(1e-3),
          min_lm_diagonal(1e-6), max_lm_diagonal(1e32), jacobi_scaling(true), minimizer_progress_to_stdout(false),
          update_state_every_iteration(false), num_threads(1), num_linear_solver_threads(1), evaluation_callback(nullptr),
          dense_limit(2400), cg_max_iterations(2000), cg_tolerance(1e-12), chain_direct(true) {}
    MinimizerType minimizer_type;
    TrustRegionStrategyType trust_region_strategy_type;
    LinearSolverType linear_solver_type;
    int max_num_iterations;
    double function_tolerance, gradient_tolerance, parameter_tolerance;
    double initial_trust_region_radius, max_trust_region_radius, min_trust_region_radius, min_relative_decrease;
    double min_lm_diagonal, max_lm_diagonal;
    bool jacobi_scaling, minimizer_progress_to_stdout, update_state_every_iteration;
    int num_threads, num_linear_solver_threads;
    EvaluationCallback* evaluation_callback;   // Ceres 1.14 placement; Problem::Options is the 2.x placement
    // stand-in linear algebra (not Ceres options): dense Cholesky up to dense_limit parameters, PCG beyond
    int dense_limit, cg_max_iterations;
    bool chain_direct;   // beyond dense_limit: block-tridiagonal systems (odometry chain + unary factors) are eliminated directly
    double cg_tolerance;
  };
  struct Summary {
    Summary() : termination_type(FAILURE), initial_cost(0), final_cost(0), num_successful_steps(0), num_unsuccessful_steps(0), num_residual_evaluations(0),
                num_jacobian_evaluations(0) {}
    TerminationType termination_type;
    double initial_cost, final_cost;
    int num_successful_steps, num_unsuccessful_steps, num_residual_evaluations, num_jacobian_evaluations;
    std::string message;
    std::string BriefReport() const;
    std::string FullReport() const { return BriefReport(); }
  };
};

void Solve(const Solver::Options& options, Problem* problem, Solver::Summary* summary);

}  // namespace ceres
}  // namespace hitl
#endif  // HITL_USE_SYSTEM_CERES
