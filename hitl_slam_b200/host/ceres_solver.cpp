// ceres_solver.cpp — Problem bookkeeping and the Levenberg-Marquardt trust-region loop behind
// hitl::ceres::Solve (stand-in for the un-vendored Ceres Solver the reference links against;
// call sites: human_in_the_loop_slam/JointOptimization.cpp:1093, :1208; EMinput.cpp:178).
//
// The loop follows Ceres' documented trust-region minimizer with the LM strategy (SURVEY.md
// Appendix C): regularised normal equations (J^T J + D^T D / radius) dx = -J^T r with
// D^2 = diag(J^T J) clamped to [min_lm_diagonal, max_lm_diagonal], Jacobi column scaling,
// step acceptance on relative_decrease > min_relative_decrease, radius update
// radius / max(1/3, 1 - (2 rho - 1)^3), halving (2, 4, 8, ...) on rejection, and the three
// tolerances.  It is not claimed to reproduce Ceres' iterates bit for bit — final poses are
// compared at convergence, within tolerance.
#ifndef HITL_USE_SYSTEM_CERES
#include "hitl_ceres.h"

#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <functional>

namespace hitl {
namespace ceres {

// ---- Problem ---------------------------------------------------------------------------------------
Problem::~Problem() {
  if (options_.cost_function_ownership == TAKE_OWNERSHIP)
    for (const ResidualBlock& rb : residuals_)
      if (--rb.cost->problem_uses_ == 0) delete rb.cost;
}

int Problem::find_block(double* values) const {
  if (!ascending_) {
    auto it = index_.find(values);
    return it == index_.end() ? -1 : it->second;
  }
  if (blocks_.empty()) return -1;
  if (blocks_.back().values == values) return (int)blocks_.size() - 1;     // a chain's block i starts where block i-1 ended
  const std::less<double*> before;
  auto it = std::lower_bound(blocks_.begin(), blocks_.end(), values, [&](const ParameterBlock& b, double* v) { return before(b.values, v); });
  return (it != blocks_.end() && it->values == values) ? (int)(it - blocks_.begin()) : -1;
}

int Problem::block_index(double* values, int size) {
  const int found = find_block(values);
  if (found >= 0) return found;
  if (ascending_ && !blocks_.empty() && !std::less<double*>()(blocks_.back().values, values)) {
    ascending_ = false;
    index_.reserve(2 * blocks_.size() + 16);
    for (size_t b = 0; b < blocks_.size(); ++b) index_[blocks_[b].values] = (int)b;
  }
  ParameterBlock b; b.values = values; b.size = size; b.constant = false;
  blocks_.push_back(b);
  if (!ascending_) index_[values] = (int)blocks_.size() - 1;
  return (int)blocks_.size() - 1;
}

void Problem::AddParameterBlock(double* values, int size) { block_index(values, size); }

ResidualBlockId Problem::add_block(CostFunction* cost, double* const* blocks, size_t n) {
  // One generous reservation instead of growth by doubling: the pose-chain problems of a correction have a few thousand blocks, and
  // re-growing (allocate, copy, free of ever larger arrays) cost more than the blocks themselves.  Untouched pages cost nothing.
  if (residuals_.empty()) { residuals_.reserve(8192); blocks_.reserve(8192); }
  residuals_.emplace_back();
  ResidualBlock& rb = residuals_.back();
  rb.cost = cost; rb.residual_offset = num_residuals_;
  const std::vector<int32_t>& sizes = cost->parameter_block_sizes();
  rb.blocks.n_ = (uint32_t)n;
  if (n > Problem::BlockList::kInline) { spill_.emplace_back(new int[n - Problem::BlockList::kInline]); rb.blocks.more_ = spill_.back().get(); }
  for (size_t i = 0; i < n; ++i) {
    const int b = block_index(blocks[i], sizes[i]);
    if (i < Problem::BlockList::kInline) rb.blocks.inl_[i] = b; else rb.blocks.more_[i - Problem::BlockList::kInline] = b;
  }
  num_residuals_ += cost->num_residuals();
  ++cost->problem_uses_;
  return (ResidualBlockId)residuals_.size() - 1;
}
ResidualBlockId Problem::AddResidualBlock(CostFunction* cost, LossFunction* loss, const std::vector<double*>& blocks) {
  (void)loss;   // the hot path passes NULL (JointOptimization.cpp:555, 820, 999, 1019, 1034, 1048)
  return add_block(cost, blocks.data(), blocks.size());
}
ResidualBlockId Problem::AddResidualBlock(CostFunction* cost, LossFunction* loss, double* x0) {
  (void)loss;
  double* b[1] = {x0};
  return add_block(cost, b, 1);
}
ResidualBlockId Problem::AddResidualBlock(CostFunction* cost, LossFunction* loss, double* x0, double* x1) {
  (void)loss;
  double* b[2] = {x0, x1};
  return add_block(cost, b, 2);
}
void Problem::SetParameterBlockConstant(double* values) {
  const int b = find_block(values);
  if (b >= 0) blocks_[b].constant = true;
}
void Problem::SetParameterBlockVariable(double* values) {
  const int b = find_block(values);
  if (b >= 0) blocks_[b].constant = false;
}
int Problem::NumParameters() const {
  int n = 0;
  for (const ParameterBlock& b : blocks_) n += b.size;
  return n;
}

namespace {

// One evaluation of every residual block at the point held by the user's parameter blocks.
// Jacobian blocks are kept per residual block: for block slot s of residual block b,
// jac[jac_off[b][s] ...] is row-major [num_residuals x size] (absent for constant blocks).
struct Linearization {
  std::vector<double> r;
  std::vector<double> jac;
  std::vector<size_t> jac_off;   // per (residual block, slot), SIZE_MAX when not requested
  std::vector<size_t> slot0;     // first slot of each residual block
};

bool evaluate_blocks(Problem* p, EvaluationCallback* cb, bool want_jac, bool new_point, Linearization* L) {
  const auto& rbs = p->residual_blocks();
  const auto& pbs = p->parameter_blocks();
  if (cb) cb->PrepareForEvaluation(want_jac, new_point);
  L->r.assign(p->NumResiduals(), 0.0);
  if (want_jac) {
    L->slot0.resize(rbs.size() + 1);
    size_t slots = 0, total = 0;
    for (size_t b = 0; b < rbs.size(); ++b) { L->slot0[b] = slots; slots += rbs[b].blocks.size(); }
    L->slot0[rbs.size()] = slots;
    L->jac_off.assign(slots, (size_t)-1);
    for (size_t b = 0; b < rbs.size(); ++b)
      for (size_t s = 0; s < rbs[b].blocks.size(); ++s) {
        const Problem::ParameterBlock& pb = pbs[rbs[b].blocks[s]];
        if (pb.constant) continue;
        L->jac_off[L->slot0[b] + s] = total;
        total += (size_t)rbs[b].cost->num_residuals() * pb.size;
      }
    L->jac.assign(total, 0.0);
  }
  std::vector<const double*> params;
  std::vector<double*> jp;
  for (size_t b = 0; b < rbs.size(); ++b) {
    const Problem::ResidualBlock& rb = rbs[b];
    params.resize(rb.blocks.size()); jp.resize(rb.blocks.size());
    for (size_t s = 0; s < rb.blocks.size(); ++s) {
      params[s] = pbs[rb.blocks[s]].values;
      jp[s] = (want_jac && L->jac_off[L->slot0[b] + s] != (size_t)-1) ? &L->jac[L->jac_off[L->slot0[b] + s]] : nullptr;
    }
    if (!rb.cost->Evaluate(params.data(), &L->r[rb.residual_offset], want_jac ? jp.data() : nullptr)) return false;
  }
  return true;
}

double half_sq_norm(const std::vector<double>& r) {
  double s = 0;
  for (double v : r) s += v * v;
  return 0.5 * s;
}

// Symmetric block-sparse matrix over the variable parameter blocks: dense diagonal blocks and
// upper off-diagonal blocks (a < b) keyed by the block pair.
struct BlockSparse {
  std::vector<int> off, size;            // scalar offset / size per variable block
  int n = 0;                             // scalar dimension
  std::vector<std::vector<double>> diag; // size x size each
  std::unordered_map<uint64_t, int> pair_index;
  std::vector<std::pair<int, int>> pairs;
  std::vector<std::vector<double>> offd; // size[a] x size[b], row-major
  int pair(int a, int b) {
    const uint64_t key = ((uint64_t)(uint32_t)a << 32) | (uint32_t)b;
    auto it = pair_index.find(key);
    if (it != pair_index.end()) return it->second;
    pairs.push_back(std::make_pair(a, b));
    offd.push_back(std::vector<double>((size_t)size[a] * size[b], 0.0));
    pair_index[key] = (int)pairs.size() - 1;
    return (int)pairs.size() - 1;
  }
  void multiply(const std::vector<double>& x, std::vector<double>* y) const {
    y->assign(n, 0.0);
    for (size_t a = 0; a < diag.size(); ++a) {
      const int o = off[a], s = size[a];
      for (int i = 0; i < s; ++i) { double t = 0; for (int j = 0; j < s; ++j) t += diag[a][i * s + j] * x[o + j]; (*y)[o + i] += t; }
    }
    for (size_t p = 0; p < pairs.size(); ++p) {
      const int a = pairs[p].first, b = pairs[p].second, oa = off[a], ob = off[b], sa = size[a], sb = size[b];
      const std::vector<double>& m = offd[p];
      for (int i = 0; i < sa; ++i) { double t = 0; for (int j = 0; j < sb; ++j) t += m[i * sb + j] * x[ob + j]; (*y)[oa + i] += t; }
      for (int j = 0; j < sb; ++j) { double t = 0; for (int i = 0; i < sa; ++i) t += m[i * sb + j] * x[oa + i]; (*y)[ob + j] += t; }
    }
  }
};

bool dense_cholesky_solve(std::vector<double>& A, int n, std::vector<double>& b) {
  for (int j = 0; j < n; ++j) {
    double d = A[(size_t)j * n + j];
    for (int k = 0; k < j; ++k) d -= A[(size_t)j * n + k] * A[(size_t)j * n + k];
    if (!(d > 0.0)) return false;
    d = ::sqrt(d);
    A[(size_t)j * n + j] = d;
    for (int i = j + 1; i < n; ++i) {
      double s = A[(size_t)i * n + j];
      const double* ri = &A[(size_t)i * n]; const double* rj = &A[(size_t)j * n];
      for (int k = 0; k < j; ++k) s -= ri[k] * rj[k];
      A[(size_t)i * n + j] = s / d;
    }
  }
  for (int i = 0; i < n; ++i) { double s = b[i]; for (int k = 0; k < i; ++k) s -= A[(size_t)i * n + k] * b[k]; b[i] = s / A[(size_t)i * n + i]; }
  for (int i = n - 1; i >= 0; --i) { double s = b[i]; for (int k = i + 1; k < n; ++k) s -= A[(size_t)k * n + i] * b[k]; b[i] = s / A[(size_t)i * n + i]; }
  return true;
}

// Inverse of a small SPD block (Cholesky based), for the block-Jacobi preconditioner.
bool small_inverse(const std::vector<double>& A, int n, std::vector<double>* inv) {
  inv->assign((size_t)n * n, 0.0);
  for (int c = 0; c < n; ++c) {
    std::vector<double> M = A, e(n, 0.0);
    e[c] = 1.0;
    if (!dense_cholesky_solve(M, n, e)) return false;
    for (int r = 0; r < n; ++r) (*inv)[(size_t)r * n + c] = e[r];
  }
  return true;
}

bool pcg_solve(const BlockSparse& H, const std::vector<double>& lm_diag2, const std::vector<double>& rhs, int max_it, double tol, std::vector<double>* x) {
  const int n = H.n;
  x->assign(n, 0.0);
  std::vector<double> r = rhs, z(n), p(n), Ap(n);
  // block-Jacobi: (D_a + lm_a)^-1 per parameter block, factored once per linear solve
  std::vector<std::vector<double>> Dinv(H.diag.size());
  for (size_t a = 0; a < H.diag.size(); ++a) {
    const int s = H.size[a], o = H.off[a];
    std::vector<double> D = H.diag[a];
    for (int i = 0; i < s; ++i) D[i * s + i] += lm_diag2[o + i];
    if (!small_inverse(D, s, &Dinv[a])) return false;
  }
  auto precond = [&](const std::vector<double>& v, std::vector<double>* out) -> bool {
    for (size_t a = 0; a < H.diag.size(); ++a) {
      const int s = H.size[a], o = H.off[a];
      const double* M = Dinv[a].data();
      for (int i = 0; i < s; ++i) { double t = 0; for (int j = 0; j < s; ++j) t += M[i * s + j] * v[o + j]; (*out)[o + i] = t; }
    }
    return true;
  };
  if (!precond(r, &z)) return false;
  p = z;
  double rz = 0, r0 = 0;
  for (int i = 0; i < n; ++i) { rz += r[i] * z[i]; r0 += r[i] * r[i]; }
  if (r0 == 0.0) return true;
  for (int it = 0; it < max_it; ++it) {
    H.multiply(p, &Ap);
    double pAp = 0;
    for (int i = 0; i < n; ++i) { Ap[i] += lm_diag2[i] * p[i]; pAp += p[i] * Ap[i]; }
    if (!(pAp > 0.0)) return false;
    const double alpha = rz / pAp;
    double rr = 0;
    for (int i = 0; i < n; ++i) { (*x)[i] += alpha * p[i]; r[i] -= alpha * Ap[i]; rr += r[i] * r[i]; }
    if (rr <= tol * tol * r0) return true;
    if (!precond(r, &z)) return false;
    double rz_new = 0;
    for (int i = 0; i < n; ++i) rz_new += r[i] * z[i];
    const double beta = rz_new / rz;
    rz = rz_new;
    for (int i = 0; i < n; ++i) p[i] = z[i] + beta * p[i];
  }
  return true;   // best effort: the trust-region ratio test guards a poor step
}

// Direct solve of (H + diag(lm)) x = rhs when H is block TRIDIAGONAL in the order of the variable blocks — the shape of the
// human-constraint problem (odometry chain + unary factors, JointOptimization.cpp:1064-1138): block Thomas elimination,
// O(#blocks), exact up to rounding.  Returns false when H has any other coupling (caller falls back to PCG) or a pivot block
// is not positive definite.
bool tridiagonal_solve(const BlockSparse& H, const std::vector<double>& lm_diag2, const std::vector<double>& rhs, std::vector<double>* x, bool* applicable) {
  const size_t m = H.diag.size();
  std::vector<int> next(m, -1);
  *applicable = true;
  for (size_t p = 0; p < H.pairs.size(); ++p) {
    const int a = H.pairs[p].first, b = H.pairs[p].second;
    if (b != a + 1 || next[(size_t)a] >= 0) { *applicable = false; return false; }
    next[(size_t)a] = (int)p;
  }
  std::vector<std::vector<double>> Sinv(m);
  std::vector<double> y = rhs, S, W;
  for (size_t a = 0; a < m; ++a) {
    const int s = H.size[a], o = H.off[a];
    S = H.diag[a];
    for (int i = 0; i < s; ++i) S[(size_t)i * s + i] += lm_diag2[o + i];
    if (a > 0 && next[a - 1] >= 0) {
      const int sp = H.size[a - 1], op = H.off[a - 1];
      const std::vector<double>& B = H.offd[(size_t)next[a - 1]];          // sp x s
      const std::vector<double>& Pinv = Sinv[a - 1];                       // sp x sp
      W.assign((size_t)s * sp, 0.0);                                       // W = B^T * Sinv_prev   (s x sp)
      for (int i = 0; i < s; ++i) for (int j = 0; j < sp; ++j) { double t = 0; for (int k = 0; k < sp; ++k) t += B[(size_t)k * s + i] * Pinv[(size_t)k * sp + j]; W[(size_t)i * sp + j] = t; }
      for (int i = 0; i < s; ++i) {
        for (int j = 0; j < s; ++j) { double t = 0; for (int k = 0; k < sp; ++k) t += W[(size_t)i * sp + k] * B[(size_t)k * s + j]; S[(size_t)i * s + j] -= t; }
        double t = 0; for (int k = 0; k < sp; ++k) t += W[(size_t)i * sp + k] * y[op + k];
        y[o + i] -= t;
      }
    }
    if (!small_inverse(S, s, &Sinv[a])) return false;
  }
  x->assign(H.n, 0.0);
  for (size_t a = m; a-- > 0;) {
    const int s = H.size[a], o = H.off[a];
    std::vector<double> t(y.begin() + o, y.begin() + o + s);
    if (a + 1 < m && next[a] >= 0) {
      const int sn = H.size[a + 1], on = H.off[a + 1];
      const std::vector<double>& B = H.offd[(size_t)next[a]];              // s x sn
      for (int i = 0; i < s; ++i) { double u = 0; for (int k = 0; k < sn; ++k) u += B[(size_t)i * sn + k] * (*x)[on + k]; t[i] -= u; }
    }
    for (int i = 0; i < s; ++i) { double u = 0; for (int k = 0; k < s; ++k) u += Sinv[a][(size_t)i * s + k] * t[k]; (*x)[o + i] = u; }
  }
  return true;
}

}  // namespace

bool Problem::Evaluate(const EvaluateOptions& options, double* cost, std::vector<double>* residuals, std::vector<double>* gradient, CRSMatrix* jacobian) {
  Linearization L;
  const bool want_jac = gradient || jacobian;
  if (!evaluate_blocks(this, options_.evaluation_callback, want_jac, true, &L)) return false;
  if (cost) *cost = half_sq_norm(L.r);
  if (residuals) *residuals = L.r;
  if (!want_jac) return true;
  // column layout: requested blocks (default: all, in insertion order).  Constant blocks keep their columns, as in Ceres
  // (ProblemEvaluateTest.ConstantParameterBlock): their gradient entries and Jacobian columns stay zero, so a consumer that indexes
  // by 3 * pose sees the same offsets whether or not pose 0 is held constant.
  std::vector<int> order;
  if (options.parameter_blocks.empty()) { for (int i = 0; i < (int)blocks_.size(); ++i) order.push_back(i); }
  else for (double* v : options.parameter_blocks) { const int b = find_block(v); if (b < 0) return false; order.push_back(b); }
  std::vector<int> col(blocks_.size(), -1);
  int n = 0;
  for (int b : order) { col[b] = n; n += blocks_[b].size; }
  if (gradient) gradient->assign(n, 0.0);
  struct Entry { int c; double v; };
  std::vector<std::vector<Entry>> rows(jacobian ? num_residuals_ : 0);
  for (size_t b = 0; b < residuals_.size(); ++b) {
    const ResidualBlock& rb = residuals_[b];
    const int nr = rb.cost->num_residuals();
    for (size_t s = 0; s < rb.blocks.size(); ++s) {
      const size_t jo = L.jac_off[L.slot0[b] + s];
      const int pb = rb.blocks[s];
      if (jo == (size_t)-1 || col[pb] < 0 || blocks_[pb].constant) continue;
      const int sz = blocks_[pb].size;
      for (int q = 0; q < nr; ++q)
        for (int c = 0; c < sz; ++c) {
          const double v = L.jac[jo + (size_t)q * sz + c];
          if (gradient) (*gradient)[col[pb] + c] += v * L.r[rb.residual_offset + q];
          if (jacobian) rows[rb.residual_offset + q].push_back(Entry{col[pb] + c, v});
        }
    }
  }
  if (jacobian) {
    jacobian->num_rows = num_residuals_; jacobian->num_cols = n;
    jacobian->rows.assign(1, 0); jacobian->cols.clear(); jacobian->values.clear();
    for (auto& row : rows) {
      std::stable_sort(row.begin(), row.end(), [](const Entry& a, const Entry& b) { return a.c < b.c; });
      for (const Entry& e : row) { jacobian->cols.push_back(e.c); jacobian->values.push_back(e.v); }
      jacobian->rows.push_back((int)jacobian->cols.size());
    }
  }
  return true;
}

std::string Solver::Summary::BriefReport() const {
  char buf[256];
  static const char* names[] = {"CONVERGENCE", "NO_CONVERGENCE", "FAILURE", "USER_SUCCESS", "USER_FAILURE"};
  snprintf(buf, sizeof(buf), "iterations: %d, initial cost: %.6e, final cost: %.6e, termination: %s%s%s", num_successful_steps + num_unsuccessful_steps,
           initial_cost, final_cost, names[termination_type], message.empty() ? "" : " — ", message.c_str());
  return buf;
}

void Solve(const Solver::Options& opt, Problem* problem, Solver::Summary* summary) {
  Solver::Summary S;
  const auto& pbs = problem->parameter_blocks();
  const auto& rbs = problem->residual_blocks();
  EvaluationCallback* cb = opt.evaluation_callback ? opt.evaluation_callback : problem->options().evaluation_callback;

  // variable blocks and the reduced state vector
  BlockSparse H;
  std::vector<int> var_of(pbs.size(), -1);
  for (size_t b = 0; b < pbs.size(); ++b)
    if (!pbs[b].constant) { var_of[b] = (int)H.size.size(); H.off.push_back(H.n); H.size.push_back(pbs[b].size); H.n += pbs[b].size; }
  const int n = H.n;
  std::vector<double> x(n), x_new(n);
  auto read_state = [&](std::vector<double>* v) { for (size_t b = 0; b < pbs.size(); ++b) if (var_of[b] >= 0) memcpy(&(*v)[H.off[var_of[b]]], pbs[b].values, sizeof(double) * pbs[b].size); };
  auto write_state = [&](const std::vector<double>& v) { for (size_t b = 0; b < pbs.size(); ++b) if (var_of[b] >= 0) memcpy(pbs[b].values, &v[H.off[var_of[b]]], sizeof(double) * pbs[b].size); };
  read_state(&x);

  Linearization L;
  if (!evaluate_blocks(problem, cb, true, true, &L)) { S.termination_type = FAILURE; S.message = "initial evaluation failed"; if (summary) *summary = S; return; }
  ++S.num_residual_evaluations; ++S.num_jacobian_evaluations;
  double cost = half_sq_norm(L.r);
  S.initial_cost = S.final_cost = cost;
  if (n == 0 || rbs.empty()) { S.termination_type = CONVERGENCE; S.message = "nothing to optimise"; if (summary) *summary = S; return; }

  std::vector<double> scale(n, 1.0), g(n), lm2(n), step(n), Hs(n);
  bool have_scale = false;
  H.diag.resize(H.size.size());

  // J^T J and J^T r of the (column-scaled) Jacobian from the per-block Jacobians.
  auto assemble = [&]() {
    for (size_t a = 0; a < H.diag.size(); ++a) H.diag[a].assign((size_t)H.size[a] * H.size[a], 0.0);
    for (auto& m : H.offd) std::fill(m.begin(), m.end(), 0.0);
    std::fill(g.begin(), g.end(), 0.0);
    if (opt.jacobi_scaling && !have_scale) {   // column norms of the first Jacobian, kept for the whole solve
      std::vector<double> cn(n, 0.0);
      for (size_t b = 0; b < rbs.size(); ++b)
        for (size_t s = 0; s < rbs[b].blocks.size(); ++s) {
          const size_t jo = L.jac_off[L.slot0[b] + s];
          if (jo == (size_t)-1) continue;
          const int v = var_of[rbs[b].blocks[s]], sz = H.size[v], nr = rbs[b].cost->num_residuals();
          for (int q = 0; q < nr; ++q) for (int c = 0; c < sz; ++c) { const double t = L.jac[jo + (size_t)q * sz + c]; cn[H.off[v] + c] += t * t; }
        }
      for (int i = 0; i < n; ++i) scale[i] = 1.0 / (1.0 + ::sqrt(cn[i]));
      have_scale = true;
    }
    for (size_t b = 0; b < rbs.size(); ++b) {
      const Problem::ResidualBlock& rb = rbs[b];
      const int nr = rb.cost->num_residuals();
      const double* r = &L.r[rb.residual_offset];
      for (size_t s = 0; s < rb.blocks.size(); ++s) {
        const size_t jo = L.jac_off[L.slot0[b] + s];
        if (jo == (size_t)-1) continue;
        const int va = var_of[rb.blocks[s]], sa = H.size[va], oa = H.off[va];
        const double* Ja = &L.jac[jo];
        for (int i = 0; i < sa; ++i) {
          double gi = 0;
          for (int q = 0; q < nr; ++q) gi += Ja[q * sa + i] * r[q];
          g[oa + i] += gi * scale[oa + i];
          for (int j = 0; j < sa; ++j) { double h = 0; for (int q = 0; q < nr; ++q) h += Ja[q * sa + i] * Ja[q * sa + j]; H.diag[va][i * sa + j] += h * scale[oa + i] * scale[oa + j]; }
        }
        for (size_t t = s + 1; t < rb.blocks.size(); ++t) {
          const size_t ko = L.jac_off[L.slot0[b] + t];
          if (ko == (size_t)-1) continue;
          int vb = var_of[rb.blocks[t]];
          const double* Jb = &L.jac[ko];
          int a = va, bb = vb; const double* A = Ja; const double* B = Jb;
          if (a == bb) {   // same block twice in one residual block: fold into the diagonal
            const int sz = H.size[a], o = H.off[a];
            for (int i = 0; i < sz; ++i) for (int j = 0; j < sz; ++j) {
              double h = 0; for (int q = 0; q < nr; ++q) h += A[q * sz + i] * B[q * sz + j] + B[q * sz + i] * A[q * sz + j];
              H.diag[a][i * sz + j] += h * scale[o + i] * scale[o + j];
            }
            continue;
          }
          if (a > bb) { std::swap(a, bb); std::swap(A, B); }
          const int p = H.pair(a, bb), sa2 = H.size[a], sb2 = H.size[bb], o1 = H.off[a], o2 = H.off[bb];
          std::vector<double>& M = H.offd[p];
          for (int i = 0; i < sa2; ++i) for (int j = 0; j < sb2; ++j) {
            double h = 0; for (int q = 0; q < nr; ++q) h += A[q * sa2 + i] * B[q * sb2 + j];
            M[i * sb2 + j] += h * scale[o1 + i] * scale[o2 + j];
          }
        }
      }
    }
  };
  auto max_abs_unscaled_gradient = [&]() { double m = 0; for (int i = 0; i < n; ++i) m = std::max(m, fabs(g[i] / scale[i])); return m; };

  assemble();
  if (max_abs_unscaled_gradient() <= opt.gradient_tolerance) { S.termination_type = CONVERGENCE; S.message = "gradient tolerance reached at the start"; if (summary) *summary = S; return; }

  double radius = opt.initial_trust_region_radius, decrease_factor = 2.0;
  S.termination_type = NO_CONVERGENCE;
  if (opt.minimizer_progress_to_stdout) printf("iter      cost      cost_change  |gradient|   |step|    tr_ratio  tr_radius\n   0  %.6e    0.00e+00    %.2e   0.00e+00   0.00e+00  %.2e\n", cost, max_abs_unscaled_gradient(), radius);
  for (int iter = 1; iter <= opt.max_num_iterations; ++iter) {
    // LM regularisation on the scaled system
    for (size_t a = 0; a < H.diag.size(); ++a)
      for (int i = 0; i < H.size[a]; ++i) lm2[H.off[a] + i] = std::min(std::max(H.diag[a][i * H.size[a] + i], opt.min_lm_diagonal), opt.max_lm_diagonal) / radius;
    bool solved;
    if (n <= opt.dense_limit) {
      std::vector<double> A((size_t)n * n, 0.0);
      for (size_t a = 0; a < H.diag.size(); ++a) { const int o = H.off[a], s = H.size[a]; for (int i = 0; i < s; ++i) for (int j = 0; j < s; ++j) A[(size_t)(o + i) * n + o + j] = H.diag[a][i * s + j]; }
      for (size_t p = 0; p < H.pairs.size(); ++p) {
        const int a = H.pairs[p].first, b = H.pairs[p].second, oa = H.off[a], ob = H.off[b], sa = H.size[a], sb = H.size[b];
        for (int i = 0; i < sa; ++i) for (int j = 0; j < sb; ++j) { A[(size_t)(oa + i) * n + ob + j] = H.offd[p][i * sb + j]; A[(size_t)(ob + j) * n + oa + i] = H.offd[p][i * sb + j]; }
      }
      for (int i = 0; i < n; ++i) A[(size_t)i * n + i] += lm2[i];
      step = g;
      solved = dense_cholesky_solve(A, n, step);
    } else {
      bool chain = false;
      solved = false;
      if (opt.chain_direct) solved = tridiagonal_solve(H, lm2, g, &step, &chain);   // odometry chain + unary factors: direct, O(#poses)
      if (!chain) solved = pcg_solve(H, lm2, g, opt.cg_max_iterations, opt.cg_tolerance, &step);
    }
    double model_change = 0;
    if (solved) {
      for (int i = 0; i < n; ++i) step[i] = -step[i];
      H.multiply(step, &Hs);
      for (int i = 0; i < n; ++i) model_change -= step[i] * (g[i] + 0.5 * Hs[i]);
    }
    if (!solved || !(model_change > 0.0)) {   // invalid step: shrink and retry
      ++S.num_unsuccessful_steps;
      radius /= decrease_factor; decrease_factor *= 2.0;
      if (radius < opt.min_trust_region_radius) { S.termination_type = CONVERGENCE; S.message = "trust region radius below minimum"; break; }
      continue;
    }
    double step_norm = 0, x_norm = 0;
    for (int i = 0; i < n; ++i) { const double d = step[i] * scale[i]; x_new[i] = x[i] + d; step_norm += d * d; x_norm += x[i] * x[i]; }
    step_norm = ::sqrt(step_norm); x_norm = ::sqrt(x_norm);
    if (step_norm <= opt.parameter_tolerance * (x_norm + opt.parameter_tolerance)) { S.termination_type = CONVERGENCE; S.message = "parameter tolerance reached"; break; }
    write_state(x_new);
    Linearization Ln;
    if (!evaluate_blocks(problem, cb, false, true, &Ln)) { S.termination_type = FAILURE; S.message = "residual evaluation failed"; write_state(x); break; }
    ++S.num_residual_evaluations;
    const double new_cost = half_sq_norm(Ln.r);
    const double cost_change = cost - new_cost;
    const double rho = cost_change / model_change;
    if (opt.minimizer_progress_to_stdout) printf("%4d  %.6e   %9.2e    %.2e   %.2e  %9.2e  %.2e\n", iter, new_cost, cost_change, max_abs_unscaled_gradient(), step_norm, rho, radius);
    if (rho > opt.min_relative_decrease) {
      ++S.num_successful_steps;
      x = x_new;
      const double old_cost = cost;
      cost = new_cost;
      if (!evaluate_blocks(problem, cb, true, false, &L)) { S.termination_type = FAILURE; S.message = "jacobian evaluation failed"; break; }
      ++S.num_jacobian_evaluations;
      assemble();
      if (fabs(cost_change) <= opt.function_tolerance * old_cost) { S.termination_type = CONVERGENCE; S.message = "function tolerance reached"; break; }
      if (max_abs_unscaled_gradient() <= opt.gradient_tolerance) { S.termination_type = CONVERGENCE; S.message = "gradient tolerance reached"; break; }
      const double t = 2.0 * rho - 1.0;
      radius = std::min(opt.max_trust_region_radius, radius / std::max(1.0 / 3.0, 1.0 - t * t * t));
      decrease_factor = 2.0;
    } else {
      ++S.num_unsuccessful_steps;
      write_state(x);
      radius /= decrease_factor; decrease_factor *= 2.0;
      if (radius < opt.min_trust_region_radius) { S.termination_type = CONVERGENCE; S.message = "trust region radius below minimum"; break; }
    }
  }
  write_state(x);
  S.final_cost = cost;
  if (summary) *summary = S;
}

}  // namespace ceres
}  // namespace hitl
#endif  // HITL_USE_SYSTEM_CERES
