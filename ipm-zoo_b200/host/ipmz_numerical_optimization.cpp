// ipm-zoo_b200/host/ipmz_numerical_optimization.cpp -- see the header.
#include "ipmz_numerical_optimization.hpp"

#include <cstring>
#include <iomanip>
#include <ostream>

namespace ipmz_host {
namespace NumericalOptimization {

namespace {
void check(int rc) {
  if (rc != IPMZ_OK) throw AssertionError(std::string("ipmz: ") + ipmz_last_error());
}
std::vector<double> flatten(const Matrix& m, size_t cols) {
  std::vector<double> out;
  out.reserve(m.size() * cols);
  for (const auto& r : m) {
    if (r.size() != cols) throw AssertionError("Assertion failed: ragged matrix row");
    out.insert(out.end(), r.begin(), r.end());
  }
  return out;
}
int bounds_code(SymbolicOptimization::Bounds b) {
  using SymbolicOptimization::Bounds;
  switch (b) {
    case Bounds::Lower: return IPMZ_BOUNDS_LOWER;
    case Bounds::Upper: return IPMZ_BOUNDS_UPPER;
    case Bounds::Both: return IPMZ_BOUNDS_BOTH;
    default: return IPMZ_BOUNDS_NONE;
  }
}
// packed-iterate slots (include/ipmz.h) and the reference's variable names
struct Slot { const char* name; int kind; };  // kind 0: n, 1: m_ineq, 2: m_eq
const Slot kSlots[] = {{"x", 0},
                       {"\\lambda_{A}", 1}, {"s", 1}, {"\\lambda_{g}", 1}, {"\\lambda_{h}", 1}, {"g", 1}, {"h", 1},
                       {"\\lambda_{C}", 2}, {"t", 2}, {"\\lambda_{v}", 2}, {"\\lambda_{w}", 2}, {"v", 2}, {"w", 2},
                       {"\\lambda_{y}", 0}, {"\\lambda_{z}", 0}, {"y", 0}, {"z", 0}};
}  // namespace

Environment build_environment(const Data& data) {
  const size_t n = data.Q.size(), mi = data.A_ineq.size(), me = data.A_eq.size();
  if (data.l_x.size() != data.u_x.size()) throw AssertionError("Assertion failed: data.l_x.size() == data.u_x.size()");
  if (data.l_A_ineq.size() != data.u_A_ineq.size())
    throw AssertionError("Assertion failed: data.l_A_ineq.size() == data.u_A_ineq.size()");
  if (data.l_x.size() < n) throw AssertionError("Assertion failed: l_x / u_x shorter than x");
  for (size_t i = 0; i < data.l_x.size(); ++i)
    if (!(data.l_x[i] < data.u_x[i])) throw AssertionError("Assertion failed: data.l_x.at(i) < data.u_x.at(i)");
  for (size_t i = 0; i < data.l_A_ineq.size(); ++i)
    if (!(data.l_A_ineq[i] <= data.u_A_ineq[i]))
      throw AssertionError("Assertion failed: data.l_A_ineq.at(i) <= data.u_A_ineq.at(i)");
  Environment env;
  const size_t len[3] = {n, mi, me};
  for (const auto& s : kSlots) env[s.name] = Vector(len[s.kind], 1.0);
  env["p"] = Vector(me, 1.0);  // EqualityHandling::Regularization (EnvironmentBuilder.cpp:56)
  for (size_t i = 0; i < n; ++i) env["x"][i] = 0.5 * (data.l_x[i] + data.u_x[i]);
  for (size_t i = 0; i < mi && i < data.l_A_ineq.size(); ++i)
    env["s"][i] = 0.5 * (data.l_A_ineq[i] + data.u_A_ineq[i]);
  return env;
}

Optimizer::Optimizer(Environment& env, const Data& data, const SymbolicOptimization::Settings& settings,
                     Reduction reduction, int device)
    : env_(env) {
  using namespace SymbolicOptimization;
  // the settings family the reference itself can solve numerically
  if (settings.inequality_handling != InequalityHandling::SlackedSlacks)
    throw AssertionError("Assertion failed: only InequalityHandling::SlackedSlacks is supported");
  // EqualityHandling::None gives the indefinite KKT matrix the reference routes to solve_indefinite_() ==
  // ASSERT(false) (Optimizer.cpp:63-75); here that hook is implemented with Bunch-Kaufman (AUGMENTED only).
  const bool hard_eq = settings.equalities && settings.equality_handling == EqualityHandling::None;
  // EqualityHandling::Regularization leaves the scalar block -delta^2 I the reference's evaluator cannot assemble
  // (Evaluation.cpp:53-60); the library treats those rows as quasi-definite rows (p travels in the `t` slot).
  reg_eq_ = settings.equalities && settings.equality_handling == EqualityHandling::Regularization;
  const bool pen_eq = settings.equalities && (settings.equality_handling == EqualityHandling::PenaltyFunction ||
                                              settings.equality_handling == EqualityHandling::PenaltyFunctionWithExtraDual);
  if (settings.equalities && !hard_eq && !reg_eq_ && !pen_eq && settings.equality_handling != EqualityHandling::SlackedSlacks)
    throw AssertionError("Assertion failed: equalities need EqualityHandling::SlackedSlacks, None, Regularization or PenaltyFunction*");
  n_ = (int)data.Q.size();
  mi_ = settings.inequalities == Bounds::None ? 0 : (int)data.A_ineq.size();
  me_ = settings.equalities ? (int)data.A_eq.size() : 0;
  const auto Q = flatten(data.Q, n_);
  const auto A = mi_ ? flatten(data.A_ineq, n_) : std::vector<double>();
  const auto C = me_ ? flatten(data.A_eq, n_) : std::vector<double>();
  ipmz_problem p;
  std::memset(&p, 0, sizeof(p));
  p.n = n_; p.m_ineq = mi_; p.m_eq = me_;
  p.Q = Q.data(); p.c = data.c.data();
  p.A = A.data(); p.l_A = data.l_A_ineq.data(); p.u_A = data.u_A_ineq.data();
  p.C = C.data(); p.d = data.b_eq.data();
  p.l_x = data.l_x.data(); p.u_x = data.u_x.data();
  p.ineq_bounds = mi_ ? bounds_code(settings.inequalities) : IPMZ_BOUNDS_NONE;
  p.var_bounds = bounds_code(settings.variable_bounds);
  p.equalities = me_ ? (hard_eq ? IPMZ_EQ_NONE : reg_eq_ ? IPMZ_EQ_REGULARIZATION : pen_eq ? IPMZ_EQ_PENALTY : IPMZ_EQ_SLACKED_SLACKS)
                     : IPMZ_EQ_OFF;
  ipmz_options opt;
  ipmz_default_options(&opt);
  opt.reduction = (int)reduction;
  opt.device = device;
  check(ipmz_create(&p, &opt, &handle_));
}

Optimizer::~Optimizer() { ipmz_destroy(handle_); }

void Optimizer::solve() {
  const int len[3] = {n_, mi_, me_};
  size_t total = 0;
  for (const auto& s : kSlots) total += len[s.kind];
  std::vector<double> packed(total, 1.0);
  size_t off = 0;
  auto key = [&](const Slot& s) { return std::string((reg_eq_ && std::strcmp(s.name, "t") == 0) ? "p" : s.name); };
  for (const auto& s : kSlots) {  // env -> device (warm start: the Environment is the state)
    auto it = env_.find(key(s));
    if (it != env_.end() && (int)it->second.size() == len[s.kind])
      std::memcpy(packed.data() + off, it->second.data(), sizeof(double) * len[s.kind]);
    off += len[s.kind];
  }
  check(ipmz_set_iterate(handle_, packed.data()));
  ipmz_result r;
  check(ipmz_solve(handle_, &r));
  log_.iterations = r.iterations;
  log_.converged = r.converged != 0;
  log_.solve_ms = r.solve_ms;
  log_.f.assign(r.iterations + 1, 0.0);
  log_.res.assign(r.iterations + 1, 0.0);
  log_.gap.assign(r.iterations + 1, 0.0);
  check(ipmz_get_trace(handle_, r.iterations + 1, log_.f.data(), log_.res.data(), log_.gap.data(), nullptr, nullptr,
                       nullptr, nullptr, nullptr));
  check(ipmz_get_iterate(handle_, packed.data()));
  off = 0;
  for (const auto& s : kSlots) {  // device -> env
    env_[key(s)] = Vector(packed.begin() + off, packed.begin() + off + len[s.kind]);
    off += len[s.kind];
  }
}

void Optimizer::print_trace(std::ostream& os) const {
  for (size_t i = 0; i < log_.f.size(); ++i)
    os << "iter: " << i << std::scientific << ", f: " << log_.f[i] << ", res: " << log_.res[i]
       << ", gap: " << log_.gap[i] << std::endl;
}

namespace LinearSolvers {
std::pair<Matrix, std::vector<double>> ldlt_decomposition(const Matrix& A) {
  const int n = (int)A.size();
  const auto flat = flatten(A, n);  // throws on non-square like the reference's ASSERT (LinearSolvers.cpp:16-17)
  std::vector<double> L((size_t)n * n), D(n);
  check(ipmz_ldlt_decomposition(n, flat.data(), L.data(), D.data()));
  Matrix Lm(n, std::vector<double>(n));
  for (int i = 0; i < n; ++i) std::memcpy(Lm[i].data(), L.data() + (size_t)i * n, sizeof(double) * n);
  return {Lm, D};
}

void overwriting_solve_ldlt(const Matrix& L, const std::vector<double>& D, std::vector<double>& b) {
  if (b.empty()) return;  // LinearSolvers.cpp:46-48
  const int n = (int)b.size();
  if ((int)D.size() != n || (int)L.size() != n) throw AssertionError("Assertion failed: D.size() == n && L.size() == n");
  const auto flat = flatten(L, n);
  check(ipmz_overwriting_solve_ldlt(n, flat.data(), D.data(), b.data()));
}

std::pair<Matrix, std::vector<int>> symmetric_indefinite_factorization(const Matrix& A) {
  const int n = (int)A.size();
  const auto flat = flatten(A, n);  // non-square throws like the reference's ASSERT (LinearSolvers.cpp:80-81)
  std::vector<double> LD((size_t)n * n);
  std::vector<int> ipiv(n, 0);
  check(ipmz_symmetric_indefinite_factorization(n, flat.data(), LD.data(), ipiv.data()));
  Matrix out(n, std::vector<double>(n));
  for (int i = 0; i < n; ++i) std::memcpy(out[i].data(), LD.data() + (size_t)i * n, sizeof(double) * n);
  return {out, ipiv};
}

void overwriting_solve_bunch_kaufman(const Matrix& L, const std::vector<int>& ipiv, std::vector<double>& b) {
  const int n = (int)b.size();
  if ((int)ipiv.size() != n || (int)L.size() != n) throw AssertionError("Assertion failed: ipiv.size() == n && L.size() == n");
  if (n == 0) return;
  const auto flat = flatten(L, n);
  check(ipmz_overwriting_solve_bunch_kaufman(n, flat.data(), ipiv.data(), b.data()));
}
}  // namespace LinearSolvers

}  // namespace NumericalOptimization
}  // namespace ipmz_host
