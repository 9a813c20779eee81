// integration/ipmz_reference_adapter.cpp -- see the header.
//
// What the adapter does, step by step, and the reference code it stands in for:
//   1. reduction selection exactly as Optimizer::solve (Optimizer.cpp:63-73): derive the
//      augmented system with the reference's own symbolic layer; a symbolic-zero diagonal block
//      is the indefinite case (solve_indefinite_, :75, ASSERT(false) in the reference): the zero
//      block of EqualityHandling::None rows goes to the Bunch-Kaufman path, anything else asserts;
//   2. classify which slack / dual groups the Newton system contains by looking up the
//      reference's variable handles in newton_system.variables (they are hash-consed, so
//      pointer equality is identity: ExprFactory.cpp:14-34) -- this replaces the tree-walking
//      evaluation of the block expressions (Evaluation.cpp:102-176) by the fused CUDA kernels;
//   3. flatten the Environment's ValMatrix / ValVector entries (Evaluation.h:12-22) into the
//      contiguous row-major buffers of ipmz_problem and the packed iterate;
//   4. ipmz_create + ipmz_set_iterate + ipmz_solve; write the final iterate back into env under
//      the variable keys (Optimizer.cpp:228) so callers read results where they always did, and the
//      rest of what the reference's solve() leaves there: the last iteration's `\Delta v` and
//      `\Delta v_affine` directions, the shorthand residuals r_{v} and mu (ipmz_get_last_iteration).
#include "ipmz_reference_adapter.h"

#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "../include/ipmz.h"
#include "Utils/Assert.h"
#include "ExprFactory.h"
#include "Utils/Helpers.h"

namespace NumericalOptimization {

namespace {
using Expression::ExprPtr;

bool has_var(const SymbolicOptimization::NewtonSystem& ns, const ExprPtr& v) {
  return std::find(ns.variables.begin(), ns.variables.end(), v) != ns.variables.end();
}

std::vector<double> flat_matrix(const Evaluation::Environment& env, const ExprPtr& key, size_t* rows) {
  std::vector<double> out;
  *rows = 0;
  if (!env.contains(key)) return out;
  const auto m = Evaluation::evaluate_matrix(key, env);
  *rows = m.size();
  for (const auto& r : m) out.insert(out.end(), r.begin(), r.end());
  return out;
}

std::vector<double> vec_of(const Evaluation::Environment& env, const ExprPtr& key) {
  if (!env.contains(key)) return {};
  return Evaluation::evaluate_vector(key, env);
}

void ipmz_check(int rc) {
  // reference convention: failures surface as Utils::AssertionError (Assert.cpp:10-31)
  ASSERT(rc == 0, std::string("ipmz: ") + ipmz_last_error());
}
}  // namespace

B200Optimizer::B200Optimizer(Evaluation::Environment& env,
                             const SymbolicOptimization::OptimizationExpressions& oe,
                             SymbolicOptimization::NewtonSystem newton_system, Reduction reduction, int device)
    : env_(env), oe_(oe), newton_system_(newton_system) {
  {
    auto ns = newton_system_;
    ns.rhs = SymbolicOptimization::get_shorthand_rhs(newton_system_).shorthand_rhs;
    augmented_system_ = SymbolicOptimization::get_augmented_system(ns);
  }
  const auto& ns = newton_system_;
  // groups present in the system (SlackedSlacks handling: SymbolicOptimization.cpp:86-103, :150-162, :224-235)
  const bool g = has_var(ns, oe_.s_A_ineq_l), h = has_var(ns, oe_.s_A_ineq_u);
  const bool v = has_var(ns, oe_.s_A_eq_l), w = has_var(ns, oe_.s_A_eq_u);
  const bool y = has_var(ns, oe_.s_x_l), z = has_var(ns, oe_.s_x_u);
  const bool ineq = has_var(ns, oe_.lambda_A_ineq), eq = has_var(ns, oe_.lambda_A_eq);
  ASSERT(!ineq || has_var(ns, oe_.s_A_ineq), "inequalities must use InequalityHandling::SlackedSlacks");
  // EqualityHandling::None (SymbolicOptimization.cpp:137-140): the multiplier exists but no equality slack does
  hard_eq_ = eq && !has_var(ns, oe_.s_A_eq) && !v && !w;
  const bool reg_probe = eq && has_var(ns, oe_.p_eq);
  ASSERT(!eq || hard_eq_ || reg_probe || (v && w && has_var(ns, oe_.s_A_eq)),
         "equalities must use EqualityHandling::SlackedSlacks, None, Regularization or PenaltyFunction*");
  // EqualityHandling::Regularization (SymbolicOptimization.cpp:184-192): p_eq is a variable of the system and the
  // multiplier's diagonal block is the scalar -delta^2 I; p travels in the `t` slot of the packed iterate
  reg_eq_ = eq && has_var(ns, oe_.p_eq);
  if (reg_eq_) hard_eq_ = false;
  // EqualityHandling::PenaltyFunction / PenaltyFunctionWithExtraDual (SymbolicOptimization.cpp:173-183): like None only the
  // multiplier exists, but its diagonal block in the augmented system is -mu, not the symbolic zero
  if (hard_eq_) {
    const auto& av = augmented_system_.variables;
    for (size_t i = 0; i < av.size(); ++i)
      if (av.at(i) == oe_.lambda_A_eq && !(augmented_system_.lhs.at(i).at(i) == Expression::zero)) pen_eq_ = true;
    if (pen_eq_) hard_eq_ = false;
  }

  size_t nq = 0, mi = 0, me = 0;
  const auto Q = flat_matrix(env_, oe_.Q, &nq);
  const auto A = ineq ? flat_matrix(env_, oe_.A_ineq, &mi) : std::vector<double>();
  const auto C = eq ? flat_matrix(env_, oe_.A_eq, &me) : std::vector<double>();
  const auto c = vec_of(env_, oe_.c), lA = vec_of(env_, oe_.l_A_ineq), uA = vec_of(env_, oe_.u_A_ineq);
  const auto d = vec_of(env_, oe_.b_eq), lx = vec_of(env_, oe_.l_x), ux = vec_of(env_, oe_.u_x);
  n_ = static_cast<int>(nq);
  mi_ = static_cast<int>(mi);
  me_ = static_cast<int>(me);
  ASSERT(c.size() == nq && lx.size() == nq && ux.size() == nq);

  ipmz_problem p;
  std::memset(&p, 0, sizeof(p));
  p.n = n_; p.m_ineq = mi_; p.m_eq = me_;
  p.Q = Q.data(); p.c = c.data();
  p.A = A.data(); p.l_A = lA.data(); p.u_A = uA.data();
  p.C = C.data(); p.d = d.data();
  p.l_x = lx.data(); p.u_x = ux.data();
  p.ineq_bounds = !ineq ? IPMZ_BOUNDS_NONE : (g && h) ? IPMZ_BOUNDS_BOTH : g ? IPMZ_BOUNDS_LOWER : IPMZ_BOUNDS_UPPER;
  p.var_bounds = (y && z) ? IPMZ_BOUNDS_BOTH : y ? IPMZ_BOUNDS_LOWER : z ? IPMZ_BOUNDS_UPPER : IPMZ_BOUNDS_NONE;
  p.equalities = !eq ? IPMZ_EQ_OFF : reg_eq_ ? IPMZ_EQ_REGULARIZATION : pen_eq_ ? IPMZ_EQ_PENALTY : hard_eq_ ? IPMZ_EQ_NONE
                     : IPMZ_EQ_SLACKED_SLACKS;
  ipmz_options opt;
  ipmz_default_options(&opt);
  opt.reduction = static_cast<int>(reduction);
  opt.device = device;
  if (reg_eq_ && env_.contains(oe_.delta_eq)) {  // EnvironmentBuilder.cpp:48
    const auto dv = Evaluation::evaluate(oe_.delta_eq, env_);
    if (std::holds_alternative<Evaluation::ValScalar>(dv)) opt.delta_eq = std::get<Evaluation::ValScalar>(dv);
  }
  ipmz_check(ipmz_create(&p, &opt, &handle_));
}

B200Optimizer::~B200Optimizer() { ipmz_destroy(handle_); }

void B200Optimizer::solve() {
  // Optimizer.cpp:63-73: a symbolic zero on the diagonal of the augmented system is the indefinite case.  The
  // reference's hook for it (solve_indefinite_, :75) is ASSERT(false); the one pattern implemented here is the
  // zero block of EqualityHandling::None rows, which the library factorizes with Bunch-Kaufman.
  const auto& lhs = augmented_system_.lhs;
  bool indefinite = false;
  for (size_t i = 0; i < lhs.size(); ++i) indefinite = indefinite || (lhs.at(i).at(i) == Expression::zero);
  ASSERT(!indefinite || hard_eq_);

  struct Slot { ExprPtr key; int len; };
  const std::vector<Slot> slots = {
      {oe_.x, n_},
      {oe_.lambda_A_ineq, mi_}, {oe_.s_A_ineq, mi_}, {oe_.lambda_sAineql, mi_}, {oe_.lambda_sAinequ, mi_},
      {oe_.s_A_ineq_l, mi_}, {oe_.s_A_ineq_u, mi_},
      {oe_.lambda_A_eq, me_}, {reg_eq_ ? oe_.p_eq : oe_.s_A_eq, me_}, {oe_.lambda_sAeql, me_}, {oe_.lambda_sAequ, me_},
      {oe_.s_A_eq_l, me_}, {oe_.s_A_eq_u, me_},
      {oe_.lambda_sxl, n_}, {oe_.lambda_sxu, n_}, {oe_.s_x_l, n_}, {oe_.s_x_u, n_}};
  size_t total = 0;
  for (const auto& s : slots) total += s.len;
  std::vector<double> packed(total, 1.0);
  size_t off = 0;
  for (const auto& s : slots) {
    if (env_.contains(s.key)) {
      const auto val = Evaluation::evaluate_vector(s.key, env_);
      if (static_cast<int>(val.size()) == s.len) std::copy(val.begin(), val.end(), packed.begin() + off);
    }
    off += s.len;
  }
  ipmz_check(ipmz_set_iterate(handle_, packed.data()));
  ipmz_result r;
  ipmz_check(ipmz_solve(handle_, &r));
  iterations_ = r.iterations;
  converged_ = r.converged != 0;
  ipmz_check(ipmz_get_iterate(handle_, packed.data()));
  auto slice = [&](const std::vector<double>& pk, size_t o, int len) {
    return Evaluation::val_vector(std::vector<double>(pk.begin() + o, pk.begin() + o + len));
  };
  off = 0;
  for (const auto& s : slots) {
    if (s.len > 0 && has_var(newton_system_, s.key)) env_.at(s.key) = slice(packed, off, s.len);
    off += s.len;
  }
  if (iterations_ == 0) return;  // converged at the first test: the reference leaves the rest of env untouched too

  // The rest of the reference's Environment contract: the last iteration's directions under the `\Delta v` keys
  // (Optimizer.cpp:369, :377), the affine directions of the complementarity variables under `\Delta v_affine`
  // (:104-112, :200), the shorthand residuals r_{v} with their corrector values (:188-209) and mu = sigma mu (:179).
  std::vector<double> delta(total), daff(total), resid(total);
  double mu_c = 0.0;
  ipmz_check(ipmz_get_last_iteration(handle_, delta.data(), daff.data(), resid.data(), &mu_c));
  using EF = Expression::ExprFactory;
  const auto shorthand = SymbolicOptimization::get_shorthand_rhs(newton_system_);
  // variables whose affine direction the reference stores: those of a shorthand definition that contains a vector of
  // ones and mu (the complementarity rows), Optimizer.cpp:191-203
  std::vector<ExprPtr> affine_vars;
  for (const auto& [vec, expr] : shorthand.vector_definitions) {
    if ((expr->contains_subexpression(oe_.e_var) || expr->contains_subexpression(oe_.e_ineq) ||
         expr->contains_subexpression(oe_.e_eq)) &&
        expr->contains_subexpression(oe_.mu))
      for (const auto& var : expr->get_variables())
        if (std::find(affine_vars.begin(), affine_vars.end(), var) == affine_vars.end()) affine_vars.push_back(var);
  }
  off = 0;
  for (const auto& s : slots) {
    if (s.len > 0 && has_var(newton_system_, s.key)) {
      const auto dv = SymbolicOptimization::get_delta_variable(s.key);
      env_[dv] = slice(delta, off, s.len);
      if (std::find(affine_vars.begin(), affine_vars.end(), s.key) != affine_vars.end()) {
        const auto& var = std::get<Expression::Variable>(dv->get_impl());
        env_[EF::variable(var.name + "_affine")] = slice(daff, off, s.len);
      }
      env_[EF::named_vector("r_{" + s.key->to_string() + "}")] = slice(resid, off, s.len);
    }
    off += s.len;
  }
  env_.at(oe_.mu) = Evaluation::val_scalar(mu_c);
}

}  // namespace NumericalOptimization
