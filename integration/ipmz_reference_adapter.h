// integration/ipmz_reference_adapter.h -- the reference-side binding: a drop-in for
// NumericalOptimization::Optimizer (include/NumericalOptimization/Optimizer.h:13-20) that keeps
// the reference's constructor signature, ownership (env by reference, in/out) and error
// convention (Utils::AssertionError), and runs the Newton/KKT hot path on the B200 through the
// C ABI of include/ipmz.h.  A maintainer adds this file + its .cpp to src/CMakeLists.txt, links
// libipmz_b200.so, and changes the one call site (src/IpmZoo.cpp:377-379):
//
//     NumericalOptimization::B200Optimizer optimizer(env, optimization_expressions, newton_system);
//     optimizer.solve();
//
// It compiles against the UNMODIFIED reference headers (see integration/Makefile).
#pragma once
#include "NumericalOptimization/Evaluation.h"
#include "SymbolicOptimization.h"

struct ipmz_solver_s;

namespace NumericalOptimization {

class B200Optimizer {
 public:
  enum class Reduction { Augmented = 0, Normal = 1, Full = 2, DualNormal = 3 };

  B200Optimizer(Evaluation::Environment& env,
                const SymbolicOptimization::OptimizationExpressions& optimization_expressions,
                SymbolicOptimization::NewtonSystem newton_system, Reduction reduction = Reduction::Augmented,
                int device = 0);
  ~B200Optimizer();
  B200Optimizer(const B200Optimizer&) = delete;
  B200Optimizer& operator=(const B200Optimizer&) = delete;

  void solve();
  int iterations() const { return iterations_; }
  bool converged() const { return converged_; }

 private:
  Evaluation::Environment& env_;
  SymbolicOptimization::OptimizationExpressions oe_;
  SymbolicOptimization::NewtonSystem newton_system_;
  SymbolicOptimization::NewtonSystem augmented_system_;
  ipmz_solver_s* handle_ = nullptr;
  int n_ = 0, mi_ = 0, me_ = 0;
  int iterations_ = 0;
  bool converged_ = false;
  bool reg_eq_ = false;   // EqualityHandling::Regularization: p_eq in the system, block -delta^2 I
  bool hard_eq_ = false;  // EqualityHandling::None: lambda_A_eq without s_A_eq -> indefinite KKT, Bunch-Kaufman
  bool pen_eq_ = false;   // EqualityHandling::PenaltyFunction*: lambda_A_eq alone with the diagonal block -mu
};

}  // namespace NumericalOptimization
