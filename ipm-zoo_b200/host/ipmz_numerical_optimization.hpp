// ipm-zoo_b200/host/ipmz_numerical_optimization.hpp
//
// Host-side C++ mirror of the reference's numerical interface for this path, written above the
// C ABI (include/ipmz.h).  Same names, argument meaning and error behaviour as
//   NumericalOptimization::Data / build_environment   include/NumericalOptimization/EnvironmentBuilder.h:7-20
//   NumericalOptimization::Optimizer                  include/NumericalOptimization/Optimizer.h:13-20
//   NumericalOptimization::LinearSolvers::*           include/NumericalOptimization/LinearSolvers.h:11-31
//   SymbolicOptimization::Settings / Bounds           include/SymbolicOptimization.h:28-64
// but without the symbolic Expression layer: the Environment is a plain map from the
// reference's variable names ("x", "\\lambda_{A}", "s", "g", ...; SymbolicOptimization.h:5-26,
// SymbolicOptimization.cpp:35-42) to vectors.  The adapter that plugs into the reference's own
// Environment / ExprPtr types is integration/ipmz_reference_adapter.cpp.
#pragma once
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ipmz.h"

namespace ipmz_host {

// reference convention: ASSERT -> Utils::AssertionError : std::logic_error (include/Utils/Assert.h:7-12)
class AssertionError : public std::logic_error {
 public:
  using std::logic_error::logic_error;
};

namespace SymbolicOptimization {
enum class Bounds { None, Lower, Upper, Both };
enum class InequalityHandling { Slacks, SlackedSlacks, NaiveSlacks };
enum class EqualityHandling { None, Slacks, SlackedSlacks, NaiveSlacks, PenaltyFunction,
                              PenaltyFunctionWithExtraDual, Regularization };
struct Settings {
  Bounds inequalities = Bounds::Both;
  Bounds variable_bounds = Bounds::Both;
  bool equalities = false;
  EqualityHandling equality_handling = EqualityHandling::None;
  InequalityHandling inequality_handling = InequalityHandling::SlackedSlacks;
};
}  // namespace SymbolicOptimization

namespace NumericalOptimization {
using Matrix = std::vector<std::vector<double>>;
using Vector = std::vector<double>;

struct Data {
  Matrix Q;
  Vector c;
  Matrix A_ineq;
  Vector l_A_ineq, u_A_ineq;
  Matrix A_eq;
  Vector b_eq;
  Vector l_x, u_x;
};

// name -> value, the numeric part of Evaluation::Environment (Evaluation.h:22)
using Environment = std::map<std::string, Vector>;

enum class Reduction { Augmented = IPMZ_REDUCTION_AUGMENTED, Normal = IPMZ_REDUCTION_NORMAL, Full = IPMZ_REDUCTION_FULL,
                       DualNormal = IPMZ_REDUCTION_DUAL_NORMAL };

// Validates the bounds like the reference (EnvironmentBuilder.cpp:10-17) and returns the
// reference's initial point keyed by variable name (EnvironmentBuilder.cpp:34-73).
Environment build_environment(const Data& data);

struct IterationLog {
  std::vector<double> f, res, gap;  // the numbers of the reference's "iter:" lines (Optimizer.cpp:131-132)
  int iterations = 0;
  bool converged = false;
  double solve_ms = 0.0;
};

class Optimizer {
 public:
  // env is held by reference and is both input (initial iterate) and output (final iterate),
  // like Optimizer.h:15 / :54.
  Optimizer(Environment& env, const Data& data, const SymbolicOptimization::Settings& settings,
            Reduction reduction = Reduction::Augmented, int device = 0);
  ~Optimizer();
  Optimizer(const Optimizer&) = delete;
  Optimizer& operator=(const Optimizer&) = delete;

  void solve();  // throws AssertionError where the reference asserts
  const IterationLog& log() const { return log_; }
  // prints the reference's "iter: k, f: ..., res: ..., gap: ..." lines for diffing
  void print_trace(std::ostream& os) const;

 private:
  Environment& env_;
  ipmz_handle handle_ = nullptr;
  int n_ = 0, mi_ = 0, me_ = 0;
  bool reg_eq_ = false;  // EqualityHandling::Regularization: the `t` slot of the packed iterate carries p
  IterationLog log_;
};

namespace LinearSolvers {
std::pair<Matrix, std::vector<double>> ldlt_decomposition(const Matrix& A);
void overwriting_solve_ldlt(const Matrix& L, const std::vector<double>& D, std::vector<double>& b);
// Bunch-Kaufman, LinearSolvers.h:24-31: {factor (lower triangle = L and D blocks, upper = A's), LAPACK-style ipiv}
std::pair<Matrix, std::vector<int>> symmetric_indefinite_factorization(const Matrix& A);
void overwriting_solve_bunch_kaufman(const Matrix& L, const std::vector<int>& ipiv, std::vector<double>& b);
}  // namespace LinearSolvers

}  // namespace NumericalOptimization
}  // namespace ipmz_host
