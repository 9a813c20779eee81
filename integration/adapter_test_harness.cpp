// integration/adapter_test_harness.cpp -- TEST INFRASTRUCTURE: builds the reference's own
// Environment with build_environment, runs the drop-in B200Optimizer on it, and hands the final
// iterate back through a C function so tests/test_adapter.py can compare it with the unmodified
// reference Optimizer run on the same Environment type.
#include <cstring>
#include <string>

#include "NumericalOptimization/EnvironmentBuilder.h"
#include "ipmz_reference_adapter.h"

extern "C" int adapter_solve(int n, int mi, int me, const double* Q, const double* c, const double* A,
                             const double* lA, const double* uA, const double* C, const double* d, const double* lx,
                             const double* ux, int ineq_bounds, int var_bounds, int equalities, int reduction,
                             double* x_out, int* iterations, int* converged, char* err, int errlen) {
  using namespace NumericalOptimization;
  using SymbolicOptimization::Bounds;
  try {
    auto mat = [](const double* a, int r, int cdim) {
      std::vector<std::vector<double>> m(r, std::vector<double>(cdim));
      for (int i = 0; i < r; ++i) std::memcpy(m[i].data(), a + (size_t)i * cdim, sizeof(double) * cdim);
      return m;
    };
    auto vec = [](const double* a, int len) { return a ? std::vector<double>(a, a + len) : std::vector<double>(); };
    const Bounds bmap[4] = {Bounds::None, Bounds::Lower, Bounds::Upper, Bounds::Both};
    Data data;
    data.Q = mat(Q, n, n); data.c = vec(c, n);
    data.A_ineq = mat(A, mi, n); data.l_A_ineq = vec(lA, mi); data.u_A_ineq = vec(uA, mi);
    data.A_eq = mat(C, me, n); data.b_eq = vec(d, me);
    data.l_x = vec(lx, n); data.u_x = vec(ux, n);
    SymbolicOptimization::Settings settings;
    settings.inequalities = bmap[ineq_bounds];
    settings.variable_bounds = bmap[var_bounds];
    settings.equalities = equalities != 0;
    settings.equality_handling = equalities == 1   ? SymbolicOptimization::EqualityHandling::SlackedSlacks
                                 : equalities == 3 ? SymbolicOptimization::EqualityHandling::Regularization
                                                   : SymbolicOptimization::EqualityHandling::None;  // 2: None
    const SymbolicOptimization::VariableNames names;
    const auto oe = SymbolicOptimization::get_optimization_expressions(names);
    auto env = build_environment(names, data);
    const auto newton = SymbolicOptimization::get_newton_system(settings, names);
    B200Optimizer optimizer(env, oe, newton, static_cast<B200Optimizer::Reduction>(reduction));
    optimizer.solve();
    const auto x = Evaluation::evaluate_vector(oe.x, env);
    std::memcpy(x_out, x.data(), sizeof(double) * n);
    *iterations = optimizer.iterations();
    *converged = optimizer.converged() ? 1 : 0;
    return 0;
  } catch (const std::exception& e) {
    std::strncpy(err, e.what(), errlen - 1);
    err[errlen - 1] = 0;
    return 1;
  }
}
