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
                                 : equalities == 4 ? SymbolicOptimization::EqualityHandling::PenaltyFunction
                                 : equalities == 5 ? SymbolicOptimization::EqualityHandling::PenaltyFunctionWithExtraDual
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


// The Environment contract: run the UNMODIFIED reference Optimizer and the drop-in on two identically built
// Environments and compare EVERY entry the reference's env holds afterwards (iterate, `\Delta v`, `\Delta v_affine`,
// r_{v}, mu, ...).  out4 = {keys in the reference env, keys missing from the adapter env, worst vector / scalar
// difference relative to max(1, |reference|_inf), index of that key}; worst_key receives the key's name.
#include <cmath>
#include <iostream>
#include <sstream>

#include "NumericalOptimization/Optimizer.h"

extern "C" int adapter_env_contract(int n, int mi, int me, const double* Q, const double* c, const double* A,
                                    const double* lA, const double* uA, const double* C, const double* d,
                                    const double* lx, const double* ux, int ineq_bounds, int var_bounds, int equalities,
                                    int reduction, double* out4, char* worst_key, int keylen, char* err, int errlen) {
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
    settings.equality_handling = SymbolicOptimization::EqualityHandling::SlackedSlacks;
    const SymbolicOptimization::VariableNames names;
    const auto oe = SymbolicOptimization::get_optimization_expressions(names);
    const auto newton = SymbolicOptimization::get_newton_system(settings, names);
    auto env_ref = build_environment(names, data);
    auto env_gpu = build_environment(names, data);
    {
      std::ostringstream sink;  // the reference prints O(N^2) text per iteration
      auto* old = std::cout.rdbuf(sink.rdbuf());
      std::cout.setstate(std::ios::failbit);
      try {
        Optimizer ref(env_ref, oe, newton);
        ref.solve();
      } catch (...) {
        std::cout.clear();
        std::cout.rdbuf(old);
        throw;
      }
      std::cout.clear();
      std::cout.rdbuf(old);
    }
    B200Optimizer gpu(env_gpu, oe, newton, static_cast<B200Optimizer::Reduction>(reduction));
    gpu.solve();
    double nkeys = 0, missing = 0, worst = 0, worst_idx = -1;
    std::string wkey;
    int idx = 0;
    for (const auto& [key, val] : env_ref) {
      ++nkeys;
      auto it = env_gpu.find(key);
      if (it == env_gpu.end()) {
        ++missing;
        if (wkey.empty()) wkey = "missing: " + key->to_string();
        ++idx;
        continue;
      }
      double diff = 0.0, scale = 1.0;
      if (std::holds_alternative<Evaluation::ValScalar>(val)) {
        const double a = std::get<Evaluation::ValScalar>(val), b = std::get<Evaluation::ValScalar>(it->second);
        diff = std::fabs(a - b);
        scale = std::max(1.0, std::fabs(a));
      } else if (!std::holds_alternative<Evaluation::ValMatrix>(val)) {
        const auto a = Evaluation::evaluate_vector(key, env_ref), b = Evaluation::evaluate_vector(key, env_gpu);
        if (a.size() != b.size()) diff = 1e300;
        for (size_t i = 0; i < a.size() && i < b.size(); ++i) {
          diff = std::max(diff, std::fabs(a[i] - b[i]));
          scale = std::max(scale, std::fabs(a[i]));
        }
      }
      if (diff / scale > worst) { worst = diff / scale; worst_idx = idx; wkey = key->to_string(); }
      ++idx;
    }
    out4[0] = nkeys; out4[1] = missing; out4[2] = worst; out4[3] = worst_idx;
    std::strncpy(worst_key, wkey.c_str(), keylen - 1);
    worst_key[keylen - 1] = 0;
    return 0;
  } catch (const std::exception& e) {
    std::strncpy(err, e.what(), errlen - 1);
    err[errlen - 1] = 0;
    return 1;
  }
}
