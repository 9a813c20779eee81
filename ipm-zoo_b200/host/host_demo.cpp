// ipm-zoo_b200/host/host_demo.cpp -- the reference's `IpmZoo -n` demo QP (src/IpmZoo.cpp:360-367,
// with the default SlackedSlacks handling) through the host mirror; prints the reference's
// "iter:" line format so the output diffs against the reference's stdout.
#include <iomanip>
#include <iostream>

#include "ipmz_numerical_optimization.hpp"

int main() {
  using namespace ipmz_host;
  using namespace ipmz_host::NumericalOptimization;
  try {
    Data data;
    data.Q = {{1.0, 0.0}, {0.0, 0.5}};
    data.c = {-10.0, 2.0};
    data.A_ineq = {{1.0, 1.0}};
    data.l_A_ineq = {1.0};
    data.u_A_ineq = {1.2};
    data.l_x = {0.0, 0.0};
    data.u_x = {10.0, 10.0};
    auto env = build_environment(data);
    SymbolicOptimization::Settings settings;
    Optimizer optimizer(env, data, settings);
    optimizer.solve();
    std::cout << std::setprecision(17);
    optimizer.print_trace(std::cout);
    std::cout << "x: " << env["x"][0] << ", " << env["x"][1] << std::endl;
    std::cout << "iterations: " << optimizer.log().iterations << " converged: " << optimizer.log().converged << std::endl;
    // LinearSolvers mirror
    auto [L, D] = LinearSolvers::ldlt_decomposition({{4.0, 2.0}, {2.0, 3.0}});
    std::vector<double> b = {2.0, 1.0};
    LinearSolvers::overwriting_solve_ldlt(L, D, b);
    std::cout << "ldlt: L10=" << L[1][0] << " D=" << D[0] << "," << D[1] << " x=" << b[0] << "," << b[1] << std::endl;
    // Bunch-Kaufman mirror: a zero diagonal forces a 2x2 pivot
    auto [F, piv] = LinearSolvers::symmetric_indefinite_factorization({{0.0, 1.0}, {1.0, 0.0}});
    std::vector<double> b2 = {3.0, 5.0};
    LinearSolvers::overwriting_solve_bunch_kaufman(F, piv, b2);
    std::cout << "bunch-kaufman: ipiv=" << piv[0] << "," << piv[1] << " x=" << b2[0] << "," << b2[1] << std::endl;
    // error convention: l >= u asserts
    try {
      Data bad = data;
      bad.u_x = bad.l_x;
      build_environment(bad);
      std::cout << "ERROR: expected AssertionError" << std::endl;
      return 1;
    } catch (const std::logic_error& e) {
      std::cout << "assertion ok: " << e.what() << std::endl;
    }
  } catch (const std::exception& e) {
    std::cerr << "FAILED: " << e.what() << std::endl;
    return 2;
  }
  return 0;
}
