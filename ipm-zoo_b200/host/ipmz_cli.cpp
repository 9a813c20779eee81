// ipm-zoo_b200/host/ipmz_cli.cpp -- caller side of the hot path: the counterpart of
// the reference's `IpmZoo -n` (src/IpmZoo.cpp:349-424, the only caller of Optimizer in the tree), able to take
// problems from files and to solve many of them as one batch.
//
//   ipmz_cli -n [options]                    the reference's built-in demo QP (IpmZoo.cpp:360-367) and its disabled
//                                            5-variable QP (IpmZoo.cpp:384-408) with -n5
//   ipmz_cli [options] file.qp [...]         one QP per file; several files of one shape are solved as ONE batch
//                                            (ipmz_batch_*), sharded by problem index over --devices GPUs
// options: --reduction augmented|normal|full   --equalities slacked|none|regularization   --devices N   --quiet
//
// Output: the reference's trace lines `iter: k, f: ..., res: ..., gap: ...` (Optimizer.cpp:131-132) with 17 digits
// for a single problem, then `x: ...`; for a batch one summary line per problem.  Errors follow the reference's
// convention (message on stderr, non-zero exit; the library has no CPU fallback).
//
// File format (whitespace separated, '#' comments): keyword followed by its numbers, in any order --
//   n <int>  m_ineq <int>  m_eq <int>
//   Q <n*n>  c <n>  A <m_ineq*n>  l_A <m_ineq>  u_A <m_ineq>  C <m_eq*n>  d <m_eq>  l_x <n>  u_x <n>
//   inequalities none|lower|upper|both   variable_bounds none|lower|upper|both
#include <algorithm>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <thread>

#include "ipmz_numerical_optimization.hpp"

using namespace ipmz_host;
using namespace ipmz_host::NumericalOptimization;
namespace SO = ipmz_host::SymbolicOptimization;

namespace {

struct Problem {
  Data data;
  SO::Settings settings;
  int n = 0, mi = 0, me = 0;
};

SO::Bounds parse_bounds(const std::string& s) {
  if (s == "none") return SO::Bounds::None;
  if (s == "lower") return SO::Bounds::Lower;
  if (s == "upper") return SO::Bounds::Upper;
  if (s == "both") return SO::Bounds::Both;
  throw AssertionError("bad bounds selector '" + s + "'");
}

Matrix to_matrix(const std::vector<double>& v, int rows, int cols, const char* what) {
  if ((int)v.size() != rows * cols) throw AssertionError(std::string("wrong number of entries for ") + what);
  Matrix m(rows, Vector(cols));
  for (int i = 0; i < rows; ++i) std::copy(v.begin() + (size_t)i * cols, v.begin() + (size_t)(i + 1) * cols, m[i].begin());
  return m;
}

Problem read_problem(const std::string& path) {
  std::ifstream in(path);
  if (!in) throw AssertionError("cannot open " + path);
  std::stringstream clean;
  for (std::string line; std::getline(in, line);) clean << line.substr(0, line.find('#')) << '\n';
  Problem p;
  std::vector<double> Q, c, A, lA, uA, C, d, lx, ux;
  auto numbers = [&](std::vector<double>& dst, size_t count) {
    dst.resize(count);
    for (size_t i = 0; i < count; ++i)
      if (!(clean >> dst[i])) throw AssertionError("unexpected end of " + path);
  };
  p.settings.inequalities = SO::Bounds::Both;
  p.settings.variable_bounds = SO::Bounds::Both;
  for (std::string key; clean >> key;) {
    if (key == "n") clean >> p.n;
    else if (key == "m_ineq") clean >> p.mi;
    else if (key == "m_eq") clean >> p.me;
    else if (key == "Q") numbers(Q, (size_t)p.n * p.n);
    else if (key == "c") numbers(c, p.n);
    else if (key == "A") numbers(A, (size_t)p.mi * p.n);
    else if (key == "l_A") numbers(lA, p.mi);
    else if (key == "u_A") numbers(uA, p.mi);
    else if (key == "C") numbers(C, (size_t)p.me * p.n);
    else if (key == "d") numbers(d, p.me);
    else if (key == "l_x") numbers(lx, p.n);
    else if (key == "u_x") numbers(ux, p.n);
    else if (key == "inequalities") { std::string s; clean >> s; p.settings.inequalities = parse_bounds(s); }
    else if (key == "variable_bounds") { std::string s; clean >> s; p.settings.variable_bounds = parse_bounds(s); }
    else throw AssertionError("unknown keyword '" + key + "' in " + path);
  }
  if (p.n <= 0) throw AssertionError("n missing in " + path);
  p.data.Q = to_matrix(Q, p.n, p.n, "Q");
  p.data.c = c;
  p.data.A_ineq = to_matrix(A, p.mi, p.n, "A");
  p.data.l_A_ineq = lA; p.data.u_A_ineq = uA;
  p.data.A_eq = to_matrix(C, p.me, p.n, "C");
  p.data.b_eq = d;
  p.data.l_x = lx; p.data.u_x = ux;
  if ((int)c.size() != p.n || (int)lx.size() != p.n || (int)ux.size() != p.n) throw AssertionError("c, l_x, u_x need n entries");
  if (p.mi == 0) p.settings.inequalities = SO::Bounds::None;
  p.settings.equalities = p.me > 0;
  p.settings.equality_handling = SO::EqualityHandling::SlackedSlacks;
  return p;
}

Problem demo(int which) {
  Problem p;
  Data& data = p.data;
  if (which == 2) {  // IpmZoo.cpp:360-367
    data.Q = {{1.0, 0.0}, {0.0, 0.5}};
    data.c = {-10.0, 2.0};
    data.A_ineq = {{1.0, 1.0}};
    data.l_A_ineq = {1.0};
    data.u_A_ineq = {1.2};
    data.l_x = {0.0, 0.0};
    data.u_x = {10.0, 10.0};
    p.n = 2; p.mi = 1;
  } else {  // IpmZoo.cpp:384-408
    const int n = 5, m = 2;
    data.Q = Matrix(n, Vector(n, 0.0));
    data.c = Vector(n);
    data.A_ineq = Matrix(m, Vector(n, 0.0));
    data.l_A_ineq = Vector(m);
    data.u_A_ineq = Vector(m);
    data.l_x = Vector(n);
    data.u_x = Vector(n, 100);
    for (int i = 0; i < n; ++i) {
      data.Q[i][n - i - 1] = 1.2 + 0.1;
      data.Q[n - i - 1][i] = 1.2 + 0.1;
      data.Q[i][i] = i + 3;
      data.c[i] = -0.5 * i;
      data.l_x[i] = i;
      data.u_x[i] = 2000;
    }
    data.A_ineq[0][1] = 1; data.A_ineq[0][2] = 1; data.l_A_ineq[0] = -10; data.u_A_ineq[0] = 5;
    data.A_ineq[1][0] = 1; data.l_A_ineq[1] = -10; data.u_A_ineq[1] = 7;
    p.n = n; p.mi = m;
  }
  return p;
}

void flatten_into(std::vector<double>& dst, const Matrix& m) {
  for (const auto& r : m) dst.insert(dst.end(), r.begin(), r.end());
}

int bounds_code(SO::Bounds b) {
  switch (b) {
    case SO::Bounds::None: return IPMZ_BOUNDS_NONE;
    case SO::Bounds::Lower: return IPMZ_BOUNDS_LOWER;
    case SO::Bounds::Upper: return IPMZ_BOUNDS_UPPER;
    default: return IPMZ_BOUNDS_BOTH;
  }
}

void check(int rc) {
  if (rc != 0) throw AssertionError(std::string("ipmz: ") + ipmz_last_error());
}

// problems [lo, hi) of one shape as one device batch (contiguous blocks of problem indices per GPU)
void solve_shard(const std::vector<Problem>& ps, int lo, int hi, int device, int reduction, int eq_mode,
                 std::vector<ipmz_result>& results, std::vector<double>& xs, std::string& err) {
  try {
    const Problem& p0 = ps[lo];
    const int count = hi - lo, n = p0.n;
    std::vector<double> Q, c, A, lA, uA, C, d, lx, ux;
    for (int i = lo; i < hi; ++i) {
      const Data& dt = ps[i].data;
      flatten_into(Q, dt.Q); c.insert(c.end(), dt.c.begin(), dt.c.end());
      flatten_into(A, dt.A_ineq); lA.insert(lA.end(), dt.l_A_ineq.begin(), dt.l_A_ineq.end());
      uA.insert(uA.end(), dt.u_A_ineq.begin(), dt.u_A_ineq.end());
      flatten_into(C, dt.A_eq); d.insert(d.end(), dt.b_eq.begin(), dt.b_eq.end());
      lx.insert(lx.end(), dt.l_x.begin(), dt.l_x.end()); ux.insert(ux.end(), dt.u_x.begin(), dt.u_x.end());
    }
    ipmz_problem ip;
    std::memset(&ip, 0, sizeof(ip));
    ip.n = n; ip.m_ineq = p0.mi; ip.m_eq = p0.me;
    ip.Q = Q.data(); ip.c = c.data(); ip.A = A.data(); ip.l_A = lA.data(); ip.u_A = uA.data();
    ip.C = C.data(); ip.d = d.data(); ip.l_x = lx.data(); ip.u_x = ux.data();
    ip.ineq_bounds = p0.mi ? bounds_code(p0.settings.inequalities) : IPMZ_BOUNDS_NONE;
    ip.var_bounds = bounds_code(p0.settings.variable_bounds);
    ip.equalities = p0.me ? eq_mode : IPMZ_EQ_OFF;
    ipmz_options opt;
    ipmz_default_options(&opt);
    opt.reduction = reduction;
    opt.device = device;
    ipmz_batch_handle h = nullptr;
    check(ipmz_batch_create(count, &ip, &opt, &h));
    double ms = 0.0;
    int rc = ipmz_batch_solve(h, results.data() + lo, &ms);
    if (rc == 0) rc = ipmz_batch_get_x(h, xs.data() + (size_t)lo * n);
    const std::string msg = rc ? ipmz_last_error() : "";
    ipmz_batch_destroy(h);
    if (rc) throw AssertionError("ipmz: " + msg);
  } catch (const std::exception& e) {
    err = e.what();
  }
}

}  // namespace

int main(int argc, char** argv) {
  try {
    int reduction = IPMZ_REDUCTION_AUGMENTED, eq_mode = IPMZ_EQ_SLACKED_SLACKS, devices = 1, demo_which = 0;
    bool quiet = false;
    std::vector<std::string> files;
    for (int i = 1; i < argc; ++i) {
      const std::string a = argv[i];
      auto next = [&]() -> std::string {
        if (i + 1 >= argc) throw AssertionError("missing value after " + a);
        return argv[++i];
      };
      if (a == "-n") demo_which = 2;
      else if (a == "-n5") demo_which = 5;
      else if (a == "--quiet") quiet = true;
      else if (a == "--devices") devices = std::max(1, std::stoi(next()));
      else if (a == "--reduction") {
        const std::string r = next();
        reduction = r == "augmented" ? IPMZ_REDUCTION_AUGMENTED : r == "normal" ? IPMZ_REDUCTION_NORMAL
                  : r == "full" ? IPMZ_REDUCTION_FULL : r == "dual" ? IPMZ_REDUCTION_DUAL_NORMAL : -1;
        if (reduction < 0) throw AssertionError("unknown reduction '" + r + "'");
      } else if (a == "--equalities") {
        const std::string r = next();
        eq_mode = r == "slacked" ? IPMZ_EQ_SLACKED_SLACKS : r == "none" ? IPMZ_EQ_NONE
                : r == "regularization" ? IPMZ_EQ_REGULARIZATION : -1;
        if (eq_mode < 0) throw AssertionError("unknown equality handling '" + r + "'");
      } else if (!a.empty() && a[0] == '-') {
        throw AssertionError("unknown option " + a);
      } else files.push_back(a);
    }
    std::vector<Problem> ps;
    if (demo_which) ps.push_back(demo(demo_which));
    for (const auto& f : files) ps.push_back(read_problem(f));
    if (ps.empty()) {
      std::cerr << "usage: ipmz_cli -n | [--reduction augmented|normal|full] [--equalities slacked|none|regularization] "
                   "[--devices N] [--quiet] file.qp [...]" << std::endl;
      return 2;
    }
    std::cout << std::setprecision(17);
    if (ps.size() == 1) {
      Problem& p = ps[0];
      if (p.me > 0 && eq_mode == IPMZ_EQ_NONE) p.settings.equality_handling = SO::EqualityHandling::None;
      if (p.me > 0 && eq_mode == IPMZ_EQ_REGULARIZATION) p.settings.equality_handling = SO::EqualityHandling::Regularization;
      auto env = build_environment(p.data);
      Optimizer optimizer(env, p.data, p.settings, static_cast<Reduction>(reduction));
      optimizer.solve();
      if (!quiet) optimizer.print_trace(std::cout);
      std::cout << "x:";
      for (double v : env["x"]) std::cout << ' ' << v;
      std::cout << std::endl;
      std::cout << "iterations: " << optimizer.log().iterations << " converged: " << optimizer.log().converged
                << " device_ms: " << std::setprecision(6) << optimizer.log().solve_ms << std::endl;
      return optimizer.log().converged ? 0 : 3;
    }
    // batch: every file must have the shape and Settings of the first
    for (const auto& p : ps)
      if (p.n != ps[0].n || p.mi != ps[0].mi || p.me != ps[0].me ||
          p.settings.inequalities != ps[0].settings.inequalities ||
          p.settings.variable_bounds != ps[0].settings.variable_bounds)
        throw AssertionError("a batch needs problems of one shape and one Settings");
    for (const auto& p : ps) build_environment(p.data);  // the reference's bound checks (EnvironmentBuilder.cpp:10-17)
    const int total = (int)ps.size();
    devices = std::min({devices, std::max(1, ipmz_device_count()), total});
    std::vector<ipmz_result> results(total);
    std::vector<double> xs((size_t)total * ps[0].n);
    std::vector<std::string> errs(devices);
    std::vector<std::thread> th;
    for (int g = 0; g < devices; ++g)
      th.emplace_back(solve_shard, std::cref(ps), g * total / devices, (g + 1) * total / devices, g, reduction, eq_mode,
                      std::ref(results), std::ref(xs), std::ref(errs[g]));
    for (auto& t : th) t.join();
    for (const auto& e : errs)
      if (!e.empty()) throw AssertionError(e);
    int bad = 0;
    for (int i = 0; i < total; ++i) {
      const ipmz_result& r = results[i];
      std::cout << "problem: " << i << ", iterations: " << r.iterations << ", converged: " << r.converged << std::scientific
                << ", f: " << r.f << ", res: " << r.res << ", gap: " << r.mu << std::defaultfloat;
      if (!quiet) {
        std::cout << ", x:";
        for (int j = 0; j < ps[0].n; ++j) std::cout << ' ' << xs[(size_t)i * ps[0].n + j];
      }
      std::cout << std::endl;
      bad += r.converged ? 0 : 1;
    }
    return bad ? 3 : 0;
  } catch (const std::exception& e) {
    std::cerr << e.what() << std::endl;
    return 1;
  }
}
