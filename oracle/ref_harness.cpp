// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// C-ABI driver around the UNMODIFIED reference (albfre/ipm-zoo), compiled by
// oracle/Makefile from the sources where they lie under /root/reference into
// oracle/_ref/libipmzoo_ref.so.  No reference source is copied; this file only calls
// the reference's public C++ API:
//   build_environment            include/NumericalOptimization/EnvironmentBuilder.h:19-20
//   get_newton_system            include/SymbolicOptimization.h:153-154
//   Optimizer(env, oe, ns)/solve include/NumericalOptimization/Optimizer.h:15-20
//   LinearSolvers::*             include/NumericalOptimization/LinearSolvers.h:11-31
//
// Optimizer::solve() returns void and reports only on stdout (Optimizer.cpp:131-132,
// :356-359).  The reference sets std::scientific but never a precision, so this harness
// sets precision 17 on std::cout and installs a filtering streambuf that keeps the
// "iter:", "b0:" (augmented RHS) and "b:" (solved augmented Newton step) lines: 17
// significant digits round-trip IEEE doubles exactly, so the trace is the reference's
// own numbers, not a re-computation.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iomanip>
#include <iostream>
#include <streambuf>
#include <string>
#include <vector>

#include "NumericalOptimization/EnvironmentBuilder.h"
#include "NumericalOptimization/Evaluation.h"
#include "NumericalOptimization/LinearSolvers.h"
#include "NumericalOptimization/Optimizer.h"
#include "SymbolicOptimization.h"
#include "oracle.h"

namespace {

std::string g_last_error;

struct StopSolve {};

class TraceTap : public std::streambuf {
 public:
  TraceTap(orc_trace* tr, int N) : tr_(tr), N_(N) {}
  int iter_lines() const { return iter_lines_; }
  double last_res() const { return last_res_; }
  double last_mu() const { return last_mu_; }

 protected:
  int_type overflow(int_type ch) override {
    if (ch != traits_type::eof()) {
      char c = static_cast<char>(ch);
      feed(&c, 1);
    }
    return ch;
  }
  std::streamsize xsputn(const char* s, std::streamsize n) override {
    feed(s, static_cast<size_t>(n));
    return n;
  }

 private:
  void feed(const char* s, size_t n) {
    size_t i = 0;
    while (i < n) {
      if (skip_) {
        const void* nl = std::memchr(s + i, '\n', n - i);
        if (!nl) return;
        i = static_cast<const char*>(nl) - s + 1;
        skip_ = false;
        line_.clear();
        continue;
      }
      const char c = s[i++];
      if (c == '\n') {
        process();
        line_.clear();
        continue;
      }
      line_.push_back(c);
      if (line_.size() == 3) {
        if (line_ != "ite" && line_ != "b: " && line_ != "b0:") {
          skip_ = true;
        }
      }
    }
  }

  void parse_vec(const char* p, double* dst) {
    int k = 0;
    while (*p && k < N_) {
      char* end = nullptr;
      const double v = std::strtod(p, &end);
      if (end == p) break;
      if (dst) dst[k] = v;
      ++k;
      p = end;
      while (*p == ',' || *p == ' ') ++p;
    }
  }

  void process() {
    if (line_.rfind("iter:", 0) == 0) {
      int it = 0;
      double f = 0, res = 0, mu = 0;
      if (std::sscanf(line_.c_str(), "iter: %d, f: %lf, res: %lf, gap: %lf", &it, &f, &res,
                      &mu) == 4) {
        if (it <= tr_->cap_iters) {
          if (tr_->f) tr_->f[it] = f;
          if (tr_->res) tr_->res[it] = res;
          if (tr_->mu) tr_->mu[it] = mu;
          tr_->n_logged = it + 1;
        }
        last_res_ = res;
        last_mu_ = mu;
        ++iter_lines_;
        b_in_iter_ = 0;
        b0_in_iter_ = 0;
        if (tr_->stop_after_cap && it >= tr_->cap_iters) throw StopSolve{};
      }
    } else if (line_.rfind("b0:", 0) == 0) {
      const int it = iter_lines_ - 1;
      if (it >= 0 && it < tr_->cap_iters) {
        double* base = b0_in_iter_ == 0 ? tr_->rhs_aff : tr_->rhs_cor;
        if (base) parse_vec(line_.c_str() + 3, base + static_cast<size_t>(it) * N_);
      }
      ++b0_in_iter_;
    } else if (line_.rfind("b: ", 0) == 0) {
      const int it = iter_lines_ - 1;
      if (it >= 0 && it < tr_->cap_iters) {
        double* base = b_in_iter_ == 0 ? tr_->step_aff : tr_->step_cor;
        if (base) parse_vec(line_.c_str() + 2, base + static_cast<size_t>(it) * N_);
      }
      ++b_in_iter_;
    }
  }

  orc_trace* tr_;
  int N_;
  std::string line_;
  bool skip_ = false;
  int iter_lines_ = 0;
  int b_in_iter_ = 0;
  int b0_in_iter_ = 0;
  double last_res_ = 1e300, last_mu_ = 1e300;
};

using Mat = std::vector<std::vector<double>>;

Mat to_mat(const double* a, int rows, int cols) {
  Mat m(rows, std::vector<double>(cols));
  for (int i = 0; i < rows; ++i)
    for (int j = 0; j < cols; ++j) m[i][j] = a[static_cast<size_t>(i) * cols + j];
  return m;
}
std::vector<double> to_vec(const double* a, int n) {
  return a ? std::vector<double>(a, a + n) : std::vector<double>();
}

SymbolicOptimization::Bounds to_bounds(int b) {
  using SymbolicOptimization::Bounds;
  switch (b) {
    case ORC_BOUNDS_LOWER: return Bounds::Lower;
    case ORC_BOUNDS_UPPER: return Bounds::Upper;
    case ORC_BOUNDS_BOTH: return Bounds::Both;
    default: return Bounds::None;
  }
}

struct Slot {
  Expression::ExprPtr key;
  int len;
};

// Packed-iterate order documented in oracle.h.
std::vector<Slot> slots(const SymbolicOptimization::OptimizationExpressions& o, int n, int mi,
                        int me) {
  return {{o.x, n},
          {o.lambda_A_ineq, mi}, {o.s_A_ineq, mi}, {o.lambda_sAineql, mi},
          {o.lambda_sAinequ, mi}, {o.s_A_ineq_l, mi}, {o.s_A_ineq_u, mi},
          {o.lambda_A_eq, me}, {o.s_A_eq, me}, {o.lambda_sAeql, me},
          {o.lambda_sAequ, me}, {o.s_A_eq_l, me}, {o.s_A_eq_u, me},
          {o.lambda_sxl, n}, {o.lambda_sxu, n}, {o.s_x_l, n}, {o.s_x_u, n}};
}

}  // namespace

extern "C" {

const char* ref_last_error(void) { return g_last_error.c_str(); }

int ref_solve(const orc_problem* p, orc_trace* tr, int quiet) {
  using namespace NumericalOptimization;
  g_last_error.clear();
  std::streambuf* old_buf = std::cout.rdbuf();
  const auto old_state = std::cout.rdstate();
  const auto old_exc = std::cout.exceptions();
  const auto old_prec = std::cout.precision();
  const auto old_flags = std::cout.flags();
  int rc = 0;
  try {
    const int n = p->n, mi = p->m_ineq, me = p->m_eq;
    Data data;
    data.Q = to_mat(p->Q, n, n);
    data.c = to_vec(p->c, n);
    data.A_ineq = to_mat(p->A, mi, n);
    data.l_A_ineq = to_vec(p->l_A, mi);
    data.u_A_ineq = to_vec(p->u_A, mi);
    data.A_eq = to_mat(p->C, me, n);
    data.b_eq = to_vec(p->d, me);
    data.l_x = to_vec(p->l_x, n);
    data.u_x = to_vec(p->u_x, n);

    SymbolicOptimization::Settings settings;
    settings.inequalities = to_bounds(p->ineq_bounds);
    settings.variable_bounds = to_bounds(p->var_bounds);
    settings.equalities = p->equalities != 0;
    settings.equality_handling = p->equalities
                                     ? SymbolicOptimization::EqualityHandling::SlackedSlacks
                                     : SymbolicOptimization::EqualityHandling::None;
    settings.inequality_handling = SymbolicOptimization::InequalityHandling::SlackedSlacks;

    const SymbolicOptimization::VariableNames names;
    const auto oe = SymbolicOptimization::get_optimization_expressions(names);
    auto env = build_environment(names, data);
    const auto sl = slots(oe, n, mi, me);
    if (tr->use_initial_iterate && tr->iterate) {
      size_t off = 0;
      for (const auto& s : sl) {
        env[s.key] = Evaluation::val_vector(
            std::vector<double>(tr->iterate + off, tr->iterate + off + s.len));
        off += s.len;
      }
    }
    const auto newton = SymbolicOptimization::get_newton_system(settings, names);

    TraceTap tap(tr, n + mi + me);
    tr->n_logged = 0;
    if (quiet) {
      std::cout.setstate(std::ios::failbit);
    } else {
      std::cout.rdbuf(&tap);
      std::cout << std::setprecision(17);
      std::cout.exceptions(std::ios::badbit);
    }
    bool stopped = false;
    const auto t0 = std::chrono::steady_clock::now();
    try {
      Optimizer optimizer(env, oe, newton);
      optimizer.solve();
    } catch (const StopSolve&) {
      stopped = true;
    }
    const auto t1 = std::chrono::steady_clock::now();
    tr->seconds = std::chrono::duration<double>(t1 - t0).count();

    if (quiet) {
      tr->iterations = -1;
      tr->converged = -1;
    } else if (stopped) {
      tr->iterations = tr->cap_iters;
      tr->converged = 0;
    } else {
      const bool conv = tap.last_res() < 1e-8 && tap.last_mu() < 1e-8;
      tr->converged = conv ? 1 : 0;
      tr->iterations = conv ? tap.iter_lines() - 1 : tap.iter_lines();
    }
    if (tr->iterate) {
      size_t off = 0;
      for (const auto& s : sl) {
        const auto v = Evaluation::evaluate_vector(s.key, env);
        for (int i = 0; i < s.len && i < static_cast<int>(v.size()); ++i)
          tr->iterate[off + i] = v[i];
        off += s.len;
      }
    }
  } catch (const std::exception& e) {
    g_last_error = e.what();
    rc = 1;
  } catch (...) {
    g_last_error = "unknown exception";
    rc = 2;
  }
  std::cout.exceptions(std::ios::goodbit);
  std::cout.clear();
  std::cout.rdbuf(old_buf);
  std::cout.flags(old_flags);
  std::cout.precision(old_prec);
  std::cout.clear(old_state);
  std::cout.exceptions(old_exc);
  return rc;
}

int ref_ldlt(int n, const double* A, double* L, double* D) {
  try {
    const auto [Lm, Dv] =
        NumericalOptimization::LinearSolvers::ldlt_decomposition(to_mat(A, n, n));
    for (int i = 0; i < n; ++i) {
      std::memcpy(L + static_cast<size_t>(i) * n, Lm[i].data(), sizeof(double) * n);
      D[i] = Dv[i];
    }
    return 0;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return 1;
  }
}

int ref_solve_ldlt(int n, const double* L, const double* D, double* b) {
  try {
    std::vector<double> bv(b, b + n);
    NumericalOptimization::LinearSolvers::overwriting_solve_ldlt(to_mat(L, n, n),
                                                                 to_vec(D, n), bv);
    std::memcpy(b, bv.data(), sizeof(double) * n);
    return 0;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return 1;
  }
}

int ref_bk_factor(int n, const double* A, double* LD, int* ipiv) {
  try {
    const auto [F, piv] =
        NumericalOptimization::LinearSolvers::symmetric_indefinite_factorization(
            to_mat(A, n, n));
    for (int i = 0; i < n; ++i) {
      std::memcpy(LD + static_cast<size_t>(i) * n, F[i].data(), sizeof(double) * n);
      ipiv[i] = piv[i];
    }
    return 0;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return 1;
  }
}

int ref_bk_solve(int n, const double* LD, const int* ipiv, double* b) {
  try {
    std::vector<double> bv(b, b + n);
    NumericalOptimization::LinearSolvers::overwriting_solve_bunch_kaufman(
        to_mat(LD, n, n), std::vector<int>(ipiv, ipiv + n), bv);
    std::memcpy(b, bv.data(), sizeof(double) * n);
    return 0;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return 1;
  }
}

}  // extern "C"
