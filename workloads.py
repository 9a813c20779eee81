"""Seeded synthetic dense QPs in the shapes of BASELINE.json's configs: plain numpy, no
dependency on the product or on the oracle.  bench.py's GPU arm and tests/problems.py (which wraps the same arrays for
the CPU oracles) both build their inputs here, so all sides see identical bytes.

numpy's PCG64 `default_rng(seed)` stream is stable across numpy versions."""
import numpy as np

NONE, LOWER, UPPER, BOTH = 0, 1, 2, 3  # Settings::inequalities / variable_bounds (SymbolicOptimization.h:42-64)


class QP:
    """Dense QP in the reference's `Data` + `Settings` vocabulary (EnvironmentBuilder.h:7-17): arrays and flags only."""

    def __init__(self, Q, c, A=None, l_A=None, u_A=None, C=None, d=None, l_x=None, u_x=None,
                 ineq_bounds=BOTH, var_bounds=BOTH, equalities=False):
        f = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
        self.Q, self.c = f(Q), f(c)
        self.A, self.l_A, self.u_A = f(A), f(l_A), f(u_A)
        self.C, self.d = f(C), f(d)
        self.l_x, self.u_x = f(l_x), f(u_x)
        self.n = self.Q.shape[0]
        self.m_ineq = 0 if self.A is None else self.A.shape[0]
        self.m_eq = 0 if self.C is None else self.C.shape[0]
        self.ineq_bounds = ineq_bounds if self.m_ineq else NONE
        self.var_bounds = var_bounds
        self.equalities = bool(equalities) and self.m_eq > 0

    @property
    def N(self):
        return self.n + self.m_ineq + self.m_eq


def _spd(rng, n, kind="gram"):
    if kind == "gram":  # Q = M M^T / n + I
        M = rng.standard_normal((n, n))
        return M @ M.T / n + np.eye(n)
    S = rng.standard_normal((n, n)) / np.sqrt(n)  # Q = 3 I + sym N(0, 1/n)
    return 3.0 * np.eye(n) + 0.5 * (S + S.T)


def ineq_box(n, m, seed, ineq_bounds=BOTH, var_bounds=BOTH, band=0.25, kind="gram"):
    """cfg2 / cfg3 / cfg4 shape: two-sided inequalities around A x0 plus a box."""
    rng = np.random.default_rng(seed)
    Q = _spd(rng, n, kind)
    c = rng.standard_normal(n)
    A = rng.standard_normal((m, n)) / np.sqrt(n)
    x0 = rng.uniform(-0.5, 0.5, n)
    mid = A @ x0
    return QP(Q=Q, c=c, A=A, l_A=mid - band, u_A=mid + band, l_x=-np.ones(n), u_x=np.ones(n),
              ineq_bounds=ineq_bounds, var_bounds=var_bounds)


def eq_box(n, m_eq, seed, var_bounds=BOTH):
    """cfg1 shape: equalities handled as SlackedSlacks plus a box."""
    rng = np.random.default_rng(seed)
    Q = _spd(rng, n)
    c = rng.standard_normal(n)
    Cm = rng.standard_normal((m_eq, n)) / np.sqrt(n)
    x0 = rng.uniform(-0.5, 0.5, n)
    return QP(Q=Q, c=c, C=Cm, d=Cm @ x0, l_x=-np.ones(n), u_x=np.ones(n),
              ineq_bounds=NONE, var_bounds=var_bounds, equalities=True)


def box_only(n, seed):
    rng = np.random.default_rng(seed)
    return QP(Q=_spd(rng, n), c=rng.standard_normal(n), l_x=-np.ones(n), u_x=np.ones(n),
              ineq_bounds=NONE, var_bounds=BOTH)


def portfolio(n, k, eps, seed):
    """cfg5 shape: Q = F F^T + eps I, budget row 1^T x = 1 (l = u), box [0, 1]."""
    rng = np.random.default_rng(seed)
    F = rng.standard_normal((n, k))
    Q = F @ F.T + eps * np.eye(n)
    c = -0.1 * np.abs(rng.standard_normal(n))
    A = np.ones((1, n))
    return QP(Q=Q, c=c, A=A, l_A=[1.0], u_A=[1.0], l_x=np.zeros(n), u_x=np.ones(n))


def toy():
    """The reference's own demo QP (src/IpmZoo.cpp:360-367)."""
    return QP(Q=[[1.0, 0.0], [0.0, 0.5]], c=[-10.0, 2.0], A=[[1.0, 1.0]], l_A=[1.0], u_A=[1.2],
              l_x=[0.0, 0.0], u_x=[10.0, 10.0])
