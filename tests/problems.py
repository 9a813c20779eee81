"""Seeded synthetic dense QPs in the shapes of BASELINE.json's configs.

numpy's PCG64 `default_rng(seed)` stream is stable across numpy versions, and the same
arrays are handed to the CPU oracles and to the CUDA path, so all sides see identical bytes.
"""
import numpy as np

try:
    from . import oracle_lib as ol
except ImportError:  # imported as a top-level module (bench.py, __graft_entry__)
    import oracle_lib as ol


def _spd(rng, n, kind="gram"):
    if kind == "gram":  # Q = M M^T / n + I
        M = rng.standard_normal((n, n))
        return M @ M.T / n + np.eye(n)
    S = rng.standard_normal((n, n)) / np.sqrt(n)  # Q = 3 I + sym N(0, 1/n)
    return 3.0 * np.eye(n) + 0.5 * (S + S.T)


def ineq_box(n, m, seed, ineq_bounds=ol.BOTH, var_bounds=ol.BOTH, band=0.25, kind="gram"):
    """cfg2 / cfg3 / cfg4 shape: two-sided inequalities around A x0 plus a box."""
    rng = np.random.default_rng(seed)
    Q = _spd(rng, n, kind)
    c = rng.standard_normal(n)
    A = rng.standard_normal((m, n)) / np.sqrt(n)
    x0 = rng.uniform(-0.5, 0.5, n)
    mid = A @ x0
    return ol.Problem(Q=Q, c=c, A=A, l_A=mid - band, u_A=mid + band,
                      l_x=-np.ones(n), u_x=np.ones(n),
                      ineq_bounds=ineq_bounds, var_bounds=var_bounds)


def eq_box(n, m_eq, seed, var_bounds=ol.BOTH):
    """cfg1 shape: equalities handled as SlackedSlacks plus a box."""
    rng = np.random.default_rng(seed)
    Q = _spd(rng, n)
    c = rng.standard_normal(n)
    Cm = rng.standard_normal((m_eq, n)) / np.sqrt(n)
    x0 = rng.uniform(-0.5, 0.5, n)
    return ol.Problem(Q=Q, c=c, Ceq=Cm, d=Cm @ x0, l_x=-np.ones(n), u_x=np.ones(n),
                      ineq_bounds=ol.NONE, var_bounds=var_bounds, equalities=True)


def box_only(n, seed):
    rng = np.random.default_rng(seed)
    return ol.Problem(Q=_spd(rng, n), c=rng.standard_normal(n), l_x=-np.ones(n), u_x=np.ones(n),
                      ineq_bounds=ol.NONE, var_bounds=ol.BOTH)


def portfolio(n, k, eps, seed):
    """cfg5 shape: Q = F F^T + eps I, budget row 1^T x = 1 (l = u), box [0, 1]."""
    rng = np.random.default_rng(seed)
    F = rng.standard_normal((n, k))
    Q = F @ F.T + eps * np.eye(n)
    c = -0.1 * np.abs(rng.standard_normal(n))
    A = np.ones((1, n))
    return ol.Problem(Q=Q, c=c, A=A, l_A=[1.0], u_A=[1.0], l_x=np.zeros(n), u_x=np.ones(n))


def toy():
    """The reference's own demo QP (src/IpmZoo.cpp:360-367)."""
    return ol.Problem(Q=[[1.0, 0.0], [0.0, 0.5]], c=[-10.0, 2.0], A=[[1.0, 1.0]], l_A=[1.0],
                      u_A=[1.2], l_x=[0.0, 0.0], u_x=[10.0, 10.0])
