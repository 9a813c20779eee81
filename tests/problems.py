"""The workloads of ../workloads.py (seeded synthetic dense QPs in the shapes of BASELINE.json's configs) wrapped in the
oracle's `Problem` container, so that the CPU oracles and the CUDA path are handed the same arrays."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import workloads as W  # noqa: E402

try:
    from . import oracle_lib as ol
except ImportError:  # imported as a top-level module (bench.py's CPU baseline, __graft_entry__)
    import oracle_lib as ol


def _wrap(q):
    return ol.Problem(Q=q.Q, c=q.c, A=q.A, l_A=q.l_A, u_A=q.u_A, Ceq=q.C, d=q.d, l_x=q.l_x, u_x=q.u_x,
                      ineq_bounds=q.ineq_bounds if q.A is not None else ol.NONE, var_bounds=q.var_bounds,
                      equalities=q.equalities)


def ineq_box(n, m, seed, ineq_bounds=ol.BOTH, var_bounds=ol.BOTH, band=0.25, kind="gram"):
    """cfg2 / cfg3 / cfg4 shape: two-sided inequalities around A x0 plus a box."""
    return _wrap(W.ineq_box(n, m, seed, ineq_bounds, var_bounds, band, kind))


def eq_box(n, m_eq, seed, var_bounds=ol.BOTH):
    """cfg1 shape: equalities handled as SlackedSlacks plus a box."""
    return _wrap(W.eq_box(n, m_eq, seed, var_bounds))


def box_only(n, seed):
    return _wrap(W.box_only(n, seed))


def portfolio(n, k, eps, seed):
    """cfg5 shape: Q = F F^T + eps I, budget row 1^T x = 1 (l = u), box [0, 1]."""
    return _wrap(W.portfolio(n, k, eps, seed))


def toy():
    """The reference's own demo QP (src/IpmZoo.cpp:360-367)."""
    return _wrap(W.toy())
