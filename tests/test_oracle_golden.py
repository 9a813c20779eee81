"""The plain-C oracle port (oracle/ipm_oracle.c) against reference-generated goldens.

tests/golden/*.npz were produced by the UNMODIFIED reference (tests/golden/make_golden.py);
the port must reproduce them to a few ulps -- in practice bit for bit -- because it
restates the same formulas with the same association order.
"""
import ctypes as C
import glob
import os

import numpy as np
import pytest

import oracle_lib as ol
from golden.make_golden import CASES, checksum

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", sorted(CASES))
def test_port_matches_reference_golden(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    p = CASES[name]()
    assert checksum(p) == float(g["checksum"]), "seeded generator drifted"
    tr = ol.port_solve(p)
    k = int(g["iterations"])
    assert tr.iterations == k
    assert tr.converged == int(g["converged"])
    for key in ("f", "res", "mu"):
        np.testing.assert_allclose(getattr(tr, key)[:k + 1], g[key], rtol=1e-13, atol=1e-300)
    keep = g["step_aff"].shape[0]
    np.testing.assert_allclose(tr.step_aff[:keep], g["step_aff"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(tr.step_cor[:keep], g["step_cor"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(tr.iterate, g["iterate"], rtol=1e-12, atol=1e-14)


def test_toy_known_answer():
    """SURVEY.md 8c KAT on the reference's demo QP (src/IpmZoo.cpp:360-367)."""
    import problems as P
    tr = ol.port_solve(P.toy())
    assert tr.iterations == 12 and tr.converged == 1
    assert abs(tr.f[12] - (-1.12799999999863552e+01)) < 1e-14
    np.testing.assert_allclose(tr.step_aff[0], [-7.02702702702702409e-01, -4.64324324324324333,
                                                6.10810810810810789], rtol=1e-15)


def test_ldlt_and_solve_golden():
    g = np.load(os.path.join(GOLD, "linear_solvers.npz"))
    L_ = ol.port()
    for nm in ("spd", "quasidef"):
        K = np.ascontiguousarray(g[nm + "_K"])
        m = K.shape[0]
        L = np.zeros((m, m)); D = np.zeros(m)
        L_.orc_ldlt(m, ol._ptr(K), ol._ptr(L), ol._ptr(D))
        np.testing.assert_array_equal(L, g[nm + "_L"])
        np.testing.assert_array_equal(D, g[nm + "_D"])
        x = g[nm + "_b"].copy()
        L_.orc_solve_ldlt(m, ol._ptr(L), ol._ptr(D), ol._ptr(x))
        np.testing.assert_array_equal(x, g[nm + "_x"])


def test_ldlt_zero_pivot_guard():
    """LinearSolvers.cpp:28: an exactly-zero pivot becomes 1e-8."""
    K = np.array([[0.0, 0.0], [0.0, 2.0]])
    L = np.zeros((2, 2)); D = np.zeros(2)
    ol.port().orc_ldlt(2, ol._ptr(K), ol._ptr(L), ol._ptr(D))
    assert D[0] == 1e-8 and D[1] == 2.0


def test_bunch_kaufman_golden():
    g = np.load(os.path.join(GOLD, "linear_solvers.npz"))
    ip = C.POINTER(C.c_int)
    for nm in ("indef", "quasidef"):
        K = np.ascontiguousarray(g[nm + "_bk_K"])
        m = K.shape[0]
        LD = np.zeros((m, m)); piv = np.zeros(m, dtype=np.int32)
        ol.port().orc_bk_factor(m, ol._ptr(K), ol._ptr(LD), piv.ctypes.data_as(ip))
        np.testing.assert_array_equal(piv, g[nm + "_bk_piv"])
        np.testing.assert_array_equal(LD, g[nm + "_bk_LD"])
        x = g[nm + "_bk_b"].copy()
        ol.port().orc_bk_solve(m, ol._ptr(LD), piv.ctypes.data_as(ip), ol._ptr(x))
        np.testing.assert_array_equal(x, g[nm + "_bk_x"])
        assert np.max(np.abs(K @ x - g[nm + "_bk_b"])) < 1e-12


def test_evaluator_known_answers():
    """Reference-owned known answers for the evaluator semantics the port relies on
    (test/Evaluation_test.cpp:107-173): x.y = 32, x^T Q x = 157, 0.5 x^T Q x + 2.5 y^T x = 158.5,
    A x = [14, 32, 50] -- checked through the port's objective path."""
    import problems as P
    x = np.array([1.0, 2.0, 3.0])
    Q = np.array([[2.0, 1.0, 0.0], [1.0, 3.0, 1.0], [0.0, 1.0, 4.0]])  # x^T Q x = 2+12+36+4+12 = 66
    p = ol.Problem(Q=Q, c=[4.0, 5.0, 6.0], l_x=x - 1.0, u_x=x + 1.0, ineq_bounds=ol.NONE)
    tr = ol.port_solve(p, cap_iters=0, stop_after_cap=True)
    assert tr.f[0] == 0.5 * float(x @ Q @ x) + 32.0


@pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("seed", [11, 12, 13])
def test_port_bitwise_vs_live_reference(seed):
    """Where the compiled reference is present, run it live on fresh seeds."""
    import problems as P
    p = P.ineq_box(24 + seed, 10 + seed % 5, seed)
    a = ol.ref_solve(p)
    b = ol.port_solve(p)
    assert a.iterations == b.iterations and a.converged == b.converged
    k = a.iterations
    np.testing.assert_allclose(b.f[:k + 1], a.f[:k + 1], rtol=1e-13)
    np.testing.assert_allclose(b.step_aff[:k], a.step_aff[:k], rtol=1e-11, atol=1e-14)
    np.testing.assert_allclose(b.step_cor[:k], a.step_cor[:k], rtol=1e-11, atol=1e-14)
    np.testing.assert_allclose(b.iterate, a.iterate, rtol=1e-11, atol=1e-14)
