"""GPU parity against the UNMODIFIED reference run live in the same process (oracle/_ref/libipmzoo_ref.so: the
reference's own sources compiled where they lie, driven through its public API by oracle/ref_harness.cpp) -- one hop,
not GPU -> port -> reference.  Skipped where the prebuilt reference library is absent.

Reference entry points exercised: Optimizer::solve (Optimizer.cpp:63-220) via ref_solve, LinearSolvers::ldlt_decomposition
/ overwriting_solve_ldlt (LinearSolvers.cpp:14-74) via ref_ldlt / ref_solve_ldlt, symmetric_indefinite_factorization /
overwriting_solve_bunch_kaufman (LinearSolvers.cpp:76-318) via ref_bk_factor / ref_bk_solve.

Tolerances (BASELINE.json north_star): Newton step 1e-9 relative, same iteration count, final objective within 1e-8.
"""
import numpy as np
import pytest

import oracle_lib as ol
import problems as P
from golden.make_golden import CASES

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref not built on this box")]

STEP_TOL = 1e-9
F_TOL = 1e-8
LIVE = ["toy", "ineq_box_20x10", "ineq_box_64x32", "eq_box_40x20", "box_30", "ineq_only_30x12",
        "ineq_lower_box_upper_30x12", "ineq_upper_box_lower_30x12", "portfolio_64"]


@pytest.fixture(scope="module")
def z():
    import ipm_zoo_b200 as z
    assert z.device_count() > 0, "no CUDA device: the product path has no CPU fallback"
    return z


def relerr(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.mark.parametrize("name", LIVE)
@pytest.mark.parametrize("reduction", [0, 1, 2, 3])
def test_solve_against_live_reference(z, name, reduction):
    p = CASES[name]()
    if reduction == 3 and p.m_ineq + p.m_eq == 0:
        pytest.skip("no constraint rows: the dual-Schur normal equations do not exist")
    tr = ol.ref_solve(p)
    k = tr.iterations
    s = z.Solver(z.Problem.from_data(p), z.Options(reduction=reduction, record_steps=True))
    r = s.solve()
    t = s.trace(r.iterations, steps=True)
    x = s.iterate()
    s.close()
    assert r.iterations == k, "iteration count differs from the reference's own solve"
    assert r.converged == bool(tr.converged)
    assert abs(r.f - tr.f[k]) <= F_TOL * max(1.0, abs(tr.f[k]))
    np.testing.assert_allclose(t["f"][:k + 1], tr.f[:k + 1], rtol=1e-7, atol=1e-8)
    np.testing.assert_allclose(t["res"][:k + 1], tr.res[:k + 1], rtol=1e-5, atol=1e-11)
    np.testing.assert_allclose(t["mu"][:k + 1], tr.mu[:k + 1], rtol=1e-5, atol=1e-13)
    # both Newton steps of the first iteration start from identical iterates
    assert relerr(t["step_aff"][0], tr.step_aff[0]) < STEP_TOL
    assert relerr(t["step_cor"][0], tr.step_cor[0]) < STEP_TOL
    assert np.max(np.abs(x[:p.n] - tr.iterate[:p.n])) < 1e-6


@pytest.mark.parametrize("name", ["ineq_box_64x32", "eq_box_40x20", "portfolio_64"])
@pytest.mark.parametrize("reduction", [0, 1, 2, 3])
@pytest.mark.parametrize("k", [1, 3])
def test_newton_step_at_the_references_own_iterate(z, name, reduction, k):
    """The reference's iterate after k of its own iterations is fed to the CUDA path; the steps of iteration k are
    compared with the ones the reference prints for it (`b:` lines, Optimizer.cpp:359)."""
    p = CASES[name]()
    full = ol.ref_solve(p)
    if full.iterations <= k:
        pytest.skip("the reference converged before iteration %d" % k)
    it = ol.ref_solve(p, cap_iters=k, stop_after_cap=True).iterate.copy()
    s = z.Solver(z.Problem.from_data(p), z.Options(reduction=reduction))
    s.set_iterate(it)
    sa, sc, aa, sg, al = s.newton_step()
    s.close()
    assert relerr(sa, full.step_aff[k]) < STEP_TOL, "affine step"
    assert relerr(sc, full.step_cor[k]) < STEP_TOL, "corrector step"


@pytest.mark.parametrize("n", [3, 33, 130, 300])
def test_ldlt_against_live_reference(z, n):
    rng = np.random.default_rng(900 + n)
    M = rng.standard_normal((n, n))
    A = np.ascontiguousarray(M @ M.T / n + np.eye(n))
    L, D = z.ldlt_decomposition(A)
    Lr, Dr = np.zeros((n, n)), np.zeros(n)
    assert ol.ref().ref_ldlt(n, ol._ptr(A), ol._ptr(Lr), ol._ptr(Dr)) == 0
    assert relerr(D, Dr) < 1e-11 and relerr(L, Lr) < 1e-11
    b = rng.standard_normal(n)
    x, xr = b.copy(), b.copy()
    z.overwriting_solve_ldlt(L, D, x)
    assert ol.ref().ref_solve_ldlt(n, ol._ptr(Lr), ol._ptr(Dr), ol._ptr(xr)) == 0
    assert relerr(x, xr) < 1e-10


@pytest.mark.parametrize("n", [5, 64, 257])
def test_bunch_kaufman_bit_exact_against_live_reference(z, n):
    rng = np.random.default_rng(1200 + n)
    K = rng.standard_normal((n, n))
    K = np.ascontiguousarray(K + K.T)
    K[0, 0] = 0.0  # forces an interchange or a 2 x 2 pivot at the first step
    F, piv = z.symmetric_indefinite_factorization(K)
    Fr, pr = np.zeros((n, n)), np.zeros(n, dtype=np.int32)
    assert ol.ref().ref_bk_factor(n, ol._ptr(K), ol._ptr(Fr), pr.ctypes.data_as(ol.C.POINTER(ol.C.c_int))) == 0
    assert np.array_equal(piv, pr), "pivot sequence"
    assert np.array_equal(np.tril(F), np.tril(Fr)), "factor is not bit-exact"


def test_batch_against_live_reference(z):
    """Eight cfg4-shaped problems through the fused batch kernel vs the reference's Optimizer::solve, one by one."""
    n, m, cnt = 96, 48, 8
    ps = [P.ineq_box(n, m, 4000 + i, kind="shift") for i in range(cnt)]
    stack = lambda k: np.ascontiguousarray(np.stack([getattr(q, k) for q in ps]))
    bp = z.Problem(stack("Q"), stack("c"), stack("A"), stack("l_A"), stack("u_A"), None, None, stack("l_x"), stack("u_x"))
    for red in (z.NORMAL, z.AUGMENTED):
        bs = z.BatchSolver(bp, cnt, z.Options(reduction=red))
        bs.upload()
        bs.solve(per_problem=False)
        res = bs.results()
        xs = np.zeros((cnt, n))
        bs.x(xs)
        bs.close()
        for i, q in enumerate(ps):
            tr = ol.ref_solve(q, steps=False)
            k = tr.iterations
            assert res[i].iterations == k and res[i].converged == bool(tr.converged), (i, res[i].iterations, k)
            assert abs(res[i].f - tr.f[k]) <= F_TOL * max(1.0, abs(tr.f[k]))
            assert np.max(np.abs(xs[i] - tr.iterate[:n])) < 1e-6


def test_equalities_and_inequalities_together_against_the_port(z):
    """SlackedSlacks inequalities AND SlackedSlacks equalities in one QP: the reference's evaluator asserts on the zero
    blocks of that 3 x 3 system (Evaluation.cpp:53-60), the plain-C port (bit-for-bit the reference on every family the
    reference can run) solves it with the same formulas; the CUDA path stacks the rows."""
    n, mi, me = 48, 20, 10
    q = P.ineq_box(n, mi, 31)
    e = P.eq_box(n, me, 32)
    # equality rows consistent with the inequality band: d = C x0 for an x0 inside the band of q
    x0 = np.linalg.lstsq(q.A, 0.5 * (q.l_A + q.u_A), rcond=None)[0]
    x0 = np.clip(x0, q.l_x + 0.1, q.u_x - 0.1)
    p = ol.Problem(q.Q, q.c, q.A, q.A @ x0 - 0.25, q.A @ x0 + 0.25, e.C, e.C @ x0, q.l_x, q.u_x,
                   ineq_bounds=ol.BOTH, var_bounds=ol.BOTH, equalities=1)
    tr = ol.port_solve(p)
    k = tr.iterations
    assert tr.converged
    for red in (0, 1, 2, 3):
        s = z.Solver(z.Problem.from_data(p), z.Options(reduction=red, record_steps=True))
        r = s.solve()
        t = s.trace(r.iterations, steps=True)
        s.close()
        assert r.iterations == k and r.converged
        assert abs(r.f - tr.f[k]) <= F_TOL * max(1.0, abs(tr.f[k]))
        assert relerr(t["step_aff"][0], tr.step_aff[0]) < STEP_TOL
        assert relerr(t["step_cor"][0], tr.step_cor[0]) < STEP_TOL
