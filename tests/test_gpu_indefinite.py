"""EqualityHandling::None on the GPU: equality rows with a multiplier only give an INDEFINITE augmented system
(zero diagonal block).  The reference routes it to solve_indefinite_() == ASSERT(false) (Optimizer.cpp:63-75), so
there is no reference solve to compare with; the path (Bunch-Kaufman factorization, bit-exact against
LinearSolvers.cpp:76-318, see test_gpu_bunch_kaufman.py) is checked through what the mathematics pins:
the Newton step solves the assembled KKT system, a full affine step lands on C x = d, and the optimum equals the
one the reference finds for the SAME QP with EqualityHandling::SlackedSlacks (objective 1e-8 relative, x 1e-6)."""
import numpy as np
import pytest

import oracle_lib as ol
import problems as P

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def z():
    import ipm_zoo_b200 as z
    assert z.device_count() > 0, "no CUDA device: the product path has no CPU fallback"
    return z


def hard(z, p):
    q = z.Problem.from_data(p)
    q.equalities = z.EQ_NONE
    return q


@pytest.mark.parametrize("n,me", [(8, 3), (40, 20), (200, 100), (300, 212)])
def test_newton_step_solves_the_indefinite_kkt_system(z, n, me):
    p = P.eq_box(n, me, 40 + n)
    s = z.Solver(hard(z, p), z.Options(reduction=z.AUGMENTED))
    K = s.assemble()
    assert K.shape == (n + me, n + me)
    assert np.all(K[n:, n:] == 0.0), "EqualityHandling::None leaves a zero diagonal block"
    assert np.array_equal(K[n:, :n], p.C) and np.array_equal(K[:n, n:], p.C.T)
    assert np.array_equal(np.diag(K)[:n], np.diag(p.Q) + 2.0)  # Y^-1 L_y + Z^-1 L_z = 1 + 1 at the initial point
    sa, sc, aa, sg, al = s.newton_step()
    s.close()
    # right-hand side of the predictor at the reference's initial point (x = mid-box, slacks = duals = 1, mu = 0)
    x = 0.5 * (p.l_x + p.u_x)
    one = np.ones(n)
    rx = p.c + one + p.Q @ x + p.C.T @ np.ones(me) - one
    ty = 1.0 * (1.0 - 1.0 * ((p.l_x + one) - x))
    tz = 1.0 * (1.0 - 1.0 * ((x + one) - p.u_x))
    b = np.concatenate([(tz - rx) - ty, -(p.C @ x - p.d)])
    assert np.max(np.abs(K @ sa - b)) <= 1e-9 * np.max(np.abs(b))
    # a full affine Newton step satisfies the linear equalities exactly
    assert np.max(np.abs(p.C @ (x + sa[:n]) - p.d)) < 1e-11
    assert 0.0 < aa <= 1.0 and 0.0 < al <= 1.0


@pytest.mark.parametrize("n,me,seed", [(20, 8, 1), (40, 20, 3), (200, 100, 1)])
def test_indefinite_path_reaches_the_reference_optimum(z, n, me, seed):
    p = P.eq_box(n, me, seed)
    ref = ol.port_solve(p)  # SlackedSlacks handling of the same QP (pinned to the reference bit for bit)
    assert ref.converged
    s = z.Solver(hard(z, p), z.Options(reduction=z.AUGMENTED))
    r = s.solve()
    it = s.iterate()
    s.close()
    assert r.converged and r.iterations <= ref.iterations + 3
    f_ref = ref.f[ref.iterations]
    assert abs(r.f - f_ref) <= 1e-8 * max(1.0, abs(f_ref))
    x = it[:n]
    assert np.max(np.abs(x - ref.iterate[:n])) < 1e-6
    assert np.max(np.abs(p.C @ x - p.d)) < 1e-8
    # stationarity with the recovered multipliers: Q x + c + C^T lam_C - lam_y + lam_z = 0
    off = p.offsets()
    g = lambda k: it[off[k][0]:off[k][0] + off[k][1]]
    rx = p.Q @ x + p.c + p.C.T @ g("lamC") - g("lamy") + g("lamz")
    assert np.max(np.abs(rx)) < 1e-7


def test_indefinite_with_inequalities_and_batch(z):
    """Inequalities (quasi-definite rows) and hard equalities (zero block) together, as a batch of 6."""
    count, n, mi, me = 6, 32, 10, 6
    probs = []
    for i in range(count):
        a = P.ineq_box(n, mi, 900 + i)
        e = P.eq_box(n, me, 950 + i)
        x0 = np.random.default_rng(i).uniform(-0.3, 0.3, n)
        mid = a.A @ x0
        probs.append(ol.Problem(Q=a.Q, c=a.c, A=a.A, l_A=mid - 0.25, u_A=mid + 0.25, Ceq=e.C, d=e.C @ x0,
                                l_x=-np.ones(n), u_x=np.ones(n), equalities=True))
    st = lambda key: np.stack([getattr(q, key) for q in probs])
    bp = z.Problem(st("Q"), st("c"), st("A"), st("l_A"), st("u_A"), st("C"), st("d"), st("l_x"), st("u_x"),
                   equalities=z.EQ_NONE)
    bs = z.BatchSolver(bp, count, z.Options(reduction=z.AUGMENTED))
    res, _ = bs.solve()
    xs = bs.x()
    bs.close()
    for i, q in enumerate(probs):
        one = z.Solver(hard(z, q))
        r1 = one.solve()
        x1 = one.iterate()[:n]
        one.close()
        assert res[i].converged and r1.converged and res[i].iterations == r1.iterations
        assert np.array_equal(xs[i], x1), "batch and single solves run the same arithmetic"
        assert np.max(np.abs(q.C @ xs[i] - q.d)) < 1e-8
        assert np.all(q.A @ xs[i] >= q.l_A - 1e-7) and np.all(q.A @ xs[i] <= q.u_A + 1e-7)


def test_hard_equalities_need_the_augmented_reduction(z):
    p = P.eq_box(12, 4, 2)
    for red in (z.NORMAL, z.FULL):
        with pytest.raises(z.IpmzError) as e:
            z.Solver(hard(z, p), z.Options(reduction=red))
        assert e.value.code == 1


# ---- EqualityHandling::Regularization: rows C x - d + delta p = 0, objective + 1/2 p^T p -------------------------
def reg(z, p):
    q = z.Problem.from_data(p)
    q.equalities = z.EQ_REGULARIZATION
    return q


@pytest.mark.parametrize("reduction", [0, 1])
@pytest.mark.parametrize("n,me", [(8, 3), (40, 20), (200, 100)])
def test_regularization_newton_step_solves_its_kkt_system(z, n, me, reduction):
    """The reference derives this system (SymbolicOptimization.cpp:184-192) but cannot assemble its scalar block
    -delta^2 I (Evaluation.cpp:53-60); the step is checked against the system written out in numpy."""
    delta = 1e-4
    p = P.eq_box(n, me, 60 + n)
    s = z.Solver(reg(z, p), z.Options(reduction=reduction, delta_eq=delta))
    sa, sc, aa, sg, al = s.newton_step()
    if reduction == 0:
        K = s.assemble()
        assert np.array_equal(K[n:, n:], -delta * delta * np.eye(me))
    s.close()
    x = 0.5 * (p.l_x + p.u_x)
    one, lam, pv = np.ones(n), np.ones(me), np.ones(me)
    K = np.block([[p.Q + 2.0 * np.eye(n), p.C.T], [p.C, -delta * delta * np.eye(me)]])
    rx = p.c + one + p.Q @ x + p.C.T @ lam - one
    ty = 1.0 * (1.0 - 1.0 * ((p.l_x + one) - x))
    tz = 1.0 * (1.0 - 1.0 * ((x + one) - p.u_x))
    r_lam = p.C @ x - p.d + delta * pv
    r_p = pv + delta * lam
    b = np.concatenate([(tz - rx) - ty, delta * r_p - r_lam])
    ref = np.linalg.solve(K, b)
    assert np.max(np.abs(sa - ref)) <= 1e-9 * np.max(np.abs(ref))


@pytest.mark.parametrize("reduction", [0, 1])
def test_regularization_reaches_the_kkt_point_of_the_regularised_qp(z, reduction):
    delta = 1e-4
    n, me = 40, 20
    p = P.eq_box(n, me, 3)
    ref = ol.port_solve(p)  # the un-regularised optimum (reference, SlackedSlacks)
    s = z.Solver(reg(z, p), z.Options(reduction=reduction, delta_eq=delta))
    r = s.solve()
    it = s.iterate()
    s.close()
    assert r.converged and r.iterations <= ref.iterations + 3
    off = p.offsets()
    g = lambda k: it[off[k][0]:off[k][0] + off[k][1]]
    x, lam, pv = it[:n], g("lamC"), g("t")  # p travels in the `t` slot
    assert np.max(np.abs(p.C @ x - p.d + delta * pv)) < 1e-8
    assert np.max(np.abs(pv + delta * lam)) < 1e-8
    assert np.max(np.abs(p.Q @ x + p.c + p.C.T @ lam - g("lamy") + g("lamz"))) < 1e-7
    # delta^2 |lambda| ~ 1e-8 perturbation of the constraint: the optimum moves by that order
    assert np.max(np.abs(x - ref.iterate[:n])) < 1e-5
    assert abs(r.f - ref.f[ref.iterations]) < 1e-5 * max(1.0, abs(r.f))
