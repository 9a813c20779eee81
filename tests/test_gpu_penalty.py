"""EqualityHandling::PenaltyFunction / PenaltyFunctionWithExtraDual on the GPU (IPMZ_EQ_PENALTY): the scalar block
-mu I of SymbolicOptimization.cpp:173-183 that the reference's evaluator cannot assemble (Evaluation.cpp:53-60).  There
is no reference run to compare with; the CUDA path is pinned against tests/penalty_model.py (numpy restatement of the
reference's loop for this system, checked on the CPU against the SlackedSlacks optimum) and the mathematics."""
import numpy as np
import pytest

import penalty_model as pm
import problems as P

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def z():
    import ipm_zoo_b200 as z
    assert z.device_count() > 0, "no CUDA device: the product path has no CPU fallback"
    return z


def relerr(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.mark.parametrize("n,me,seed", [(40, 20, 3), (200, 100, 1), (30, 5, 8)])
@pytest.mark.parametrize("reduction", ["AUGMENTED"])
def test_penalty_rows_match_the_model(z, n, me, seed, reduction):
    p = P.eq_box(n, me, seed)
    tr = pm.solve(p.Q, p.c, p.C, p.d, p.l_x, p.u_x)
    k = tr["iterations"]
    zp = z.Problem(p.Q, p.c, None, None, None, p.C, p.d, p.l_x, p.u_x, z.NONE, z.BOTH, z.EQ_PENALTY)
    s = z.Solver(zp, z.Options(reduction=getattr(z, reduction), record_steps=True))
    r = s.solve()
    t = s.trace(r.iterations, steps=True)
    it = s.iterate()
    s.close()
    assert r.iterations == k and r.converged == tr["converged"]
    assert abs(r.f - tr["f"][k]) <= 1e-8 * max(1.0, abs(tr["f"][k]))
    np.testing.assert_allclose(t["f"][:k + 1], tr["f"], rtol=1e-7, atol=1e-8)
    np.testing.assert_allclose(t["res"][:k + 1], tr["res"], rtol=1e-5, atol=1e-11)
    np.testing.assert_allclose(t["mu"][:k + 1], tr["mu"], rtol=1e-5, atol=1e-13)
    assert relerr(t["step_aff"][0], tr["step_aff"][0]) < 1e-9
    assert relerr(t["step_cor"][0], tr["step_cor"][0]) < 1e-9
    for j in range(1, min(k, 4)):  # later iterations start from iterates that agree to ~1e-9 only
        assert relerr(t["step_aff"][j], tr["step_aff"][j]) < 1e-6
        assert relerr(t["step_cor"][j], tr["step_cor"][j]) < 1e-6
    np.testing.assert_allclose(t["alpha_aff"], tr["alpha_aff"], rtol=1e-7)
    np.testing.assert_allclose(t["sigma"], tr["sigma"], rtol=1e-6, atol=1e-12)
    np.testing.assert_allclose(t["alpha"], tr["alpha"], rtol=1e-7)
    x = it[:n]
    assert np.max(np.abs(x - tr["x"])) < 1e-6
    assert np.max(np.abs(p.C @ x - p.d)) < 1e-8  # mu -> 0: the penalty rows become C x = d


def test_penalty_rows_are_refused_where_not_built(z):
    p = P.eq_box(20, 6, 5)
    zp = z.Problem(p.Q, p.c, None, None, None, p.C, p.d, p.l_x, p.u_x, z.NONE, z.BOTH, z.EQ_PENALTY)
    for red in (z.NORMAL, z.FULL, z.DUAL_NORMAL):
        with pytest.raises(Exception):
            z.Solver(zp, z.Options(reduction=red))


def test_penalty_rows_in_a_batch(z):
    """the fused batch kernel runs the same row formulas: every problem of a batch equals its single solve"""
    n, me, cnt = 48, 12, 6
    ps = [P.eq_box(n, me, 600 + i) for i in range(cnt)]
    stack = lambda k: np.ascontiguousarray(np.stack([getattr(q, k) for q in ps]))
    bp = z.Problem(stack("Q"), stack("c"), None, None, None, stack("C"), stack("d"), stack("l_x"), stack("u_x"),
                   z.NONE, z.BOTH, z.EQ_PENALTY)
    bs = z.BatchSolver(bp, cnt, z.Options(reduction=z.AUGMENTED))
    bs.upload()
    bs.solve(per_problem=False)
    res = bs.results()
    bs.close()
    for i, q in enumerate(ps):
        tr = pm.solve(q.Q, q.c, q.C, q.d, q.l_x, q.u_x)
        assert res[i].iterations == tr["iterations"] and res[i].converged
        assert abs(res[i].f - tr["f"][-1]) <= 1e-8 * max(1.0, abs(tr["f"][-1]))
