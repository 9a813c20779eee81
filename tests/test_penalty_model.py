"""CPU checks of tests/penalty_model.py (the numpy restatement of the reference's loop for EqualityHandling::
PenaltyFunction*, which the reference derives but cannot evaluate): it converges, drives C x - d to zero as mu -> 0 and
lands on the optimum the reference's own SlackedSlacks handling of the same QP reaches (oracle port, bit-for-bit the
reference)."""
import numpy as np
import pytest

import oracle_lib as ol
import penalty_model as pm
import problems as P


@pytest.mark.parametrize("n,me,seed", [(40, 20, 3), (200, 100, 1), (30, 5, 8)])
def test_penalty_iteration_reaches_the_slacked_optimum(n, me, seed):
    p = P.eq_box(n, me, seed)
    tr = pm.solve(p.Q, p.c, p.C, p.d, p.l_x, p.u_x)
    assert tr["converged"] and tr["iterations"] < 20
    assert np.max(np.abs(p.C @ tr["x"] - p.d)) < 1e-8
    ref = ol.port_solve(p, steps=False)
    assert ref.converged
    assert abs(tr["f"][-1] - ref.f[ref.iterations]) < 1e-6 * max(1.0, abs(ref.f[ref.iterations]))
    assert np.max(np.abs(tr["x"] - ref.iterate[:n])) < 1e-5
