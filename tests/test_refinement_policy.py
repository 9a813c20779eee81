"""CPU study behind the refinement policy of the NORMAL reduction (ipm-zoo_b200/csrc/solver.cu, select_refinement):
problems with mu >= 1e-3 skip iterative refinement.  For every golden case and every iterate of the oracle's
trajectory the condensed solve WITHOUT refinement is compared with the augmented solve the reference performs: the
relative step difference must stay two orders below the 1e-9 parity bound while mu >= 1e-3 (it grows like eps / mu
afterwards, which is why the later iterations refine)."""
import numpy as np
import pytest

import oracle_lib as ol
from golden.make_golden import CASES

THRESHOLD = 1e-3  # = thr in select_refinement


@pytest.mark.parametrize("name", [k for k in sorted(CASES) if CASES[k]().N > CASES[k]().n])
def test_condensed_step_without_refinement_is_accurate_above_threshold(name):
    p = CASES[name]()
    full = ol.port_solve(p)
    n, N = p.n, p.N
    worst_above, worst_below = 0.0, 0.0
    for k in range(full.iterations):
        it = np.zeros(p.iterate_len)
        if k:
            it = ol.port_solve(p, cap_iters=k, stop_after_cap=True).iterate.copy()
        else:
            ol.port().orc_initial_iterate(p.c_struct(), ol._ptr(it))
        K, rhs = np.zeros((N, N)), np.zeros(N)
        ol.port().orc_assemble_kkt(p.c_struct(), ol._ptr(it), ol._ptr(K), ol._ptr(rhs))
        H, M, W = K[:n, :n], K[n:, :n], -1.0 / np.diag(K)[n:]
        b0, b1 = rhs[:n], rhs[n:]
        L = np.linalg.cholesky(H + M.T @ (W[:, None] * M))
        dx = np.linalg.solve(L.T, np.linalg.solve(L, b0 + M.T @ (W * b1)))
        step = np.concatenate([dx, W * (M @ dx - b1)])
        ref = np.linalg.solve(K, rhs)
        err = np.max(np.abs(step - ref)) / np.max(np.abs(ref))
        if full.mu[k] >= THRESHOLD:
            worst_above = max(worst_above, err)
        else:
            worst_below = max(worst_below, err)
    assert worst_above < 1e-11, worst_above
    # and refinement is not optional later on: the unrefined step leaves the parity bound
    if full.mu[full.iterations - 1] < 1e-7:
        assert worst_below > 1e-9, worst_below
