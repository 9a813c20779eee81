"""CPU check of the FULL-reduction design (no GPU): solving the symmetrised un-reduced Newton
system in the FullLayout order gives the oracle's augmented Newton step, and unpivoted LDL^T in
that order meets only nonzero pivots with the signs of the reference's block elimination
(SymbolicOptimization.cpp:529-567): + for the slack rows, - for their multipliers, + for s."""
import numpy as np
import pytest

import full_model as fm
import oracle_lib as ol
from golden.make_golden import CASES

SMALL = [k for k in sorted(CASES) if CASES[k]().n <= 64]


def _unpivoted_ldlt(K):
    N = K.shape[0]
    L, D = np.eye(N), np.zeros(N)
    A = K.copy()
    for k in range(N):
        D[k] = A[k, k]
        assert D[k] != 0.0
        L[k + 1:, k] = A[k + 1:, k] / D[k]
        A[k + 1:, k + 1:] -= np.outer(L[k + 1:, k], A[k, k + 1:])
    return L, D


@pytest.mark.parametrize("name", SMALL)
@pytest.mark.parametrize("k", [0, 3])
def test_full_system_reproduces_augmented_step(name, k):
    p = CASES[name]()
    it = ol.port_solve(p, cap_iters=k, stop_after_cap=True).iterate.copy() if k else None
    if it is None:
        it = np.zeros(p.iterate_len)
        ol.port().orc_initial_iterate(p.c_struct(), ol._ptr(it))
    tr = ol.port_solve(p, cap_iters=1, stop_after_cap=True, iterate=it)
    K, b, L = fm.full_system(p, it)
    assert np.array_equal(K, K.T)
    Lf, D = _unpivoted_ldlt(K)
    u = np.linalg.solve(Lf.T, np.linalg.solve(Lf, b) / D)
    step = fm.augmented_part(p, u, L)
    err = np.max(np.abs(step - tr.step_aff[0])) / np.max(np.abs(tr.step_aff[0]))
    assert err < 1e-9, err
    # pivot signs: slack groups > 0, multiplier groups < 0, s > 0, x > 0, lambda < 0
    sign = np.sign(D)
    n, m = p.n, (L["x"] - L["s"])
    assert np.all(sign[:L["ly"]] > 0) and np.all(sign[L["s"]:L["lam"]] > 0) and np.all(sign[L["lam"]:] < 0)
    mult = sign[L["ly"]:L["s"]]
    # decoupled identity rows of absent sides have pivot +1; every real multiplier pivot is negative
    real = np.abs(np.diag(K)[L["ly"]:L["s"]]) == 0.0
    assert np.all(mult[real] < 0) and np.all(mult[~real] > 0)


def test_full_layout_size_matches_reference_count():
    """Both-sided rows and box: 5n + 6m unknowns ('full 1600 unknowns' for cfg1)."""
    p = CASES["cfg1_eq_box_200x100"]()
    assert fm.layout(p)["N"] == 5 * 200 + 6 * 100 == 1600
    q = CASES["ineq_only_30x12"]()
    assert fm.layout(q)["N"] == 30 + 6 * 12


@pytest.mark.parametrize("name", sorted(CASES))
def test_library_full_layout_matches_the_model(name):
    """The FullLayout the CUDA kernels index with (host function of ipmz_device.cuh, through the C ABI, no GPU) is
    the layout of the numpy model for every Settings family of the golden cases."""
    import ipm_zoo_b200 as z
    p = CASES[name]()
    assert z.full_layout(z.Problem.from_data(p)) == fm.layout(p)
