"""GPU parity tests proper: the CUDA path (through the C ABI, ipm-zoo_b200/capi.py) against the
CPU oracle (oracle/ipm_oracle.c, itself pinned bit-for-bit to the unmodified reference) and the
reference-generated goldens, on identical seeded inputs.

Tolerances (BASELINE.json north_star): Newton step 1e-9 relative, same iteration count, final
objective within 1e-8.  Steps are compared norm-wise: max|d_gpu - d_ref| / max|d_ref|.
"""
import os

import numpy as np
import pytest

import oracle_lib as ol
import problems as P
from golden.make_golden import CASES

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
STEP_TOL = 1e-9
F_TOL = 1e-8


@pytest.fixture(scope="module")
def z():
    import ipm_zoo_b200 as z
    assert z.device_count() > 0, "no CUDA device: the product path has no CPU fallback"
    return z


def relerr(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def quasidef(rng, n, m):
    M = rng.standard_normal((n, n))
    H = M @ M.T / n + np.eye(n)
    A = rng.standard_normal((m, n)) / np.sqrt(n)
    return np.block([[H, A.T], [A, -np.diag(rng.uniform(0.5, 2.0, m))]])


@pytest.mark.parametrize("n", [1, 2, 5, 31, 32, 33, 64, 100, 127, 128, 129, 200, 257, 300, 515])
def test_ldlt_matches_oracle_spd(z, n):
    rng = np.random.default_rng(100 + n)
    M = rng.standard_normal((n, n))
    A = M @ M.T / n + np.eye(n)
    L, D = z.ldlt_decomposition(A)
    Lo, Do = np.zeros((n, n)), np.zeros(n)
    ol.port().orc_ldlt(n, ol._ptr(np.ascontiguousarray(A)), ol._ptr(Lo), ol._ptr(Do))
    assert relerr(D, Do) < 1e-11
    assert relerr(L, Lo) < 1e-11
    assert np.all(np.triu(L, 1) == 0.0) and np.all(np.diag(L) == 1.0)
    b = rng.standard_normal(n)
    x = b.copy()
    z.overwriting_solve_ldlt(L, D, x)
    xo = b.copy()
    ol.port().orc_solve_ldlt(n, ol._ptr(Lo), ol._ptr(Do), ol._ptr(xo))
    assert relerr(x, xo) < 1e-10
    assert np.max(np.abs(A @ x - b)) < 1e-9 * max(1.0, np.max(np.abs(b)))


@pytest.mark.parametrize("n,m", [(3, 2), (40, 17), (128, 64), (200, 100), (300, 140)])
def test_ldlt_matches_oracle_quasidefinite(z, n, m):
    rng = np.random.default_rng(7 * n + m)
    K = np.ascontiguousarray(quasidef(rng, n, m))
    N = n + m
    L, D = z.ldlt_decomposition(K)
    Lo, Do = np.zeros((N, N)), np.zeros(N)
    ol.port().orc_ldlt(N, ol._ptr(K), ol._ptr(Lo), ol._ptr(Do))
    assert np.all(D[:n] > 0) and np.all(D[n:] < 0)
    assert relerr(D, Do) < 1e-10
    assert relerr(L, Lo) < 1e-10
    b = rng.standard_normal(N)
    x = b.copy()
    z.overwriting_solve_ldlt(L, D, x)
    assert np.max(np.abs(K @ x - b)) < 1e-9


def test_ldlt_golden_reference_vectors(z):
    g = np.load(os.path.join(GOLD, "linear_solvers.npz"))
    for nm in ("spd", "quasidef"):
        L, D = z.ldlt_decomposition(g[nm + "_K"])
        assert relerr(L, g[nm + "_L"]) < 1e-11
        assert relerr(D, g[nm + "_D"]) < 1e-11
        x = g[nm + "_b"].copy()
        z.overwriting_solve_ldlt(g[nm + "_L"], g[nm + "_D"], x)
        assert relerr(x, g[nm + "_x"]) < 1e-10


def test_ldlt_zero_pivot_guard(z):
    """LinearSolvers.cpp:28: exactly-zero pivot -> 1e-8."""
    L, D = z.ldlt_decomposition(np.array([[0.0, 0.0], [0.0, 2.0]]))
    assert D[0] == 1e-8 and D[1] == 2.0


def test_solve_empty_rhs_is_noop(z):
    """LinearSolvers.cpp:46-48."""
    b = np.zeros(0)
    z.overwriting_solve_ldlt(np.zeros((0, 0)), np.zeros(0), b)


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("reduction", ["augmented", "normal"])
def test_assemble_matches_oracle(z, name, reduction):
    p = CASES[name]()
    if p.n > 100:
        pytest.skip("covered by the solve tests")
    it0 = np.zeros(p.iterate_len)
    ps = p.c_struct()
    ol.port().orc_initial_iterate(ps, ol._ptr(it0))
    N = p.N
    K = np.zeros((N, N)); rhs = np.zeros(N)
    ol.port().orc_assemble_kkt(ps, ol._ptr(it0), ol._ptr(K), ol._ptr(rhs))
    s = z.Solver(z.Problem.from_data(p), z.Options(reduction=z.NORMAL if reduction == "normal" else z.AUGMENTED))
    Kg = s.assemble()
    if reduction == "augmented":
        assert Kg.shape == (N, N)
        assert relerr(Kg, K) < 1e-14
    else:
        n = p.n
        H, A, Wd = K[:n, :n], K[n:, :n], -1.0 / np.diag(K)[n:]
        ref = H + A.T @ (Wd[:, None] * A) if N > n else H
        assert relerr(Kg, ref) < 1e-12
    s.close()


@pytest.mark.parametrize("name", [k for k in sorted(CASES) if CASES[k]().n <= 64])
@pytest.mark.parametrize("k", [0, 3])
def test_assemble_full_matches_model(z, name, k):
    """FULL reduction: the CUDA assembly of the symmetrised un-reduced Newton system equals the numpy model
    (tests/full_model.py, itself checked against the oracle's step on the CPU) entry by entry."""
    import full_model as fm
    p = CASES[name]()
    it = np.zeros(p.iterate_len)
    if k:
        it = ol.port_solve(p, cap_iters=k, stop_after_cap=True).iterate.copy()
    else:
        ol.port().orc_initial_iterate(p.c_struct(), ol._ptr(it))
    K, _, L = fm.full_system(p, it)
    s = z.Solver(z.Problem.from_data(p), z.Options(reduction=z.FULL))
    s.set_iterate(it)
    Kg = s.assemble()
    s.close()
    assert Kg.shape == K.shape == (L["N"], L["N"])
    assert np.array_equal(Kg != 0.0, K != 0.0), "sparsity pattern"
    assert relerr(Kg, K) < 1e-15
    assert np.array_equal(Kg, Kg.T)


def _needs_rows(p, reduction):
    """DUAL_NORMAL (3) eliminates dx onto the constraint rows: systems without rows have no such reduction."""
    if reduction == 3 and p.m_ineq + p.m_eq == 0:
        pytest.skip("no constraint rows: the dual-Schur normal equations do not exist")


def _step_parity(z, p, iterate, reduction, tol=STEP_TOL):
    _needs_rows(p, reduction)
    tr = ol.port_solve(p, cap_iters=1, stop_after_cap=True, iterate=iterate)
    s = z.Solver(z.Problem.from_data(p), z.Options(reduction=reduction))
    if iterate is not None:
        s.set_iterate(iterate)
    sa, sc, aa, sg, al = s.newton_step()
    s.close()
    assert relerr(sa, tr.step_aff[0]) < tol, "affine step"
    assert relerr(sc, tr.step_cor[0]) < tol, "corrector step"
    assert abs(aa - tr.alpha_aff[0]) < 1e-9 * max(1.0, abs(tr.alpha_aff[0]))
    assert abs(sg - tr.sigma[0]) < 1e-8
    assert abs(al - tr.alpha[0]) < 1e-9 * max(1.0, abs(tr.alpha[0]))


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("reduction", [0, 1, 2, 3])
def test_newton_step_at_initial_point(z, name, reduction):
    _step_parity(z, CASES[name](), None, reduction)


@pytest.mark.parametrize("name", ["ineq_box_64x32", "eq_box_40x20", "cfg1_eq_box_200x100", "cfg4_unit_256x128"])
@pytest.mark.parametrize("reduction", [0, 1, 2, 3])
@pytest.mark.parametrize("k", [2, 4])
def test_newton_step_at_reference_iterate(z, name, reduction, k):
    """Feed the CUDA path the oracle's iterate after k iterations, compare that iteration's steps."""
    p = CASES[name]()
    it = ol.port_solve(p, cap_iters=k, stop_after_cap=True).iterate.copy()
    _step_parity(z, p, it, reduction)


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("reduction", [0, 1, 2, 3])
def test_full_solve_matches_reference_golden(z, name, reduction):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    p = CASES[name]()
    _needs_rows(p, reduction)
    s = z.Solver(z.Problem.from_data(p), z.Options(reduction=reduction, record_steps=True))
    r = s.solve()
    k = int(g["iterations"])
    assert r.iterations == k, "iteration count differs from the reference"
    assert r.converged == bool(g["converged"])
    assert abs(r.f - g["f"][k]) <= F_TOL * max(1.0, abs(g["f"][k]))
    tr = s.trace(k, steps=True)
    np.testing.assert_allclose(tr["f"][:k + 1], g["f"], rtol=1e-7, atol=1e-8)
    # residual norm and mean complementarity of every iteration (Optimizer.cpp:129-130) against the reference's trace:
    # the trajectories share every Newton step to ~1e-9, entries near the 1e-8 tolerance get an absolute allowance
    np.testing.assert_allclose(tr["res"][:k + 1], g["res"], rtol=1e-5, atol=1e-11)
    np.testing.assert_allclose(tr["mu"][:k + 1], g["mu"], rtol=1e-5, atol=1e-13)
    # first iteration's steps start from identical iterates
    assert relerr(tr["step_aff"][0], g["step_aff"][0]) < STEP_TOL
    assert relerr(tr["step_cor"][0], g["step_cor"][0]) < STEP_TOL
    x = s.iterate()
    n = p.n
    assert np.max(np.abs(x[:n] - g["iterate"][:n])) < 1e-6
    s.close()


def test_warm_start_second_solve_is_immediate(z):
    p = CASES["ineq_box_20x10"]()
    s = z.Solver(z.Problem.from_data(p))
    r1 = s.solve()
    r2 = s.solve()
    assert r1.converged and r2.converged and r2.iterations == 0
    s.close()


@pytest.mark.parametrize("reduction", [0, 1, 2, 3])
def test_batch_matches_oracle_per_problem(z, reduction):
    count, n, m = 12, 48, 20
    probs = [P.ineq_box(n, m, 2000 + i, kind="shift") for i in range(count)]
    st = lambda key: np.stack([getattr(q, key) for q in probs])
    bp = z.Problem(st("Q"), st("c"), st("A"), st("l_A"), st("u_A"), None, None, st("l_x"), st("u_x"))
    bs = z.BatchSolver(bp, count, z.Options(reduction=reduction))
    res, ms = bs.solve()
    xs = bs.x()
    for i, q in enumerate(probs):
        tr = ol.port_solve(q)
        assert res[i].iterations == tr.iterations, i
        assert res[i].converged == bool(tr.converged)
        assert abs(res[i].f - tr.f[tr.iterations]) <= F_TOL * max(1.0, abs(tr.f[tr.iterations]))
        assert np.max(np.abs(xs[i] - tr.iterate[:n])) < 1e-6
    bs.close()


@pytest.mark.parametrize("n,m", [(37, 15), (1, 1), (65, 0), (130, 61), (257, 127), (300, 211)])
@pytest.mark.parametrize("reduction", [0, 1])
def test_batch_ragged_shapes_match_oracle(z, n, m, reduction):
    """Shapes that are not multiples of the fused batch kernel's 64 x 64 factor tiles, 32-wide panels, 16-row assembly
    slices or 4-row matvec groups (odd N, N = 1, no constraint rows, N just above a tile / panel edge, N = 511 in the
    AUGMENTED reduction = the largest the kernel takes), fewer problems than resident CTAs."""
    count = 5
    if m > 0:
        probs = [P.ineq_box(n, m, 3000 + 7 * n + i, kind="shift") for i in range(count)]
        st = lambda key: np.stack([getattr(q, key) for q in probs])
        bp = z.Problem(st("Q"), st("c"), st("A"), st("l_A"), st("u_A"), None, None, st("l_x"), st("u_x"))
    else:
        probs = [P.box_only(n, 3000 + i) for i in range(count)]
        st = lambda key: np.stack([getattr(q, key) for q in probs])
        bp = z.Problem(st("Q"), st("c"), None, None, None, None, None, st("l_x"), st("u_x"))
    bs = z.BatchSolver(bp, count, z.Options(reduction=reduction))
    res, ms = bs.solve()
    xs = bs.x()
    bs.close()
    for i, q in enumerate(probs):
        tr = ol.port_solve(q, steps=False)
        assert res[i].iterations == tr.iterations, (i, res[i].iterations, tr.iterations)
        assert res[i].converged == bool(tr.converged)
        assert abs(res[i].f - tr.f[tr.iterations]) <= F_TOL * max(1.0, abs(tr.f[tr.iterations]))
        assert np.max(np.abs(xs[i] - tr.iterate[:n])) < 1e-6


@pytest.mark.parametrize("cap", [0, 1, 3])
def test_batch_iteration_cap(z, cap):
    """The iteration cap inside the persistent batch kernel (Optimizer.cpp:125: the loop just ends): every problem
    stops after `cap` iterations, unconverged, with the iterate the oracle has at that point, and the kernel's queue
    drains (a problem that hits the cap leaves it like one that converged)."""
    count, n, m = 9, 40, 16
    probs = [P.ineq_box(n, m, 5000 + i, kind="shift") for i in range(count)]
    st = lambda key: np.stack([getattr(q, key) for q in probs])
    bp = z.Problem(st("Q"), st("c"), st("A"), st("l_A"), st("u_A"), None, None, st("l_x"), st("u_x"))
    bs = z.BatchSolver(bp, count, z.Options(reduction=z.AUGMENTED, max_iter=cap))
    res, ms = bs.solve()
    xs = bs.x()
    bs.close()
    for i, q in enumerate(probs):
        assert res[i].iterations == cap and not res[i].converged
        if cap > 0:
            tr = ol.port_solve(q, cap_iters=cap, stop_after_cap=True, steps=False)
            assert np.max(np.abs(xs[i] - tr.iterate[:n])) < 1e-9


def test_error_conventions(z):
    """Reference: ASSERT(l < u) (EnvironmentBuilder.cpp:10-17) and solve_indefinite_ == ASSERT(false)."""
    p = CASES["box_30"]()
    bad = z.Problem.from_data(p)
    bad.u_x = bad.l_x.copy()
    with pytest.raises(z.IpmzError) as e:
        z.Solver(bad)
    assert e.value.code == 5
    q = CASES["eq_box_40x20"]()
    zp = z.Problem.from_data(q)
    zp.equalities = False
    with pytest.raises(z.IpmzError) as e:
        z.Solver(zp)
    assert e.value.code == 4


def test_batch_solve_group_matches_single_batch():
    """ipmz_batch_solve_group (several handles of one device on their own host threads / streams) gives
    exactly the results of one batch over the same problems."""
    import ipm_zoo_b200 as z
    import problems as P
    n, m, count = 48, 20, 12
    probs = [P.ineq_box(n, m, 700 + i) for i in range(count)]
    st = lambda key, lo, hi: np.stack([getattr(q, key) for q in probs[lo:hi]])
    mk = lambda lo, hi: z.Problem(st("Q", lo, hi), st("c", lo, hi), st("A", lo, hi), st("l_A", lo, hi),
                                  st("u_A", lo, hi), None, None, st("l_x", lo, hi), st("u_x", lo, hi))
    one = z.BatchSolver(mk(0, count), count)
    res, _ = one.solve()
    x_one = one.x()
    one.close()
    cuts = [0, 5, 9, 12]
    subs = [z.BatchSolver(mk(cuts[g], cuts[g + 1]), cuts[g + 1] - cuts[g]) for g in range(3)]
    ms = z.solve_group(subs)
    assert ms > 0
    got = [r for s in subs for r in s.results()]
    x_grp = np.concatenate([s.x() for s in subs])
    for s in subs:
        s.close()
    assert [r.iterations for r in got] == [r.iterations for r in res]
    assert all(r.converged for r in got)
    assert np.array_equal(x_grp, x_one)


def test_reset_iterate_after_set_iterate_is_the_fresh_initial_pack(z):
    """ipmz_reset_iterate after ipmz_set_iterate reproduces the fresh initial pack bit for bit -- including the slots
    of EqualityHandling::Regularization rows that no kernel reads (they used to keep what set_iterate wrote)."""
    for p, eq in ((CASES["ineq_box_20x10"](), None), (CASES["eq_box_40x20"](), 3)):
        zp = z.Problem.from_data(p)
        if eq is not None:
            zp.equalities = eq  # IPMZ_EQ_REGULARIZATION
        s = z.Solver(zp)
        fresh = s.iterate()
        s.set_iterate(np.full(zp.iterate_len, 7.25))
        assert not np.array_equal(s.iterate(), fresh)
        s.reset_iterate()
        assert np.array_equal(s.iterate(), fresh)
        s.close()


@pytest.mark.parametrize("reduction", [0, 1])
def test_batch_streamed_solve_matches_upload_then_solve(z, reduction):
    """ipmz_batch_solve_streamed (kernel launched first, problems consumed as the copy stream delivers them, initial
    point and M^T built inside the kernel) gives bitwise the results of ipmz_batch_upload + ipmz_batch_solve; a second
    solve without a new upload is a warm start (0 iterations), not a stale copy of the first."""
    count, n, m = 40, 48, 20
    probs = [P.ineq_box(n, m, 900 + i, kind="shift") for i in range(count)]
    st = lambda key: np.stack([getattr(q, key) for q in probs])
    bp = z.Problem(st("Q"), st("c"), st("A"), st("l_A"), st("u_A"), None, None, st("l_x"), st("u_x"))
    bs = z.BatchSolver(bp, count, z.Options(reduction=reduction))
    res, ms = bs.solve()
    x_ref = bs.x()
    assert all(r.converged for r in res) and ms > 0
    again, _ = bs.solve()
    assert all(r.converged and r.iterations == 0 for r in again)
    for chunks in (1, 7):
        got, _ = bs.solve_streamed(chunks=chunks)
        assert [r.iterations for r in got] == [r.iterations for r in res]
        assert [r.f for r in got] == [r.f for r in res]
        assert np.array_equal(bs.x(), x_ref)
    bs.close()


def test_batch_iteration_queue_is_bitwise_the_problem_granular_schedule(z, tmp_path):
    """k_ipm_batch's default work distribution hands a problem from CTA to CTA between iterations (device FIFO, hand-over
    by __threadfence + queue slot).  A problem's arithmetic does not depend on who runs it: 700 problems (more than twice
    the resident CTAs, so every problem migrates between SMs) must give bitwise the iterates of the schedule that keeps
    a problem on one CTA (IPMZ_FUSED_QUEUE=0, read once per process: run in a child), twice in a row."""
    import subprocess
    import sys
    count, n, m = 700, 40, 16
    code = (
        "import sys, numpy as np\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import ipm_zoo_b200 as z, problems as P\n"
        "count, n, m = %d, %d, %d\n"
        "probs = [P.ineq_box(n, m, 7000 + i, kind='shift') for i in range(count)]\n"
        "st = lambda key: np.stack([getattr(q, key) for q in probs])\n"
        "bp = z.Problem(st('Q'), st('c'), st('A'), st('l_A'), st('u_A'), None, None, st('l_x'), st('u_x'))\n"
        "bs = z.BatchSolver(bp, count, z.Options(reduction=z.NORMAL))\n"
        "out = []\n"
        "for rep in range(2):\n"
        "    bs.upload(); res, ms = bs.solve(); out.append(bs.x().copy())\n"
        "    assert all(r.converged for r in res)\n"
        "assert np.array_equal(out[0], out[1])\n"
        "np.save(sys.argv[1], out[0]); np.save(sys.argv[2], np.array([r.iterations for r in res]))\n"
    ) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)), count, n, m)
    outs = {}
    for mode in ("1", "0"):
        env = dict(os.environ, IPMZ_FUSED_QUEUE=mode)
        fx, fi = str(tmp_path / ("x%s.npy" % mode)), str(tmp_path / ("i%s.npy" % mode))
        r = subprocess.run([sys.executable, "-c", code, fx, fi], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs[mode] = (np.load(fx), np.load(fi))
    assert np.array_equal(outs["1"][1], outs["0"][1]), "iteration counts"
    assert np.array_equal(outs["1"][0], outs["0"][0]), "iterates differ between the two work distributions"


def test_batch_upload_rejects_a_different_shape(z):
    count, n, m = 4, 24, 8
    probs = [P.ineq_box(n, m, 50 + i) for i in range(count)]
    st = lambda key: np.stack([getattr(q, key) for q in probs])
    bp = z.Problem(st("Q"), st("c"), st("A"), st("l_A"), st("u_A"), None, None, st("l_x"), st("u_x"))
    bs = z.BatchSolver(bp, count)
    other = [P.ineq_box(n + 4, m, 60 + i) for i in range(count)]
    so = lambda key: np.stack([getattr(q, key) for q in other])
    bad = z.Problem(so("Q"), so("c"), so("A"), so("l_A"), so("u_A"), None, None, so("l_x"), so("u_x"))
    with pytest.raises(z.IpmzError) as e:
        bs.upload(bad)
    assert e.value.code == 1
    bs.close()


@pytest.mark.parametrize("name", ["ineq_box_64x32", "eq_box_40x20", "cfg4_unit_256x128", "ineq_only_30x12"])
def test_dual_schur_matrix_matches_the_oracle_kkt(z, name):
    """IPMZ_REDUCTION_DUAL_NORMAL: the matrix the CUDA path factorizes is the reference's normal-equations block
    -(W^-1 + M Hx^-1 M^T) (SymbolicOptimization.cpp:465-478), sign flipped: checked against the Schur complement of the
    oracle's assembled augmented KKT matrix at the oracle's iterate after two iterations."""
    p = CASES[name]()
    it = ol.port_solve(p, cap_iters=2, stop_after_cap=True).iterate.copy()
    N, n = p.N, p.n
    K = np.zeros((N, N))
    ol.port().orc_assemble_kkt(p.c_struct(), ol._ptr(it), ol._ptr(K), None)
    S_ref = -(K[n:, n:] - K[n:, :n] @ np.linalg.solve(K[:n, :n], K[:n, n:]))
    s = z.Solver(z.Problem.from_data(p), z.Options(reduction=z.DUAL_NORMAL))
    s.set_iterate(it)
    S = s.assemble()
    s.close()
    assert S.shape == (N - n, N - n)
    assert relerr(S, S_ref) < 1e-11
    assert np.all(np.linalg.eigvalsh(S) > 0)
