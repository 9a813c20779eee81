"""Parity at BASELINE.json's sizes through size-independent properties (the CPU oracle would take
minutes there): KKT residuals of the converged iterate, agreement of the two reductions with each
other (same iteration count, objective within 1e-8), factor residuals ||L D L^T v - A v||, and a
sample of the cfg4 batch against the reference-generated golden."""
import os

import numpy as np
import pytest

import oracle_lib as ol
import problems as P

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def z():
    import ipm_zoo_b200 as z
    assert z.device_count() > 0
    return z


def kkt_residuals(p, it):
    """Primal/dual residuals and complementarity of a packed iterate (ineq + box, two-sided)."""
    o = p.offsets()
    g = lambda k: it[o[k][0]:o[k][0] + o[k][1]]
    x, lamA, s = g("x"), g("lamA"), g("s")
    r_x = p.c + g("lamz") + p.Q @ x + p.A.T @ lamA - g("lamy")
    r_p = np.concatenate([p.A @ x - s, p.l_A + g("g") - s, g("h") + s - p.u_A, p.l_x + g("y") - x, x + g("z") - p.u_x])
    comp = np.concatenate([g("g") * g("lamg"), g("h") * g("lamh"), g("y") * g("lamy"), g("z") * g("lamz")])
    nonneg = np.concatenate([g("g"), g("h"), g("y"), g("z"), g("lamg"), g("lamh"), g("lamy"), g("lamz")])
    return np.linalg.norm(r_x), np.linalg.norm(r_p), np.mean(np.abs(comp)), nonneg.min()


def test_cfg2_augmented_vs_normal_vs_full_n2048(z):
    """cfg2 at full size in all three reductions (FULL: 5n + 6m = 16384 unknowns, one 2.1 GB LDL^T per iteration)."""
    p = P.ineq_box(2048, 1024, 2, kind="shift")
    out = {}
    for red in (z.AUGMENTED, z.NORMAL, z.FULL):
        s = z.Solver(z.Problem.from_data(p), z.Options(reduction=red))
        r = s.solve()
        it = s.iterate()
        s.close()
        assert r.converged and r.res < 1e-8 and r.mu < 1e-8
        rd, rp, mu, mn = kkt_residuals(p, it)
        assert rd < 1e-8 and rp < 1e-8 and mu < 1e-8 and mn > 0.0
        out[red] = (r, it)
    ra = out[z.AUGMENTED][0]
    n = p.n
    for red in (z.NORMAL, z.FULL):
        rn = out[red][0]
        assert ra.iterations == rn.iterations
        assert abs(ra.f - rn.f) <= 1e-8 * max(1.0, abs(ra.f))
        assert np.max(np.abs(out[z.AUGMENTED][1][:n] - out[red][1][:n])) < 1e-7


@pytest.mark.parametrize("n", [1000, 4096])
def test_factor_solve_residuals_large(z, n):
    rng = np.random.default_rng(n)
    S = rng.standard_normal((n, n)) / np.sqrt(n)
    A = 3.0 * np.eye(n) + 0.5 * (S + S.T)
    f = z.Factor(n)
    f.set_matrix(A)
    b = rng.standard_normal(n)
    f.set_rhs(b)
    f.run(1, 2)
    x = f.solution()
    assert np.max(np.abs(A @ x - b)) / np.max(np.abs(b)) < 1e-12
    L, D = f.ld()
    v = rng.standard_normal(n)
    assert np.max(np.abs(L @ (D * (L.T @ v)) - A @ v)) / np.max(np.abs(A @ v)) < 1e-12
    assert np.all(D > 0)
    f.close()


def test_quasidefinite_factor_n3072(z):
    """cfg2's augmented KKT shape (N = 3072): signs of the pivots and the solve residual."""
    rng = np.random.default_rng(5)
    n, m = 2048, 1024
    S = rng.standard_normal((n, n)) / np.sqrt(n)
    H = 3.0 * np.eye(n) + 0.5 * (S + S.T)
    A = rng.standard_normal((m, n)) / np.sqrt(n)
    K = np.block([[H, A.T], [A, -np.diag(rng.uniform(0.1, 10.0, m))]])
    f = z.Factor(n + m)
    f.set_matrix(K)
    b = rng.standard_normal(n + m)
    f.set_rhs(b)
    f.run(1, 1)
    x = f.solution()
    _, D = f.ld()
    assert np.all(D[:n] > 0) and np.all(D[n:] < 0)
    assert np.max(np.abs(K @ x - b)) / np.max(np.abs(b)) < 1e-11
    f.close()


def test_cfg5_portfolio_vs_oracle(z):
    p = P.portfolio(512, 32, 1e-6, 5)
    tr = ol.port_solve(p, steps=False)
    for red in (z.AUGMENTED, z.NORMAL):
        s = z.Solver(z.Problem.from_data(p), z.Options(reduction=red))
        r = s.solve()
        x = s.iterate()[:p.n]
        s.close()
        assert r.converged == bool(tr.converged)
        assert r.iterations == tr.iterations
        assert abs(r.f - tr.f[tr.iterations]) <= 1e-8 * max(1.0, abs(r.f))
        assert abs(x.sum() - 1.0) < 1e-7 and x.min() > -1e-9


def test_cfg4_batch_sample_vs_golden_and_oracle(z):
    """32 QPs of the cfg4 shape; problem 0 is the reference-generated golden (seed 1000)."""
    count, n, m = 32, 256, 128
    probs = [P.ineq_box(n, m, 1000 + i, kind="shift") for i in range(count)]
    st = lambda key: np.stack([getattr(q, key) for q in probs])
    bp = z.Problem(st("Q"), st("c"), st("A"), st("l_A"), st("u_A"), None, None, st("l_x"), st("u_x"))
    g = np.load(os.path.join(GOLD, "cfg4_unit_256x128.npz"))
    for red in (z.NORMAL, z.AUGMENTED):
        bs = z.BatchSolver(bp, count, z.Options(reduction=red))
        res, _ = bs.solve()
        xs = bs.x()
        bs.close()
        k = int(g["iterations"])
        assert res[0].iterations == k and res[0].converged
        assert abs(res[0].f - g["f"][k]) <= 1e-8 * max(1.0, abs(g["f"][k]))
        assert np.max(np.abs(xs[0] - g["iterate"][:n])) < 1e-6
        for i in (5, 17, 31):
            tr = ol.port_solve(probs[i], steps=False)
            assert res[i].iterations == tr.iterations
            assert abs(res[i].f - tr.f[tr.iterations]) <= 1e-8 * max(1.0, abs(res[i].f))
        assert all(r.converged for r in res)


def test_batch_ragged_convergence(z):
    """Problems that converge in different iteration counts exercise the active-list compaction."""
    n, m = 40, 16
    probs = [P.ineq_box(n, m, 300 + i, band=(0.05 if i % 3 == 0 else 0.5)) for i in range(9)]
    st = lambda key: np.stack([getattr(q, key) for q in probs])
    bp = z.Problem(st("Q"), st("c"), st("A"), st("l_A"), st("u_A"), None, None, st("l_x"), st("u_x"))
    bs = z.BatchSolver(bp, len(probs))
    res, _ = bs.solve()
    its = []
    for i, q in enumerate(probs):
        tr = ol.port_solve(q, steps=False)
        its.append(tr.iterations)
        assert res[i].iterations == tr.iterations and res[i].converged == bool(tr.converged)
        assert abs(res[i].f - tr.f[tr.iterations]) <= 1e-8 * max(1.0, abs(res[i].f))
    assert len(set(its)) > 1, "test needs differing iteration counts"
    bs.close()


@pytest.mark.parametrize("n", [130, 257, 1000, 1536])
def test_dataflow_path_ragged_sizes(n):
    """The persistent dataflow factorization + streaming solves at sizes below their default
    threshold (ragged last tile, 64-row half tiles), in a subprocess with IPMZ_DATAFLOW_MIN_N=128,
    against numpy: factor residual, solve residual, and the quasi-definite sign pattern."""
    import subprocess
    import sys
    code = r'''
import numpy as np, ipm_zoo_b200 as z
n = %d
rng = np.random.default_rng(n)
m = n // 3
S = rng.standard_normal((n - m, n - m)) / np.sqrt(n)
H = 2.0 * np.eye(n - m) + 0.5 * (S + S.T)
A = rng.standard_normal((m, n - m)) / np.sqrt(n)
K = np.block([[H, A.T], [A, -np.diag(rng.uniform(0.5, 2.0, m))]])
f = z.Factor(n)
assert f.info()["dataflow"], f.info()
f.set_matrix(K); b = rng.standard_normal(n); f.set_rhs(b); f.run(1, 2)
x = f.solution(); L, D = f.ld()
v = rng.standard_normal(n)
r1 = np.max(np.abs(K @ x - b)) / np.max(np.abs(b))
r2 = np.max(np.abs(L @ (D * (L.T @ v)) - K @ v)) / np.max(np.abs(K @ v))
assert r1 < 1e-11 and r2 < 1e-12, (r1, r2)
assert np.all(D[:n - m] > 0) and np.all(D[n - m:] < 0)
print("ok", r1, r2)
''' % n
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, IPMZ_DATAFLOW_MIN_N="128", PYTHONPATH=root)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ok" in out.stdout, out.stdout + out.stderr


def test_cfg5_portfolio_full_size_vs_golden(z):
    """cfg5 at BASELINE.json's size (n = 4096, N = 4097 augmented: a one-row ragged tile in the
    dataflow factorization and the streaming solves) against the oracle trace committed by
    tests/golden/make_golden_cfg5.py: same iteration count, objective within 1e-8, same x."""
    g = np.load(os.path.join(GOLD, "cfg5_portfolio_4096.npz"))
    p = P.portfolio(4096, 32, 1e-6, 5)
    k = int(g["iterations"])
    for red in (z.AUGMENTED, z.NORMAL):
        s = z.Solver(z.Problem.from_data(p), z.Options(reduction=red))
        r = s.solve()
        tr = s.trace(r.iterations)
        x = s.iterate()[:p.n]
        s.close()
        assert r.converged and bool(g["converged"])
        assert r.iterations == k
        assert abs(r.f - g["f"][k]) <= 1e-8 * max(1.0, abs(g["f"][k]))
        f_dev = np.asarray(tr["f"][:k + 1])
        assert np.max(np.abs(f_dev - g["f"]) / np.maximum(1.0, np.abs(g["f"]))) < 1e-8
        assert np.max(np.abs(x - g["x"])) < 1e-7
        assert abs(x.sum() - 1.0) < 1e-7 and x.min() > -1e-9


@pytest.mark.parametrize("n", [1000, 3001])
def test_dataflow_factor_is_bitwise_reproducible(n):
    """The dataflow kernel hands tiles to whichever CTA is free, but every tile receives its updates
    in the fixed order and K-grouping of the task list: repeated runs must agree bit for bit (a race
    on a dependency flag or a tile would show up as run-to-run differences)."""
    import subprocess
    import sys
    code = r'''
import numpy as np, ipm_zoo_b200 as z
n = %d
rng = np.random.default_rng(3)
S = rng.standard_normal((n, n)) / np.sqrt(n)
A = 2.5 * np.eye(n) + 0.5 * (S + S.T)
f = z.Factor(n); assert f.info()["dataflow"]
f.set_matrix(A); f.set_rhs(rng.standard_normal(n))
ref = None
for rep in range(12):
    f.run(1, 2)
    L, D = f.ld(); x = f.solution()
    cur = (L.tobytes(), D.tobytes(), x.tobytes())
    if ref is None: ref = cur
    assert cur == ref, "run %%d differs" %% rep
print("ok")
''' % n
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, IPMZ_DATAFLOW_MIN_N="128", PYTHONPATH=root)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "ok" in out.stdout, out.stdout + out.stderr


def test_cfg3_full_size_kkt_and_reduction_agreement(z):
    """cfg3 at BASELINE.json's size (n = 8192, m = 4096): the NORMAL reduction (condensed 8192^2 matrix, the bench
    workload) and the AUGMENTED one (N = 12288) take the same number of iterations, agree on the objective to 1e-8
    and on x to 1e-7, and the converged iterate satisfies the KKT conditions to the solver's tolerance."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench", os.path.join(root, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    d = b.make_cfg3(8192, 4096, b.CFG3["seed"])
    p = ol.Problem(Q=d["Q"], c=d["c"], A=d["A"], l_A=d["l_A"], u_A=d["u_A"], l_x=d["l_x"], u_x=d["u_x"])
    out = {}
    for red in (z.NORMAL, z.AUGMENTED):
        s = z.Solver(z.Problem.from_data(p), z.Options(reduction=red))
        r = s.solve()
        it = s.iterate()
        s.close()
        assert r.converged and r.res < 1e-8 and r.mu < 1e-8
        rd, rp, mu, mn = kkt_residuals(p, it)
        assert rd < 1e-8 and rp < 1e-8 and mu < 1e-8 and mn > 0.0
        out[red] = (r, it)
    rn, ra = out[z.NORMAL][0], out[z.AUGMENTED][0]
    assert rn.iterations == ra.iterations
    assert abs(rn.f - ra.f) <= 1e-8 * max(1.0, abs(ra.f))
    assert np.max(np.abs(out[z.NORMAL][1][:p.n] - out[z.AUGMENTED][1][:p.n])) < 1e-7
