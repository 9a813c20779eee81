"""GPU Bunch-Kaufman (ipm-zoo_b200/csrc/bunch_kaufman.cu, through the C ABI) against the oracle port of
LinearSolvers.cpp:76-318 and the reference-generated goldens.  The factorization is index and
order-preserving FP64 work: the bar is BIT-EXACT (same pivots, same bits in L and D).  The solve uses
parallel sums and is compared to the reference's sequential solve within 1e-10 relative."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ip = C.POINTER(C.c_int)


@pytest.fixture(scope="module")
def z():
    import ipm_zoo_b200 as z
    assert z.device_count() > 0, "no CUDA device: the product path has no CPU fallback"
    return z


def oracle_bk(K):
    n = K.shape[0]
    LD = np.zeros((n, n)); piv = np.zeros(n, dtype=np.int32)
    assert ol.port().orc_bk_factor(n, ol._ptr(K), ol._ptr(LD), piv.ctypes.data_as(ip)) == 0
    return LD, piv


def sym_indef(rng, n):
    S = rng.standard_normal((n, n))
    S = S + S.T
    if n > 1:
        S[0, 0] = 0.0  # forces an interchange at the first pivot
    return np.ascontiguousarray(S)


def saddle(rng, n, m):
    """KKT matrix of an equality-constrained QP with EqualityHandling::None: a zero (2,2) block."""
    M = rng.standard_normal((n, n))
    H = M @ M.T / n + np.eye(n)
    A = rng.standard_normal((m, n)) / np.sqrt(n)
    return np.ascontiguousarray(np.block([[H, A.T], [A, np.zeros((m, m))]]))


def check(z, K, solve=True, seed=0):
    n = K.shape[0]
    LDo, pivo = oracle_bk(K)
    LD, piv = z.symmetric_indefinite_factorization(K)
    np.testing.assert_array_equal(piv, pivo)
    np.testing.assert_array_equal(LD, LDo)  # bit-exact, upper triangle (= the input's) included
    if solve:
        b = np.random.default_rng(seed).standard_normal(n)
        x = b.copy()
        z.overwriting_solve_bunch_kaufman(LD, piv, x)
        xo = b.copy()
        ol.port().orc_bk_solve(n, ol._ptr(LDo), pivo.ctypes.data_as(ip), ol._ptr(xo))
        assert np.max(np.abs(x - xo)) <= 1e-10 * max(1.0, np.max(np.abs(xo)))
    return piv


@pytest.mark.parametrize("n", [1, 2, 3, 5, 17, 33, 64, 100, 255, 256, 257, 300, 515])
def test_bk_indefinite_bit_exact(z, n):
    piv = check(z, sym_indef(np.random.default_rng(300 + n), n), seed=n)
    if n >= 17:
        assert np.any(piv < 0), "expected 2x2 pivots in a random indefinite matrix"


@pytest.mark.parametrize("n,m", [(3, 2), (40, 17), (128, 64), (300, 140), (700, 324)])
def test_bk_saddle_point_bit_exact(z, n, m):
    """n+m >= 256 runs the multi-CTA team (cooperative launch), below one CTA."""
    check(z, saddle(np.random.default_rng(11 * n + m), n, m), seed=n + m)


def test_bk_large_team_bit_exact(z):
    check(z, sym_indef(np.random.default_rng(5), 1500), seed=1)


def test_bk_solve_beyond_the_prefetching_kernel(z):
    """n > 4096 takes the plain one-CTA solve (x alone in shared memory, opt-in above 48 KB): residual check
    against the matrix, the factor comes from the device as well."""
    rng = np.random.default_rng(12)
    K = saddle(rng, 3000, 1400)
    n = K.shape[0]
    LD, piv = z.symmetric_indefinite_factorization(K)
    b = rng.standard_normal(n)
    x = b.copy()
    z.overwriting_solve_bunch_kaufman(LD, piv, x)
    assert np.max(np.abs(K @ x - b)) <= 1e-8 * max(1.0, np.max(np.abs(x)))


def test_bk_golden_reference_vectors(z):
    g = np.load(os.path.join(GOLD, "linear_solvers.npz"))
    for nm in ("indef", "quasidef"):
        K = np.ascontiguousarray(g[nm + "_bk_K"])
        LD, piv = z.symmetric_indefinite_factorization(K)
        np.testing.assert_array_equal(piv, g[nm + "_bk_piv"])
        np.testing.assert_array_equal(LD, g[nm + "_bk_LD"])
        x = g[nm + "_bk_b"].copy()
        z.overwriting_solve_bunch_kaufman(g[nm + "_bk_LD"], g[nm + "_bk_piv"], x)
        assert np.max(np.abs(x - g[nm + "_bk_x"])) <= 1e-10 * np.max(np.abs(g[nm + "_bk_x"]))
        assert np.max(np.abs(K @ x - g[nm + "_bk_b"])) < 1e-11


def test_bk_zero_column_and_nonsymmetric_upper(z):
    """LinearSolvers.cpp:111-117: a zero column is skipped (ipiv[k] = k); only the lower triangle is read, the
    caller's upper triangle is returned untouched."""
    rng = np.random.default_rng(9)
    n = 40
    K = sym_indef(rng, n)
    K[:, 7] = 0.0; K[7, :] = 0.0
    K[np.triu_indices(n, 1)] = rng.standard_normal(n * (n - 1) // 2)  # garbage above the diagonal
    check(z, np.ascontiguousarray(K), solve=False)


def test_bk_several_zero_columns_keep_the_reference_quirk(z):
    """LinearSolvers.cpp:103,111-117: only the first zero column (info == 0) records kp = k; later zero columns leave
    kp = 0, so ipiv[k] = 0 there.  Index parity includes that."""
    rng = np.random.default_rng(10)
    n = 48
    K = sym_indef(rng, n)
    for j in (5, 20, 33):
        K[:, j] = 0.0; K[j, :] = 0.0
    K = np.ascontiguousarray(K)
    LDo, pivo = oracle_bk(K)
    assert pivo[5] == 5 and pivo[20] == 0 and pivo[33] == 0
    check(z, K, solve=False)


@pytest.mark.parametrize("n", [64, 300, 1000])
def test_bk_tied_magnitudes_first_index_wins(z, n):
    """Entries from {-1, 0, 1}: the pivot searches meet exact ties all the time, and the reference takes the FIRST
    index of the maximum (strict > scan, LinearSolvers.cpp:86-101)."""
    rng = np.random.default_rng(n)
    S = rng.integers(-1, 2, size=(n, n)).astype(np.float64)
    S = np.tril(S) + np.tril(S, -1).T
    check(z, np.ascontiguousarray(S), solve=False)


def test_bk_empty(z):
    LD, piv = z.symmetric_indefinite_factorization(np.zeros((0, 0)))
    assert LD.shape == (0, 0) and piv.size == 0
