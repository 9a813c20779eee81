"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libipmzoo_ref.so).

Run in the build container (needs /root/reference to build oracle/_ref):
    make -C oracle reference && python tests/golden/make_golden.py
Each file stores the reference's own trace of one solve -- per-iteration f / res / gap as
printed by Optimizer.cpp:131-132, the solved augmented Newton steps printed at
Optimizer.cpp:359, and the final iterate read back from the Environment -- plus a checksum of
the seeded inputs so generator drift is detected.  LinearSolvers goldens store L, D and the
Bunch-Kaufman factor / pivots of small seeded matrices (LinearSolvers.cpp:14-318).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402
import problems as P  # noqa: E402

CASES = {
    "toy": lambda: P.toy(),
    "ineq_box_20x10": lambda: P.ineq_box(20, 10, 1),
    "ineq_box_64x32": lambda: P.ineq_box(64, 32, 2),
    "eq_box_40x20": lambda: P.eq_box(40, 20, 3),
    "box_30": lambda: P.box_only(30, 4),
    "ineq_only_30x12": lambda: P.ineq_box(30, 12, 5, var_bounds=ol.NONE),
    "ineq_lower_box_upper_30x12": lambda: P.ineq_box(30, 12, 6, ineq_bounds=ol.LOWER, var_bounds=ol.UPPER),
    "ineq_upper_box_lower_30x12": lambda: P.ineq_box(30, 12, 7, ineq_bounds=ol.UPPER, var_bounds=ol.LOWER),
    "portfolio_64": lambda: P.portfolio(64, 8, 1e-6, 9),
    "cfg1_eq_box_200x100": lambda: P.eq_box(200, 100, 1),
    "cfg4_unit_256x128": lambda: P.ineq_box(256, 128, 1000, kind="shift"),
}


def checksum(p):
    parts = [p.Q, p.c, p.A, p.l_A, p.u_A, p.C, p.d, p.l_x, p.u_x]
    return float(sum(np.sum(np.abs(a)) for a in parts if a is not None))


def main():
    for name, make in CASES.items():
        p = make()
        tr = ol.ref_solve(p)
        k = tr.iterations
        keep = min(k, 100)
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            checksum=checksum(p), iterations=k, converged=tr.converged,
            f=tr.f[:k + 1], res=tr.res[:k + 1], mu=tr.mu[:k + 1],
            step_aff=tr.step_aff[:keep], step_cor=tr.step_cor[:keep],
            iterate=tr.iterate)
        print(name, "iters", k, "converged", tr.converged, "f*", repr(tr.f[min(k, tr.n_logged - 1)]))
    rng = np.random.default_rng(77)
    n = 24
    M = rng.standard_normal((n, n))
    spd = M @ M.T + n * np.eye(n)
    A = rng.standard_normal((10, n))
    qd = np.block([[spd, A.T], [A, -np.diag(rng.uniform(0.5, 2.0, 10))]])
    indef = rng.standard_normal((n, n)); indef = indef + indef.T
    indef[0, 0] = 0.0  # forces pivoting
    out = {}
    for nm, K in (("spd", spd), ("quasidef", qd)):
        m = K.shape[0]
        L = np.zeros((m, m)); D = np.zeros(m)
        K = np.ascontiguousarray(K)
        assert ol.ref().ref_ldlt(m, ol._ptr(K), ol._ptr(L), ol._ptr(D)) == 0
        b = rng.standard_normal(m); x = b.copy()
        assert ol.ref().ref_solve_ldlt(m, ol._ptr(L), ol._ptr(D), ol._ptr(x)) == 0
        out.update({nm + "_K": K, nm + "_L": L, nm + "_D": D, nm + "_b": b, nm + "_x": x})
    for nm, K in (("indef", indef), ("quasidef", qd)):
        m = K.shape[0]
        K = np.ascontiguousarray(K)
        LD = np.zeros((m, m)); piv = np.zeros(m, dtype=np.int32)
        import ctypes as C
        assert ol.ref().ref_bk_factor(m, ol._ptr(K), ol._ptr(LD), piv.ctypes.data_as(C.POINTER(C.c_int))) == 0
        b = rng.standard_normal(m); x = b.copy()
        assert ol.ref().ref_bk_solve(m, ol._ptr(LD), piv.ctypes.data_as(C.POINTER(C.c_int)), ol._ptr(x)) == 0
        out.update({nm + "_bk_K": K, nm + "_bk_LD": LD, nm + "_bk_piv": piv, nm + "_bk_b": b, nm + "_bk_x": x})
        print(nm, "BK residual", np.max(np.abs(K @ x - b)))
    np.savez_compressed(os.path.join(HERE, "linear_solvers.npz"), **out)


if __name__ == "__main__":
    main()
