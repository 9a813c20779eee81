"""Per-iteration device time of the indefinite (EqualityHandling::None, Bunch-Kaufman) path next to the
quasi-definite (SlackedSlacks, unpivoted LDL^T) path on the same equality-constrained QP."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import ipm_zoo_b200 as z  # noqa: E402
import problems as P  # noqa: E402

for n, me in [(256, 128), (1024, 512), (2048, 1024), (4096, 2048)]:
    p = P.eq_box(n, me, 7)
    for name, eq in (("SlackedSlacks/LDL^T", z.EQ_SLACKED_SLACKS), ("None/Bunch-Kaufman", z.EQ_NONE)):
        q = z.Problem.from_data(p)
        q.equalities = eq
        s = z.Solver(q, z.Options(reduction=z.AUGMENTED))
        r = s.solve()
        s.close()
        print("N=%5d %-20s iterations %2d converged %d f %.10f  %.2f ms/iteration (factor flops N^3/3 = %.2e)"
              % (n + me, name, r.iterations, r.converged, r.f, r.solve_ms / max(1, r.iterations), (n + me) ** 3 / 3))
# raw factorization through the C ABI (includes H2D / D2H of the matrix)
rng = np.random.default_rng(0)
for n in (1024, 3072):
    S = rng.standard_normal((n, n)); S = S + S.T
    t0 = time.perf_counter()
    LD, piv = z.symmetric_indefinite_factorization(S)
    t1 = time.perf_counter()
    print("BK C-ABI n=%d: %.1f ms wall, 2x2 pivots: %d" % (n, 1e3 * (t1 - t0), int(np.sum(piv < 0)) // 2))
