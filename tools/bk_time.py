"""Device time of the Bunch-Kaufman factorization on two matrix kinds: a KKT saddle point [[H, C^T],[C, 0]] (few
interchanges until the zero block) and a random symmetric indefinite matrix (interchanges and 2x2 pivots throughout)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ipm_zoo_b200 as z
sizes = [int(a) for a in sys.argv[1:]] or [1536, 3072]
for N in sizes:
    rng = np.random.default_rng(N)
    n = 2 * N // 3; m = N - n
    M = rng.standard_normal((n, n)); H = M @ M.T / n + np.eye(n); Cm = rng.standard_normal((m, n)) / np.sqrt(n)
    K = np.block([[H, Cm.T], [Cm, np.zeros((m, m))]])
    S = rng.standard_normal((N, N)); S = S + S.T
    for name, A in (("saddle", K), ("random", S)):
        ms = z.bk_factor_time(A, 2)
        print("BK N=%d %-7s %8.2f ms  %.2f TB/s of the reference's (N^3/3) x 16 B" % (N, name, ms, N ** 3 / 3 * 16 / ms * 1e-9))
