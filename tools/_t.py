import sys, numpy as np
sys.path.insert(0, '/root/repo')
import ipm_zoo_b200 as z
n = int(sys.argv[1])
rng = np.random.default_rng(100 + n)
M = rng.standard_normal((n, n)); A = M @ M.T / n + np.eye(n)
print("factor...", flush=True)
f = z.Factor(n); f.set_matrix(A); f.set_rhs(np.ones(n))
ms = f.run(1, 0); print("factored", ms, flush=True)
L, D = f.ld(); print("D ok", np.isfinite(D).all(), np.isfinite(L).all(), flush=True)
ms = f.run(1, 1); print("solved", ms, flush=True)
x = f.solution(); print(np.max(np.abs(A @ x - 1)))
