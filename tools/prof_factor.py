"""Small driver for ncu: one LDL^T factorization + solves of an n x n SPD matrix (default 8192)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ipm_zoo_b200 as z  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(0)
S = rng.standard_normal((n, n)) / np.sqrt(n)
A = 3.0 * np.eye(n) + 0.5 * (S + S.T)
f = z.Factor(n)
f.set_matrix(A)
b = rng.standard_normal(n)
f.set_rhs(b)
ms = f.run(reps, 2)
x = f.solution()
print("n=%d ms/step=%.3f TFLOP/s=%.2f resid=%.2e" % (n, ms / reps, reps * (n ** 3 / 3 + 4 * n * n) / ms * 1e-9,
                                                     np.max(np.abs(A @ x - b))))
