import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ipm_zoo_b200 as z
n = 128
rng = np.random.default_rng(0)
S = rng.standard_normal((n, n)); A = S @ S.T / n + np.eye(n)
f = z.Factor(n); f.set_matrix(A); f.run(1, 0); f.run(1, 0)
out = (C.c_longlong * 16)()
z.lib().ipmz_debug_phase_clocks(out)
v = list(out)
print("phase deltas (cycles):", [v[i + 1] - v[i] for i in range(7)], "total", v[7] - v[0])
