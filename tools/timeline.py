import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ipm_zoo_b200 as z
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
rng = np.random.default_rng(0)
S = rng.standard_normal((n, n)) / np.sqrt(n); A = 3.0 * np.eye(n) + 0.5 * (S + S.T)
f = z.Factor(n); f.set_matrix(A); f.run(2, 0)
cap = 4096
out = np.zeros(3 * cap); nrec = C.c_int()
z.lib().ipmz_debug_factor_timeline(f._h, out.ctypes.data_as(C.POINTER(C.c_double)), cap, C.byref(nrec))
rows = out[:3 * nrec.value].reshape(-1, 3)
names = {0: "diag", 1: "trsm", 2: "syrk"}
print("launches", nrec.value, "total ms", rows[:, 2].max())
for k, a, b in rows[:40]:
    print("%-5s %8.3f -> %8.3f  (%.3f)" % (names[int(k)], a, b, b - a))
print("...")
for k, a, b in rows[-30:]:
    print("%-5s %8.3f -> %8.3f  (%.3f)" % (names[int(k)], a, b, b - a))
