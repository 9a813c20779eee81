"""Small driver for ncu: one batched IPM solve of cfg4-shaped QPs (default 1024 problems)."""
import os, sys, time
import numpy as np
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
import ipm_zoo_b200 as z
import problems as P
cnt = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n, m = 256, 128
keys = ("Q", "c", "A", "l_A", "u_A", "l_x", "u_x")
shapes = dict(Q=(cnt, n, n), c=(cnt, n), A=(cnt, m, n), l_A=(cnt, m), u_A=(cnt, m), l_x=(cnt, n), u_x=(cnt, n))
pin = {k: np.empty(shapes[k]) for k in keys}
for i in range(cnt):
    q = P.ineq_box(n, m, 1000 + i, kind="shift")
    for k in keys:
        pin[k][i] = getattr(q, k)
bp = z.Problem(pin["Q"], pin["c"], pin["A"], pin["l_A"], pin["u_A"], None, None, pin["l_x"], pin["u_x"])
red = z.NORMAL if (len(sys.argv) < 3 or sys.argv[2] == "normal") else z.AUGMENTED
bs = z.BatchSolver(bp, cnt, z.Options(reduction=red))
for rep in range(2):
    bs.upload()
    t0 = time.perf_counter()
    res, ms = bs.solve(per_problem=False)
    print("batch %d: device %.2f ms, wall %.2f ms -> %.0f solves/s" % (cnt, ms, (time.perf_counter() - t0) * 1e3, cnt / ms * 1e3))
bs.close()
import ctypes as C
clk = (C.c_ulonglong * 16)()
if hasattr(z.lib(), "ipmz_debug_fused_clocks") and z.lib().ipmz_debug_fused_clocks(clk) == 0 and sum(clk) > 0:
    names = ["matvecs", "residuals", "assembly", "ldlt", "predictor", "mu+rhs", "corrector", "update"]
    tot = float(sum(clk))
    if clk[12]:
        print("  solve tiles (cycles per step): wait+barrier+issue %.0f; diagonal chain %.0f (x%d); off-diagonal %.0f (x%d)"
              % (clk[9] / (clk[12] + clk[13]), clk[10] / clk[12], clk[12], clk[11] / max(1, clk[13]), clk[13]))
    tot = float(sum(clk[:8])) or 1.0
    print("phase clocks (both reps): " + ", ".join("%s %.1f%%" % (names[i], 100.0 * clk[i] / tot) for i in range(8)))
    print("  inside predictor/corrector: " + ", ".join("sub%d %.1f%%" % (i, 100.0 * clk[i] / tot) for i in range(8, 9) if clk[i]))
