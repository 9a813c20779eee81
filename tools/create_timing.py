"""Host wall time of create / solve / get_iterate / destroy for one large QP (cfg3 by default), pinned host buffers."""
import os, sys, time
import numpy as np
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
import ipm_zoo_b200 as z
import problems as P
n, m = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (8192, 4096)
q = P.ineq_box(n, m, 3, kind="shift")
pin = {}
for k in ("Q", "c", "A", "l_A", "u_A", "l_x", "u_x"):
    pin[k] = z.pinned_empty(getattr(q, k).shape); pin[k][...] = getattr(q, k)
pr = z.Problem(pin["Q"], pin["c"], pin["A"], pin["l_A"], pin["u_A"], None, None, pin["l_x"], pin["u_x"])
for rep in range(int(os.environ.get("REPS", "3"))):
    t0 = time.perf_counter(); s = z.Solver(pr, z.Options(reduction=z.NORMAL))
    t1 = time.perf_counter(); r = s.solve()
    t2 = time.perf_counter(); it = s.iterate()
    t3 = time.perf_counter(); s.close()
    t4 = time.perf_counter()
    print("rep %d: create %.1f ms, solve %.1f ms (device loop %.1f, %d it), get_iterate %.1f ms, destroy %.1f ms; e2e %.1f ms"
          % (rep, (t1 - t0) * 1e3, (t2 - t1) * 1e3, r.solve_ms, r.iterations, (t3 - t2) * 1e3, (t4 - t3) * 1e3, (t3 - t0) * 1e3))
