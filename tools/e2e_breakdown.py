"""Wall-clock breakdown of the end-to-end cfg3 call sequence (create = H2D + allocation + plan, solve, read-back, destroy)."""
import importlib.util, os, sys, time
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
import numpy as np
import ipm_zoo_b200 as z
spec = importlib.util.spec_from_file_location("bench", os.path.join(root, "bench.py")); b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
d = b.make_cfg3(8192, 4096, b.CFG3["seed"])
pin = {}
for k in ("Q", "c", "A", "l_A", "u_A", "l_x", "u_x"):
    pin[k] = z.pinned_empty(d[k].shape); pin[k][...] = d[k]
prob = z.Problem(pin["Q"], pin["c"], pin["A"], pin["l_A"], pin["u_A"], None, None, pin["l_x"], pin["u_x"])
opt = z.Options(reduction=z.NORMAL)
for rep in range(3):
    t0 = time.perf_counter(); s = z.Solver(prob, opt)
    t1 = time.perf_counter(); r = s.solve()
    t2 = time.perf_counter(); it = s.iterate()
    t3 = time.perf_counter(); s.close()
    t4 = time.perf_counter()
    print("create %.1f ms | solve %.1f ms (device loop %.1f) | get_iterate %.1f | destroy %.1f | total %.1f" %
          (1e3 * (t1 - t0), 1e3 * (t2 - t1), r.solve_ms, 1e3 * (t3 - t2), 1e3 * (t4 - t3), 1e3 * (t4 - t0)))
