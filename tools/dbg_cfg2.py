import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import ipm_zoo_b200 as z
import problems as P
p = P.ineq_box(2048, 1024, 2, kind="shift")
for red in (z.AUGMENTED, z.NORMAL):
    s = z.Solver(z.Problem.from_data(p), z.Options(reduction=red))
    r = s.solve()
    tr = s.trace(r.iterations)
    print("red", red, "iters", r.iterations, "conv", r.converged, "res %.3e mu %.3e f %.12f" % (r.res, r.mu, r.f))
    print("  res trace", ["%.1e" % v for v in tr["res"][:r.iterations + 1]])
    s.close()
