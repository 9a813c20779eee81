"""cfg4-shaped batch split over G concurrent BatchSolver handles (one host thread + stream each)."""
import os, sys, time, threading
import numpy as np
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
import ipm_zoo_b200 as z
import problems as P
cnt = int(sys.argv[1]); G = int(sys.argv[2])
n, m = 256, 128
keys = ("Q", "c", "A", "l_A", "u_A", "l_x", "u_x")
probs = [P.ineq_box(n, m, 1000 + i, kind="shift") for i in range(min(cnt, 256))]
def make(lo, hi):
    st = lambda key: np.stack([getattr(probs[i % len(probs)], key) for i in range(lo, hi)])
    bp = z.Problem(st("Q"), st("c"), st("A"), st("l_A"), st("u_A"), None, None, st("l_x"), st("u_x"))
    return z.BatchSolver(bp, hi - lo, z.Options(reduction=z.NORMAL))
bounds = [(g * cnt // G, (g + 1) * cnt // G) for g in range(G)]
solvers = [make(lo, hi) for lo, hi in bounds]
import torch
for rep in range(3):
    for s in solvers: s.upload()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    th = [threading.Thread(target=lambda s=s: s.solve(per_problem=False)) for s in solvers]
    for t in th: t.start()
    for t in th: t.join()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("batch %d as %d x %d: wall %.2f ms -> %.0f solves/s" % (cnt, G, cnt // G, dt * 1e3, cnt / dt))
for s in solvers: s.close()
