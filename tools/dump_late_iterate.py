"""Dump a late iterate of the CUDA path for the CPU oracle (tests/golden/make_golden_large.py late):
    python tools/dump_late_iterate.py cfg3 8 gpurun_out/cfg3_iter8.npz
runs the NORMAL-reduction solve with max_iter = k and saves the packed iterate the next iteration would start from."""
import os, sys
import numpy as np
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
import ipm_zoo_b200 as z
import problems as P
cfg, k, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
p = P.ineq_box(2048, 1024, 2, kind="shift") if cfg == "cfg2" else P.ineq_box(8192, 4096, 3, kind="shift")
s = z.Solver(z.Problem.from_data(p), z.Options(reduction=z.AUGMENTED, max_iter=k))
r = s.solve()
np.savez_compressed(out, iterate=s.iterate(), iterations=r.iterations, mu=r.mu, res=r.res)
print(cfg, "iterations", r.iterations, "mu", r.mu, "res", r.res)
