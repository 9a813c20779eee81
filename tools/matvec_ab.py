"""k_matvec variants (IPMZ_MATVEC_RU = rows per warp, trips unrolled) on the cfg3 shapes: achieved GB/s."""
import os, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
import ipm_zoo_b200 as z
import problems as P
p = P.ineq_box(8192, 4096, 3, kind="shift")
s = z.Solver(z.Problem.from_data(p), z.Options(reduction=z.NORMAL))
out = []
for k, ms, by in s.probe_kernels(20)[:3]:
    out.append("%s %.0f GB/s" % (k.replace("k_matvec ", ""), by / ms * 1e-6))
print(os.environ.get("IPMZ_MATVEC_RU", "auto"), " | ".join(out))
s.close()
