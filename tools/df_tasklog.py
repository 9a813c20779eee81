"""Per-task log of the persistent dataflow LDL^T: where the time of one factorization goes."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ipm_zoo_b200 as z
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
rng = np.random.default_rng(0)
S = rng.standard_normal((n, n)) / np.sqrt(n); A = 3.0 * np.eye(n) + 0.5 * (S + S.T)
f = z.Factor(n); f.set_matrix(A)
b = rng.standard_normal(n); f.set_rhs(b)
ms = f.run(1, 1); ms = f.run(3, 0) / 3
x = f.solution()
print("n=%d factor ms=%.3f (%.2f TF) resid=%.2e" % (n, ms, n ** 3 / 3 / ms * 1e-9, np.max(np.abs(A @ x - b))))
L = z.lib()
nt = L.ipmz_debug_factor_ntasks(f._h)
if nt == 0:
    print("dataflow plan not active"); sys.exit(0)
log = np.zeros((nt, 8), dtype=np.int64); got = C.c_int(); sim = C.c_double()
rc = L.ipmz_debug_factor_tasklog(f._h, log.ctypes.data_as(C.POINTER(C.c_longlong)), nt, C.byref(got), C.byref(sim))
print("tasklog rc", rc, "tasks", got.value, "simulated makespan us", sim.value)
t0 = log[:, 0].min(); st = (log[:, 0] - t0) * 1e-3; en = (log[:, 1] - t0) * 1e-3
typ = log[:, 3] & 0xff; kw = (log[:, 3] >> 32); K = (kw >> 16) - (kw & 0xffff)
print("makespan us %.1f" % en.max())
dur = en - st
for t, name in ((0, "DIAG"), (3, "DIAGU"), (1, "TRSM"), (2, "UPD")):
    m = typ == t
    if m.any():
        print("%s: n=%d mean %.2f us  p10 %.2f p50 %.2f p90 %.2f max %.2f  total %.1f ms-SM" % (
            name, m.sum(), dur[m].mean(), *np.percentile(dur[m], [10, 50, 90]), dur[m].max(), dur[m].sum() * 1e-3))
dm = typ == 0
print("DIAG phases (cycles, median): load %d | 4 x ldlt32 %d | elimination loop %d | store %d" % tuple(
    np.median(log[dm, 4 + i]) for i in range(4)))
tm = typ == 1
print("TRSM phases (cycles, median): L_kk issue %d | acc loads %d | wait+barrier %d | solve %d | (task total %d)" % (
    *(np.median(log[tm, 4 + i]) for i in range(4)), np.median(dur[tm]) * 1965))
um = typ == 2
for kk in (1, 8):
    mk = um & (K == kk)
    if mk.any():
        print("UPD K=%d phases (cycles, median): C tile + first slice %d | other slices %d (%.0f per slice) | store issue %d | barrier+fence+release %d | task total %d" % (
            kk, *(np.median(log[mk, 4 + i]) for i in range(2)), np.median(log[mk, 5]) / (8 * kk - 1),
            *(np.median(log[mk, 4 + i]) for i in (2, 3)), np.median(dur[mk]) * 1965))
# time from the end of the previous task on the SM to the start of this one (task switch)
m = typ == 2
for k in range(1, 9):
    mk = m & (K == k)
    if mk.any():
        print("  UPD K=%d: n=%d mean %.2f us (%.2f us/panel)" % (k, mk.sum(), dur[mk].mean(), dur[mk].mean() / k))
nsm = int(log[:, 2].max()) + 1
busy = np.zeros(nsm)
np.add.at(busy, log[:, 2], dur)
print("SMs %d, busy fraction mean %.3f min %.3f" % (nsm, busy.mean() / en.max(), busy.min() / en.max()))
# per-SM idle gaps between consecutive tasks = dependency waits + task switch
order = np.lexsort((st, log[:, 2]))
gaps = []
for a, b_ in zip(order[:-1], order[1:]):
    if log[a, 2] == log[b_, 2]:
        gaps.append(st[b_] - en[a])
gaps = np.array(gaps)
print("gaps between tasks on an SM: mean %.2f us p50 %.2f p90 %.2f p99 %.2f max %.1f; total %.1f ms-SM" % (
    gaps.mean(), *np.percentile(gaps, [50, 90, 99]), gaps.max(), gaps.sum() * 1e-3))
fu = typ == 3
for k in range(1, 9):
    mk = fu & (K == k)
    if mk.any():
        print("  DIAGU K=%d: n=%d median %.2f us" % (k, mk.sum(), np.median(dur[mk])))
d = np.where((typ == 0) | (typ == 3))[0]
ds = st[d]; print("DIAG start times (us) every 8th:", np.round(np.sort(ds)[::8], 0))
if len(sys.argv) > 2:
    np.save(sys.argv[2], log)
