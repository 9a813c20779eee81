"""Timestamps of the streaming triangular solves: where one block hop spends its time."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ipm_zoo_b200 as z
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
rng = np.random.default_rng(0)
S = rng.standard_normal((n, n)) / np.sqrt(n); A = 3.0 * np.eye(n) + 0.5 * (S + S.T)
f = z.Factor(n); f.set_matrix(A); b = rng.standard_normal(n); f.set_rhs(b)
f.run(30, 2)  # ~0.3 s of work: boost clocks before the logged (single) solve
nblk = (n + 127) // 128
out = np.zeros((2, nblk, 16), dtype=np.int64)
rc = z.lib().ipmz_debug_trsv_log(f._h, out.ctypes.data_as(C.POINTER(C.c_longlong)), nblk)
print("rc", rc)
names = ["start", "prep done", "last tile begins", "chunks 0-2 done", "last chunk done", "rhs ready", "end"]
ghz = float(os.environ.get("SM_GHZ", "1.965"))
for sw, nm in ((0, "forward"), (1, "backward")):
    raw = out[sw].astype(np.float64)
    # cycle stamps -> us on the global time axis, anchored at each CTA's (stamp 6, globaltimer) pair
    t = raw[:, 7:8] + (raw[:, :7] - raw[:, 6:7]) / ghz
    rounds = raw[:, 7:8] + (raw[:, 10:14] - raw[:, 6:7]) / ghz
    t0 = min(out[0][:, 7].min(), out[1][:, 7].min()) - 1e6 if sw == 0 else t0
    if sw == 0:
        t0 = t[:, 0].min()
    t = (t - t0) * 1e-3; rounds = (rounds - t0) * 1e-3
    order = np.argsort(t[:, 6])
    pub = rounds[order, 3]  # last chunk of the block published
    print(nm, "sweep: first block published at %.1f us, last at %.1f us" % (pub.min(), pub.max()))
    print("  hop (last-chunk publish to last-chunk publish) us: mean %.2f p50 %.2f p90 %.2f" % (
        np.diff(pub).mean(), *np.percentile(np.diff(pub), [50, 90])))
    d = raw[order[1:], :]
    print("  cycles: last chunk wait+process %.0f | rhs (reduce, barrier) %.0f | init v + round 0 %.0f | round 1 %.0f | round 2 %.0f | round 3 %.0f" % (
        (d[:, 4] - d[:, 3]).mean(), (d[:, 5] - d[:, 4]).mean(), (d[:, 10] - d[:, 5]).mean(), (d[:, 11] - d[:, 10]).mean(),
        (d[:, 12] - d[:, 11]).mean(), (d[:, 13] - d[:, 12]).mean()))
    # absolute time (us) of every stamp
    T = (raw[:, 7:8] + (raw - raw[:, 6:7]) / ghz - t0) * 1e-3
    cons = T[order[1:], :]; prod = T[order[:-1], :]
    lat = [(cons[:, c] - prod[:, 10 + k]).mean() for k, c in enumerate((8, 9, 3, 4))]
    print("  consumer chunk k done - producer round k barrier (us):", " ".join("%.2f" % v for v in lat))
    print("  consumer: tile begins -> chunk0 %.2f us; producer round spacing %.2f %.2f %.2f us" % (
        (cons[:, 8] - cons[:, 2]).mean(), (prod[:, 11] - prod[:, 10]).mean(), (prod[:, 12] - prod[:, 11]).mean(), (prod[:, 13] - prod[:, 12]).mean()))
    print("  prep (start -> real pass) us: %.1f" % ((d[:, 1] - d[:, 0]).mean() / ghz * 1e-3))
