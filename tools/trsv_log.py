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
names = ["start", "prep done", "last tile: poll", "x arrived", "tile landed+bar", "rhs ready", "published"]
for sw, nm in ((0, "forward"), (1, "backward")):
    raw = out[sw].astype(np.float64)
    ghz = float(os.environ.get("SM_GHZ", "1.965"))
    # cycle stamps -> ns on the global time axis, anchored at each CTA's (stamp 6, globaltimer) pair
    t = raw[:, 7:8] + (raw[:, :7] - raw[:, 6:7]) / ghz
    t0 = t[:, 0].min(); t = (t - t0) * 1e-3
    print(nm, "sweep total us %.1f" % t[:, 6].max())
    order = np.argsort(t[:, 6])
    pub = t[order, 6]
    print("  hop (publish-to-publish) us: mean %.2f p50 %.2f p90 %.2f" % (np.diff(pub).mean(), *np.percentile(np.diff(pub), [50, 90])))
    for k in (1, 8, 24, 40, 56, nblk - 1):
        r = order[k]
        print("  block#%d (row %d): " % (k, r) + "  ".join("%s %.1f" % (names[i], t[r, i]) for i in range(7)))
    d = t[order[1:], :]
    print("  mean us: prep %.1f | wait x %.2f | land+bar %.2f | tile+reduce %.2f | diag step %.2f" % (
        (d[:, 1] - d[:, 0]).mean(), (d[:, 3] - d[:, 2]).mean(), (d[:, 4] - d[:, 3]).mean(), (d[:, 5] - d[:, 4]).mean(),
        (d[:, 6] - d[:, 5]).mean()))
    dd = raw[order[1:], :]
    print("  diag step cycles: init v %.0f | warp0 substeps %.0f | bar0 %.0f | round1 %.0f | round2 %.0f | round3 %.0f | tail %.0f" % (
        (dd[:, 8] - dd[:, 5]).mean(), (dd[:, 9] - dd[:, 8]).mean(), (dd[:, 10] - dd[:, 9]).mean(), (dd[:, 11] - dd[:, 10]).mean(),
        (dd[:, 12] - dd[:, 11]).mean(), (dd[:, 13] - dd[:, 12]).mean(), (dd[:, 6] - dd[:, 13]).mean()))
    print("  tile+reduce cycles %.0f" % (dd[:, 5] - dd[:, 4]).mean())
    prev_pub = pub[:-1]
    print("  mean (x arrived - previous publish) %.2f us" % (d[:, 3] - prev_pub).mean())
