#!/bin/bash
# schedule-parameter sweep of the dataflow LDL^T on the TMA build (n=8192): factor-only ms
run() { echo -n "$* : "; env "$@" timeout 120 python - <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, ipm_zoo_b200 as z
n = int(os.environ.get("SWEEP_N", "8192"))
rng = np.random.default_rng(0)
S = rng.standard_normal((n, n)) / np.sqrt(n); A = 3.0 * np.eye(n) + 0.5 * (S + S.T)
f = z.Factor(n); f.set_matrix(A); f.set_rhs(rng.standard_normal(n))
f.run(2, 0)
ms = min(f.run(5, 0) / 5 for _ in range(3))
print("factor %.3f ms (%.2f TF) sim %.0f us" % (ms, n ** 3 / 3 / ms * 1e-9, f.info()["simulated_us"]))
PY
}
for kb in 6 8 10 12; do for la in 2 3 4 6; do run IPMZ_DF_KB=$kb IPMZ_DF_LA=$la; done; done
run IPMZ_DF_UPD_BASE_US=5 A=1
run IPMZ_DF_UPD_BASE_US=9 A=1
run IPMZ_DF_DIAG_US=60 A=1
run IPMZ_DF_TRSM_US=20 A=1
run IPMZ_DF_FUSE_DIAG=0 A=1
