#!/bin/bash
# schedule-parameter sweep of the dataflow LDL^T (n=8192): factor-only ms
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
print("factor %.3f ms (%.2f TF)" % (ms, n ** 3 / 3 / ms * 1e-9))
PY
}
run IPMZ_DF_KB=8 IPMZ_DF_KMAX=15
run IPMZ_DF_KB=8 IPMZ_DF_KMAX=15 IPMZ_DF_LA=3
run IPMZ_DF_KB=10 IPMZ_DF_KMAX=15
run IPMZ_DF_KB=12 IPMZ_DF_KMAX=15
run IPMZ_DF_KB=15 IPMZ_DF_KMAX=15
run IPMZ_DF_KB=10 IPMZ_DF_KMAX=15 IPMZ_DF_LA=3
run SWEEP_N=4096 A=1
run SWEEP_N=4096 IPMZ_DF_KB=8 IPMZ_DF_KMAX=15
run SWEEP_N=4096 IPMZ_DF_KB=12 IPMZ_DF_KMAX=15
run SWEEP_N=12288 A=1
run SWEEP_N=12288 IPMZ_DF_KB=8 IPMZ_DF_KMAX=15
run SWEEP_N=12288 IPMZ_DF_KB=12 IPMZ_DF_KMAX=15
