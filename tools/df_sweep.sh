#!/bin/bash
# schedule-parameter sweep of the dataflow LDL^T (n=8192): factor-only ms
run() { echo -n "$* : "; env "$@" timeout 120 python - <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, ipm_zoo_b200 as z
n = 8192
rng = np.random.default_rng(0)
S = rng.standard_normal((n, n)) / np.sqrt(n); A = 3.0 * np.eye(n) + 0.5 * (S + S.T)
f = z.Factor(n); f.set_matrix(A); f.set_rhs(rng.standard_normal(n))
f.run(2, 0)
ms = min(f.run(5, 0) / 5 for _ in range(3))
print("factor %.3f ms (%.2f TF)" % (ms, n ** 3 / 3 / ms * 1e-9))
PY
}
run A=1
run IPMZ_DF_UPD_BASE_US=7 IPMZ_DF_UPD_PANEL_US=17.5
run IPMZ_DF_UPD_BASE_US=7 IPMZ_DF_UPD_PANEL_US=17.5 IPMZ_DF_KB=3
run IPMZ_DF_UPD_BASE_US=7 IPMZ_DF_UPD_PANEL_US=17.5 IPMZ_DF_KB=4
run IPMZ_DF_UPD_BASE_US=7 IPMZ_DF_UPD_PANEL_US=17.5 IPMZ_DF_KB=3 IPMZ_DF_KMAX=6
run IPMZ_DF_UPD_BASE_US=7 IPMZ_DF_UPD_PANEL_US=17.5 IPMZ_DF_KB=4 IPMZ_DF_KMAX=8
run IPMZ_DF_UPD_BASE_US=7 IPMZ_DF_UPD_PANEL_US=17.5 IPMZ_DF_LA=1
run IPMZ_DF_UPD_BASE_US=7 IPMZ_DF_UPD_PANEL_US=17.5 IPMZ_DF_LA=3
run IPMZ_DF_UPD_BASE_US=7 IPMZ_DF_UPD_PANEL_US=17.5 IPMZ_DF_LA=4 IPMZ_DF_KB=3 IPMZ_DF_KMAX=6
run IPMZ_DF_UPD_BASE_US=7 IPMZ_DF_UPD_PANEL_US=17.5 IPMZ_DF_DIAG_US=60 IPMZ_DF_TRSM_US=22
run IPMZ_DF_UPD_BASE_US=7 IPMZ_DF_UPD_PANEL_US=17.5 IPMZ_DF_DIAG_US=35 IPMZ_DF_TRSM_US=14
run IPMZ_DF_UPD_BASE_US=10 IPMZ_DF_UPD_PANEL_US=19 IPMZ_DF_KMAX=2
