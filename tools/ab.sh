#!/bin/bash
# A/B of two library builds inside one GPU call: IPMZ_LIB=<path> selects the build
OLD=$PWD/ipm-zoo_b200/ab/libipmz_b200_old.so
for rep in 1 2 3; do
  echo -n "old: "; IPMZ_LIB=$OLD timeout 120 python tools/prof_factor.py 8192 5
  echo -n "new: "; timeout 120 python tools/prof_factor.py 8192 5
done
cat > /tmp/e2e.py <<'PY'
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, ipm_zoo_b200 as z
import importlib.util
spec = importlib.util.spec_from_file_location("bench", "bench.py"); b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
d = b.make_cfg3(8192, 4096, b.CFG3["seed"])
prob = z.Problem(d["Q"], d["c"], d["A"], d["l_A"], d["u_A"], None, None, d["l_x"], d["u_x"])
for rep in range(3):
    s = z.Solver(prob, z.Options(reduction=z.NORMAL)); r = s.solve(); s.close()
    print("  solve: %d iterations, device loop %.1f ms, f=%.10f" % (r.iterations, r.solve_ms, r.f))
PY
echo "old e2e:"; IPMZ_LIB=$OLD timeout 300 python /tmp/e2e.py
echo "new e2e:"; timeout 300 python /tmp/e2e.py
