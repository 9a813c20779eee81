"""cfg5 at BASELINE.json's full size (n = 4096 portfolio QP, eps = 1e-6): the plain-C oracle port
(pinned bit-for-bit against the unmodified reference on every smaller case, test_oracle_golden.py)
solves it once on the CPU -- about 5 minutes, the LDL^T of the 4097 x 4097 augmented system 14 times --
and the trace is committed as tests/golden/cfg5_portfolio_4096.npz for the GPU parity test.

    python tests/golden/make_golden_cfg5.py
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402
import problems as P  # noqa: E402


def main():
    t0 = time.time()
    p = P.portfolio(4096, 32, 1e-6, 5)
    tr = ol.port_solve(p, steps=False)
    print("iterations", tr.iterations, "converged", tr.converged, "f", tr.f[tr.iterations], "seconds", time.time() - t0)
    it = np.asarray(tr.iterate)
    np.savez_compressed(os.path.join(HERE, "cfg5_portfolio_4096.npz"), iterations=tr.iterations,
                        converged=int(tr.converged), f=np.asarray(tr.f[:tr.iterations + 1]),
                        res=np.asarray(tr.res[:tr.iterations + 1]), x=it[:p.n])


if __name__ == "__main__":
    main()
