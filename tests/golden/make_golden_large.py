"""Oracle goldens at BASELINE.json's full sizes (cfg2, cfg3, cfg5) for the GPU step-parity tests.

    python tests/golden/make_golden_large.py cfg2        # reference itself (+ port, asserted bitwise equal)
    python tests/golden/make_golden_large.py cfg3        # port, first iteration at N = 12288 (~6-10 min)
    python tests/golden/make_golden_large.py cfg5 1e-6   # port, last three iterations (mu <~ 1e-6)
    python tests/golden/make_golden_large.py cfg5 1e-10  # port, full solve + last three iterations
    python tests/golden/make_golden_large.py late <npz with 'iterate'> <cfg2|cfg3> <out name>
                                                         # one Newton iteration of the port from a given iterate
                                                         # (e.g. a late iterate dumped by the GPU path)

What is stored per "step record": the packed iterate the iteration starts from, f / res / mu there,
both solved augmented Newton steps (Optimizer.cpp:359), alpha_aff / sigma / alpha, and the iterate
after the update (Optimizer.cpp:222-238).  The GPU tests upload the stored iterate, run ONE iteration
(ipmz_newton_step) and compare step for step, so rounding drift between the trajectories never enters.
cfg2 runs the UNMODIFIED reference (oracle/_ref) for the whole solve and asserts that the port reproduces its
trace bit for bit at N = 3072 before the port is used for the intermediate iterates.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402
import problems as P  # noqa: E402


def checksum(p):
    parts = [p.Q, p.c, p.A, p.l_A, p.u_A, p.C, p.d, p.l_x, p.u_x]
    return float(sum(np.sum(np.abs(a)) for a in parts if a is not None))


def one_iteration(p, iterate):
    """One Newton iteration of the port from `iterate` (None = the reference's initial point)."""
    tr = ol.port_solve(p, cap_iters=1, stop_after_cap=True, iterate=iterate)
    start = np.array(iterate) if iterate is not None else None
    if start is None:
        start = np.zeros(p.iterate_len)
        ps = p.c_struct()
        ol.port().orc_initial_iterate(ps, ol._ptr(start))
    return dict(start=start, f=tr.f[0], res=tr.res[0], mu=tr.mu[0], step_aff=tr.step_aff[0].copy(),
                step_cor=tr.step_cor[0].copy(), alpha_aff=tr.alpha_aff[0], sigma=tr.sigma[0], alpha=tr.alpha[0],
                after=np.array(tr.iterate), f1=tr.f[1], res1=tr.res[1], mu1=tr.mu[1])


def pack(prefix, rec, out):
    for k, v in rec.items():
        out[prefix + k] = v


def iterate_at(p, k):
    """Packed iterate at the start of iteration k (port)."""
    if k == 0:
        return None
    tr = ol.port_solve(p, cap_iters=k, stop_after_cap=True, steps=False)
    assert tr.iterations == k
    return np.array(tr.iterate)


def cfg2():
    p = P.ineq_box(2048, 1024, 2, kind="shift")
    t0 = time.time()
    tr = ol.ref_solve(p)
    k = tr.iterations
    print("reference: iterations", k, "converged", tr.converged, "f", repr(tr.f[k]), "wall", time.time() - t0, flush=True)
    tp = ol.port_solve(p)
    assert tp.iterations == k and tp.converged == tr.converged
    for a in ("f", "res", "mu"):
        assert np.array_equal(getattr(tp, a)[:k + 1], getattr(tr, a)[:k + 1]), a
    assert np.array_equal(tp.step_aff[:k], tr.step_aff[:k]) and np.array_equal(tp.step_cor[:k], tr.step_cor[:k])
    assert np.array_equal(tp.iterate, tr.iterate)
    print("port == reference bit for bit at N = 3072", flush=True)
    out = dict(checksum=checksum(p), iterations=k, converged=int(tr.converged), f=tr.f[:k + 1], res=tr.res[:k + 1],
               mu=tr.mu[:k + 1], iterate=np.array(tr.iterate), alpha_aff=tp.alpha_aff[:k], sigma=tp.sigma[:k],
               alpha=tp.alpha[:k])
    its = sorted({0, 3, k - 2, k - 1})
    out["step_iters"] = np.array(its)
    for it in its:
        rec = one_iteration(p, iterate_at(p, it))
        # the one-iteration record must be the reference's own iteration `it`
        assert np.array_equal(rec["step_aff"], tr.step_aff[it]) and np.array_equal(rec["step_cor"], tr.step_cor[it])
        assert rec["f"] == tr.f[it] and rec["res1"] == tr.res[it + 1]
        pack("it%d_" % it, rec, out)
        print("iteration", it, "mu", rec["mu"], "alpha", rec["alpha"], flush=True)
    np.savez_compressed(os.path.join(HERE, "cfg2_ineq_box_2048x1024.npz"), **out)


def cfg3():
    p = P.ineq_box(8192, 4096, 3, kind="shift")
    t0 = time.time()
    rec = one_iteration(p, None)
    print("cfg3 first iteration: f", rec["f"], "res", rec["res"], "alpha", rec["alpha"], "wall", time.time() - t0, flush=True)
    out = dict(checksum=checksum(p))
    pack("it0_", rec, out)
    del out["it0_start"]  # the reference's initial point: rebuilt by ipmz_reset_iterate
    np.savez_compressed(os.path.join(HERE, "cfg3_first_iteration_8192x4096.npz"), **out)


def cfg5(eps):
    p = P.portfolio(4096, 32, eps, 5)
    t0 = time.time()
    tr = ol.port_solve(p, steps=False)
    k = tr.iterations
    print("cfg5 eps", eps, "iterations", k, "converged", tr.converged, "f", repr(tr.f[k]), "wall", time.time() - t0, flush=True)
    tag = ("%g" % eps).replace("-", "m")
    it = np.asarray(tr.iterate)
    out = dict(checksum=checksum(p), iterations=k, converged=int(tr.converged), f=tr.f[:k + 1], res=tr.res[:k + 1],
               mu=tr.mu[:k + 1], x=it[:p.n])
    its = [k - 3, k - 2, k - 1]
    out["step_iters"] = np.array(its)
    cur = iterate_at(p, its[0])
    for j in its:
        rec = one_iteration(p, cur)
        assert rec["f"] == tr.f[j] and rec["res"] == tr.res[j] and rec["f1"] == tr.f[j + 1]
        pack("it%d_" % j, rec, out)
        cur = rec["after"]
        print("iteration", j, "mu", rec["mu"], "alpha", rec["alpha"], "wall", time.time() - t0, flush=True)
    np.savez_compressed(os.path.join(HERE, "cfg5_portfolio_4096_eps%s_steps.npz" % tag), **out)


def late(src, cfg, name):
    p = {"cfg2": lambda: P.ineq_box(2048, 1024, 2, kind="shift"),
         "cfg3": lambda: P.ineq_box(8192, 4096, 3, kind="shift")}[cfg]()
    start = np.load(src)["iterate"]
    t0 = time.time()
    rec = one_iteration(p, start)
    print(cfg, "late iteration: mu", rec["mu"], "res", rec["res"], "alpha", rec["alpha"], "wall", time.time() - t0, flush=True)
    out = dict(checksum=checksum(p))
    pack("late_", rec, out)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


if __name__ == "__main__":
    what = sys.argv[1]
    if what == "cfg2":
        cfg2()
    elif what == "cfg3":
        cfg3()
    elif what == "cfg5":
        cfg5(float(sys.argv[2]))
    elif what == "late":
        late(sys.argv[2], sys.argv[3], sys.argv[4])
