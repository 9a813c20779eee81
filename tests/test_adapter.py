"""Drop-in check against the reference's own types: integration/ipmz_reference_adapter.cpp
(class NumericalOptimization::B200Optimizer, the reference's constructor signature) runs on an
Evaluation::Environment produced by the reference's build_environment, and must leave in it
the iterate the unmodified reference Optimizer produces.  The test library is prebuilt here
(integration/Makefile, needs the reference headers) and travels to the GPU box."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as ol
from golden.make_golden import CASES

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "integration", "_build", "libipmz_adapter_test.so")
GOLD = os.path.join(ROOT, "tests", "golden")
dp = C.POINTER(C.c_double)


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(SO), reason="adapter test library not prebuilt")
@pytest.mark.parametrize("name", ["toy", "ineq_box_64x32", "eq_box_40x20", "box_30", "ineq_lower_box_upper_30x12"])
@pytest.mark.parametrize("reduction", [0, 1])
def test_reference_adapter_drop_in(name, reduction):
    L = C.CDLL(SO)
    p = CASES[name]()
    g = np.load(os.path.join(GOLD, name + ".npz"))
    x = np.zeros(p.n)
    it, conv = C.c_int(), C.c_int()
    err = C.create_string_buffer(512)
    P = lambda a: a.ctypes.data_as(dp) if a is not None and a.size else None
    rc = L.adapter_solve(p.n, p.m_ineq, p.m_eq, P(p.Q), P(p.c), P(p.A), P(p.l_A), P(p.u_A), P(p.C), P(p.d),
                         P(p.l_x), P(p.u_x), p.ineq_bounds, p.var_bounds, int(p.equalities), reduction,
                         P(x), C.byref(it), C.byref(conv), err, 512)
    assert rc == 0, err.value.decode()
    assert it.value == int(g["iterations"]) and conv.value == int(g["converged"])
    assert np.max(np.abs(x - g["iterate"][:p.n])) < 1e-6


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(SO), reason="adapter test library not prebuilt")
def test_reference_adapter_equality_handling_none():
    """EqualityHandling::None through the reference's own symbolic layer: get_newton_system yields a system with
    lambda_A_eq but no equality slack, its augmented form has a symbolic-zero diagonal block (the reference's
    solve_indefinite_ == ASSERT(false)), and the adapter routes it to the Bunch-Kaufman path.  Same optimum as the
    reference's SlackedSlacks solve of the same QP."""
    L = C.CDLL(SO)
    p = CASES["eq_box_40x20"]()
    g = np.load(os.path.join(GOLD, "eq_box_40x20.npz"))
    x = np.zeros(p.n)
    it, conv = C.c_int(), C.c_int()
    err = C.create_string_buffer(512)
    P = lambda a: a.ctypes.data_as(dp) if a is not None and a.size else None
    rc = L.adapter_solve(p.n, 0, p.m_eq, P(p.Q), P(p.c), None, None, None, P(p.C), P(p.d), P(p.l_x), P(p.u_x),
                         0, p.var_bounds, 2, 0, P(x), C.byref(it), C.byref(conv), err, 512)
    assert rc == 0, err.value.decode()
    assert conv.value == 1
    assert np.max(np.abs(x - g["iterate"][:p.n])) < 1e-6
    assert np.max(np.abs(p.C @ x - p.d)) < 1e-8


@pytest.mark.gpu
def test_host_cpp_mirror_demo():
    """ipm-zoo_b200/host/host_demo: the reference's demo QP through the C++ host mirror."""
    import subprocess
    exe = os.path.join(ROOT, "ipm-zoo_b200", "host", "host_demo")
    if not os.path.exists(exe):
        pytest.skip("host_demo not built")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "iterations: 12 converged: 1" in out.stdout
    assert "assertion ok" in out.stdout
    bk = [l for l in out.stdout.splitlines() if l.startswith("bunch-kaufman:")][0]  # [[0,1],[1,0]] x = (3,5): one 2x2 pivot
    assert "ipiv=-1,-1" in bk and [float(t) for t in bk.split("x=")[1].split(",")] == [5.0, 3.0]
    last_iter = [l for l in out.stdout.splitlines() if l.startswith("iter: 12")][0]
    f = float(last_iter.split("f: ")[1].split(",")[0])
    assert abs(f - (-1.12799999999863552e+01)) < 1e-8
