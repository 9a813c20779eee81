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
@pytest.mark.skipif(not os.path.exists(SO), reason="adapter test library not prebuilt")
@pytest.mark.parametrize("handling", [4, 5])
def test_reference_adapter_equality_handling_penalty(handling):
    """EqualityHandling::PenaltyFunction (4) / PenaltyFunctionWithExtraDual (5) through the reference's own symbolic
    layer: the adapter recognises the diagonal block -mu of the multiplier row in get_augmented_system's output and
    routes it to IPMZ_EQ_PENALTY; the result is the iteration of tests/penalty_model.py."""
    import penalty_model as pm
    L = C.CDLL(SO)
    p = CASES["eq_box_40x20"]()
    tr = pm.solve(p.Q, p.c, p.C, p.d, p.l_x, p.u_x)
    x = np.zeros(p.n)
    it, conv = C.c_int(), C.c_int()
    err = C.create_string_buffer(512)
    P = lambda a: a.ctypes.data_as(dp) if a is not None and a.size else None
    rc = L.adapter_solve(p.n, 0, p.m_eq, P(p.Q), P(p.c), None, None, None, P(p.C), P(p.d), P(p.l_x), P(p.u_x),
                         0, p.var_bounds, handling, 0, P(x), C.byref(it), C.byref(conv), err, 512)
    assert rc == 0, err.value.decode()
    assert conv.value == 1 and it.value == tr["iterations"]
    assert np.max(np.abs(x - tr["x"])) < 1e-6
    assert np.max(np.abs(p.C @ x - p.d)) < 1e-8


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(SO), reason="adapter test library not prebuilt")
def test_reference_adapter_equality_handling_regularization():
    """EqualityHandling::Regularization through the reference's symbolic layer: p_eq is a variable of the Newton
    system and delta_eq comes from the reference's Environment (EnvironmentBuilder.cpp:48); the reference's own
    evaluator asserts on the scalar block (Evaluation.cpp:53-60)."""
    L = C.CDLL(SO)
    p = CASES["eq_box_40x20"]()
    g = np.load(os.path.join(GOLD, "eq_box_40x20.npz"))
    x = np.zeros(p.n)
    it, conv = C.c_int(), C.c_int()
    err = C.create_string_buffer(512)
    P = lambda a: a.ctypes.data_as(dp) if a is not None and a.size else None
    rc = L.adapter_solve(p.n, 0, p.m_eq, P(p.Q), P(p.c), None, None, None, P(p.C), P(p.d), P(p.l_x), P(p.u_x),
                         0, p.var_bounds, 3, 0, P(x), C.byref(it), C.byref(conv), err, 512)
    assert rc == 0, err.value.decode()
    assert conv.value == 1
    assert np.max(np.abs(x - g["iterate"][:p.n])) < 1e-5  # delta^2 |lambda| perturbation of the constraint
    assert np.max(np.abs(p.C @ x - p.d)) < 1e-6


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


def _write_qp(path, p):
    with open(path, "w") as f:
        f.write("# written by tests/test_adapter.py\nn %d m_ineq %d m_eq %d\n" % (p.n, p.m_ineq, p.m_eq))
        names = {ol.NONE: "none", ol.LOWER: "lower", ol.UPPER: "upper", ol.BOTH: "both"}
        f.write("inequalities %s\nvariable_bounds %s\n" % (names[p.ineq_bounds], names[p.var_bounds]))
        for key, arr in (("Q", p.Q), ("c", p.c), ("A", p.A), ("l_A", p.l_A), ("u_A", p.u_A), ("C", p.C), ("d", p.d),
                         ("l_x", p.l_x), ("u_x", p.u_x)):
            if arr is not None and arr.size:
                f.write(key + "\n" + " ".join(repr(float(v)) for v in np.asarray(arr).ravel()) + "\n")


@pytest.mark.gpu
def test_cli_demo_files_and_batch(tmp_path):
    """ipm-zoo_b200/host/ipmz_cli: the reference's `IpmZoo -n` demo, a problem file in every
    reduction, and several files of one shape as one batch -- compared with the reference-generated goldens."""
    import subprocess
    exe = os.path.join(ROOT, "ipm-zoo_b200", "host", "ipmz_cli")
    if not os.path.exists(exe):
        pytest.skip("ipmz_cli not built")
    run = lambda *a: subprocess.run([exe] + list(a), capture_output=True, text=True, timeout=120)
    out = run("-n")
    assert out.returncode == 0, out.stdout + out.stderr
    assert "iterations: 12 converged: 1" in out.stdout
    f12 = float([l for l in out.stdout.splitlines() if l.startswith("iter: 12")][0].split("f: ")[1].split(",")[0])
    assert abs(f12 - (-1.12799999999863552e+01)) < 1e-8
    # one file, all three reductions, against the golden of the same case
    p = CASES["ineq_box_20x10"]()
    g = np.load(os.path.join(GOLD, "ineq_box_20x10.npz"))
    k = int(g["iterations"])
    path = str(tmp_path / "p0.qp")
    _write_qp(path, p)
    for red in ("augmented", "normal", "full"):
        out = run("--reduction", red, path)
        assert out.returncode == 0, out.stdout + out.stderr
        assert ("iterations: %d converged: 1" % k) in out.stdout
        x = np.array([float(t) for t in [l for l in out.stdout.splitlines() if l.startswith("x:")][0].split()[1:]])
        assert np.max(np.abs(x - g["iterate"][:p.n])) < 1e-6
    # equality rows, both handlings
    q = CASES["eq_box_40x20"]()
    pq = str(tmp_path / "eq.qp")
    _write_qp(pq, q)
    ge = np.load(os.path.join(GOLD, "eq_box_40x20.npz"))
    for mode in ("slacked", "none"):
        out = run("--equalities", mode, "--quiet", pq)
        assert out.returncode == 0, out.stdout + out.stderr
        x = np.array([float(t) for t in [l for l in out.stdout.splitlines() if l.startswith("x:")][0].split()[1:]])
        assert np.max(np.abs(x - ge["iterate"][:q.n])) < 1e-6
    # a batch of three files of one shape
    import problems as P
    paths, probs = [], []
    for i in range(3):
        probs.append(P.ineq_box(20, 10, 300 + i))
        paths.append(str(tmp_path / ("b%d.qp" % i)))
        _write_qp(paths[-1], probs[-1])
    out = run("--reduction", "normal", *paths)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = [l for l in out.stdout.splitlines() if l.startswith("problem:")]
    assert len(lines) == 3
    for i, l in enumerate(lines):
        tr = ol.port_solve(probs[i])
        assert ("iterations: %d, converged: 1" % tr.iterations) in l
        x = np.array([float(t) for t in l.split("x:")[1].split()])
        assert np.max(np.abs(x - tr.iterate[:20])) < 1e-6
    # error convention: unknown keyword -> message on stderr, non-zero exit
    bad = str(tmp_path / "bad.qp")
    open(bad, "w").write("n 2 bogus 1\n")
    out = run(bad)
    assert out.returncode == 1 and "unknown keyword" in out.stderr


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(SO), reason="adapter test library not prebuilt")
@pytest.mark.parametrize("name", ["toy", "ineq_box_20x10", "eq_box_40x20", "box_30", "ineq_upper_box_lower_30x12"])
@pytest.mark.parametrize("reduction", [0, 1])
def test_reference_adapter_environment_contract(name, reduction):
    """Everything the UNMODIFIED reference Optimizer leaves in its Environment -- the iterate, the last iteration's
    `\\Delta v` and `\\Delta v_affine` directions (Optimizer.cpp:369, :377, :200), the shorthand residuals r_{v}
    (:147-157, :188-209) and mu (:179) -- is present in the adapter's Environment with the same values: the harness
    runs both on identically built Environments and compares every key of the reference's."""
    L = C.CDLL(SO)
    p = CASES[name]()
    out = np.zeros(4)
    key = C.create_string_buffer(256)
    err = C.create_string_buffer(512)
    P = lambda a: a.ctypes.data_as(dp) if a is not None and a.size else None
    rc = L.adapter_env_contract(p.n, p.m_ineq, p.m_eq, P(p.Q), P(p.c), P(p.A), P(p.l_A), P(p.u_A), P(p.C), P(p.d),
                                P(p.l_x), P(p.u_x), p.ineq_bounds, p.var_bounds, int(p.equalities), reduction,
                                P(out), key, 256, err, 512)
    assert rc == 0, err.value.decode()
    nkeys, missing, worst = int(out[0]), int(out[1]), out[2]
    assert nkeys > 20
    assert missing == 0, key.value.decode()
    # directions of the last iteration are O(1e-9) themselves; entries are compared relative to max(1, |ref|_inf)
    assert worst < 1e-7, "%s differs by %.3e" % (key.value.decode(), worst)
