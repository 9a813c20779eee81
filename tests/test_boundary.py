"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/ipmz.h declares, and refuses to compute without a GPU (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def z():
    import ipm_zoo_b200 as z
    if not os.path.exists(z.lib_path()):
        z.build()
    return z


def test_header_symbols_are_exported(z):
    hdr = open(os.path.join(ROOT, "include", "ipmz.h")).read()
    declared = set(re.findall(r"\b(ipmz_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations found"
    L = z.lib()
    for sym in sorted(declared):
        assert hasattr(L, sym), sym
    assert declared == set(z.EXPORTED_SYMBOLS)


def test_no_cpu_fallback(z):
    if z.device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(z.IpmzError) as e:
        z.ldlt_decomposition(np.eye(4))
    assert e.value.code == 2
    with pytest.raises(z.IpmzError):
        z.Solver(z.Problem(np.eye(2), np.zeros(2), l_x=-np.ones(2), u_x=np.ones(2), ineq_bounds=z.NONE))


def test_product_does_not_reference_oracle():
    """The product tree must never include, link, load or execute anything under oracle/."""
    prod = os.path.join(ROOT, "ipm-zoo_b200")
    banned = ("oracle/", "oracle.h", "libipmzoo_oracle", "libipmzoo_ref", "import oracle", "oracle_lib", "orc_", "ref_solve")
    for dirpath, _, files in os.walk(prod):
        if "_obj" in dirpath:
            continue
        for fn in files:
            if fn.endswith((".cu", ".cuh", ".h", ".cpp", ".py", ".sh")):
                txt = open(os.path.join(dirpath, fn)).read()
                for b in banned:
                    assert b not in txt, (fn, b)


def test_default_options_match_reference_constants(z):
    import ctypes as C
    from importlib import import_module
    capi = import_module("ipm_zoo_b200.capi")
    o = capi._Options()
    z.lib().ipmz_default_options(C.byref(o))
    assert (o.tolerance, o.max_iter, o.fraction_to_boundary, o.sigma_power) == (1e-8, 100, 0.995, 3.0)
    assert o.delta_eq == 1e-4  # EnvironmentBuilder.cpp:48


def test_bunch_kaufman_entry_points_refuse_without_gpu(z):
    if z.device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(z.IpmzError) as e:
        z.symmetric_indefinite_factorization(np.array([[0.0, 1.0], [1.0, 0.0]]))
    assert e.value.code == 2
    with pytest.raises(z.IpmzError):
        z.bk_factor_time(np.eye(3))


def test_cli_host_logic_without_gpu(tmp_path):
    """ipmz_cli: argument / file parsing is host logic (runs anywhere); a compute request without a CUDA device
    fails loudly with the library's message -- there is no CPU path behind the CLI either."""
    import subprocess
    exe = os.path.join(ROOT, "ipm-zoo_b200", "host", "ipmz_cli")
    if not os.path.exists(exe):
        pytest.skip("ipmz_cli not built")
    run = lambda *a: subprocess.run([exe] + list(a), capture_output=True, text=True, timeout=60)
    out = run()
    assert out.returncode == 2 and "usage:" in out.stderr
    bad = str(tmp_path / "bad.qp")
    open(bad, "w").write("n 2 # comment\nbogus 1\n")
    out = run(bad)
    assert out.returncode == 1 and "unknown keyword 'bogus'" in out.stderr
    short = str(tmp_path / "short.qp")
    open(short, "w").write("n 2 m_ineq 0 m_eq 0\nQ 1 0 0\n")
    out = run(short)
    assert out.returncode == 1 and "unexpected end" in out.stderr
    out = run("--reduction", "cholesky", "-n")
    assert out.returncode == 1 and "unknown reduction" in out.stderr
    import ipm_zoo_b200 as zz
    if zz.device_count() == 0:
        out = run("-n")
        assert out.returncode == 1 and "no CUDA device" in out.stderr
