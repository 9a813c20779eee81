"""ctypes bindings of oracle/oracle.h (TEST INFRASTRUCTURE).

Loads oracle/_build/libipmzoo_oracle.so (plain-C port, `orc_*`) and, when present,
oracle/_ref/libipmzoo_ref.so (the unmodified reference behind oracle/ref_harness.cpp,
`ref_*`).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
PORT_SO = os.path.join(ORACLE_DIR, "_build", "libipmzoo_oracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libipmzoo_ref.so")

NONE, LOWER, UPPER, BOTH = 0, 1, 2, 3
dp = C.POINTER(C.c_double)


class OrcProblem(C.Structure):
    _fields_ = [("n", C.c_int), ("m_ineq", C.c_int), ("m_eq", C.c_int),
                ("Q", dp), ("c", dp), ("A", dp), ("l_A", dp), ("u_A", dp),
                ("C", dp), ("d", dp), ("l_x", dp), ("u_x", dp),
                ("ineq_bounds", C.c_int), ("var_bounds", C.c_int), ("equalities", C.c_int)]


class OrcTrace(C.Structure):
    _fields_ = [("cap_iters", C.c_int), ("stop_after_cap", C.c_int),
                ("iterations", C.c_int), ("converged", C.c_int), ("n_logged", C.c_int),
                ("f", dp), ("res", dp), ("mu", dp),
                ("rhs_aff", dp), ("step_aff", dp), ("rhs_cor", dp), ("step_cor", dp),
                ("alpha_aff", dp), ("sigma", dp), ("alpha", dp),
                ("iterate", dp), ("use_initial_iterate", C.c_int), ("seconds", C.c_double)]


def _ptr(a):
    return a.ctypes.data_as(dp) if a is not None and a.size else None


class Problem:
    """Dense QP in the reference's `Data` + `Settings` vocabulary (EnvironmentBuilder.h:7-17)."""

    def __init__(self, Q, c, A=None, l_A=None, u_A=None, Ceq=None, d=None, l_x=None, u_x=None,
                 ineq_bounds=BOTH, var_bounds=BOTH, equalities=False):
        f = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
        self.Q, self.c = f(Q), f(c)
        self.n = self.Q.shape[0]
        self.A, self.l_A, self.u_A = f(A), f(l_A), f(u_A)
        self.C, self.d = f(Ceq), f(d)
        self.l_x, self.u_x = f(l_x), f(u_x)
        self.m_ineq = 0 if self.A is None else self.A.shape[0]
        self.m_eq = 0 if self.C is None else self.C.shape[0]
        self.ineq_bounds = ineq_bounds if self.m_ineq else NONE
        self.var_bounds = var_bounds
        self.equalities = bool(equalities) and self.m_eq > 0

    @property
    def N(self):
        return self.n + self.m_ineq + self.m_eq

    @property
    def iterate_len(self):
        return 5 * self.n + 6 * self.m_ineq + 6 * self.m_eq

    def c_struct(self):
        return OrcProblem(self.n, self.m_ineq, self.m_eq, _ptr(self.Q), _ptr(self.c),
                          _ptr(self.A), _ptr(self.l_A), _ptr(self.u_A), _ptr(self.C),
                          _ptr(self.d), _ptr(self.l_x), _ptr(self.u_x),
                          self.ineq_bounds, self.var_bounds, int(self.equalities))

    # offsets of the named vectors inside the packed iterate (oracle.h)
    def offsets(self):
        n, mi, me = self.n, self.m_ineq, self.m_eq
        names = [("x", n), ("lamA", mi), ("s", mi), ("lamg", mi), ("lamh", mi), ("g", mi), ("h", mi),
                 ("lamC", me), ("t", me), ("lamv", me), ("lamw", me), ("v", me), ("w", me),
                 ("lamy", n), ("lamz", n), ("y", n), ("z", n)]
        out, off = {}, 0
        for k, ln in names:
            out[k] = (off, ln)
            off += ln
        return out


class Trace:
    def __init__(self, prob, cap_iters=100, stop_after_cap=False, iterate=None, steps=True):
        N = prob.N
        self.cap = cap_iters
        self.f = np.zeros(cap_iters + 1)
        self.res = np.zeros(cap_iters + 1)
        self.mu = np.zeros(cap_iters + 1)
        z = lambda: np.zeros((cap_iters, N)) if steps else None
        self.rhs_aff, self.step_aff, self.rhs_cor, self.step_cor = z(), z(), z(), z()
        self.alpha_aff = np.zeros(cap_iters)
        self.sigma = np.zeros(cap_iters)
        self.alpha = np.zeros(cap_iters)
        self.iterate = np.zeros(prob.iterate_len)
        use = 0
        if iterate is not None:
            self.iterate[:] = iterate
            use = 1
        self.c = OrcTrace(cap_iters, int(stop_after_cap), 0, 0, 0, _ptr(self.f), _ptr(self.res),
                          _ptr(self.mu), _ptr(self.rhs_aff), _ptr(self.step_aff),
                          _ptr(self.rhs_cor), _ptr(self.step_cor), _ptr(self.alpha_aff),
                          _ptr(self.sigma), _ptr(self.alpha), _ptr(self.iterate), use, 0.0)

    iterations = property(lambda s: s.c.iterations)
    converged = property(lambda s: s.c.converged)
    n_logged = property(lambda s: s.c.n_logged)
    seconds = property(lambda s: s.c.seconds)


def build_oracles(reference=True):
    targets = ["port"] + (["reference"] if reference and os.path.isdir("/root/reference") else [])
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "-j8"] + targets, check=True)


_port = None
_ref = None


def port():
    global _port
    if _port is None:
        if not os.path.exists(PORT_SO):
            build_oracles(reference=False)
        L = C.CDLL(PORT_SO)
        L.orc_solve.argtypes = [C.POINTER(OrcProblem), C.POINTER(OrcTrace)]
        L.orc_initial_iterate.argtypes = [C.POINTER(OrcProblem), dp]
        L.orc_assemble_kkt.argtypes = [C.POINTER(OrcProblem), dp, dp, dp]
        L.orc_ldlt.argtypes = [C.c_int, dp, dp, dp]
        L.orc_solve_ldlt.argtypes = [C.c_int, dp, dp, dp]
        L.orc_bk_factor.argtypes = [C.c_int, dp, dp, C.POINTER(C.c_int)]
        L.orc_bk_solve.argtypes = [C.c_int, dp, C.POINTER(C.c_int), dp]
        _port = L
    return _port


def have_ref():
    return os.path.exists(REF_SO)


def ref():
    global _ref
    if _ref is None:
        L = C.CDLL(REF_SO)
        L.ref_solve.argtypes = [C.POINTER(OrcProblem), C.POINTER(OrcTrace), C.c_int]
        L.ref_ldlt.argtypes = [C.c_int, dp, dp, dp]
        L.ref_solve_ldlt.argtypes = [C.c_int, dp, dp, dp]
        L.ref_bk_factor.argtypes = [C.c_int, dp, dp, C.POINTER(C.c_int)]
        L.ref_bk_solve.argtypes = [C.c_int, dp, C.POINTER(C.c_int), dp]
        L.ref_last_error.restype = C.c_char_p
        _ref = L
    return _ref


def ref_solve(prob, cap_iters=100, stop_after_cap=False, iterate=None, quiet=False, steps=True):
    tr = Trace(prob, cap_iters, stop_after_cap, iterate, steps)
    ps = prob.c_struct()
    rc = ref().ref_solve(C.byref(ps), C.byref(tr.c), int(quiet))
    if rc != 0:
        raise RuntimeError("reference: " + ref().ref_last_error().decode())
    return tr


def port_solve(prob, cap_iters=100, stop_after_cap=False, iterate=None, steps=True):
    tr = Trace(prob, cap_iters, stop_after_cap, iterate, steps)
    ps = prob.c_struct()
    rc = port().orc_solve(C.byref(ps), C.byref(tr.c))
    if rc != 0:
        raise RuntimeError("oracle port failed rc=%d" % rc)
    return tr
