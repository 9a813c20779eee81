"""ctypes binding of include/ipmz.h (the C ABI of libipmz_b200.so)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
AUGMENTED, NORMAL, FULL, DUAL_NORMAL = 0, 1, 2, 3
NONE, LOWER, UPPER, BOTH = 0, 1, 2, 3
EQ_OFF, EQ_SLACKED_SLACKS, EQ_NONE, EQ_REGULARIZATION, EQ_PENALTY = 0, 1, 2, 3, 4  # ipmz_problem.equalities (EqualityHandling)
dp = C.POINTER(C.c_double)

EXPORTED_SYMBOLS = [
    "ipmz_last_error", "ipmz_version", "ipmz_device_count", "ipmz_default_options", "ipmz_iterate_len", "ipmz_full_layout",
    "ipmz_host_alloc", "ipmz_host_free", "ipmz_launch_count", "ipmz_fp64_peak_probe",
    "ipmz_create", "ipmz_destroy", "ipmz_set_iterate", "ipmz_get_iterate", "ipmz_reset_iterate",
    "ipmz_solve", "ipmz_newton_step", "ipmz_get_trace", "ipmz_assemble", "ipmz_probe_kernels",
    "ipmz_ldlt_decomposition", "ipmz_overwriting_solve_ldlt",
    "ipmz_symmetric_indefinite_factorization", "ipmz_overwriting_solve_bunch_kaufman", "ipmz_bk_factor_time",
    "ipmz_factor_create", "ipmz_factor_destroy", "ipmz_factor_set_matrix", "ipmz_factor_set_rhs",
    "ipmz_factor_run", "ipmz_factor_profile", "ipmz_factor_get_solution", "ipmz_factor_get_ld",
    "ipmz_factor_info", "ipmz_schedule_check", "ipmz_assembly_schedule_check",
    "ipmz_batch_create", "ipmz_batch_destroy", "ipmz_batch_upload", "ipmz_batch_solve",
    "ipmz_batch_get_iterates", "ipmz_batch_get_x", "ipmz_batch_solve_group", "ipmz_batch_results",
    "ipmz_batch_solve_streamed", "ipmz_get_last_iteration",
]


class IpmzError(RuntimeError):
    """Non-zero ipmz_status (the C++ adapter raises std::logic_error for the same codes)."""

    def __init__(self, code, msg):
        super().__init__("ipmz status %d: %s" % (code, msg))
        self.code = code


class _Problem(C.Structure):
    _fields_ = [("n", C.c_int), ("m_ineq", C.c_int), ("m_eq", C.c_int),
                ("Q", dp), ("c", dp), ("A", dp), ("l_A", dp), ("u_A", dp),
                ("C", dp), ("d", dp), ("l_x", dp), ("u_x", dp),
                ("ineq_bounds", C.c_int), ("var_bounds", C.c_int), ("equalities", C.c_int)]


class _Options(C.Structure):
    _fields_ = [("tolerance", C.c_double), ("max_iter", C.c_int),
                ("fraction_to_boundary", C.c_double), ("sigma_power", C.c_double),
                ("reduction", C.c_int), ("device", C.c_int), ("record_steps", C.c_int),
                ("refine_steps", C.c_int), ("delta_eq", C.c_double)]


class _Result(C.Structure):
    _fields_ = [("iterations", C.c_int), ("converged", C.c_int), ("f", C.c_double),
                ("res", C.c_double), ("mu", C.c_double), ("solve_ms", C.c_double),
                ("factor_flops", C.c_double)]


def lib_path():
    # IPMZ_LIB: A/B runs of two builds inside one GPU call (tools/); the product is the in-tree library
    return os.environ.get("IPMZ_LIB") or os.path.join(HERE, "libipmz_b200.so")


def build(verbose=False):
    """Compile csrc/*.cu for sm_100a into libipmz_b200.so (nvcc cross-compiles without a GPU)."""
    env = dict(os.environ)
    if verbose:
        env["VERBOSE"] = "1"
    subprocess.run(["sh", os.path.join(HERE, "build.sh")], check=True, env=env)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(lib_path()):
            raise IpmzError(2, "libipmz_b200.so is not built (run ipm-zoo_b200/build.sh); "
                               "there is no CPU fallback")
        L = C.CDLL(lib_path())
        L.ipmz_last_error.restype = C.c_char_p
        L.ipmz_version.restype = C.c_char_p
        L.ipmz_host_alloc.restype = C.c_void_p
        L.ipmz_host_alloc.argtypes = [C.c_size_t]
        L.ipmz_host_free.argtypes = [C.c_void_p]
        vp = C.c_void_p
        L.ipmz_default_options.argtypes = [C.POINTER(_Options)]
        L.ipmz_iterate_len.argtypes = [C.POINTER(_Problem)]
        L.ipmz_full_layout.argtypes = [C.POINTER(_Problem), C.POINTER(C.c_int)]
        L.ipmz_assembly_schedule_check.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.ipmz_create.argtypes = [C.POINTER(_Problem), C.POINTER(_Options), C.POINTER(vp)]
        L.ipmz_destroy.argtypes = [vp]
        L.ipmz_set_iterate.argtypes = [vp, dp]
        L.ipmz_get_iterate.argtypes = [vp, dp]
        L.ipmz_reset_iterate.argtypes = [vp]
        L.ipmz_solve.argtypes = [vp, C.POINTER(_Result)]
        L.ipmz_newton_step.argtypes = [vp, dp, dp, dp, dp, dp]
        L.ipmz_get_trace.argtypes = [vp, C.c_int, dp, dp, dp, dp, dp, dp, dp, dp]
        L.ipmz_assemble.argtypes = [vp, dp, C.POINTER(C.c_int)]
        L.ipmz_probe_kernels.argtypes = [vp, C.c_int, dp, dp]
        L.ipmz_ldlt_decomposition.argtypes = [C.c_int, dp, dp, dp]
        L.ipmz_overwriting_solve_ldlt.argtypes = [C.c_int, dp, dp, dp]
        L.ipmz_symmetric_indefinite_factorization.argtypes = [C.c_int, dp, dp, C.POINTER(C.c_int)]
        L.ipmz_overwriting_solve_bunch_kaufman.argtypes = [C.c_int, dp, C.POINTER(C.c_int), dp]
        L.ipmz_bk_factor_time.argtypes = [C.c_int, dp, C.c_int, dp]
        L.ipmz_factor_create.argtypes = [C.c_int, C.c_int, C.POINTER(vp)]
        L.ipmz_factor_destroy.argtypes = [vp]
        L.ipmz_factor_set_matrix.argtypes = [vp, dp]
        L.ipmz_factor_set_rhs.argtypes = [vp, dp]
        L.ipmz_factor_run.argtypes = [vp, C.c_int, C.c_int, dp]
        L.ipmz_factor_get_solution.argtypes = [vp, dp]
        L.ipmz_factor_profile.argtypes = [vp, dp, dp, C.POINTER(C.c_int)]
        L.ipmz_launch_count.restype = C.c_ulonglong
        L.ipmz_fp64_peak_probe.argtypes = [C.c_int, dp]
        L.ipmz_factor_get_ld.argtypes = [vp, dp, dp]
        L.ipmz_batch_create.argtypes = [C.c_int, C.POINTER(_Problem), C.POINTER(_Options), C.POINTER(vp)]
        L.ipmz_batch_destroy.argtypes = [vp]
        L.ipmz_batch_upload.argtypes = [vp, C.POINTER(_Problem)]
        L.ipmz_batch_solve.argtypes = [vp, C.POINTER(_Result), dp]
        L.ipmz_batch_solve_streamed.argtypes = [vp, C.POINTER(_Problem), C.c_int, C.POINTER(_Result), dp]
        L.ipmz_batch_get_iterates.argtypes = [vp, dp]
        L.ipmz_batch_get_x.argtypes = [vp, dp]
        L.ipmz_batch_solve_group.argtypes = [C.c_int, C.POINTER(vp), dp]
        L.ipmz_batch_results.argtypes = [vp, C.POINTER(_Result)]
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise IpmzError(rc, lib().ipmz_last_error().decode())


def device_count():
    return lib().ipmz_device_count()


def shard_range(total, world_size, rank):
    """Contiguous block of problem indices [lo, hi) owned by `rank` when `total` independent QPs
    are sharded over `world_size` GPUs; no data-path collective."""
    return rank * total // world_size, (rank + 1) * total // world_size


def launch_count():
    return int(lib().ipmz_launch_count())


def fp64_peak_tflops(device=0):
    t = C.c_double()
    _check(lib().ipmz_fp64_peak_probe(device, C.byref(t)))
    return t.value


def _ptr(a):
    return a.ctypes.data_as(dp) if a is not None and a.size else None


def pinned_empty(shape, dtype=np.float64):
    """numpy array over page-locked host memory (ipmz_host_alloc); lives until process exit."""
    count = int(np.prod(shape))
    nbytes = max(count * np.dtype(dtype).itemsize, 8)
    ptr = lib().ipmz_host_alloc(nbytes)
    if not ptr:
        raise IpmzError(3, "ipmz_host_alloc failed")
    buf = (C.c_char * nbytes).from_address(ptr)
    _PINNED.append(ptr)
    return np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)


_PINNED = []


class Problem:
    """Dense QP(s) in the reference's Data + Settings vocabulary (EnvironmentBuilder.h:7-17,
    SymbolicOptimization.h:58-64).  For a batch every array carries a leading `count` axis."""

    def __init__(self, Q, c, A=None, l_A=None, u_A=None, Ceq=None, d=None, l_x=None, u_x=None,
                 ineq_bounds=BOTH, var_bounds=BOTH, equalities=False, count=None):
        f = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
        self.Q, self.c = f(Q), f(c)
        self.A, self.l_A, self.u_A = f(A), f(l_A), f(u_A)
        self.C, self.d = f(Ceq), f(d)
        self.l_x, self.u_x = f(l_x), f(u_x)
        self.count = count
        self.n = self.Q.shape[-1]
        self.m_ineq = 0 if self.A is None else self.A.shape[-2]
        self.m_eq = 0 if self.C is None else self.C.shape[-2]
        self.ineq_bounds = ineq_bounds if self.m_ineq else NONE
        self.var_bounds = var_bounds
        self.equalities = int(equalities) if self.m_eq > 0 else 0

    @classmethod
    def from_data(cls, p):
        """Same QP as another Problem-like object with the reference's field names (identical bytes)."""
        return cls(p.Q, p.c, p.A, p.l_A, p.u_A, p.C, p.d, p.l_x, p.u_x,
                   p.ineq_bounds, p.var_bounds, p.equalities)

    @property
    def N(self):
        """Rows of the solved augmented steps the library writes (n + the constraint rows the Settings keep):
        inequality rows count only when `ineq_bounds` is not NONE, equality rows only when `equalities` is set
        (fill_shape in csrc/solver.cu)."""
        mi = self.m_ineq if self.ineq_bounds != NONE else 0
        me = self.m_eq if self.equalities else 0
        return self.n + mi + me

    @property
    def iterate_len(self):
        return 5 * self.n + 6 * self.m_ineq + 6 * self.m_eq

    def c_struct(self):
        return _Problem(self.n, self.m_ineq, self.m_eq, _ptr(self.Q), _ptr(self.c), _ptr(self.A),
                        _ptr(self.l_A), _ptr(self.u_A), _ptr(self.C), _ptr(self.d), _ptr(self.l_x),
                        _ptr(self.u_x), self.ineq_bounds, self.var_bounds, int(self.equalities))


class Options:
    def __init__(self, reduction=AUGMENTED, device=0, record_steps=False, tolerance=1e-8, max_iter=100,
                 fraction_to_boundary=0.995, sigma_power=3.0, refine_steps=-1, delta_eq=1e-4):
        self.c = _Options(tolerance, max_iter, fraction_to_boundary, sigma_power, reduction, device,
                          int(record_steps), refine_steps, delta_eq)


class Result:
    def __init__(self, r):
        self.iterations, self.converged = r.iterations, bool(r.converged)
        self.f, self.res, self.mu = r.f, r.res, r.mu
        self.solve_ms, self.factor_flops = r.solve_ms, r.factor_flops

    def __repr__(self):
        return ("Result(iterations=%d, converged=%s, f=%.17g, res=%.3e, mu=%.3e, solve_ms=%.3f)" %
                (self.iterations, self.converged, self.f, self.res, self.mu, self.solve_ms))


class Solver:
    """Mirror of NumericalOptimization::Optimizer for one QP (Optimizer.h:13-20): construct,
    solve(); the iterate lives on the device and is read back with iterate()."""

    def __init__(self, problem, options=None):
        self.p = problem
        self.opt = options or Options()
        self._h = C.c_void_p()
        ps = problem.c_struct()
        _check(lib().ipmz_create(C.byref(ps), C.byref(self.opt.c), C.byref(self._h)))

    def close(self):
        if self._h:
            lib().ipmz_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def solve(self):
        r = _Result()
        _check(lib().ipmz_solve(self._h, C.byref(r)))
        return Result(r)

    def iterate(self):
        out = np.zeros(self.p.iterate_len)
        _check(lib().ipmz_get_iterate(self._h, _ptr(out)))
        return out

    def set_iterate(self, packed):
        packed = np.ascontiguousarray(packed, dtype=np.float64)
        assert packed.size == self.p.iterate_len
        _check(lib().ipmz_set_iterate(self._h, _ptr(packed)))

    def reset_iterate(self):
        _check(lib().ipmz_reset_iterate(self._h))

    def newton_step(self):
        N = self.p.N
        sa, sc = np.zeros(N), np.zeros(N)
        aa, sg, al = C.c_double(), C.c_double(), C.c_double()
        _check(lib().ipmz_newton_step(self._h, _ptr(sa), _ptr(sc), C.byref(aa), C.byref(sg), C.byref(al)))
        return sa, sc, aa.value, sg.value, al.value

    def trace(self, iterations, steps=False):
        cap = iterations + 1
        f, res, mu = np.zeros(cap), np.zeros(cap), np.zeros(cap)
        aa, sg, al = np.zeros(cap), np.zeros(cap), np.zeros(cap)
        sa = np.zeros((cap, self.p.N)) if steps else None
        sc = np.zeros((cap, self.p.N)) if steps else None
        _check(lib().ipmz_get_trace(self._h, cap, _ptr(f), _ptr(res), _ptr(mu), _ptr(sa), _ptr(sc),
                                    _ptr(aa), _ptr(sg), _ptr(al)))
        return dict(f=f, res=res, mu=mu, alpha_aff=aa[:iterations], sigma=sg[:iterations],
                    alpha=al[:iterations], step_aff=None if sa is None else sa[:iterations],
                    step_cor=None if sc is None else sc[:iterations])

    PROBE_SLOTS = ("k_matvec Q x", "k_matvec M x", "k_matvec M^T lambda", "k_assemble", "k_residuals_rhs<0>",
                   "k_backsub_step<0>", "k_update")

    def probe_kernels(self, reps=20):
        """[(kernel, ms per launch, algorithmic bytes per launch)] of the streaming kernels (CUDA events)."""
        ms, by = np.zeros(7), np.zeros(7)
        _check(lib().ipmz_probe_kernels(self._h, reps, _ptr(ms), _ptr(by)))
        return [(self.PROBE_SLOTS[i], float(ms[i]), float(by[i])) for i in range(7)]

    def assemble(self):
        N = self.p.N if self.opt.c.reduction != FULL else 5 * self.p.n + 6 * (self.p.N - self.p.n)  # DUAL_NORMAL: m <= N
        K = np.zeros((N, N))
        nout = C.c_int()
        _check(lib().ipmz_assemble(self._h, _ptr(K), C.byref(nout)))
        n = nout.value
        return K.ravel()[:n * n].reshape(n, n).copy()


class BatchSolver:
    """`count` independent QPs of one shape (cfg4), solved in lock-step on one device."""

    def __init__(self, problem, count, options=None):
        self.p, self.count = problem, count
        self.opt = options or Options()
        self._h = C.c_void_p()
        ps = problem.c_struct()
        _check(lib().ipmz_batch_create(count, C.byref(ps), C.byref(self.opt.c), C.byref(self._h)))

    def close(self):
        if self._h:
            lib().ipmz_batch_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def upload(self, problem=None):
        ps = (problem or self.p).c_struct()
        _check(lib().ipmz_batch_upload(self._h, C.byref(ps)))

    def solve(self, per_problem=True):
        arr = (_Result * self.count)() if per_problem else None
        ms = C.c_double()
        _check(lib().ipmz_batch_solve(self._h, arr, C.byref(ms)))
        return ([Result(r) for r in arr] if per_problem else None), ms.value

    def solve_streamed(self, problem=None, chunks=8, per_problem=True):
        """Upload + solve pipelined (ipmz_batch_solve_streamed): the persistent kernel takes problems as the copy stream
        delivers the `chunks` groups of host data."""
        ps = (problem or self.p).c_struct()
        arr = (_Result * self.count)() if per_problem else None
        ms = C.c_double()
        _check(lib().ipmz_batch_solve_streamed(self._h, C.byref(ps), chunks, arr, C.byref(ms)))
        return ([Result(r) for r in arr] if per_problem else None), ms.value

    def results(self):
        """Per-problem outcome of the last solve (also after solve_group)."""
        arr = (_Result * self.count)()
        _check(lib().ipmz_batch_results(self._h, arr))
        return [Result(r) for r in arr]

    def x(self, out=None):
        out = np.zeros((self.count, self.p.n)) if out is None else out
        _check(lib().ipmz_batch_get_x(self._h, _ptr(out)))
        return out

    def iterates(self):
        out = np.zeros((self.count, self.p.iterate_len))
        _check(lib().ipmz_batch_get_iterates(self._h, _ptr(out)))
        return out


def solve_group(solvers):
    """Solve several BatchSolver handles of one device concurrently (one host thread + CUDA stream each,
    native threads inside the library); returns the device time in ms from the earliest start to the
    latest end.  Per-problem outcomes: BatchSolver.results()."""
    hs = (C.c_void_p * len(solvers))(*[s._h for s in solvers])
    ms = C.c_double()
    _check(lib().ipmz_batch_solve_group(len(solvers), hs, C.byref(ms)))
    return ms.value


class Factor:
    """Device-resident LDL^T factor + solves of one dense symmetric matrix (bench path)."""

    def __init__(self, n, device=0):
        self.n = n
        self._h = C.c_void_p()
        _check(lib().ipmz_factor_create(n, device, C.byref(self._h)))

    def close(self):
        if self._h:
            lib().ipmz_factor_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def set_matrix(self, A):
        A = np.ascontiguousarray(A, dtype=np.float64)
        assert A.shape == (self.n, self.n)
        _check(lib().ipmz_factor_set_matrix(self._h, _ptr(A)))

    def set_rhs(self, b):
        b = np.ascontiguousarray(b, dtype=np.float64)
        _check(lib().ipmz_factor_set_rhs(self._h, _ptr(b)))

    def run(self, reps=1, nrhs=1):
        ms = C.c_double()
        _check(lib().ipmz_factor_run(self._h, reps, nrhs, C.byref(ms)))
        return ms.value

    def profile(self):
        """One factorization with events around every launch -> dict of per-kernel-class ms."""
        ms = np.zeros(3)
        fl, ns = C.c_double(), C.c_int()
        _check(lib().ipmz_factor_profile(self._h, _ptr(ms), C.byref(fl), C.byref(ns)))
        return dict(diag_ms=ms[0], panel_ms=ms[1], syrk_ms=ms[2], syrk_flops=fl.value, syrk_launches=ns.value)

    def info(self):
        """Factorization path of this handle: dict(dataflow, ntasks, simulated_us)."""
        d, nt, us = C.c_int(), C.c_int(), C.c_double()
        _check(lib().ipmz_factor_info(self._h, C.byref(d), C.byref(nt), C.byref(us)))
        return dict(dataflow=bool(d.value), ntasks=nt.value, simulated_us=us.value)

    def solution(self):
        x = np.zeros(self.n)
        _check(lib().ipmz_factor_get_solution(self._h, _ptr(x)))
        return x

    def ld(self):
        L, D = np.zeros((self.n, self.n)), np.zeros(self.n)
        _check(lib().ipmz_factor_get_ld(self._h, _ptr(L), _ptr(D)))
        return L, D


def schedule_check(n, workers=148):
    """Host-only: the dataflow task list for an n x n matrix is a valid topological order.
    Returns dict(valid, diag, trsm, upd, makespan_us, work_us)."""
    cnt = (C.c_int * 3)()
    mk, wk = C.c_double(), C.c_double()
    rc = lib().ipmz_schedule_check(int(n), int(workers), cnt, C.byref(mk), C.byref(wk))
    return dict(valid=rc == 0, diag=cnt[0], trsm=cnt[1], upd=cnt[2], makespan_us=mk.value, work_us=wk.value)


def ldlt_decomposition(A):
    """LinearSolvers::ldlt_decomposition (LinearSolvers.h:11): returns (L, D)."""
    A = np.ascontiguousarray(A, dtype=np.float64)
    n = A.shape[0]
    L, D = np.zeros((n, n)), np.zeros(n)
    _check(lib().ipmz_ldlt_decomposition(n, _ptr(A), _ptr(L), _ptr(D)))
    return L, D


def symmetric_indefinite_factorization(A):
    """LinearSolvers::symmetric_indefinite_factorization (LinearSolvers.h:24-25): returns (LD, ipiv)."""
    A = np.ascontiguousarray(A, dtype=np.float64)
    n = A.shape[0]
    LD = np.zeros((n, n))
    ipiv = np.zeros(n, dtype=np.int32)
    _check(lib().ipmz_symmetric_indefinite_factorization(n, _ptr(A), _ptr(LD), ipiv.ctypes.data_as(C.POINTER(C.c_int))))
    return LD, ipiv


def bk_factor_time(A, reps=3):
    """Device ms of one Bunch-Kaufman factorization of A (CUDA events)."""
    A = np.ascontiguousarray(A, dtype=np.float64)
    ms = C.c_double()
    _check(lib().ipmz_bk_factor_time(A.shape[0], _ptr(A), reps, C.byref(ms)))
    return ms.value


def overwriting_solve_bunch_kaufman(LD, ipiv, b):
    """LinearSolvers::overwriting_solve_bunch_kaufman (LinearSolvers.h:29-31): b is overwritten."""
    LD = np.ascontiguousarray(LD, dtype=np.float64)
    ipiv = np.ascontiguousarray(ipiv, dtype=np.int32)
    assert b.dtype == np.float64 and b.flags.c_contiguous
    _check(lib().ipmz_overwriting_solve_bunch_kaufman(LD.shape[0], _ptr(LD), ipiv.ctypes.data_as(C.POINTER(C.c_int)),
                                                     _ptr(b)))
    return b


def full_layout(problem):
    """Offsets of the FULL reduction's unknown groups (host only): dict name -> start, plus 'N'."""
    ps = problem.c_struct()
    o = (C.c_int * 12)()
    _check(lib().ipmz_full_layout(C.byref(ps), o))
    names = ("y", "z", "sl", "su", "ly", "lz", "ll", "lu", "s", "x", "lam", "N")
    return dict(zip(names, list(o)))


def assembly_schedule_check(n, m):
    """(valid, ntasks) of the condensed-assembly task list (host only, no GPU needed)."""
    nt = C.c_int()
    rc = lib().ipmz_assembly_schedule_check(n, m, C.byref(nt))
    return rc == 0, nt.value


def overwriting_solve_ldlt(L, D, b):
    """LinearSolvers::overwriting_solve_ldlt (LinearSolvers.h:16-17): b is overwritten."""
    L = np.ascontiguousarray(L, dtype=np.float64)
    D = np.ascontiguousarray(D, dtype=np.float64)
    assert b.dtype == np.float64 and b.flags.c_contiguous
    _check(lib().ipmz_overwriting_solve_ldlt(L.shape[0], _ptr(L), _ptr(D), _ptr(b)))
    return b
