"""ctypes loader for oracle/liboracle.so (TEST INFRASTRUCTURE: the checker, never the product).

Builds the library on demand with oracle/Makefile.  Used by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs only.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "liboracle.so")

OBJ_QUARTIC, OBJ_ROSENBROCK, OBJ_DIAGQUAD, OBJ_QUARTIC_SHIFTED = 0, 1, 2, 3
START_QUARTIC_U, START_ROSEN_STD, START_ROSEN_PERT, START_ZERO = 0, 1, 2, 3

F_T = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int))
FD_T = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int))
FFD_T = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                    C.POINTER(C.c_int))
TRACE_T = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double),
                      C.POINTER(C.c_double), C.c_double, C.c_double, C.c_double, C.c_long)


class Stats(C.Structure):
    _fields_ = [("n_f", C.c_long), ("n_fd", C.c_long), ("n_ffd", C.c_long), ("n_trials", C.c_long),
                ("n_linesearch", C.c_long), ("n_iter", C.c_long), ("n_quirk_f9", C.c_long),
                ("status", C.c_int)]


# FLGPU_ORACLE_VARIANT=omp (set by bench.py's CPU-sample subprocess, nowhere else) loads liboracle_omp.so: the same
# source with parallel loops, -O3 -march=native -fopenmp -- the GENEROUS CPU baseline, never a checker (its sums are
# reassociated).  It is rebuilt on the machine it runs on because of -march=native.
OMP_VARIANT = os.environ.get("FLGPU_ORACLE_VARIANT", "") == "omp"
if OMP_VARIANT:
    LIB_PATH = os.path.join(ORACLE_DIR, "liboracle_omp.so")


def build(force=False):
    target = os.path.basename(LIB_PATH)
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("oracle.c", "objectives.c", "oracle.h", "Makefile")]
    stale = (not os.path.exists(LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if force or stale or OMP_VARIANT:
        subprocess.run(["make", "-C", ORACLE_DIR, "-B", target], check=True, stdout=subprocess.DEVNULL,
                       stderr=subprocess.DEVNULL if OMP_VARIANT else None)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        _lib.orc_diag_coeff.restype = C.c_double
        _lib.orc_diag_coeff.argtypes = [C.c_longlong, C.c_longlong]
        _lib.orc_obj_select.argtypes = [C.c_int, C.c_longlong, C.c_longlong]
        _lib.orc_obj_start.argtypes = [C.c_int, C.c_ulonglong, C.c_void_p, C.c_longlong, C.c_longlong,
                                       C.c_longlong]
        _lib.orc_set_trace.argtypes = [C.c_void_p, C.c_void_p]
        _lib.orc_set_sum_mode.argtypes = [C.c_int]
        _lib.orc_set_line_search.argtypes = [C.c_int]
        _lib.orc_set_eval_budget.argtypes = [C.c_long]
    return _lib


class eval_budget:
    """Context manager: the oracle gives up (status 9) after n callback invocations inside line searches.  The
    reference never terminates once a step is NaN (f90:1518-1546) and a ctypes callback cannot unwind C, so tests that
    generate objectives at random run the oracle under a budget and skip the cases that exhaust it."""

    def __init__(self, n):
        self.n = n

    def __enter__(self):
        lib().orc_set_eval_budget(self.n)
        return self

    def __exit__(self, *exc):
        lib().orc_set_eval_budget(0)
        return False


class fast_line_search:
    """Context manager: the oracle's drivers search with its restatement of the product's FLGPU_LS_FAST policy
    (NOT a reference routine, see oracle.c) instead of the reference's searchers."""

    def __enter__(self):
        lib().orc_set_line_search(1)
        return self

    def __exit__(self, *exc):
        lib().orc_set_line_search(0)
        return False


def _opt(ctype, v):
    return None if v is None else C.byref(ctype(v))


def start_vector(start_kind, n, seed=0, offset=0, n_global=None):
    x = np.empty(n, dtype=np.float64)
    lib().orc_obj_start(start_kind, seed, x.ctypes.data, offset, n, n if n_global is None else n_global)
    return x


def builtin_callbacks(kind, offset=0, n_global=0):
    L = lib()
    L.orc_obj_select(kind, offset, n_global)
    return (C.cast(L.orc_obj_f, C.c_void_p), C.cast(L.orc_obj_fd, C.c_void_p),
            C.cast(L.orc_obj_f_fd, C.c_void_p))


class Trace:
    """Collects (iter, p, x, g, a, f, phid0, trials) after every line search."""

    def __init__(self, keep_vectors=True, max_vec_iters=10**9):
        self.rows = []
        self.p, self.x, self.g = [], [], []
        self.keep = keep_vectors
        self.max_vec_iters = max_vec_iters
        self._cb = TRACE_T(self._on)

    def _on(self, user, it, dim, p, x, g, a, fx, phid0, trials):
        self.rows.append((it, a, fx, phid0, trials))
        if self.keep and it < self.max_vec_iters:
            self.p.append(np.ctypeslib.as_array(p, (dim,)).copy())
            self.x.append(np.ctypeslib.as_array(x, (dim,)).copy())
            self.g.append(np.ctypeslib.as_array(g, (dim,)).copy())

    def install(self):
        lib().orc_set_trace(C.cast(self._cb, C.c_void_p), None)

    @staticmethod
    def uninstall():
        lib().orc_set_trace(None, None)


def stats():
    s = Stats()
    lib().orc_get_stats(C.byref(s))
    return s


def lbfgs(cbs, x, Memory=None, use_ffd=False, Strong=None, Warning=None, MaxIteration=None,
          Precision=None, MinStepLength=None, WolfeConst1=None, WolfeConst2=None, Increment=None,
          trace=None, sum_mode=0):
    """orc_lbfgs with the reference's optional-argument semantics (None = absent)."""
    L = lib()
    f, fd, ffd = cbs
    x = np.ascontiguousarray(x, dtype=np.float64)
    n = C.c_int(x.size)
    L.orc_set_sum_mode(sum_mode)
    if trace is not None:
        trace.install()
    try:
        L.orc_lbfgs(f, fd, x.ctypes.data_as(C.c_void_p), C.byref(n), _opt(C.c_int, Memory),
                    ffd if use_ffd else None, _opt(C.c_int, None if Strong is None else int(Strong)),
                    _opt(C.c_int, None if Warning is None else int(Warning)),
                    _opt(C.c_int, MaxIteration), _opt(C.c_double, Precision),
                    _opt(C.c_double, MinStepLength), _opt(C.c_double, WolfeConst1),
                    _opt(C.c_double, WolfeConst2), _opt(C.c_double, Increment))
    finally:
        Trace.uninstall()
        L.orc_set_sum_mode(0)
    return x, stats()


def cg(cbs, x, Method=None, use_ffd=False, Strong=None, Warning=None, MaxIteration=None,
       Precision=None, MinStepLength=None, WolfeConst1=None, WolfeConst2=None, Increment=None,
       trace=None, sum_mode=0):
    L = lib()
    f, fd, ffd = cbs
    x = np.ascontiguousarray(x, dtype=np.float64)
    n = C.c_int(x.size)
    m = None if Method is None else Method.encode()
    L.orc_set_sum_mode(sum_mode)
    if trace is not None:
        trace.install()
    try:
        L.orc_conjugategradient(f, fd, x.ctypes.data_as(C.c_void_p), C.byref(n), m,
                                ffd if use_ffd else None,
                                _opt(C.c_int, None if Strong is None else int(Strong)),
                                _opt(C.c_int, None if Warning is None else int(Warning)),
                                _opt(C.c_int, MaxIteration), _opt(C.c_double, Precision),
                                _opt(C.c_double, MinStepLength), _opt(C.c_double, WolfeConst1),
                                _opt(C.c_double, WolfeConst2), _opt(C.c_double, Increment),
                                C.c_int(0 if m is None else len(m)))
    finally:
        Trace.uninstall()
        L.orc_set_sum_mode(0)
    return x, stats()


def sd(cbs, x, use_ffd=False, Strong=None, Warning=None, MaxIteration=None, Precision=None,
       MinStepLength=None, WolfeConst1=None, WolfeConst2=None, Increment=None, trace=None, sum_mode=0):
    """orc_steepestdescent (f90:55-188) with the reference's optional-argument semantics."""
    L = lib()
    f, fd, ffd = cbs
    x = np.ascontiguousarray(x, dtype=np.float64)
    n = C.c_int(x.size)
    L.orc_set_sum_mode(sum_mode)
    if trace is not None:
        trace.install()
    try:
        L.orc_steepestdescent(f, fd, x.ctypes.data_as(C.c_void_p), C.byref(n), ffd if use_ffd else None,
                              _opt(C.c_int, None if Strong is None else int(Strong)),
                              _opt(C.c_int, None if Warning is None else int(Warning)),
                              _opt(C.c_int, MaxIteration), _opt(C.c_double, Precision),
                              _opt(C.c_double, MinStepLength), _opt(C.c_double, WolfeConst1),
                              _opt(C.c_double, WolfeConst2), _opt(C.c_double, Increment))
    finally:
        Trace.uninstall()
        L.orc_set_sum_mode(0)
    return x, stats()


class ALStats(C.Structure):
    _fields_ = [("outer_iterations", C.c_long), ("inner_iterations", C.c_long), ("trials", C.c_long),
                ("status", C.c_int), ("cnorm2", C.c_double), ("miu", C.c_double)]


def sphere_constraint():
    L = lib()
    return C.cast(L.orc_con_sphere_c, C.c_void_p), C.cast(L.orc_con_sphere_cd, C.c_void_p)


def al(cbs, con, x, M=1, UnconstrainedSolver="LBFGS", lambda0=None, miu0=None, Memory=None, Method=None, use_ffd=False,
       Strong=None, Warning=None, MaxIteration=None, Precision=None, MinStepLength=None, WolfeConst1=None,
       WolfeConst2=None, Increment=None):
    """orc_augmentedlagrangian (f90:2005-2241; LBFGS / ConjugateGradient branches)."""
    L = lib()
    f, fd, ffd = cbs
    c, cd = con
    x = np.ascontiguousarray(x, dtype=np.float64)
    n, m = C.c_int(x.size), C.c_int(M)
    solver = UnconstrainedSolver.encode()
    meth = None if Method is None else Method.encode()
    lam = None if lambda0 is None else np.ascontiguousarray(lambda0, dtype=np.float64)
    L.orc_augmentedlagrangian(f, fd, c, cd, x.ctypes.data_as(C.c_void_p), C.byref(n), C.byref(m), solver,
                              None if lam is None else lam.ctypes.data_as(C.c_void_p), _opt(C.c_double, miu0), None, None,
                              None, _opt(C.c_int, Memory), meth, ffd if use_ffd else None,
                              _opt(C.c_int, None if Strong is None else int(Strong)),
                              _opt(C.c_int, None if Warning is None else int(Warning)),
                              _opt(C.c_int, MaxIteration), _opt(C.c_double, Precision),
                              _opt(C.c_double, MinStepLength), _opt(C.c_double, WolfeConst1),
                              _opt(C.c_double, WolfeConst2), _opt(C.c_double, Increment),
                              C.c_int(len(solver)), C.c_int(0 if meth is None else len(meth)))
    st = ALStats()
    L.orc_get_al_stats(C.byref(st))
    return x, st


def cg_basic(cbs, x, Method="DY", Strong=True, Warning=True, MaxIteration=1000, Precision=1e-15,
             MinStepLength=1e-15, WolfeConst1=1e-4, WolfeConst2=0.45, Increment=1.05, trace=None):
    L = lib()
    f, fd, _ = cbs
    x = np.ascontiguousarray(x, dtype=np.float64)
    n = C.c_int(x.size)
    m = Method.encode()
    if trace is not None:
        trace.install()
    try:
        L.orc_conjugategradient_basic(f, fd, x.ctypes.data_as(C.c_void_p), C.byref(n), m,
                                      C.byref(C.c_int(-1 if Strong else 0)),
                                      C.byref(C.c_int(-1 if Warning else 0)),
                                      C.byref(C.c_int(MaxIteration)), C.byref(C.c_double(Precision)),
                                      C.byref(C.c_double(MinStepLength)),
                                      C.byref(C.c_double(WolfeConst1)), C.byref(C.c_double(WolfeConst2)),
                                      C.byref(C.c_double(Increment)), C.c_int(len(m)))
    finally:
        Trace.uninstall()
    return x, stats()
