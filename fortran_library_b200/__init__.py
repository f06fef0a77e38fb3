"""fortran_library_b200 -- Python host side of libflgpu.so (ctypes, like the reference's own
FortranLibrary/*.py which binds libFL.so with ctypes.CDLL).

The package mirrors the reference's optimizer interface for the one hot path that is built:
``LBFGS`` (NonlinearOptimization.f90:398) and ``ConjugateGradient`` (f90:193), same argument
names, defaults and meaning; vectors live on the GPU.  All arithmetic happens in
hand-written CUDA kernels inside ``libflgpu.so``; there is no CPU or PyTorch fallback -- a
missing library or GPU raises immediately.
"""
import ctypes as C
import os

from . import _capi as capi
from ._capi import (CG_DY, CG_PR, LS_REFERENCE, LS_FAST, SPACE_HOST, SPACE_DEVICE, OBJ_QUARTIC, OBJ_ROSENBROCK, OBJ_DIAGQUAD,  # noqa: F401
                    OBJ_QUARTIC_SHIFTED,
                    START_QUARTIC_U, START_ROSEN_STD, START_ROSEN_PERT, START_ZERO, CONVERGED,
                    STEP_CONVERGED, MAX_ITERATION, INITIAL_CONVERGED, STOPPED_BY_OBSERVER, AL_LBFGS, AL_CG,
                    CON_SPHERE)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libflgpu.so")
_lib = None


class FlgpuError(RuntimeError):
    pass


def lib():
    """Load libflgpu.so (built in-tree by __graft_entry__.build() / csrc/Makefile)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("FLGPU_LIBRARY", LIB_PATH)   # development: an alternative build of the same library
    if not os.path.exists(path):
        raise FlgpuError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                         "g.build()'` (nvcc, sm_100a). There is no CPU fallback.")
    L = C.CDLL(path, mode=C.RTLD_GLOBAL)
    L.flgpu_version.restype = C.c_char_p
    L.flgpu_malloc.restype = C.c_void_p
    L.flgpu_malloc.argtypes = [C.c_size_t]
    L.flgpu_free.argtypes = [C.c_void_p]
    L.flgpu_memcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p]
    L.flgpu_fill_start.argtypes = [C.c_int, C.c_uint64, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p]
    L.flgpu_vec_dot.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    L.flgpu_vec_trial.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int64, C.c_void_p]
    L.flgpu_comm_create.restype = C.c_void_p
    L.flgpu_comm_create.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.flgpu_comm_destroy.argtypes = [C.c_void_p]
    L.flgpu_comm_unique_id.argtypes = [C.c_void_p]
    L.flgpu_current_stream.restype = C.c_void_p
    for name in ("flgpu_lbfgs", "flgpu_conjugate_gradient", "flgpu_steepest_descent"):
        getattr(L, name).argtypes = [C.POINTER(capi.Problem), C.POINTER(capi.Options), C.c_void_p, C.c_int64,
                                     C.c_int, C.POINTER(capi.Stats)]
    L.flgpu_kernel_times.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_double), C.POINTER(C.c_int64),
                                     C.POINTER(C.c_double), C.c_int]
    _lib = L
    return L


def device_count():
    return int(lib().flgpu_device_count())


def require_gpu():
    if device_count() <= 0:
        raise FlgpuError("no usable CUDA device: fortran_library_b200 is CUDA-only (sm_100a), no CPU fallback")


# ----------------------------------------------------------------------------- device vectors
class DeviceVector:
    """n float64 values in device memory owned by this object."""

    def __init__(self, n):
        require_gpu()
        self.n = int(n)
        self.ptr = lib().flgpu_malloc(max(self.n, 1) * 8)

    @classmethod
    def from_numpy(cls, a):
        import numpy as np
        a = np.ascontiguousarray(a, dtype=np.float64)
        v = cls(a.size)
        lib().flgpu_memcpy(v.ptr, a.ctypes.data, a.size * 8, SPACE_DEVICE, SPACE_HOST, None)
        return v

    @classmethod
    def start(cls, kind, n, seed=0, offset=0, n_global=None):
        v = cls(n)
        lib().flgpu_fill_start(kind, seed, v.ptr, offset, n, n if n_global is None else n_global, None)
        lib().flgpu_memcpy(v.ptr, v.ptr, 0, SPACE_DEVICE, SPACE_DEVICE, None)  # sync the null stream
        return v

    def numpy(self):
        import numpy as np
        out = np.empty(self.n, dtype=np.float64)
        lib().flgpu_memcpy(out.ctypes.data, self.ptr, self.n * 8, SPACE_HOST, SPACE_DEVICE, None)
        return out

    def free(self):
        if self.ptr:
            lib().flgpu_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def copy_to_numpy(dev_ptr, n, stream=None):
    import numpy as np
    out = np.empty(int(n), dtype=np.float64)
    lib().flgpu_memcpy(out.ctypes.data, dev_ptr, int(n) * 8, SPACE_HOST, SPACE_DEVICE, stream)
    return out


# ----------------------------------------------------------------------------- problems
def builtin_problem(kind):
    """The CUDA objective kernels of the benchmark configs (quartic / Rosenbrock / diagonal quadratic)."""
    p = capi.Problem()
    if lib().flgpu_builtin_problem(kind, C.byref(p)) != 0:
        raise ValueError(f"unknown built-in objective {kind}")
    return p


def make_problem(f, fd, f_fd=None, user=None, fused=None):
    """Wrap Python callables (ctx, f_dev, x_dev, n) / (ctx, g_dev, x_dev, n) / (ctx, f_dev, g_dev, x_dev, n)
    [/ fused: (ctx, flags, f_dev, gp_dev, x_out, g_out, x0_dev, p_dev, a, n)] as device callbacks.
    The returned object keeps the ctypes thunks alive."""
    p = capi.Problem()
    keep = [capi.F_FN(f), capi.FD_FN(fd), capi.F_FD_FN(f_fd) if f_fd is not None else None,
            capi.FUSED_FN(fused) if fused is not None else None]
    p.f = C.cast(keep[0], C.c_void_p)
    p.fd = C.cast(keep[1], C.c_void_p)
    p.f_fd = C.cast(keep[2], C.c_void_p) if keep[2] is not None else None
    p.fused = C.cast(keep[3], C.c_void_p) if keep[3] is not None else None
    p.user = user
    p._keep = keep
    return p


class Observer:
    """Per-iteration observer: records (iteration, step, f, phi'(0), trials) and optionally copies
    p / x / f' of the first `max_vec_iters` iterations to the host."""

    def __init__(self, keep_vectors=False, max_vec_iters=10**9, stop_after=None, on_iteration=None):
        self.rows, self.p, self.x, self.g = [], [], [], []
        self.keep, self.max_vec_iters, self.stop_after = keep_vectors, max_vec_iters, stop_after
        self.on_iteration = on_iteration
        self.cb = capi.OBSERVER_FN(self._on)

    def _on(self, user, info):
        i = info.contents
        self.rows.append((i.iteration, i.step, i.f, i.phid0, i.trials))
        if self.keep and i.iteration < self.max_vec_iters:
            self.p.append(copy_to_numpy(i.p_dev, i.n_local, i.stream))
            self.x.append(copy_to_numpy(i.x_dev, i.n_local, i.stream))
            self.g.append(copy_to_numpy(i.g_dev, i.n_local, i.stream))
        stop = 0
        if self.on_iteration is not None:
            stop = int(bool(self.on_iteration(i)))
        if self.stop_after is not None and i.iteration + 1 >= self.stop_after:
            stop = 1
        return stop


def set_line_search(policy):
    """Line-search policy of later Fortran-ABI calls (`__nonlinearoptimization_MOD_*`) on this thread: "reference"
    (default), "fast" (FLGPU_LS_FAST, not a reference routine) or None = let FLGPU_LINE_SEARCH decide."""
    code = {None: -1, "reference": LS_REFERENCE, "fast": LS_FAST}.get(policy, policy)
    lib().flgpu_set_line_search(int(code))


def default_options(for_cg=False):
    o = capi.Options()
    lib().flgpu_options_default(C.byref(o), int(for_cg))
    return o


def _run(fn, for_cg, problem, x, n, x_space, observer, stream, comm, offset, n_global, time_kernels, kw):
    require_gpu()
    o = default_options(for_cg)
    o.no_fused = int(not kw.pop("fused", True))
    ds = kw.pop("device_search", None)                    # None = auto (by size), True / False = force
    o.device_search = 2 if ds is None else int(bool(ds))
    capi.apply_options(o, **kw)
    o.stream = stream
    o.comm = comm
    o.offset, o.n_global = offset, n_global
    o.time_kernels = int(time_kernels)
    if observer is not None:
        o.observer = C.cast(observer.cb, C.c_void_p)
    st = capi.Stats()
    rc = fn(C.byref(problem), C.byref(o), x, n, x_space, C.byref(st))
    if rc == capi.ERR_MEMORY_LIMIT:
        raise FlgpuError(f"LBFGS Memory = {o.memory} exceeds FLGPU_MAX_MEMORY = {capi.MAX_MEMORY}: nothing was done")
    if rc != 0:
        raise FlgpuError(f"libflgpu returned error {rc}")
    return st


def _resolve_x(x):
    """Accept DeviceVector, numpy array (host, updated in place) or (ptr, n, space)."""
    if isinstance(x, DeviceVector):
        return x.ptr, x.n, SPACE_DEVICE
    if isinstance(x, tuple):
        return x
    import numpy as np
    if isinstance(x, np.ndarray):
        if x.dtype != np.float64 or not x.flags.c_contiguous:
            raise TypeError("x must be a C-contiguous float64 array (it is updated in place)")
        return x.ctypes.data, x.size, SPACE_HOST
    if hasattr(x, "data_ptr"):  # torch tensor
        return x.data_ptr(), x.numel(), (SPACE_DEVICE if x.is_cuda else SPACE_HOST)
    raise TypeError(f"unsupported x: {type(x)}")


def LBFGS(problem, x, Memory=None, Strong=None, Warning=None, MaxIteration=None, Precision=None,
          MinStepLength=None, WolfeConst1=None, WolfeConst2=None, Increment=None, observer=None, stream=None,
          comm=None, offset=0, n_global=0, time_kernels=False, fused=True, device_search=None,
          line_search=None):
    """Limited-memory BFGS (reference: LBFGS, NonlinearOptimization.f90:398-625).  x is updated in
    place with the minimiser; returns the run statistics.  `problem.f_fd` present selects the
    _fdwithf line searcher exactly as the reference's optional f_fd does.  fused=False ignores
    `problem.fused` (trial points are then materialised and the plain callbacks called).
    line_search="fast" (default "reference") selects FLGPU_LS_FAST, an accept-at-first-Wolfe-point searcher that is
    NOT a reference routine (include/flgpu.h); available on every optimizer here."""
    ptr, n, space = _resolve_x(x)
    return _run(lib().flgpu_lbfgs, False, problem, ptr, n, space, observer, stream, comm, offset, n_global,
                time_kernels, dict(Memory=Memory, Strong=Strong, Warning=Warning, MaxIteration=MaxIteration,
                                   Precision=Precision, MinStepLength=MinStepLength, WolfeConst1=WolfeConst1,
                                   WolfeConst2=WolfeConst2, Increment=Increment, fused=fused, device_search=device_search,
                     line_search=line_search))


def ConjugateGradient(problem, x, Method=None, Strong=None, Warning=None, MaxIteration=None, Precision=None,
                      MinStepLength=None, WolfeConst1=None, WolfeConst2=None, Increment=None, observer=None,
                      stream=None, comm=None, offset=0, n_global=0, time_kernels=False, no_clamp=False, fused=True, device_search=None,
          line_search=None):
    """Nonlinear conjugate gradient, Method 'DY' (default) or 'PR' (reference: ConjugateGradient,
    f90:193-394; no_clamp=True gives ConjugateGradient_basic, f90:2249-2346)."""
    if Method is not None and Method not in ("DY", "PR", CG_DY, CG_PR):
        raise SystemExit("Program abort: unsupported conjugate gradient method " + str(Method))  # f90:345
    ptr, n, space = _resolve_x(x)
    return _run(lib().flgpu_conjugate_gradient, True, problem, ptr, n, space, observer, stream, comm, offset,
                n_global, time_kernels,
                dict(Method=Method, Strong=Strong, Warning=Warning, MaxIteration=MaxIteration,
                     Precision=Precision, MinStepLength=MinStepLength, WolfeConst1=WolfeConst1,
                     WolfeConst2=WolfeConst2, Increment=Increment, no_clamp=int(no_clamp), fused=fused, device_search=device_search,
                     line_search=line_search))


def SteepestDescent(problem, x, Strong=None, Warning=None, MaxIteration=None, Precision=None, MinStepLength=None,
                    WolfeConst1=None, WolfeConst2=None, Increment=None, observer=None, stream=None, comm=None,
                    offset=0, n_global=0, time_kernels=False, fused=True, device_search=None,
          line_search=None):
    """Steepest descent (reference: SteepestDescent, f90:55-188): p = -f'(x) through the same line searchers."""
    ptr, n, space = _resolve_x(x)
    return _run(lib().flgpu_steepest_descent, False, problem, ptr, n, space, observer, stream, comm, offset,
                n_global, time_kernels,
                dict(Strong=Strong, Warning=Warning, MaxIteration=MaxIteration, Precision=Precision,
                     MinStepLength=MinStepLength, WolfeConst1=WolfeConst1, WolfeConst2=WolfeConst2,
                     Increment=Increment, fused=fused, device_search=device_search,
                     line_search=line_search))


def builtin_constraints(kind=CON_SPHERE):
    """The reference's test constraint (test.f90:692-705): unit sphere, c = x.x - 1, as CUDA kernels."""
    c = capi.Constraints()
    if lib().flgpu_builtin_constraints(kind, C.byref(c)) != 0:
        raise ValueError(f"unknown built-in constraint {kind}")
    return c


def AugmentedLagrangian(problem, constraints, x, UnconstrainedSolver="LBFGS", lambda0=None, miu0=None, Memory=None,
                        Method=None, Strong=None, Warning=None, MaxIteration=None, Precision=None, MinStepLength=None,
                        WolfeConst1=None, WolfeConst2=None, Increment=None, stream=None, comm=None, offset=0, n_global=0,
                        line_search=None, fused=True):
    """Equality-constrained minimisation (reference: AugmentedLagrangian, f90:2005-2241) with the hot path as inner
    solver: UnconstrainedSolver 'LBFGS' or 'ConjugateGradient' (the dense-Hessian solvers are outside the GPU path).
    line_search="fast" runs every inner solve with FLGPU_LS_FAST (not a reference routine).  fused=False ignores
    `problem.fused` / `constraints.fused` (every trial point of the inner solves is then materialised)."""
    import numpy as np
    if UnconstrainedSolver not in ("LBFGS", "ConjugateGradient"):
        raise SystemExit("Program abort: unsupported unconstrained solver " + str(UnconstrainedSolver))
    require_gpu()
    L = lib()
    L.flgpu_augmented_lagrangian.argtypes = [C.POINTER(capi.Problem), C.POINTER(capi.Constraints),
                                             C.POINTER(capi.ALOptions), C.c_void_p, C.c_int64, C.c_int,
                                             C.POINTER(capi.ALStats)]
    o = capi.ALOptions()
    L.flgpu_al_options_default(C.byref(o), capi.AL_CG if UnconstrainedSolver == "ConjugateGradient" else capi.AL_LBFGS)
    capi.apply_options(o.inner, Memory=Memory, Method=Method, Strong=Strong, Warning=Warning, MaxIteration=MaxIteration,
                       Precision=Precision, MinStepLength=MinStepLength, WolfeConst1=WolfeConst1,
                       WolfeConst2=WolfeConst2, Increment=Increment, line_search=line_search)
    o.inner.stream, o.inner.comm, o.inner.offset, o.inner.n_global = stream, comm, offset, n_global
    o.inner.no_fused = int(not fused)
    lam = None
    if lambda0 is not None:
        lam = np.ascontiguousarray(lambda0, dtype=np.float64)
        o.lambda0 = lam.ctypes.data
    if miu0 is not None:
        o.miu0 = miu0
    ptr, n, space = _resolve_x(x)
    st = capi.ALStats()
    L.flgpu_augmented_lagrangian(C.byref(problem), C.byref(constraints), C.byref(o), ptr, n, space, C.byref(st))
    return st


class History:
    """The two-loop recursion as a standalone operator (flgpu_history_*; reference LBFGS::Before /
    After, f90:586-624).  push()/direction() take host arrays or DeviceVectors; direction() returns
    host copies of p and of the first trial point x1 + p together with f'.p and p.p."""

    def __init__(self, n, memory, stream=None, comm=None):
        require_gpu()
        L = lib()
        L.flgpu_history_create.restype = C.c_void_p
        L.flgpu_history_create.argtypes = [C.c_int64, C.c_int, C.c_void_p, C.c_void_p]
        L.flgpu_history_push.argtypes = [C.c_void_p] * 5
        L.flgpu_history_direction.argtypes = [C.c_void_p] * 7
        L.flgpu_history_destroy.argtypes = [C.c_void_p]
        self.n = int(n)
        self.h = L.flgpu_history_create(self.n, int(memory), stream, comm)
        self._p, self._xt = DeviceVector(self.n), DeviceVector(self.n)

    @staticmethod
    def _dev(v):
        return v if isinstance(v, DeviceVector) else DeviceVector.from_numpy(v)

    def push(self, x1, x0, g1, g0):
        vs = [self._dev(v) for v in (x1, x0, g1, g0)]
        lib().flgpu_history_push(self.h, *[v.ptr for v in vs])
        lib().flgpu_memcpy(None, None, 0, SPACE_DEVICE, SPACE_DEVICE, None)
        return vs

    def direction(self, g1, x1):
        g1, x1 = self._dev(g1), self._dev(x1)
        gp, pp = C.c_double(), C.c_double()
        lib().flgpu_history_direction(self.h, g1.ptr, x1.ptr, self._p.ptr, self._xt.ptr, C.addressof(gp),
                                      C.addressof(pp))
        return self._p.numpy(), self._xt.numpy(), gp.value, pp.value

    def close(self):
        if self.h:
            lib().flgpu_history_destroy(self.h)
            self.h = None


def set_workspace_cache(on):
    """Keep the device work space of finished calls for the next call (flgpu_set_workspace_cache)."""
    lib().flgpu_set_workspace_cache(int(bool(on)))


def release_workspace():
    lib().flgpu_release_workspace()


def kernel_times():
    """Per-kernel CUDA-event totals of the last call made with time_kernels=True."""
    cap = 64
    names = (C.c_char_p * cap)()
    ms = (C.c_double * cap)()
    launches = (C.c_int64 * cap)()
    nbytes = (C.c_double * cap)()
    k = lib().flgpu_kernel_times(names, ms, launches, nbytes, cap)
    return {names[i].decode(): {"ms": ms[i], "launches": launches[i], "bytes": nbytes[i]} for i in range(k)}


def comm_create(rank, nranks, broadcast_bytes):
    """Create the row-shard communicator.  broadcast_bytes(buf: bytes|None) -> bytes distributes rank 0's
    128-byte id to every rank (e.g. over torch.distributed)."""
    ident = (C.c_char * 128)()
    if rank == 0:
        lib().flgpu_comm_unique_id(ident)
    data = broadcast_bytes(bytes(ident) if rank == 0 else None)
    buf = (C.c_char * 128).from_buffer_copy(data)
    return lib().flgpu_comm_create(buf, rank, nranks)
