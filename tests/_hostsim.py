"""Loader for tests/hostsim/libflgpu_hostsim.so (TEST INFRASTRUCTURE).

The host simulator runs the product's driver.cpp over host memory; it exists so the host
control flow can be tested without a GPU.  It is never imported by the package.
"""
import ctypes as C
import importlib.util
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HS_DIR = os.path.join(ROOT, "tests", "hostsim")
LIB_PATH = os.path.join(HS_DIR, "libflgpu_hostsim.so")

_spec = importlib.util.spec_from_file_location("_flgpu_capi", os.path.join(ROOT, "fortran_library_b200", "_capi.py"))
capi = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(capi)

ALLGATHER_FN = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int)

_lib = None


def lib():
    global _lib
    if _lib is None:
        subprocess.run(["make", "-C", HS_DIR, "libflgpu_hostsim.so"], check=True, stdout=subprocess.DEVNULL)
        _lib = C.CDLL(LIB_PATH)
    return _lib


class Observer:
    """Collects per-iteration rows and (optionally) copies of p, x, g (host pointers here)."""

    def __init__(self, keep_vectors=True, max_vec_iters=10**9, stop_after=None):
        self.rows, self.p, self.x, self.g = [], [], [], []
        self.keep, self.max_vec_iters, self.stop_after = keep_vectors, max_vec_iters, stop_after
        self.cb = capi.OBSERVER_FN(self._on)

    def _on(self, user, info):
        i = info.contents
        self.rows.append((i.iteration, i.step, i.f, i.phid0, i.trials))
        if self.keep and i.iteration < self.max_vec_iters:
            n = i.n_local
            for lst, ptr in ((self.p, i.p_dev), (self.x, i.x_dev), (self.g, i.g_dev)):
                lst.append(np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), (n,)).copy())
        if self.stop_after is not None and i.iteration + 1 >= self.stop_after:
            return 1
        return 0


def _run(fn_name, for_cg, kind, x, observer=None, use_ffd=True, offset=0, n_global=0, **kw):
    L = lib()
    prob = capi.Problem()
    L.flgpu_hostsim_builtin_problem(kind, C.byref(prob))
    if not use_ffd:
        prob.f_fd = None
    o = capi.Options()
    L.flgpu_hostsim_options_default(C.byref(o), int(for_cg))
    o.no_fused = int(not kw.pop("fused", True))
    ds = kw.pop("device_search", False)          # the simulator's eager "device-resident" search: off unless asked
    o.device_search = 2 if ds is None else int(bool(ds))
    capi.apply_options(o, **kw)
    o.offset, o.n_global = offset, n_global
    if observer is not None:
        o.observer = C.cast(observer.cb, C.c_void_p)
    x = np.ascontiguousarray(x, dtype=np.float64).copy()
    st = capi.Stats()
    getattr(L, fn_name)(C.byref(prob), C.byref(o), x.ctypes.data_as(C.c_void_p), C.c_int64(x.size), C.byref(st))
    return x, st


def lbfgs(kind, x, **kw):
    return _run("flgpu_hostsim_lbfgs", False, kind, x, **kw)


def cg(kind, x, **kw):
    return _run("flgpu_hostsim_cg", True, kind, x, **kw)


def sd(kind, x, **kw):
    return _run("flgpu_hostsim_sd", False, kind, x, **kw)


class History:
    """flgpu_hostsim_history_*: the two-loop recursion as an operator (host pointers)."""

    def __init__(self, n, memory):
        L = lib()
        L.flgpu_hostsim_history_create.restype = C.c_void_p
        L.flgpu_hostsim_history_create.argtypes = [C.c_int64, C.c_int]
        L.flgpu_hostsim_history_push.argtypes = [C.c_void_p] * 5
        L.flgpu_hostsim_history_direction.argtypes = [C.c_void_p] * 7
        L.flgpu_hostsim_history_destroy.argtypes = [C.c_void_p]
        self.n, self.h = n, L.flgpu_hostsim_history_create(n, memory)

    def push(self, x1, x0, g1, g0):
        self._keep = [np.ascontiguousarray(v, dtype=np.float64) for v in (x1, x0, g1, g0)]
        lib().flgpu_hostsim_history_push(self.h, *[v.ctypes.data for v in self._keep])

    def direction(self, g1, x1):
        g1 = np.ascontiguousarray(g1, dtype=np.float64)
        x1 = np.ascontiguousarray(x1, dtype=np.float64)
        p, xt = np.empty(self.n), np.empty(self.n)
        gp, pp = C.c_double(), C.c_double()
        lib().flgpu_hostsim_history_direction(self.h, g1.ctypes.data, x1.ctypes.data, p.ctypes.data, xt.ctypes.data,
                                              C.addressof(gp), C.addressof(pp))
        return p, xt, gp.value, pp.value

    def close(self):
        lib().flgpu_hostsim_history_destroy(self.h)


def set_comm(fn, rank, nranks):
    lib().flgpu_hostsim_set_comm(fn, None, rank, nranks)
