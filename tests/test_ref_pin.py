"""Oracle #0: oracle.c against the reference's OWN Fortran (oracle/_ref/libref.so, built by oracle/make_ref.sh from the
sources under /root/reference wherever `gfortran` exists).  SKIPPED in this image -- there is no Fortran compiler
(SURVEY.md F1), which is why DESIGN.md says PARITY UNPINNED -- and written so that the day the library can be built,
the pin is one `python -m pytest tests/test_ref_pin.py` away.  Same calling convention as the oracle (gfortran's: every
argument by reference, absent OPTIONAL = NULL, 4-byte LOGICAL, trailing hidden CHARACTER length)."""
import ctypes as C
import os

import numpy as np
import pytest

import _cases
import _golden as G
import _oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "libref.so")
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="oracle #0 needs gfortran + /root/reference "
                                                                "(oracle/make_ref.sh); neither exists in this image")


def _ref():
    return C.CDLL(REF)


def _opt(ctype, v):
    return None if v is None else C.byref(ctype(v))


def _logical(v):
    return None if v is None else C.byref(C.c_int32(1 if v else 0))


def ref_lbfgs(cbs, x, Memory=None, use_ffd=False, Strong=None, MaxIteration=None, WolfeConst1=None, WolfeConst2=None,
              Increment=None):
    f, fd, ffd = cbs
    x = np.ascontiguousarray(x, dtype=np.float64).copy()
    getattr(_ref(), "__nonlinearoptimization_MOD_lbfgs")(
        f, fd, x.ctypes.data_as(C.c_void_p), C.byref(C.c_int(x.size)), _opt(C.c_int, Memory), ffd if use_ffd else None,
        _logical(Strong), _logical(False), _opt(C.c_int, MaxIteration), None, None, _opt(C.c_double, WolfeConst1),
        _opt(C.c_double, WolfeConst2), _opt(C.c_double, Increment))
    return x


def ref_cg(cbs, x, Method=None, use_ffd=False, Strong=None, MaxIteration=None, WolfeConst1=None, WolfeConst2=None,
           Increment=None):
    f, fd, ffd = cbs
    x = np.ascontiguousarray(x, dtype=np.float64).copy()
    m = None if Method is None else Method.encode()
    getattr(_ref(), "__nonlinearoptimization_MOD_conjugategradient")(
        f, fd, x.ctypes.data_as(C.c_void_p), C.byref(C.c_int(x.size)), m, ffd if use_ffd else None, _logical(Strong),
        _logical(False), _opt(C.c_int, MaxIteration), None, None, _opt(C.c_double, WolfeConst1),
        _opt(C.c_double, WolfeConst2), _opt(C.c_double, Increment), C.c_size_t(0 if m is None else len(m)))
    return x


def ref_sd(cbs, x, use_ffd=False, Strong=None, MaxIteration=None, WolfeConst1=None, WolfeConst2=None, Increment=None):
    f, fd, ffd = cbs
    x = np.ascontiguousarray(x, dtype=np.float64).copy()
    getattr(_ref(), "__nonlinearoptimization_MOD_steepestdescent")(
        f, fd, x.ctypes.data_as(C.c_void_p), C.byref(C.c_int(x.size)), ffd if use_ffd else None, _logical(Strong),
        _logical(False), _opt(C.c_int, MaxIteration), None, None, _opt(C.c_double, WolfeConst1),
        _opt(C.c_double, WolfeConst2), _opt(C.c_double, Increment))
    return x


@pytest.mark.parametrize("case", sorted(_cases.TORTURE_1D))
@pytest.mark.parametrize("method", ["DY", "PR"])
def test_reference_fortran_evaluates_the_oracles_trial_points(case, method):
    """dim = 1 through Python callbacks: the real Fortran and oracle.c must request the SAME sequence of evaluation
    points, bit for bit, through every branch of the Strong-Wolfe searchers (incl. the fall-through of f90:1511-1512)."""
    x0, (f, g) = _cases.TORTURE_1D[case]
    for use in (False, True):
        if use and case in _cases.TORTURE_NO_FFD:
            continue
        runs = []
        for which in ("oracle", "reference"):
            fu = _cases.Fuse(f, g)
            cf, cfd, cffd = _cases.make_ref_callbacks(fu.f, fu.g, fu.fg)
            keep = (O.F_T(cf), O.FD_T(cfd), O.FFD_T(cffd))
            cbs = tuple(C.cast(k, C.c_void_p) for k in keep)
            if which == "oracle":
                x, _ = O.cg(cbs, np.array([x0]), Method=method, use_ffd=use, Warning=False, MaxIteration=30)
            else:
                x = ref_cg(cbs, np.array([x0]), Method=method, use_ffd=use, MaxIteration=30)
            runs.append((fu.xs, x))
        assert runs[0][0] == runs[1][0], "oracle.c and the reference Fortran evaluated different points"
        assert np.array_equal(runs[0][1], runs[1][1], equal_nan=True)


@pytest.mark.parametrize("name", G.names())
def test_reference_fortran_reproduces_the_golden_vectors(name):
    """The committed known-answer vectors (outputs of oracle.c) against the real Fortran on the same starts, with the
    oracle's C objectives as the callbacks: identical minimisers."""
    d = G.load(name)
    kw = dict(d["options"])
    use = kw.pop("use_ffd", False)
    kind = _cases.OBJECTIVES[d["objective"]][0]
    cbs = O.builtin_callbacks(kind, 0, d["n"])
    run = {"lbfgs": ref_lbfgs, "cg": ref_cg, "sd": ref_sd}[d["algorithm"]]
    x = run(cbs, d["x0"].copy(), use_ffd=use, **kw)
    assert np.array_equal(x, d["x_final"]), f"{name}: the reference's result differs from the stored oracle output"
