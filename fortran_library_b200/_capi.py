"""ctypes mirror of include/flgpu.h (structs, callback types, constants)."""
import ctypes as C

# enums
CG_DY, CG_PR = 0, 1
LS_REFERENCE, LS_FAST = 0, 1
SPACE_HOST, SPACE_DEVICE = 0, 1
CONVERGED, STEP_CONVERGED, MAX_ITERATION, INITIAL_CONVERGED, STOPPED_BY_OBSERVER, INVALID_ARGUMENT = 0, 1, 2, 3, 4, 5
ERR_MEMORY_LIMIT, MAX_MEMORY = 2, 64
OBJ_QUARTIC, OBJ_ROSENBROCK, OBJ_DIAGQUAD, OBJ_QUARTIC_SHIFTED = 0, 1, 2, 3
START_QUARTIC_U, START_ROSEN_STD, START_ROSEN_PERT, START_ZERO = 0, 1, 2, 3


class EvalCtx(C.Structure):
    _fields_ = [("user", C.c_void_p), ("stream", C.c_void_p), ("offset", C.c_int64),
                ("n_global", C.c_int64), ("rank", C.c_int), ("nranks", C.c_int), ("device", C.c_int)]


F_FN = C.CFUNCTYPE(None, C.POINTER(EvalCtx), C.c_void_p, C.c_void_p, C.c_int64)
FD_FN = C.CFUNCTYPE(None, C.POINTER(EvalCtx), C.c_void_p, C.c_void_p, C.c_int64)
F_FD_FN = C.CFUNCTYPE(None, C.POINTER(EvalCtx), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64)

# reference callback ABI (f90:33-38)
REF_F_FN = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.c_void_p, C.POINTER(C.c_int))
REF_FD_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.POINTER(C.c_int))
REF_F_FD_FN = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_double), C.c_void_p, C.c_void_p, C.POINTER(C.c_int))


# fused line-search evaluation (flgpu_fused_fn)
WANT_F, WANT_GP, WRITE_X, WRITE_G = 1, 2, 4, 8
FUSED_FN = C.CFUNCTYPE(None, C.POINTER(EvalCtx), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                       C.c_void_p, C.c_void_p, C.c_double, C.c_int64)


class Problem(C.Structure):
    _fields_ = [("f", C.c_void_p), ("fd", C.c_void_p), ("f_fd", C.c_void_p), ("user", C.c_void_p),
                ("fused", C.c_void_p), ("search", C.c_void_p), ("search_caps", C.c_int),
                ("update", C.c_void_p), ("direction", C.c_void_p), ("fused_multi", C.c_void_p)]


class IterInfo(C.Structure):
    _fields_ = [("iteration", C.c_int64), ("n_local", C.c_int64), ("step", C.c_double),
                ("f", C.c_double), ("phid0", C.c_double), ("trials", C.c_int64),
                ("p_dev", C.c_void_p), ("x_dev", C.c_void_p), ("g_dev", C.c_void_p),
                ("stream", C.c_void_p), ("gpu_launches", C.c_int64), ("callbacks", C.c_int64),
                ("total_trials", C.c_int64)]


OBSERVER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(IterInfo))


class Options(C.Structure):
    _fields_ = [("memory", C.c_int), ("method", C.c_int), ("strong", C.c_int), ("warning", C.c_int),
                ("max_iteration", C.c_int), ("precision", C.c_double), ("min_step_length", C.c_double),
                ("wolfe_c1", C.c_double), ("wolfe_c2", C.c_double), ("increment", C.c_double),
                ("no_clamp", C.c_int), ("stream", C.c_void_p), ("comm", C.c_void_p),
                ("offset", C.c_int64), ("n_global", C.c_int64), ("observer", C.c_void_p),
                ("observer_user", C.c_void_p), ("time_kernels", C.c_int), ("no_fused", C.c_int),
                ("device_search", C.c_int), ("line_search", C.c_int)]


class Stats(C.Structure):
    _fields_ = [("iterations", C.c_int64), ("status", C.c_int), ("n_f", C.c_int64), ("n_fd", C.c_int64),
                ("n_f_fd", C.c_int64), ("n_trials", C.c_int64), ("n_f_only_trials", C.c_int64),
                ("n_linesearch", C.c_int64), ("gpu_launches", C.c_int64), ("host_syncs", C.c_int64),
                ("f", C.c_double), ("gnorm2", C.c_double), ("n_batched_passes", C.c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


# AugmentedLagrangian (flgpu_constraints, flgpu_al_options, flgpu_al_stats)
AL_LBFGS, AL_CG = 0, 1
CON_SPHERE = 0
C_FN = C.CFUNCTYPE(None, C.POINTER(EvalCtx), C.c_void_p, C.c_void_p, C.c_int, C.c_int64)
CD_FN = C.CFUNCTYPE(None, C.POINTER(EvalCtx), C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int64)
REF_C_FN = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int))
REF_CD_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int))


class Constraints(C.Structure):
    _fields_ = [("c", C.c_void_p), ("cd", C.c_void_p), ("m", C.c_int), ("fused", C.c_void_p)]


class ALOptions(C.Structure):
    _fields_ = [("solver", C.c_int), ("lambda0", C.c_void_p), ("miu0", C.c_double), ("inner", Options)]


class ALStats(C.Structure):
    _fields_ = [("outer_iterations", C.c_int64), ("inner_iterations", C.c_int64), ("trials", C.c_int64),
                ("status", C.c_int), ("cnorm2", C.c_double), ("miu", C.c_double), ("f", C.c_double),
                ("gpu_launches", C.c_int64)]


def apply_options(o, **kw):
    """Set Options fields from reference-style keyword names (None = leave default)."""
    names = {"Memory": "memory", "Method": "method", "Strong": "strong", "Warning": "warning",
             "MaxIteration": "max_iteration", "Precision": "precision", "MinStepLength": "min_step_length",
             "WolfeConst1": "wolfe_c1", "WolfeConst2": "wolfe_c2", "Increment": "increment"}
    for k, v in kw.items():
        if v is None:
            continue
        field = names.get(k, k)
        if field == "method" and isinstance(v, str):
            v = {"DY": CG_DY, "PR": CG_PR}[v]
        if field == "line_search" and isinstance(v, str):
            v = {"reference": LS_REFERENCE, "fast": LS_FAST}[v]
        if field in ("strong", "warning"):
            v = int(bool(v))
        setattr(o, field, v)
    return o
