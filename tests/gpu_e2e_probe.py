"""Developer probe (not a test): where the end-to-end time of one Fortran-ABI LBFGS call goes (FLGPU_TRACE_PHASES)."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["FLGPU_TRACE_PHASES"] = "1"
import numpy as np  # noqa: E402
import torch  # noqa: E402
import fortran_library_b200 as fl  # noqa: E402

n = 1 << 28
L = fl.lib()
f, fd, ffd = fl.capi.REF_F_FN(), fl.capi.REF_FD_FN(), fl.capi.REF_F_FD_FN()
L.flgpu_builtin_ref_callbacks(fl.OBJ_ROSENBROCK, C.byref(f), C.byref(fd), C.byref(ffd))
xh = torch.empty(n, dtype=torch.float64).pin_memory()
for rep in range(3):
    x0 = fl.DeviceVector.start(fl.START_ROSEN_PERT, n, seed=7)
    L.flgpu_memcpy(xh.data_ptr(), x0.ptr, n * 8, fl.SPACE_HOST, fl.SPACE_DEVICE, None)
    x0.free()
    for maxit in (0, 30):
        t = time.time()
        L.__getattr__("__nonlinearoptimization_MOD_lbfgs")(
            f, fd, C.c_void_p(xh.data_ptr()), C.byref(C.c_int(n)), C.byref(C.c_int(10)), ffd, None,
            C.byref(C.c_int32(0)), C.byref(C.c_int(maxit)), None, None, None, None, None)
        st = fl.capi.Stats()
        L.flgpu_last_stats(C.byref(st))
        print(f"rep {rep} MaxIteration={maxit}: {1e3 * (time.time() - t):.1f} ms wall, {st.iterations} iterations, "
              f"{st.n_trials} trials", flush=True)
