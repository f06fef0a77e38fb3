"""Developer probe: one size, device-resident search forced on (for ncu)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fortran_library_b200 as fl  # noqa: E402
n = 1 << int(sys.argv[1])
x = fl.DeviceVector.start(fl.START_ROSEN_PERT, n, seed=7)
st = fl.LBFGS(fl.builtin_problem(fl.OBJ_ROSENBROCK), x, Memory=10, Warning=False, MaxIteration=20, device_search=True, time_kernels=True)
print({k: (round(v["ms"] / max(v["launches"], 1), 4), v["launches"]) for k, v in fl.kernel_times().items()}, st.n_trials)
