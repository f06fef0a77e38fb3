"""Developer probe: per-kernel rates of LBFGS m=10 at n=2^28 with plain callbacks (K3 then also writes the trial point)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fortran_library_b200 as fl  # noqa: E402
n, mem = 1 << 28, 10
x = fl.DeviceVector.start(fl.START_ROSEN_PERT, n, seed=7)


def on_iter(i):
    if i.iteration == mem + 2:
        fl.lib().flgpu_reset_kernel_times()
    return False


st = fl.LBFGS(fl.builtin_problem(fl.OBJ_ROSENBROCK), x, Memory=mem, Warning=False, MaxIteration=12, time_kernels=True,
              fused=False, observer=fl.Observer(on_iteration=on_iter))
for name, v in fl.kernel_times().items():
    if v["launches"]:
        print(f"{name}: {v['ms'] / v['launches']:.3f} ms {v['bytes'] / v['ms'] / 1e6:.0f} GB/s x{v['launches']}")
