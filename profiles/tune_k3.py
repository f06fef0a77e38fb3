"""profiles/tune_k3.py [log2n] [mem] -- K3 per-launch time for one FLGPU_K3 setting (read from the environment):
LBFGS on Rosenbrock with per-kernel CUDA events, 12 main-loop iterations.  Run once per setting (the mode is read once
per process):  for s in regs 8,3 7,3; do FLGPU_K3=$s python profiles/tune_k3.py; done"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fortran_library_b200 as fl  # noqa: E402

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
mem = int(sys.argv[2]) if len(sys.argv) > 2 else 10
n = 1 << log2n
x = fl.DeviceVector.start(fl.START_ROSEN_PERT, n, seed=7)


def on_iter(i):
    if i.iteration == mem + 2:
        fl.lib().flgpu_reset_kernel_times()
    return False


st = fl.LBFGS(fl.builtin_problem(fl.OBJ_ROSENBROCK), x, Memory=mem, Warning=False, MaxIteration=15, time_kernels=True,
              observer=fl.Observer(on_iteration=on_iter))
kt = fl.kernel_times()
out = [os.environ.get("FLGPU_K3", "default")]
for name in ("k3_direction", "k1_update_dots", "k1_update_dots_fused", "callback:fused_probe", "callback:fused_store"):
    if name in kt and kt[name]["launches"]:
        v = kt[name]
        out.append(f"{name}: {v['ms'] / v['launches']:.3f} ms {v['bytes'] / v['ms'] / 1e6:.0f} GB/s x{v['launches']}")
print(" | ".join(out), flush=True)
