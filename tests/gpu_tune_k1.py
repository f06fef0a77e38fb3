"""Developer tuning run (not a test): K1 (columns per group, groups) shapes and per-kernel GB/s at m = 10 / 30.
python tests/gpu_tune_k1.py > gpurun_out/tune_k1.log"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fortran_library_b200 as fl  # noqa: E402


def run(n, mem, kind, start, seed, shape, iters=8):
    fl.lib().flgpu_debug_set_k1_shape(*shape)
    x = fl.DeviceVector.start(start, n, seed=seed)
    first = mem + 1

    def on_iter(i):
        if i.iteration == first:
            fl.lib().flgpu_reset_kernel_times()
        return False
    st = fl.LBFGS(fl.builtin_problem(kind), x, Memory=mem, Warning=False, MaxIteration=iters + 2, time_kernels=True,
                  observer=fl.Observer(on_iteration=on_iter))
    x.free()
    kt = fl.kernel_times()
    out = []
    for name in ("k1_update_dots", "k3_direction", "callback:fused_probe", "callback:fused_store"):
        v = kt.get(name)
        if v and v["ms"] > 0:
            out.append(f"{name} {v['bytes'] / v['ms'] / 1e6:7.0f} GB/s ({v['ms'] / v['launches']:6.2f} ms)")
    print(f"n=2^{n.bit_length() - 1} m={mem} shape={shape}: " + " | ".join(out), flush=True)


if __name__ == "__main__":
    n = 1 << 27
    for shape in ((5, 2), (8, 1), (6, 2)):
        run(n, 10, fl.OBJ_ROSENBROCK, fl.START_ROSEN_PERT, 7, shape)
    for shape in ((5, 2), (6, 2), (7, 2), (8, 2), (8, 1)):
        run(n, 30, fl.OBJ_DIAGQUAD, fl.START_ZERO, 0, shape)
    for shape in ((5, 2), (8, 2)):
        run(n, 20, fl.OBJ_DIAGQUAD, fl.START_ZERO, 0, shape)
    fl.lib().flgpu_debug_set_k1_shape(0, 0)
