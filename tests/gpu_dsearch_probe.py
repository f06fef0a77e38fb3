"""Developer probe (not a test): host-driven vs device-resident line search, wall-clock iteration rate by size."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fortran_library_b200 as fl  # noqa: E402

for log2n in (14, 18, 20, 22, 24, 26, 28):
    n = 1 << log2n
    for dev in (False, True):
        x = fl.DeviceVector.start(fl.START_ROSEN_PERT, n, seed=7)
        mark = {}

        def on_iter(i):
            if i.iteration == 12:
                fl.lib().flgpu_memcpy(None, None, 0, 1, 1, i.stream)
                mark["t0"], mark["tr0"] = time.perf_counter(), i.total_trials
            if i.iteration == 42:
                fl.lib().flgpu_memcpy(None, None, 0, 1, 1, i.stream)
                mark["t1"], mark["tr1"] = time.perf_counter(), i.total_trials
                return True
            return False
        st = fl.LBFGS(fl.builtin_problem(fl.OBJ_ROSENBROCK), x, Memory=10, Warning=False, MaxIteration=40,
                      observer=fl.Observer(on_iteration=on_iter), device_search=dev)
        x.free()
        dt = mark["t1"] - mark["t0"]
        print(f"n=2^{log2n} device_search={int(dev)}: {30 / dt:9.1f} it/s  {1e3 * dt / 30:8.3f} ms/it  "
              f"{(mark['tr1'] - mark['tr0']) / 30:.1f} trials/it  syncs={st.host_syncs}", flush=True)
