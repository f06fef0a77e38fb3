"""Exploratory GPU run (not a pytest file): parity vs oracle and first timings.  Writes gpurun_out/explore.log."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import _oracle as O  # noqa: E402
import fortran_library_b200 as fl  # noqa: E402


def compare(name, tr, ob, xa, xb, x0):
    n_or, n_g = len(tr.rows), len(ob.rows)
    worst = 0.0
    flip = None
    for k in range(min(len(tr.p), len(ob.p), 20)):
        worst = max(worst, np.linalg.norm(tr.p[k] - ob.p[k]) / np.linalg.norm(tr.p[k]))
    for k in range(min(n_or, n_g)):
        if tr.rows[k][4] != ob.rows[k][4]:
            flip = k
            break
    sc = max(np.linalg.norm(xa), np.linalg.norm(x0))
    print(f"{name:50s} iters {n_or:4d}/{n_g:4d} dir_err20 {worst:.2e} xerr {np.linalg.norm(xa - xb) / sc:.2e} "
          f"first trial diff @ {flip}", flush=True)


def parity(n=10_000):
    cases = ((O.OBJ_ROSENBROCK, O.START_ROSEN_STD, 0, "rosenR0"), (O.OBJ_ROSENBROCK, O.START_ROSEN_PERT, 7, "rosenR1"),
             (O.OBJ_QUARTIC, O.START_QUARTIC_U, 12345, "quartic"), (O.OBJ_DIAGQUAD, O.START_ZERO, 0, "diag"))
    for kind, start, seed, nm in cases:
        x0 = O.start_vector(start, n, seed=seed)
        for kw in (dict(Memory=10), dict(Memory=5), dict(Memory=30, MaxIteration=100), dict(Memory=1, MaxIteration=60),
                   dict(Memory=10, use_ffd=False), dict(Memory=10, Strong=False, MaxIteration=40)):
            kw = dict(kw)
            use = kw.pop("use_ffd", True)
            tr = O.Trace(max_vec_iters=25)
            xa, s = O.lbfgs(O.builtin_callbacks(kind, 0, n), x0.copy(), use_ffd=use, Warning=False, trace=tr, **kw)
            ob = fl.Observer(keep_vectors=True, max_vec_iters=25)
            prob = fl.builtin_problem(kind)
            if not use:
                prob.f_fd = None
            x = fl.DeviceVector.start(start, n, seed=seed)
            assert np.array_equal(x.numpy(), x0), "start vectors differ between host and device"
            st = fl.LBFGS(prob, x, observer=ob, Warning=False, **kw)
            compare(f"lbfgs {nm} {kw} ffd={int(use)}", tr, ob, xa, x.numpy(), x0)
        for M in ("DY", "PR"):
            for kw in (dict(), dict(use_ffd=False), dict(Strong=False)):
                kw = dict(kw)
                use = kw.pop("use_ffd", True)
                tr = O.Trace(max_vec_iters=25)
                xa, s = O.cg(O.builtin_callbacks(kind, 0, n), x0.copy(), Method=M, use_ffd=use, Warning=False, trace=tr,
                             MaxIteration=100, **kw)
                ob = fl.Observer(keep_vectors=True, max_vec_iters=25)
                prob = fl.builtin_problem(kind)
                if not use:
                    prob.f_fd = None
                x = fl.DeviceVector.start(start, n, seed=seed)
                st = fl.ConjugateGradient(prob, x, Method=M, observer=ob, Warning=False, MaxIteration=100, **kw)
                compare(f"cg {M} {nm} {kw} ffd={int(use)}", tr, ob, xa, x.numpy(), x0)


def timing(n, mem=10, iters=30, kind=fl.OBJ_ROSENBROCK, start=fl.START_ROSEN_PERT, cg=None, fused=True):
    x = fl.DeviceVector.start(start, n, seed=7)
    ob = fl.Observer()
    t = time.time()
    if cg:
        st = fl.ConjugateGradient(fl.builtin_problem(kind), x, Method=cg, observer=ob, Warning=False,
                                  MaxIteration=iters, time_kernels=True, fused=fused)
    else:
        st = fl.LBFGS(fl.builtin_problem(kind), x, Memory=mem, observer=ob, Warning=False, MaxIteration=iters,
                      time_kernels=True, fused=fused)
    dt = time.time() - t
    kt = fl.kernel_times()
    print(f"--- n={n} mem={mem} cg={cg} fused={fused}: iterations={st.iterations} trials={st.n_trials} wall={dt:.3f}s "
          f"it/s={st.iterations / dt:.2f} launches={st.gpu_launches} syncs={st.host_syncs}", flush=True)
    tot = sum(v["ms"] for v in kt.values())
    for name, v in sorted(kt.items(), key=lambda kv: -kv[1]["ms"]):
        gbs = v["bytes"] / (v["ms"] * 1e-3) / 1e9 if v["ms"] > 0 else 0.0
        print(f"    {name:18s} launches {v['launches']:6d} ms {v['ms']:10.3f} share {v['ms'] / tot:6.1%} "
              f"avg_us {1e3 * v['ms'] / max(v['launches'], 1):9.1f} GB/s {gbs:8.1f}", flush=True)
    x.free()


if __name__ == "__main__":
    print(fl.lib().flgpu_version().decode(), "devices", fl.device_count(), flush=True)
    if "parity" in sys.argv:
        parity()
    for n in (1 << 20, 1 << 24):
        timing(n)
    timing(1 << 26, mem=30, kind=fl.OBJ_DIAGQUAD, start=fl.START_ZERO)
    timing(1 << 26, mem=5)
    timing(1 << 26, kind=fl.OBJ_QUARTIC, start=fl.START_QUARTIC_U, cg="DY")
    timing(1 << 26, kind=fl.OBJ_QUARTIC, start=fl.START_QUARTIC_U, cg="PR", fused=False)
    timing(1 << 28, fused=False)
    timing(1 << 28)
