"""Prints the measured margins behind the tolerances of the GPU suite (run on the GPU box; not a pytest file):
how far each asserted quantity actually is from its bound.  `python tests/gpu_calibrate.py > gpurun_out/calibrate.log`"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _cases  # noqa: E402
import _oracle as O  # noqa: E402
import fortran_library_b200 as fl  # noqa: E402


def problem(name, use=True):
    p = fl.builtin_problem(_cases.OBJECTIVES[name][0])
    if not use:
        p.f_fd = None
    return p


def dev_start(name, n):
    kind, st, seed = _cases.OBJECTIVES[name]
    return fl.DeviceVector.start(st, n, seed=seed)


def envelope_ratio(traces, p_got):
    ref, ld, worst = 0.0, traces[1].p, 0.0
    for k in range(min(len(p_got), 20, *(len(t.p) for t in traces))):
        ref = max(ref, _cases.rel(traces[0].p[k], ld[k]), _cases.rel(traces[2].p[k], ld[k]))
        err = _cases.rel(p_got[k], ld[k])
        if err > _cases.FLOOR:
            worst = max(worst, err / max(ref, 1e-300))
    return worst, ref


n = 10_000
print("== trajectory envelopes: worst err/oracle-noise over the first 20 directions (only where err > 1e-12)")
for name, kw in [("rosenR0", dict(Memory=10)), ("rosenR1", dict(Memory=10)), ("rosenR1", dict(Memory=5)), ("quartic", dict(Memory=10)),
                 ("diag", dict(Memory=30, MaxIteration=40)), ("rosenR1", dict(Memory=1, MaxIteration=30)),
                 ("rosenR1", dict(Memory=10, Strong=False, MaxIteration=30)), ("rosenR1", dict(Memory=10, use_ffd=False)),
                 ("quartic1", dict(Memory=10))]:
    for fused in (True, False):
        kw2 = dict(kw)
        use = kw2.pop("use_ffd", True)
        traces, _ = _cases.oracle_envelope(name, n, lambda cbs, x, **k: O.lbfgs(cbs, x, use_ffd=use, **k), **kw2)
        ob = fl.Observer(keep_vectors=True, max_vec_iters=20)
        fl.LBFGS(problem(name, use), dev_start(name, n), observer=ob, Warning=False, fused=fused, **kw2)
        w, ref = envelope_ratio(traces, ob.p)
        print(f"lbfgs {name:9s} {str(kw):60s} fused={fused!s:5s} worst ratio {w:8.2f} (final oracle noise {ref:.1e})", flush=True)
for M in ("DY", "PR"):
    for name, kw in [("quartic", dict()), ("rosenR1", dict(MaxIteration=60)), ("diag", dict(MaxIteration=60)),
                     ("quartic", dict(Strong=False, MaxIteration=60)), ("quartic", dict(use_ffd=False)), ("quartic1", dict())]:
        kw2 = dict(kw)
        use = kw2.pop("use_ffd", True)
        traces, _ = _cases.oracle_envelope(name, n, lambda cbs, x, **k: O.cg(cbs, x, Method=M, use_ffd=use, **k), **kw2)
        ob = fl.Observer(keep_vectors=True, max_vec_iters=20)
        fl.ConjugateGradient(problem(name, use), dev_start(name, n), Method=M, observer=ob, Warning=False, **kw2)
        w, ref = envelope_ratio(traces, ob.p)
        print(f"cg {M} {name:9s} {str(kw):60s} worst ratio {w:8.2f} (final oracle noise {ref:.1e})", flush=True)
for name, kw in [("quartic", dict(MaxIteration=40)), ("rosenR1", dict(MaxIteration=40)), ("diag", dict(MaxIteration=40, Strong=False))]:
    traces, _ = _cases.oracle_envelope(name, n, lambda cbs, x, **k: O.sd(cbs, x, use_ffd=True, **k), **kw)
    ob = fl.Observer(keep_vectors=True, max_vec_iters=20)
    fl.SteepestDescent(problem(name), dev_start(name, n), observer=ob, Warning=False, **kw)
    w, ref = envelope_ratio(traces, ob.p)
    print(f"sd {name:9s} {str(kw):60s} worst ratio {w:8.2f} (final oracle noise {ref:.1e})", flush=True)

print("== one-step directions on the oracle's history: worst error, and the oracle's own distance from the exact two-loop")
for name, mem in [("rosenR1", 10), ("rosenR1", 3), ("quartic", 10), ("diag", 30), ("rosenR0", 5), ("quartic", 1), ("rosenR1", 17),
                  ("rosenR0", 10), ("quartic1", 10)]:
    kind = _cases.OBJECTIVES[name][0]
    x0 = _cases.start(name, n)
    tr = O.Trace(max_vec_iters=22)
    O.lbfgs(O.builtin_callbacks(kind, 0, n), x0.copy(), Memory=mem, use_ffd=True, Warning=False, MaxIteration=22, trace=tr)
    g0 = np.empty(n)
    O.lib().orc_obj_select(kind, 0, n)
    O.lib().orc_obj_fd(g0.ctypes.data_as(C.c_void_p), x0.ctypes.data_as(C.c_void_p), C.byref(C.c_int(n)))
    xs, gs = [x0] + tr.x, [g0] + tr.g
    h = fl.History(n, mem)
    pairs, worst, wnoise = [], 0.0, 0.0
    for k in range(min(20, len(tr.p) - 1)):
        h.push(xs[k + 1], xs[k], gs[k + 1], gs[k])
        p, xt, gp, pp = h.direction(gs[k + 1], xs[k + 1])
        pairs = (pairs + [(xs[k + 1] - xs[k], gs[k + 1] - gs[k])])[-mem:]
        exact = _cases.two_loop_extended(pairs, gs[k + 1])
        worst = max(worst, _cases.rel(p, exact)); wnoise = max(wnoise, _cases.rel(tr.p[k + 1], exact))
    h.close()
    print(f"one-step {name:9s} m={mem:2d}: ours {worst:.2e}   oracle (sequential double) {wnoise:.2e}", flush=True)

print("== minimisers / iteration counts")
for name in ("rosenR0", "rosenR1", "quartic1"):
    xr, sr = O.lbfgs(O.builtin_callbacks(_cases.OBJECTIVES[name][0], 0, n), _cases.start(name, n), use_ffd=True, Warning=False)
    for fused in (True, False):
        x = dev_start(name, n)
        st = fl.LBFGS(problem(name), x, Warning=False, fused=fused)
        print(f"lbfgs {name} fused={fused}: rel dx {_cases.rel(x.numpy(), xr):.2e} iterations {st.iterations} vs oracle {sr.n_iter}, "
              f"status {st.status}/{sr.status}", flush=True)
for name in ("quartic1", "quartic"):
    x0 = _cases.start(name, n)
    for M in ("DY", "PR"):
        xr, sr = O.cg(O.builtin_callbacks(_cases.OBJECTIVES[name][0], 0, n), x0.copy(), Method=M, use_ffd=True, Warning=False)
        x = dev_start(name, n)
        st = fl.ConjugateGradient(problem(name), x, Method=M, Warning=False)
        xn = x.numpy()
        print(f"cg {M} {name}: |x-xr|/|xr| {_cases.rel(xn, xr):.2e} |x-xr|/|x0| {np.linalg.norm(xn - xr) / np.linalg.norm(x0):.2e} "
              f"|x|/|x0| {np.linalg.norm(xn) / np.linalg.norm(x0):.2e} (oracle {np.linalg.norm(xr) / np.linalg.norm(x0):.2e}) "
              f"iterations {st.iterations} vs {sr.n_iter} status {st.status}/{sr.status}", flush=True)
print("== edge cases (Memory=4, MaxIteration=5, quartic, f / fd callbacks)")
for nn in (1, 2, 3, 7, 33, 1001):
    x0 = _cases.start("quartic", nn)
    xr, sr = O.lbfgs(O.builtin_callbacks(O.OBJ_QUARTIC, 0, nn), x0.copy(), Memory=4, Warning=False, MaxIteration=5)
    x = fl.DeviceVector.from_numpy(x0)
    st = fl.LBFGS(problem("quartic", False), x, Memory=4, Warning=False, MaxIteration=5)
    print(f"n={nn}: rel dx {_cases.rel(x.numpy(), xr):.2e} iterations {st.iterations}/{sr.n_iter}", flush=True)
