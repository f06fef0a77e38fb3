"""Developer probe (not a test): per-iteration direction error of the host simulator vs the oracle's
summation-order envelope.  python tests/dev_envelope.py [n]"""
import sys
import numpy as np
import _cases, _hostsim as H, _oracle as O

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
CASES = [("rosenR0", dict(Memory=10)), ("rosenR1", dict(Memory=10)), ("rosenR1", dict(Memory=5)),
         ("quartic", dict(Memory=10)), ("diag", dict(Memory=30, MaxIteration=40)),
         ("rosenR1", dict(Memory=1, MaxIteration=30)), ("rosenR1", dict(Memory=30, MaxIteration=40)),
         ("quartic", dict(Memory=30, MaxIteration=40))]
for name, kw in CASES:
    kind = _cases.OBJECTIVES[name][0]
    traces, env = _cases.oracle_envelope(name, n, lambda cbs, x, **k: O.lbfgs(cbs, x, use_ffd=True, **k), **kw)
    ob = H.Observer(max_vec_iters=20)
    x, st = H.lbfgs(kind, _cases.start(name, n), observer=ob, use_ffd=True, Warning=False, n_global=n, **kw)
    ld = traces[1].p
    ref = 0.0
    worst = 0.0
    rows = []
    for k in range(min(len(ob.p), 20, *(len(t.p) for t in traces))):
        ref = max(ref, _cases.rel(traces[0].p[k], ld[k]), _cases.rel(traces[2].p[k], ld[k]))
        err = _cases.rel(ob.p[k], ld[k])
        rows.append((k, err, ref))
        worst = max(worst, err / max(ref, 1e-12 / 64))
    print(f"{name:8s} {str(kw):40s} worst err/noise = {worst:9.2f}   " +
          " ".join(f"{e:.0e}/{r:.0e}" for _, e, r in rows[::3]))
