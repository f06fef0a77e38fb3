"""Generates tests/golden/*.json -- known-answer vectors for the hot path.

The reference (Fortran + MKL) cannot be compiled or run in this image and its own tests store no
numbers for this path, so these vectors are outputs of the CPU oracle
(oracle/oracle.c: strict-IEEE C restatement of NonlinearOptimization.f90, sequential sums), which is
itself cross-checked bit for bit against the independent NumPy transcription (oracle/oracle_np.py)
when this script runs.  They pin the oracle against regressions and give the GPU tests fixed numbers
to hit on a machine where nothing under /root/reference exists.  PARITY UNPINNED applies.

    python tests/golden/make_golden.py        # rewrites tests/golden/*.json

Floats are stored as C99 hex strings (exact).  The option mixes are the ones the reference's own
smoke test exercises (test/test.f90:350-388: CG DY non-strong/strong/f_fd, CG PR, LBFGS default /
Strong / f_fd+Memory=5 on f=sum x^4, dim=10) plus the benchmark objectives of BASELINE.json.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)

import _cases            # noqa: E402
import _oracle as O      # noqa: E402
import oracle_np as N    # noqa: E402

CASES = [
    # name, algorithm, objective, n, kwargs
    ("lbfgs_quartic10_default", "lbfgs", "quartic", 10, dict()),                         # test.f90:375-378
    ("lbfgs_quartic10_strong", "lbfgs", "quartic", 10, dict(Strong=True)),               # test.f90:380-383
    ("lbfgs_quartic10_ffd_mem5", "lbfgs", "quartic", 10, dict(use_ffd=True, Memory=5)),  # test.f90:385-388
    ("cg_dy_quartic10_weak", "cg", "quartic", 10, dict(Method="DY", Strong=False)),      # test.f90:350-353
    ("cg_dy_quartic10_strong", "cg", "quartic", 10, dict(Method="DY")),                  # test.f90:354-357
    ("cg_dy_quartic10_ffd", "cg", "quartic", 10, dict(Method="DY", use_ffd=True)),       # test.f90:358-361
    ("cg_pr_quartic10", "cg", "quartic", 10, dict(Method="PR")),                         # test.f90:363-367
    ("cg_pr_quartic10_ffd", "cg", "quartic", 10, dict(Method="PR", use_ffd=True)),       # test.f90:369-373
    ("lbfgs_rosenR0_100_m10", "lbfgs", "rosenR0", 100, dict(use_ffd=True)),              # BASELINE configs[0] shape
    ("lbfgs_rosenR1_64_m5", "lbfgs", "rosenR1", 64, dict(use_ffd=True, Memory=5)),
    ("lbfgs_diag_60_m30", "lbfgs", "diag", 60, dict(use_ffd=True, Memory=30, MaxIteration=40)),
    ("cg_dy_quartic_200", "cg", "quartic", 200, dict(Method="DY", use_ffd=True)),        # BASELINE configs[2] shape
    ("cg_pr_quartic_200", "cg", "quartic", 200, dict(Method="PR", use_ffd=True)),
    ("sd_quartic10", "sd", "quartic", 10, dict(MaxIteration=40)),                         # test.f90:336-339 shape
    ("sd_quartic10_ffd", "sd", "quartic", 10, dict(use_ffd=True, MaxIteration=40)),       # test.f90:341-344 shape
    ("sd_rosenR1_64", "sd", "rosenR1", 64, dict(use_ffd=True, MaxIteration=40)),
]


def _np_objective(name, n):
    if name.startswith("rosen"):
        return N.rosenbrock()
    if name == "quartic":
        return N.quartic()
    return N.diagquad(np.array([O.lib().orc_diag_coeff(i, n) for i in range(n)]))


def run_case(algo, name, n, kw):
    kw = dict(kw)
    use = kw.pop("use_ffd", False)
    kind = _cases.OBJECTIVES[name][0]
    x0 = _cases.start(name, n)
    tr = O.Trace()
    run = {"lbfgs": O.lbfgs, "cg": O.cg, "sd": O.sd}[algo]
    x, st = run(O.builtin_callbacks(kind, 0, n), x0.copy(), use_ffd=use, Warning=False, trace=tr, **kw)
    return x0, x, st, tr, use


def cross_check(algo, name, n, kw, x, tr):
    kw = dict(kw)
    use = kw.pop("use_ffd", False)
    f, fd, ffd = _np_objective(name, n)
    run = {"lbfgs": N.lbfgs, "cg": N.conjugate_gradient, "sd": N.steepest_descent}[algo]
    with np.errstate(all="ignore"):
        xb, c = run(f, fd, _cases.start(name, n), f_fd=ffd if use else None, Warning=False, **kw)
    assert np.array_equal(x, xb), "oracle.c and oracle_np.py disagree"
    assert len(c.history) == len(tr.rows)


def main():
    for fname, algo, name, n, kw in CASES:
        x0, x, st, tr, use = run_case(algo, name, n, kw)
        cross_check(algo, name, n, kw, x, tr)
        doc = {
            "generator": "tests/golden/make_golden.py (oracle/oracle.c, sequential sums; cross-checked with oracle_np.py)",
            "algorithm": algo, "objective": name, "n": n, "options": kw,
            "x0": [float(v).hex() for v in x0],
            "x_final": [float(v).hex() for v in x],
            "iterations": int(st.n_iter), "status": int(st.status),
            "n_f": int(st.n_f), "n_fd": int(st.n_fd), "n_ffd": int(st.n_ffd), "n_trials": int(st.n_trials),
            # per outer iteration: accepted step, f, phi'(0), trials of that search
            "rows": [[int(r[0]), float(r[1]).hex(), float(r[2]).hex(), float(r[3]).hex(), int(r[4])] for r in tr.rows],
            # first three search directions (exact)
            "p_first": [[float(v).hex() for v in p] for p in tr.p[:3]],
        }
        with open(os.path.join(HERE, fname + ".json"), "w") as fh:
            json.dump(doc, fh, indent=0)
        print(f"{fname}: {st.n_iter} iterations, {st.n_trials} trials, status {st.status}")


if __name__ == "__main__":
    main()
