"""Generates tests/golden_fast/*.json -- known-answer vectors for the optional FLGPU_LS_FAST line-search policy.

FLGPU_LS_FAST is NOT a reference routine (include/flgpu.h), so these vectors pin the PRODUCT's published algorithm, as
restated by oracle/oracle.c (fast_impl) and cross-checked bit for bit against the second restatement in
oracle/oracle_np.py (fast_search) when this script runs.  Their job is regression protection: a change to
SearchCore::fast that alters a single trial point shows up here.  Same file format as tests/golden/ (hex floats).

    python tests/golden_fast/make_golden_fast.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)

import _cases            # noqa: E402
import _oracle as O      # noqa: E402
import oracle_np as N    # noqa: E402
import make_golden as MG  # noqa: E402

CASES = [
    ("fast_lbfgs_quartic10", "lbfgs", "quartic", 10, dict()),
    ("fast_lbfgs_quartic10_weak_ffd", "lbfgs", "quartic", 10, dict(Strong=False, use_ffd=True, Memory=5)),
    ("fast_lbfgs_rosenR0_100_m10", "lbfgs", "rosenR0", 100, dict(use_ffd=True)),
    ("fast_lbfgs_rosenR1_64_m5_c2_01", "lbfgs", "rosenR1", 64, dict(use_ffd=True, Memory=5, WolfeConst2=0.1)),
    ("fast_lbfgs_diag_60_m30", "lbfgs", "diag", 60, dict(use_ffd=True, Memory=30, MaxIteration=40)),
    ("fast_cg_dy_quartic10", "cg", "quartic", 10, dict(Method="DY")),
    ("fast_cg_pr_quartic_200", "cg", "quartic", 200, dict(Method="PR", use_ffd=True)),
    ("fast_cg_dy_rosenR1_64", "cg", "rosenR1", 64, dict(Method="DY", use_ffd=True, MaxIteration=200)),
    ("fast_sd_quartic10", "sd", "quartic", 10, dict(MaxIteration=40)),
]


def main():
    N.LINE_SEARCH_POLICY = 1
    with O.fast_line_search():
        for fname, algo, name, n, kw in CASES:
            x0, x, st, tr, use = MG.run_case(algo, name, n, kw)
            MG.cross_check(algo, name, n, kw, x, tr)
            doc = {
                "generator": "tests/golden_fast/make_golden_fast.py (oracle/oracle.c fast_impl; cross-checked with "
                             "oracle_np.py fast_search); line_search = fast, NOT a reference routine",
                "algorithm": algo, "objective": name, "n": n, "options": kw,
                "x0": [float(v).hex() for v in x0],
                "x_final": [float(v).hex() for v in x],
                "iterations": int(st.n_iter), "status": int(st.status),
                "n_f": int(st.n_f), "n_fd": int(st.n_fd), "n_ffd": int(st.n_ffd), "n_trials": int(st.n_trials),
                "rows": [[int(r[0]), float(r[1]).hex(), float(r[2]).hex(), float(r[3]).hex(), int(r[4])] for r in tr.rows],
                "p_first": [[float(v).hex() for v in p] for p in tr.p[:3]],
            }
            with open(os.path.join(HERE, fname + ".json"), "w") as fh:
                json.dump(doc, fh, indent=0)
            print(f"{fname}: {st.n_iter} iterations, {st.n_trials} trials, status {st.status}")
    N.LINE_SEARCH_POLICY = 0


if __name__ == "__main__":
    main()
