"""Oracle trajectories at production launch geometry, computed in worker processes (TEST INFRASTRUCTURE).

One oracle run at n = 2^22 takes ~20 s of one host core, and a trajectory check needs the oracle under its three
summation orders, so the runs are made concurrently, one process each (the C oracle keeps its objective selection,
summation mode and trace hook in globals, hence processes and not threads), while the GPU tests that need no oracle
run in the main process.  Each worker stores the first `iters` directions (and, when asked, iterates and gradients)
as .npy files; the test maps them read-only.

    python tests/_oracle_traj.py <outdir> <algo:lbfgs|cg> <objective name of _cases.OBJECTIVES> <n> <sum_mode>
                                 <iters> <keep_xg:0|1> <json kwargs>
"""
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(argv):
    sys.path.insert(0, HERE)
    import _cases
    import _oracle as O
    out, algo, name, n, mode, iters, keep_xg, kw = (argv[0], argv[1], argv[2], int(argv[3]), int(argv[4]),
                                                     int(argv[5]), int(argv[6]), json.loads(argv[7]))
    kind = _cases.OBJECTIVES[name][0]
    x0 = _cases.start(name, n)
    tr = O.Trace(max_vec_iters=iters)
    run = O.lbfgs if algo == "lbfgs" else O.cg
    x, st = run(O.builtin_callbacks(kind, 0, n), x0.copy(), trace=tr, sum_mode=mode, Warning=False, **kw)
    np.save(os.path.join(out, "p.npy"), np.stack(tr.p))
    if keep_xg:
        np.save(os.path.join(out, "x.npy"), np.stack(tr.x))
        np.save(os.path.join(out, "g.npy"), np.stack(tr.g))
    json.dump({"rows": tr.rows, "n_iter": st.n_iter, "status": st.status, "n_trials": st.n_trials},
              open(os.path.join(out, "meta.json"), "w"))


class Job:
    """One oracle trajectory being computed in a subprocess."""

    def __init__(self, root, algo, name, n, mode, iters, keep_xg=False, **kw):
        self.dir = os.path.join(root, f"{algo}_{name}_{n}_{mode}_" + "_".join(f"{k}{v}" for k, v in sorted(kw.items())))
        os.makedirs(self.dir, exist_ok=True)
        self.proc = subprocess.Popen([sys.executable, os.path.abspath(__file__), self.dir, algo, name, str(n), str(mode),
                                      str(iters), str(int(keep_xg)), json.dumps(kw)],
                                     stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)

    def result(self, timeout=900):
        out, _ = self.proc.communicate(timeout=timeout)
        if self.proc.returncode != 0:
            raise RuntimeError("oracle worker failed:\n" + out)
        meta = json.load(open(os.path.join(self.dir, "meta.json")))
        load = lambda f: np.load(os.path.join(self.dir, f), mmap_mode="r")   # noqa: E731
        meta["p"] = load("p.npy")
        if os.path.exists(os.path.join(self.dir, "x.npy")):
            meta["x"], meta["g"] = load("x.npy"), load("g.npy")
        return meta


if __name__ == "__main__":
    _worker(sys.argv[1:])
