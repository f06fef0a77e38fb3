"""Loader for tests/golden/*.json (known-answer vectors made by tests/golden/make_golden.py)."""
import glob
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(HERE, "golden", "*.json")))


def load(name):
    with open(os.path.join(HERE, "golden", name + ".json")) as fh:
        d = json.load(fh)
    d["x0"] = np.array([float.fromhex(v) for v in d["x0"]])
    d["x_final"] = np.array([float.fromhex(v) for v in d["x_final"]])
    d["rows"] = [(r[0], float.fromhex(r[1]), float.fromhex(r[2]), float.fromhex(r[3]), r[4]) for r in d["rows"]]
    d["p_first"] = [np.array([float.fromhex(v) for v in p]) for p in d["p_first"]]
    return d
