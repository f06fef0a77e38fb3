#!/usr/bin/env python
"""profiles/make_sweep_table.py out.md sweep_n1.json sweep_n2.json ... -- BASELINE.json configs[4]: one table over
n = 2^26..2^31 x workload x {1,2,4,8} GPUs from the JSON files profiles/sweep.py wrote, with the strong-scaling
efficiency it/s(N) / (N it/s(1)) where the 1-GPU cell exists and the per-GPU HBM rate over the whole step."""
import json
import sys

out, files = sys.argv[1], sys.argv[2:]
data = {}
for f in files:
    d = json.load(open(f))
    for r in d["rows"]:
        data[(r["workload"], r["log2n"], d["gpus"])] = r
gpus = sorted({k[2] for k in data})
keys = sorted({(k[0], k[1]) for k in data}, key=lambda k: (k[0], k[1]))
lines = ["| workload | n | " + " | ".join(f"{g} GPU{'s' if g > 1 else ''}: it/s (trials/it, GB/s per GPU over the step, efficiency)" for g in gpus) + " |",
         "|---|---|" + "---|" * len(gpus)]
for w, l in keys:
    base = data.get((w, l, 1))
    cells = []
    for g in gpus:
        r = data.get((w, l, g))
        if r is None:
            cells.append("not run")
        elif r.get("skipped"):
            cells.append("does not fit")
        else:
            eff = f", {r['it_per_s'] / (g * base['it_per_s']):.2f}" if base and not base.get("skipped") and g > 1 else ""
            cells.append(f"{r['it_per_s']:.1f} ({r['trials_per_it']:.1f}, {r['step_GBps']:.0f}{eff})")
    lines.append(f"| {w} | 2^{l} | " + " | ".join(cells) + " |")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
