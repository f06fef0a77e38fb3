#!/usr/bin/env python
"""profiles/summarise.py <tag> -- turn the raw ncu outputs of profiles/run_ncu.sh (gpurun_out/<tag>_*) into the
committed summaries: profiles/<tag>_launches.md (per-kernel launch counts, time and share of the step),
profiles/<tag>_kernels.md (per-kernel --set full metrics) and profiles/ncu_traffic.json (DRAM bytes per launch,
read by bench.py for roofline.traffic).  Runs here (no GPU): `ncu -i` only reads the report."""
import csv
import io
import json
import os
import re
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")


def short(name):
    m = re.search(r"(k1_update_dots_kernel|k3_direction_tma_kernel|k3_direction_kernel|tree_kernel|search_kernel|k2_solve_kernel|trial_kernel|dot_kernel|neg_kernel|"
                  r"cg_dots_kernel|cg_update_kernel|objective_multi_kernel|objective_kernel|start_kernel|combine_kernel|set_scalar_kernel)(<[^(]*>)?", name)
    if m:
        return m.group(1) + (m.group(2) or "")
    return name[:60]


def launches(tag):
    path = os.path.join(OUT, f"{tag}_launches.csv")
    rows = []
    with open(path) as fh:
        text = fh.read()
    start = text.index('"ID"')
    rd = csv.DictReader(io.StringIO(text[start:]))
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ns = v * {"nsecond": 1, "ns": 1, "usecond": 1e3, "us": 1e3, "msecond": 1e6, "ms": 1e6, "second": 1e9}.get(unit, 1)
        rows.append((int(r["ID"]), short(r["Kernel Name"]), ns))
    agg = defaultdict(lambda: [0, 0.0])
    for _, k, ns in rows:
        agg[k][0] += 1
        agg[k][1] += ns
    total = sum(v[1] for v in agg.values())
    lines = [f"# ncu launch list `{tag}` (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: read the SHARES)",
             "", f"launches captured: {len(rows)}, total kernel time {total / 1e6:.1f} ms", "",
             "| kernel | launches | total ms | avg us | share |", "|---|---:|---:|---:|---:|"]
    for k, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{k}` | {c} | {ns / 1e6:.2f} | {ns / c / 1e3:.1f} | {ns / total:.1%} |")
    open(os.path.join(PROF, f"{tag}_launches.md"), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
        "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]


def kernels(tag, parts):
    traffic = {}
    tpath = os.path.join(PROF, "ncu_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath))
    lines = [f"# ncu --set full summaries `{tag}` (per launch; --clock-control none)", ""]
    for part in parts:
        rep = os.path.join(OUT, f"{tag}_{part}.ncu-rep")
        raw = os.path.join(OUT, f"{tag}_{part}.raw.csv")          # exported on the GPU box by run_ncu.sh
        if os.path.exists(raw):
            txt = open(raw).read()
        elif os.path.exists(rep):
            txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        else:
            continue
        rd = list(csv.reader(io.StringIO(txt)))
        hdr, units, data = rd[0], rd[1], rd[2:]
        idx = {h: i for i, h in enumerate(hdr)}
        for row in data:
            name = short(row[idx["Kernel Name"]])
            lines.append(f"## `{name}`  (launch id {row[idx['ID']]}, grid {row[idx.get('launch__grid_size', 0)]} x block {row[idx.get('launch__block_size', 0)]})")
            lines.append("")
            lines.append("| metric | value | unit |")
            lines.append("|---|---:|---|")
            vals = {}
            for k in KEYS:
                if k in idx:
                    vals[k] = row[idx[k]]
                    lines.append(f"| {k} | {row[idx[k]]} | {units[idx[k]]} |")
            try:
                def num(k):
                    v = float(vals[k].replace(",", ""))
                    u = units[idx[k]]
                    return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1}.get(u, 1)
                tot = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
                dur = float(vals["gpu__time_duration.sum"].replace(",", ""))
                du = units[idx["gpu__time_duration.sum"]]
                dur_s = dur * {"nsecond": 1e-9, "ns": 1e-9, "usecond": 1e-6, "us": 1e-6, "msecond": 1e-3, "ms": 1e-3, "second": 1}.get(du, 1e-9)
                lines.append(f"| **DRAM bytes (read+write)** | {tot / 1e9:.3f} | GB |")
                lines.append(f"| **DRAM GB/s under ncu** | {tot / dur_s / 1e9:.0f} | GB/s |")
                key = re.sub(r"<.*", "", name)
                traffic.setdefault(key, {})
                # n and memory of the capture (run_ncu.sh profiles bench.py's default workload): bench.py scales the
                # figure by rows per GPU and ignores it for another memory
                traffic[key] = {"dram_bytes_per_launch": tot, "duration_s_under_ncu": dur_s, "tag": tag, "kernel": name,
                                "n": 1 << 28, "memory": 10}
            except (KeyError, ValueError) as e:
                lines.append(f"| (derived metrics unavailable: {e}) | | |")
            lines.append("")
    open(os.path.join(PROF, f"{tag}_kernels.md"), "w").write("\n".join(lines) + "\n")
    json.dump(traffic, open(tpath, "w"), indent=1)
    print("\n".join(lines[:80]))


if __name__ == "__main__":
    tag = sys.argv[1]
    if os.path.exists(os.path.join(OUT, f"{tag}_launches.csv")):
        launches(tag)
    kernels(tag, sys.argv[2:] or ["k1k3", "multi", "ls", "tree"])
