#!/usr/bin/env python
"""profiles/make_scaling.py -- builds profiles/r01_scaling.md from the committed bench.py lines (profiles/r01_bench_*.json)."""
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def load(f):
    return json.loads(open(os.path.join(HERE, f)).read().strip().splitlines()[-1])


rows = [(1, load('r01_bench_n1.json')), (2, load('r01_bench_n2.json')), (4, load('r01_bench_n4.json')), (8, load('r01_bench_n8.json'))]
base = rows[0][1]['value']
out = ["# Scaling of the headline workload (bench.py, this round's builder runs; the driver re-measures at round end)", "",
       "## Strong scaling: n = 2^28 global (rows per GPU = 2^28 / N)", "",
       "LBFGS m=10, extended Rosenbrock, fused line search, 30 timed iterations after 3 warm-up, CUDA events, max over ranks.",
       "Exchange per reduction: one kernel over IPC-mapped peer memory (DESIGN.md section 5). All four lines are from the final",
       "build; device-resident line search in auto mode = on at N = 8 (2^25 rows per GPU), host-driven at N = 1, 2, 4.", "",
       "| GPUs | it/s | ms/step | speed-up | efficiency | kernel GB/s per GPU (whole step) | e2e it/s (host x) | plain-callback it/s |",
       "|---:|---:|---:|---:|---:|---:|---:|---:|"]
for n, d in rows:
    out.append(f"| {n} | {d['value']:.2f} | {d['ms_per_step']:.3f} | {d['value'] / base:.2f}x | {d['value'] / base / n:.1%} | "
               f"{d['roofline']['whole_step']['GBps']:.0f} | {d['e2e']['value']:.2f} | {d['other_line_search_mode']['value']:.2f} |")
w = load('r01_bench_n8_weak_m10_2p31.json')
out += ["", "## Weak scaling: n = 2^31 on 8 GPUs (2^28 rows per GPU, the same per-GPU work as the 1-GPU headline)", "",
        f"LBFGS m=10, extended Rosenbrock, n = 2^31: **{w['value']:.2f} it/s**, {w['ms_per_step']:.2f} ms/step, "
        f"{w['config']['trials_per_iteration']:.1f} trials/iteration, whole-step {w['roofline']['whole_step']['GBps']:.0f} GB/s per GPU "
        f"-- against {base:.2f} it/s for n = 2^28 on one GPU: {w['value'] / base:.3f} of the 1-GPU iteration rate on 8x the problem = "
        f"**{8 * w['value'] / base:.2f}x** the work rate (north_star asks >= 6.5x).", ""]
d = load('r01_bench_n8_diag_m30_2p31.json')
out += ["BASELINE.json configs[3] -- LBFGS m=30, diagonal quadratic (condition 1e6), n = 2^31 row-sharded over 8 GPUs (2^28 rows and 130 GiB per GPU):",
        f"{d['value']:.2f} it/s, {d['ms_per_step']:.1f} ms/step, {d['config']['trials_per_iteration']:.1f} trials/iteration, whole-step "
        f"{d['roofline']['whole_step']['GBps']:.0f} GB/s per GPU ({d['roofline']['whole_step']['frac']:.1%} of measured 6467.7), "
        f"K1 {d['roofline']['kernels']['k1_update_dots']['GBps']:.0f} GB/s (3 passes of <5,2>), K3 {d['roofline']['kernels']['k3_direction']['GBps']:.0f} GB/s.",
        "", "## Per-kernel averages, strong scaling (ms, algorithmic GB/s per GPU)", "",
        "| GPUs | K1 | K3 | line search per step | sum of kernels per step | wall per step | gap |", "|---:|---|---|---:|---:|---:|---:|"]
for n, d in rows:
    k = d['roofline']['kernels']
    tp = d['config']['trials_per_iteration']
    if 'callback:device_search' in k:
        ls = k['callback:device_search']['avg_ms']
    else:
        ls = k['callback:fused_store']['avg_ms'] + tp * k['callback:fused_probe']['avg_ms']
    tot = k['k1_update_dots']['avg_ms'] + k['k3_direction']['avg_ms'] + ls
    out.append(f"| {n} | {k['k1_update_dots']['avg_ms']:.3f} / {k['k1_update_dots']['GBps']:.0f} | {k['k3_direction']['avg_ms']:.3f} / "
               f"{k['k3_direction']['GBps']:.0f} | {ls:.3f} | {tot:.3f} | {d['ms_per_step']:.3f} | {d['ms_per_step'] - tot:.3f} |")
out += ["", "Line search per step = probes (8.3 per iteration) + the store of the accepted point, or the one search kernel when the",
        "device-resident search is on. The gap is host round trips (per trial when host-driven, per search otherwise, plus one per",
        "iteration) and rank skew; it is what separates 8-GPU strong scaling from 100 %.", "",
        "Multi-GPU correctness on the same boxes (`tests/gpu_multi.py`, 2 and 8 ranks, `r01_multi8.log`): all ranks see bitwise",
        "identical scalars; device-resident search, host-driven search and the ncclAllGather fallback give bitwise identical",
        "minimisers; first directions agree with the 1-GPU run to <= 1e-13; the sharded two-loop operator agrees to 5e-16;",
        "AugmentedLagrangian over LBFGS / CG takes the same number of outer iterations sharded and unsharded."]
open(os.path.join(HERE, 'r01_scaling.md'), 'w').write("\n".join(out) + "\n")
print("\n".join(out))
