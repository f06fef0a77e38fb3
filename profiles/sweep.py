#!/usr/bin/env python
"""profiles/sweep.py -- BASELINE.json configs[2] and [4]: n = 2^26..2^28 (GLOBAL n; under torchrun the rows are
sharded over the ranks: strong scaling, raise --max-log2n for the weak-scaling cells) x {LBFGS m=5,10,30 on
Rosenbrock / diag-quadratic, CG-DY and CG-PR on the quartic, SteepestDescent}, K timed iterations each after the
warm-up iterations, per-kernel CUDA-event totals of rank 0 -> achieved GB/s per GPU over the algorithmic bytes
(DESIGN.md section 3).

    python profiles/sweep.py [--out gpurun_out/sweep.md] [--max-log2n 28] [--line-search fast]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 profiles/sweep.py ...

Every rate is timed on the device (CUDA events on the library's stream around the K timed iterations) and is the
MAX over ranks; cells whose work space does not fit 180 GB per GPU are listed as skipped.  --json writes the rows for
profiles/make_sweep_table.py, which combines the 1/2/4/8-GPU files into the efficiency table.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fortran_library_b200 as fl  # noqa: E402

PEAK = 6467.7


RANK = int(os.environ.get("RANK", "0"))
WORLD = int(os.environ.get("WORLD_SIZE", "1"))
COMM = None          # row-shard communicator under torchrun


def setup_comm():
    """One process per GPU; the 128-byte communicator id travels over torch.distributed (as in bench.py)."""
    global COMM
    if WORLD == 1:
        return None
    import torch
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def bcast(data):
        t = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if RANK == 0:
            t.copy_(torch.frombuffer(bytearray(data), dtype=torch.uint8))
        dist.broadcast(t, 0)
        return bytes(t.cpu().numpy().tobytes())
    COMM = fl.comm_create(RANK, WORLD, bcast)
    return dist


def one(algo, kind, start, seed, n, K, W, **kw):
    lo = (n * RANK // WORLD) // 2 * 2            # even boundaries: Rosenbrock pairs never straddle shards
    hi = n if RANK == WORLD - 1 else (n * (RANK + 1) // WORLD) // 2 * 2
    x = fl.DeviceVector.start(start, hi - lo, seed=seed, offset=lo, n_global=n)
    kw = dict(kw)
    if COMM is not None:
        kw.update(comm=COMM, offset=lo, n_global=n)
    mem = kw.get("Memory", 0)
    first = (mem if algo == "lbfgs" else 1) + W - 1
    last = first + K
    mark = {}

    import torch
    ev = [torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)]

    def on_iter(i):
        if i.iteration == first:
            fl.lib().flgpu_reset_kernel_times()
            ev[0].record(torch.cuda.ExternalStream(i.stream))
            mark["t0"], mark["tr0"] = time.perf_counter(), i.total_trials
        elif i.iteration == last:
            ev[1].record(torch.cuda.ExternalStream(i.stream))
            fl.lib().flgpu_memcpy(None, None, 0, 1, 1, i.stream)       # drain the stream
            mark["t1"], mark["tr1"] = time.perf_counter(), i.total_trials
            return True
        return False
    ob = fl.Observer(on_iteration=on_iter)
    run = {"lbfgs": fl.LBFGS, "cg": fl.ConjugateGradient, "sd": fl.SteepestDescent}[algo]
    st = run(fl.builtin_problem(kind), x, observer=ob, Warning=False, MaxIteration=W + K + 1, time_kernels=True, **kw)
    x.free()
    if "t1" not in mark:
        return None
    kt = fl.kernel_times()
    ms = sum(v["ms"] for v in kt.values())
    gb = sum(v["bytes"] for v in kt.values()) / 1e9
    dev_ms = ev[0].elapsed_time(ev[1])
    if WORLD > 1:                                    # the slowest rank counts
        import torch.distributed as dist
        t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms = float(t.item())
    return {"it_per_s": K / (dev_ms * 1e-3), "trials_per_it": (mark["tr1"] - mark["tr0"]) / K, "kernel_ms_per_it": ms / K,
            "GBps": gb / (ms * 1e-3), "frac": gb / (ms * 1e-3) / PEAK, "GB_per_it": gb / K, "ms_per_it": dev_ms / K,
            "step_GBps": gb / (dev_ms * 1e-3)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.md"))
    ap.add_argument("--min-log2n", type=int, default=26)
    ap.add_argument("--max-log2n", type=int, default=28)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--line-search", default="reference", choices=["reference", "fast"],
                    help="flgpu_options.line_search for every case (fast = FLGPU_LS_FAST, not a reference routine)")
    ap.add_argument("--skip", default="", help="comma-separated substrings of case labels to leave out")
    ap.add_argument("--json", default=None, help="also write the rows as JSON (for profiles/make_sweep_table.py)")
    a = ap.parse_args()
    fl.require_gpu()
    dist = setup_comm()
    rows = []
    cases = [("LBFGS m=5 Rosenbrock", "lbfgs", fl.OBJ_ROSENBROCK, fl.START_ROSEN_PERT, 7, dict(Memory=5)),
             ("LBFGS m=10 Rosenbrock", "lbfgs", fl.OBJ_ROSENBROCK, fl.START_ROSEN_PERT, 7, dict(Memory=10)),
             ("LBFGS m=10 Rosenbrock, plain callbacks", "lbfgs", fl.OBJ_ROSENBROCK, fl.START_ROSEN_PERT, 7, dict(Memory=10, fused=False)),
             ("LBFGS m=10 Rosenbrock, Increment=2 (a reference tunable; NOT the default)", "lbfgs", fl.OBJ_ROSENBROCK,
              fl.START_ROSEN_PERT, 7, dict(Memory=10, Increment=2.0)),
             ("LBFGS m=30 diag-quadratic", "lbfgs", fl.OBJ_DIAGQUAD, fl.START_ZERO, 0, dict(Memory=30)),
             ("CG DY quartic", "cg", fl.OBJ_QUARTIC, fl.START_QUARTIC_U, 12345, dict(Method="DY")),
             ("CG PR quartic", "cg", fl.OBJ_QUARTIC, fl.START_QUARTIC_U, 12345, dict(Method="PR")),
             ("CG DY quartic, plain callbacks", "cg", fl.OBJ_QUARTIC, fl.START_QUARTIC_U, 12345, dict(Method="DY", fused=False)),
             ("SteepestDescent Rosenbrock", "sd", fl.OBJ_ROSENBROCK, fl.START_ROSEN_PERT, 7, dict())]
    for log2n in range(a.min_log2n, a.max_log2n + 1):
        n = 1 << log2n
        for label, algo, kind, start, seed, kw in cases:
            if any(t and t in label for t in a.skip.split(",")):
                continue
            if a.line_search != "reference":
                kw = dict(kw, line_search=a.line_search)
                label += f", line_search={a.line_search}"
            mem = kw.get("Memory", 0)
            if (2 * mem + 6) * 8 * n / WORLD > 170e9:
                if RANK == 0:
                    rows.append((log2n, label, None))
                    print(f"2^{log2n} {label}: skipped, {(2 * mem + 6) * 8 * n / WORLD / 2**30:.0f} GiB per GPU do not fit", flush=True)
                continue
            r = one(algo, kind, start, seed, n, a.steps, a.warmup, **kw)
            if RANK != 0:
                continue
            if r is None:
                print(f"2^{log2n} {label}: converged before the timed window", flush=True)
                continue
            r["memory"], r["algo"] = mem, algo
            rows.append((log2n, label, r))
            print(f"2^{log2n} {label:48.48s} {r['it_per_s']:8.2f} it/s  {r['trials_per_it']:5.1f} trials/it  "
                  f"{r['GB_per_it']:7.1f} GB/it  {r['GBps']:7.0f} GB/s ({r['frac']:.0%} of measured)", flush=True)
    if dist is not None:
        fl.lib().flgpu_comm_destroy(COMM)
        dist.destroy_process_group()
    if RANK != 0:
        return
    with open(a.out, "w") as fh:
        fh.write("| n | workload | it/s (device time, max over ranks) | trials/it | algorithmic GB/it | kernel ms/it | achieved GB/s | of measured 6467.7 |\n")
        fh.write("|---|---|---:|---:|---:|---:|---:|---:|\n")
        for log2n, label, r in rows:
            if r is None:
                fh.write(f"| 2^{log2n} | {label} | skipped: work space exceeds 180 GB per GPU | | | | | |\n")
                continue
            fh.write(f"| 2^{log2n} | {label} | {r['it_per_s']:.2f} | {r['trials_per_it']:.1f} | {r['GB_per_it']:.1f} | "
                     f"{r['kernel_ms_per_it']:.2f} | {r['GBps']:.0f} | {r['frac']:.1%} |\n")
    if a.json:
        json.dump({"gpus": WORLD, "line_search": a.line_search, "steps": a.steps, "warmup": a.warmup,
                   "rows": [{"log2n": l, "workload": w, **(r or {"skipped": True})} for l, w, r in rows]},
                  open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
