#!/usr/bin/env python
"""bench.py -- the reference's headline metric on its headline config (BASELINE.json):
L-BFGS (m = 10) outer iterations per second on extended Rosenbrock at n = 2^28, fp64, row-sharded over
--gpus N B200s (strong scaling: the global n is fixed), plus the HBM roofline of the dominant kernel.

A "step" is one main-loop L-BFGS iteration = Before + line search + After of the reference
(NonlinearOptimization.f90:514-518), i.e. one accepted step: K1 (accepted point, ring update and all dots in one
pass) -> K2 -> K3 (direction) -> Strong-Wolfe trials (fused probes of x0 + a p).  Iteration 0 and the
m-1 pre-iterations (f90:442-510) always run first and are never timed.

  python bench.py [--gpus N --steps K --warmup W]         our arm (one JSON line on rank 0)
  python bench.py --impl reference [...]                   the reference's CPU algorithm (oracle port): one thread
                                                           (its own semantics) and, as the line's value, all host cores

Beside the headline the line carries (same run, same box):
  secondary   BASELINE.json configs[2] (CG Dai-Yuan / Polak-Ribiere+ on the separable quartic, n = 2^28) and configs[3]
              (LBFGS m = 30 on the diagonal quadratic, 2^28 rows per GPU: n = 2^31 on 8 GPUs -- the weak-scaling series)
  parity      (N > 1) a 2^20-row L-BFGS run on the shards against the same problem on rank 0 alone, before any timing

Timing: CUDA events on the library's stream, barrier + synchronize on both sides, max over ranks.
Inputs are 2 GiB per vector (>> 126 MB L2), so no L2 flush is needed between iterations.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

# rank 0 prints exactly ONE line on stdout (the JSON); NCCL's banner ("NCCL version ...", printed on stdout when
# NCCL_DEBUG is VERSION or WARN) and any NCCL warning go to stderr instead
# (NCCL honours NCCL_DEBUG_FILE only above the VERSION level, so VERSION is raised to WARN)
if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION", "WARN"):
    os.environ["NCCL_DEBUG"] = "WARN"
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

LOG2_N = 28
MEM = 10
SEED = 7
METRIC = "lbfgs_iterations_per_sec"
UNIT = "it/s"


def workload(n, mem, objective="rosenbrock"):
    if objective == "diag":
        return (f"LBFGS m={mem} diagonal-scaled quadratic (condition 1e6) n=2^{n.bit_length() - 1} fp64, start x=0, "
                f"f_fd present, default tunables (Strong, c1=1e-4, c2=0.9, Increment=1.05)")
    return (f"LBFGS m={mem} extended Rosenbrock n=2^{n.bit_length() - 1} fp64, start R1 = (-1.2,1)+0.1(u-0.5) "
            f"seed {SEED}, f_fd present, default tunables (Strong, c1=1e-4, c2=0.9, Increment=1.05)")


def static_config(n, mem, objective):
    """What is measured -- IDENTICAL in our arm and in the reference arm (everything about HOW a run went lives in
    the line's `run` object instead)."""
    return {"workload": workload(n, mem, objective), "n_global": n, "memory": mem, "objective": objective,
            "l2": "inputs (2 GiB/vector) exceed L2; no flush needed"}


POLICIES = {"reference": "reference (StrongWolfe_fdwithf f90:1582-1698, statement by statement)",
            "fast": "fast (FLGPU_LS_FAST: first trial satisfying the strong Wolfe conditions is accepted; not a "
                    "reference routine)"}
LS_MODES = {True: "fused (flgpu_fused_fn / flgpu_fused_multi_fn + flgpu_update_fn + flgpu_direction_fn: the objective kernel "
                  "forms x0+a*p; 2n doubles per PASS, a pass evaluating up to four trials of a bracketing walk; the first four "
                  "trials of a search are evaluated by K3 while it writes p; the accepted point is formed and stored by K1)",
            False: "plain (opaque f/fd/f_fd device callbacks; 7n doubles per f+g trial)"}
NCU_NAMES = {"k1_update_dots": "k1_update_dots_kernel", "k1_update_dots_fused": "k1_update_dots_kernel",
             "k3_direction": "k3_direction_tma_kernel", "k3_direction_probe": "k3_direction_tma_kernel",
             "trial_x": "trial_kernel",
             "dot": "dot_kernel", "cg_dots": "cg_dots_kernel", "cg_update": "cg_update_kernel"}


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        """Clocks and throttle reasons sampled inside [t0, t1] (the timed region); when the region is shorter than a
        few sampling periods (multi-GPU runs: ~0.1 s) the window is widened to the 0.5 s of load that precede it."""
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        inside = [r for r in self.rows if t0 - 0.02 <= r[0] <= t1 + 0.05]
        if len(inside) < 3:
            inside = [r for r in self.rows if t0 - 0.5 <= r[0] <= t1 + 0.1]
        sm, mx, reasons = [], [], set()
        for t, line in inside:
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------- CPU baseline (oracle = checker, timed)
def cpu_lbfgs(n_sample, mem, warmup, steps, n_target, objective="rosenbrock"):
    """Times the oracle (C restatement of the reference, strict IEEE, sequential sums, ONE thread -- the
    reference has no threading on this path) on a bounded sample of the same workload.  The workload is
    streaming, so rates scale with 1/n: `value` is the sample's rate x n_sample/n_target and says so
    (`extrapolated`); `ms_per_trial_at_n` is the size-normalised figure that does not depend on how many trials the
    timed iterations happened to need."""
    import _oracle as O
    marks = {}

    class T(O.Trace):
        def _on(self, user, it, dim, p, x, g, a, fx, phid0, trials):
            marks[it] = (time.perf_counter(), trials)

    obj, start = (O.OBJ_DIAGQUAD, O.START_ZERO) if objective == "diag" else (O.OBJ_ROSENBROCK, O.START_ROSEN_PERT)
    x0 = O.start_vector(start, n_sample, seed=SEED)
    tr = T(keep_vectors=False)
    x, st = O.lbfgs(O.builtin_callbacks(obj, 0, n_sample), x0, Memory=mem, use_ffd=True, Warning=False,
                    MaxIteration=warmup + steps, trace=tr)
    first, last = mem + warmup - 1, mem + warmup + steps - 1
    if last not in marks:                      # converged early (never at these sizes)
        last = max(marks)
    dt = marks[last][0] - marks[first][0]
    its = last - first
    trials = sum(marks[i][1] for i in range(first + 1, last + 1))
    rate_sample = its / dt
    scale = n_sample / n_target
    threads = int(O.lib().orc_threads())
    build = ("oracle/liboracle_omp.so (the same source with OpenMP-parallel loops and dots, gcc -O3 -march=native "
             f"-fopenmp, {threads} threads: a GENEROUS baseline, the reference itself is serial)" if O.OMP_VARIANT else
             "oracle/liboracle.so (gcc -O2 -ffp-contract=off, 1 thread = the reference's semantics)")
    return {"value": rate_sample * scale, "unit": UNIT, "cores": threads, "kind": "port",
            "extrapolated": n_sample != n_target, "n_sample": n_sample, "n_target": n_target,
            "sample": (f"{build} on n=2^{n_sample.bit_length() - 1}: "
                       f"{its} main-loop iterations, {trials} trials, {dt:.2f} s = {rate_sample:.3f} it/s; scaled by "
                       f"n_sample/n (streaming)"),
            "sample_it_per_s": rate_sample, "sample_trials_per_iteration": trials / max(its, 1),
            "trials_per_sec": trials / dt * scale, "ms_per_trial_at_n": dt / max(trials, 1) / scale * 1e3,
            "sample_seconds": dt}


def cpu_lbfgs_all_cores(n_sample, mem, warmup, steps, n_target, objective="rosenbrock"):
    """The generous CPU row of BASELINE.md section 3: the oracle's OpenMP build on every host core, in a subprocess
    (the in-process oracle is the strict single-thread library).  None if that build is not possible here."""
    env = dict(os.environ, FLGPU_ORACLE_VARIANT="omp")
    env.pop("OMP_NUM_THREADS", None)           # torchrun sets it to 1; this row is about all the cores
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "cpu-sample", "--cpu-log2n", str(n_sample.bit_length() - 1),
           "--log2n", str(n_target.bit_length() - 1), "--mem", str(mem), "--warmup", str(warmup), "--steps", str(steps),
           "--objective", objective]
    try:
        r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
        return json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else None
    except (subprocess.SubprocessError, ValueError, IndexError):
        return None


def run_cpu_sample(args):
    """bench.py --impl cpu-sample: one cpu_lbfgs() record as JSON (helper of cpu_lbfgs_all_cores)."""
    n = 1 << args.log2n
    print(json.dumps(cpu_lbfgs(1 << min(args.log2n, args.cpu_log2n), args.mem, args.warmup, args.steps, n,
                               args.objective)), flush=True)
    return 0


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n = 1 << args.log2n
    n_sample = 1 << min(args.log2n, args.cpu_log2n)
    t0 = time.time()
    # the reference's algorithm on ONE thread (its own semantics: it has no threading on this path) ...
    single = cpu_lbfgs(n_sample, args.mem, args.warmup, args.steps, n, args.objective)
    # ... and on all host cores (OpenMP build of the same port, a larger sample).  The line's value is the faster,
    # all-cores one, so the driver's ours/reference ratio is the conservative reading; the single-thread figure -- what
    # the reference itself would do -- is `value_reference_semantics`.
    n_omp = 1 << min(args.log2n, args.cpu_log2n + 2)
    cb = cpu_lbfgs_all_cores(n_omp, args.mem, args.warmup, args.steps, n, args.objective) or single
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / cb["value"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": static_config(n, args.mem, args.objective),
            "kind": "port", "extrapolated": cb["extrapolated"], "n_sample": cb["n_sample"],
            "value_reference_semantics": single["value"],
            "trials_per_sec": cb["trials_per_sec"], "ms_per_trial": cb["ms_per_trial_at_n"],
            "run": {"timing": "host perf_counter around oracle iterations; rates measured on n_sample rows and scaled by "
                              "n_sample/n (streaming workload): ms_per_step is DERIVED (1e3/value), not measured at n",
                    "sample_seconds": cb["sample_seconds"], "trials_per_iteration": cb["sample_trials_per_iteration"],
                    "note": "the reference cannot be compiled here (Fortran + MKL, SURVEY F1/F2): this is the C "
                            "restatement oracle/oracle.c, parity-unpinned"},
            "cpu_baseline": cb, "single_thread": single,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.time() - t0}
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------- our arm
def shard_bounds(n, rank, world):
    lo = (n * rank // world) // 2 * 2            # even boundaries: Rosenbrock pairs never straddle shards
    hi = n if rank == world - 1 else (n * (rank + 1) // world) // 2 * 2
    return lo, hi


def load_peak():
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    return peak, ("MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback)")


def kernel_table(kt, peak):
    kernels = {}
    for name, v in kt.items():
        if v["ms"] > 0 and v["bytes"] > 0:
            kernels[name] = {"launches": v["launches"], "ms": round(v["ms"], 3), "avg_ms": v["ms"] / v["launches"],
                             "GBps": v["bytes"] / v["ms"] / 1e6, "frac": v["bytes"] / v["ms"] / 1e6 / peak}
    return kernels


def roofline_of(kt, peak, peak_src, n_local, mem):
    """Roofline object of the dominant LIBRARY kernel of the timed region (CUDA-event times of pass 2)."""
    kernels = kernel_table(kt, peak)
    own = {k: v for k, v in kernels.items() if not k.startswith("callback:")}
    top = max(own, key=lambda k: own[k]["ms"])
    # DRAM bytes per launch from the committed ncu --set full capture of the same kernel; the capture was taken at one
    # size (recorded beside it), so it is scaled by rows per GPU and only used when the memory matches
    traffic, traffic_note = None, None
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        ent = prof.get(NCU_NAMES.get(top, top), {})
        if ent.get("dram_bytes_per_launch") and ent.get("n") and ent.get("memory") == mem:
            traffic = ent["dram_bytes_per_launch"] * n_local / ent["n"]
            traffic_note = (f"ncu --set full capture {ent.get('tag')} at n=2^{int(ent['n']).bit_length() - 1}, "
                            f"m={ent['memory']}" + ("" if n_local == ent["n"] else ", scaled by rows per GPU"))
    except (OSError, ValueError):
        pass
    total_bytes = sum(v["bytes"] for v in kt.values())
    total_kernel_ms = sum(v["ms"] for v in kt.values())
    return {"bound": "hbm", "kernel": top, "achieved": own[top]["GBps"], "peak": peak, "unit": "GB/s",
            "frac": own[top]["frac"], "traffic": traffic, "traffic_source": traffic_note, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": kt[top]["bytes"] / kt[top]["launches"],
            "share_of_step": kt[top]["ms"] / total_kernel_ms,
            "whole_step": {"algorithmic_GB": total_bytes / 1e9, "kernel_ms": total_kernel_ms,
                           "GBps": total_bytes / total_kernel_ms / 1e6, "frac": total_bytes / total_kernel_ms / 1e6 / peak,
                           "note": "all kernels of the K timed iterations (per-kernel CUDA events)"},
            "kernels": kernels}


def run_ours(args):
    import numpy as np
    import torch
    import fortran_library_b200 as fl

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    fl.require_gpu()
    dist = None
    comm = comm_nccl = None
    L = fl.lib()
    L.flgpu_comm_uses_peer_memory.argtypes = [C.c_void_p]
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

        def bcast(data):
            t = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                t.copy_(torch.frombuffer(bytearray(data), dtype=torch.uint8))
            dist.broadcast(t, 0)
            return bytes(t.cpu().numpy().tobytes())
        comm = fl.comm_create(rank, world, bcast)
        os.environ["FLGPU_EXCHANGE"] = "nccl"            # a second communicator on the ncclAllGather fallback (parity)
        comm_nccl = fl.comm_create(rank, world, bcast)
        del os.environ["FLGPU_EXCHANGE"]

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    def allmax(v):
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allmin_flag(ok):
        t = torch.tensor([int(bool(ok))], device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    n = 1 << args.log2n
    mem, W, K = args.mem, args.warmup, args.steps
    lo, hi = shard_bounds(n, rank, world)
    n_local = hi - lo
    peak, peak_src = load_peak()
    DS = {"auto": None, "on": True, "off": False}[args.device_search]
    OBJS = {"rosenbrock": (fl.OBJ_ROSENBROCK, fl.START_ROSEN_PERT, SEED), "diag": (fl.OBJ_DIAGQUAD, fl.START_ZERO, 0),
            "quartic": (fl.OBJ_QUARTIC, fl.START_QUARTIC_U, 12345)}

    def timed_run(algo, objective, n_glob, m, time_kernels, fused=True, policy="reference", method=None):
        """K iterations of `algo` bracketed by CUDA events on the library's stream (observer callbacks mark the
        window): LBFGS skips iteration 0 and the m-1 pre-iterations, CG its first W iterations."""
        kind, start, seed = OBJS[objective]
        a, b = shard_bounds(n_glob, rank, world)
        first = max((m if algo == "lbfgs" else 0) + W - 1, 0)
        last = first + K
        x = fl.DeviceVector.start(start, b - a, seed=seed, offset=a, n_global=n_glob)
        ev = [torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)]
        mark = {}
        per = mark["per"] = []          # one event per iteration boundary inside the window (median)

        def stamp(i):
            try:
                e = torch.cuda.Event(enable_timing=True)
                e.record(torch.cuda.ExternalStream(i.stream))
                per.append(e)
            except Exception:           # the per-iteration figures are a by-product: never let them break the metric
                mark["per"] = None

        def on_iter(i):
            if first < i.iteration < last and mark["per"] is not None:
                stamp(i)
            if i.iteration == first or i.iteration == last:
                s = torch.cuda.ExternalStream(i.stream)
                if i.iteration == first:
                    barrier()
                    mark["t0"] = time.time()
                    mark["c0"] = (i.gpu_launches, i.callbacks, i.total_trials)
                    if time_kernels:
                        fl.lib().flgpu_reset_kernel_times()
                    ev[0].record(s)
                    if mark["per"] is not None:
                        stamp(i)
                else:
                    if mark["per"] is not None:
                        stamp(i)
                    ev[1].record(s)
                    barrier()
                    mark["t1"] = time.time()
                    mark["c1"] = (i.gpu_launches, i.callbacks, i.total_trials)
                    return True
            return False
        ob = fl.Observer(on_iteration=on_iter)
        prob = fl.builtin_problem(kind)
        common = dict(Warning=False, MaxIteration=W + K + 1, observer=ob, comm=comm, offset=a, n_global=n_glob,
                      time_kernels=time_kernels, fused=fused, device_search=DS, line_search=policy)
        if algo == "lbfgs":
            st = fl.LBFGS(prob, x, Memory=m, **common)
        else:
            st = fl.ConjugateGradient(prob, x, Method=method, **common)
        if "t1" not in mark:
            raise SystemExit(f"bench.py: {algo} on {objective} stopped after {st.iterations} iterations (status "
                             f"{st.status}) before {last + 1}; lower --steps")
        ms = ev[0].elapsed_time(ev[1])
        try:
            evs = mark.get("per") or []
            mark["iter_ms"] = sorted(evs[k].elapsed_time(evs[k + 1]) for k in range(len(evs) - 1))
        except Exception:
            mark["iter_ms"] = []
        x.free()
        return ms, mark, st, fl.kernel_times() if time_kernels else None

    def record(algo, objective, n_glob, m, method=None):
        """A secondary record: one pass with per-kernel events -> rate, trials and the roofline of its top kernel."""
        ms, mark, st, kt = timed_run(algo, objective, n_glob, m, True, True, "reference", method)
        ms = allmax(ms)
        trials = mark["c1"][2] - mark["c0"][2]
        a, b = shard_bounds(n_glob, rank, world)
        rf = roofline_of(kt, peak, peak_src, b - a, m)
        return {"workload": (f"{'LBFGS m=' + str(m) if algo == 'lbfgs' else 'ConjugateGradient ' + method} on {objective}, "
                             f"n=2^{n_glob.bit_length() - 1} ({b - a} rows per GPU), fused line search, reference policy"),
                "value": K / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / K, "trials_per_iteration": trials / K,
                "trials_per_sec": trials / (ms * 1e-3),
                "roofline": {k: rf[k] for k in ("kernel", "achieved", "peak", "frac", "share_of_step", "unit")},
                "whole_step_GBps": rf["whole_step"]["GBps"], "whole_step_frac": rf["whole_step"]["frac"],
                "kernels": {k: {"avg_ms": v["avg_ms"], "GBps": v["GBps"], "frac": v["frac"], "launches": v["launches"]}
                            for k, v in rf["kernels"].items()}}

    # ---- parity (N > 1), before any timing: a 2^20-row problem on the shards against rank 0 alone.  The reductions are
    # partition-independent (include/flgpu_reduce.cuh) and these shards are aligned subtrees, so the comparison is
    # BITWISE: every rank's scalars, the gathered iterate and the first directions must equal the single-GPU run's.
    parity = None
    if world > 1:
        parity = {}
        n_par = 1 << args.parity_log2n
        pa, pb = shard_bounds(n_par, rank, world)

        def gather(v):
            t = torch.from_numpy(np.ascontiguousarray(v)).cuda()
            outs = []
            for r in range(world):
                ra, rb = shard_bounds(n_par, r, world)
                buf = t if r == rank else torch.empty(rb - ra, dtype=torch.float64, device="cuda")
                dist.broadcast(buf, r)
                outs.append(buf.clone())
            return torch.cat(outs).cpu().numpy()

        def sharded(policy, c, ds=None):
            x = fl.DeviceVector.start(fl.START_ROSEN_PERT, pb - pa, seed=SEED, offset=pa, n_global=n_par)
            ob = fl.Observer(keep_vectors=True, max_vec_iters=6)
            st = fl.LBFGS(fl.builtin_problem(fl.OBJ_ROSENBROCK), x, Memory=mem, Warning=False, MaxIteration=max(25 - mem, 2),
                          observer=ob, comm=c, offset=pa, n_global=n_par, line_search=policy, device_search=ds)
            out = (x.numpy(), ob, st)
            x.free()
            return out

        for policy in ("reference", "fast"):
            xs, ob, st = sharded(policy, comm, ds=True)      # device-resident search, in-kernel exchange
            mine = torch.tensor([v for r in ob.rows for v in (r[1], r[2], r[3], float(r[4]))] + [float(st.iterations)],
                                dtype=torch.float64, device="cuda")
            ref = mine.clone()
            dist.broadcast(ref, 0)
            same = allmin_flag(ref.numel() == mine.numel() and torch.equal(ref.view(torch.int64), mine.view(torch.int64)))
            x2, _, st2 = sharded(policy, comm_nccl)
            x3, _, st3 = sharded(policy, comm, ds=False)     # host-driven search over the same exchange
            modes = allmin_flag(np.array_equal(xs, x2) and np.array_equal(xs, x3) and st2.iterations == st.iterations
                                and st3.iterations == st.iterations)
            xg = gather(xs)
            pg = [gather(p) for p in ob.p[:6]]
            rec = {"ranks_identical_scalars": same, "device_search==nccl_fallback==host_driven_search": modes,
                   "iterations": st.iterations}
            if rank == 0:
                x1 = fl.DeviceVector.start(fl.START_ROSEN_PERT, n_par, seed=SEED)
                ob1 = fl.Observer(keep_vectors=True, max_vec_iters=6)
                st1 = fl.LBFGS(fl.builtin_problem(fl.OBJ_ROSENBROCK), x1, Memory=mem, Warning=False,
                               MaxIteration=max(25 - mem, 2), observer=ob1, line_search=policy)
                x1n = x1.numpy()
                rec["bitwise_equal_to_single_gpu"] = bool(np.array_equal(xg, x1n) and ob1.rows == ob.rows and
                                                          all(np.array_equal(a, b) for a, b in zip(pg, ob1.p[:6])))
                rec["rel_dx"] = float(np.linalg.norm(xg - x1n) / np.linalg.norm(x1n))
                rec["max_rel_dp_first6"] = float(max(np.linalg.norm(a - b) / np.linalg.norm(b) for a, b in zip(pg, ob1.p[:6])))
                rec["iterations_single_gpu"] = st1.iterations
                rec["ok"] = bool(same and modes and rec["rel_dx"] < 1e-8 and rec["max_rel_dp_first6"] < 1e-9
                                 and st1.iterations == st.iterations)
                x1.free()
            parity[policy] = rec
        parity["n_global"] = n_par
        parity["what"] = (f"LBFGS m={mem}, Rosenbrock R1, n=2^{args.parity_log2n} over {world} row shards vs rank 0 alone, "
                          "25 accepted steps; asserted before timing")
        ok = allmin_flag(rank != 0 or all(parity[p]["ok"] for p in ("reference", "fast")))
        parity["ok"] = ok

    # ---- pass 1: the metric (device-resident inputs, no per-kernel events)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    fused = not args.plain
    ds_active = fused and (DS is True or (DS is None and n_local <= (1 << 18))) and \
        (comm is None or bool(L.flgpu_comm_uses_peer_memory(comm)))
    ms, mark, st, _ = timed_run("lbfgs", args.objective, n, mem, False, fused, args.line_search)
    clocks = sampler.stop(mark["t0"], mark["t1"]) if rank == 0 else None
    ms_max = allmax(ms)
    launches = (mark["c1"][0] - mark["c0"][0]) + (mark["c1"][1] - mark["c0"][1])   # library kernels + objective kernels
    trials = mark["c1"][2] - mark["c0"][2]
    it_ms = mark.get("iter_ms") or []
    per_iteration = ({"median_ms": it_ms[len(it_ms) // 2], "min_ms": it_ms[0], "max_ms": it_ms[-1], "n": len(it_ms),
                      "note": "this rank's CUDA-event time between consecutive accepted steps inside the timed window "
                              "(iterations differ by their trial counts)"} if it_ms else None)

    # ---- pass 2: per-kernel CUDA-event times over the same timed region -> roofline of the dominant kernel
    ms2, mark2, st2, kt = timed_run("lbfgs", args.objective, n, mem, True, fused, args.line_search)
    roofline = roofline_of(kt, peak, peak_src, n_local, mem)
    # ---- pass 3: the other line-search mode, for the record (same iterates up to summation order)
    ms3, mark3, _, _ = timed_run("lbfgs", args.objective, n, mem, False, not fused, args.line_search)
    other = {"line_search": LS_MODES[not fused], "value": K / (allmax(ms3) * 1e-3), "unit": UNIT,
             "trials_in_timed_region": mark3["c1"][2] - mark3["c0"][2]}
    # ---- pass 4: the optional FLGPU_LS_FAST policy (NOT the reference's searcher: iterates differ, so this is a
    # separate record and never the headline)
    fast = None
    if args.line_search == "reference":
        ms4, mark4, _, _ = timed_run("lbfgs", args.objective, n, mem, False, fused, "fast")
        t4 = allmax(ms4)
        tr4 = mark4["c1"][2] - mark4["c0"][2]
        fast = {"line_search_policy": POLICIES["fast"], "value": K / (t4 * 1e-3), "unit": UNIT,
                "ms_per_step": t4 / K, "trials_in_timed_region": tr4, "trials_per_iteration": tr4 / K,
                "note": "same K-iteration window of a run made with line_search=fast; iteration rate, not time to "
                        "solution (DESIGN.md 3.2: with WolfeConst2 = 0.9 this policy needs more iterations on Rosenbrock)"}

    # ---- secondary records: BASELINE.json configs[2] and [3], driver-run beside the headline
    secondary = {}
    if not args.no_secondary and args.objective == "rosenbrock" and args.line_search == "reference" and fused:
        secondary["cg_dy_quartic"] = record("cg", "quartic", n, 0, "DY")
        secondary["cg_pr_quartic"] = record("cg", "quartic", n, 0, "PR")
        # configs[3]: 2^28 rows per GPU (130 GiB of work space) -- the weak-scaling series 2^28 x N, n = 2^31 on 8 GPUs
        n30 = (1 << 28) * world
        secondary[f"lbfgs_m30_diag_2p{n30.bit_length() - 1}"] = dict(
            record("lbfgs", "diag", n30, 30), scaling="weak: 2^28 rows per GPU at every N (n = 2^31 on 8 GPUs)")

    # ---- e2e: the reference-facing call with HOST buffers (pinned): H2D of x, the whole optimisation (iteration 0,
    # the m-1 pre-iterations, Ke main iterations) and D2H of x inside the timed region.  Called twice: cold (work space
    # allocated and returned to the driver inside the call, as the reference does) and warm (flgpu_set_workspace_cache:
    # buffers parked by an untimed warm-up call are reused) -- `value` is the warm call, the cold one is reported beside it.
    Ke = args.e2e_steps
    OBJ, START, _ = OBJS[args.objective]
    prob = fl.builtin_problem(OBJ)
    xh = torch.empty(n_local, dtype=torch.float64).pin_memory()
    if world == 1 and not fused:
        os.environ["FLGPU_NO_FUSED"] = "1"
    L.flgpu_set_line_search(fl.LS_FAST if args.line_search == "fast" else fl.LS_REFERENCE)   # Fortran-ABI calls
    ref_cbs = (fl.capi.REF_F_FN(), fl.capi.REF_FD_FN(), fl.capi.REF_F_FD_FN())
    L.flgpu_builtin_ref_callbacks(OBJ, C.byref(ref_cbs[0]), C.byref(ref_cbs[1]), C.byref(ref_cbs[2]))

    def e2e_call(maxit):
        x0 = fl.DeviceVector.start(START, n_local, seed=SEED, offset=lo, n_global=n)
        L.flgpu_memcpy(xh.data_ptr(), x0.ptr, n_local * 8, fl.SPACE_HOST, fl.SPACE_DEVICE, None)
        x0.free()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        te = time.time()
        if world == 1:
            L.__getattr__("__nonlinearoptimization_MOD_lbfgs")(
                ref_cbs[0], ref_cbs[1], C.c_void_p(xh.data_ptr()), C.byref(C.c_int(n_local)), C.byref(C.c_int(mem)),
                ref_cbs[2], None, C.byref(C.c_int32(0)), C.byref(C.c_int(maxit)), None, None, None, None, None)
            ste = fl.capi.Stats()
            L.flgpu_last_stats(C.byref(ste))
        else:
            ste = fl.LBFGS(prob, xh, Memory=mem, Warning=False, MaxIteration=maxit, comm=comm, offset=lo, n_global=n,
                           fused=fused, line_search=args.line_search)
        e1.record()
        barrier()
        return allmax(max(e0.elapsed_time(e1), (time.time() - te) * 1e3)), ste

    cold_ms, cold_st = e2e_call(Ke)
    L.flgpu_set_workspace_cache(1)
    e2e_call(0)                                  # untimed warm-up: parks the work space
    e2e_ms, ste = e2e_call(Ke)
    L.flgpu_set_workspace_cache(0)               # releases the parked buffers
    e2e = {"value": ste.iterations / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 8 * n / ste.iterations,
           "d2h_bytes_per_step": 8 * n / ste.iterations, "iterations": ste.iterations, "ms": e2e_ms,
           "trials": ste.n_trials, "trials_per_sec": ste.n_trials / (e2e_ms * 1e-3),
           "call": ("__nonlinearoptimization_MOD_lbfgs (host x, built-in CUDA objective in reference-ABI form)" if world == 1
                    else "flgpu_lbfgs (host x shard, row-shard communicator)"),
           "cold_call": {"value": cold_st.iterations / (cold_ms * 1e-3), "ms": cold_ms,
                         "note": "first call: work space cudaMalloc'ed and cudaFree'd inside the call"},
           "note": "one optimizer call incl. H2D of x, iteration 0 + m-1 pre-iterations (never part of `value` above: "
                   f"their first line searches take 100-200 trials) + {Ke} main iterations, D2H of x; work space reused "
                   "from an untimed warm-up call (flgpu_set_workspace_cache); bytes/step = 8n/iterations (x crosses "
                   "PCIe once per call)"}

    cpu = cpu_all = None
    if rank == 0 and world == 1 and not args.no_cpu and args.objective == "rosenbrock":
        cpu = cpu_lbfgs(1 << min(args.log2n, args.cpu_log2n), mem, min(W, 3), min(K, 10), n)
        cpu_all = cpu_lbfgs_all_cores(1 << min(args.log2n, args.cpu_log2n + 2), mem, min(W, 3), min(K, 10), n)

    rc = 0
    if rank == 0:
        line = {"metric": METRIC, "value": K / (ms_max * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": static_config(n, mem, args.objective),
                "hbm_GBps_whole_step": roofline["whole_step"]["GBps"], "hbm_frac_whole_step": roofline["whole_step"]["frac"],
                "trials_per_sec": trials / (ms_max * 1e-3), "ms_per_trial": ms_max / max(trials, 1),
                "run": {"rows_per_gpu": n_local, "parallelism": f"row-shard x{world}" if world > 1 else "1 GPU",
                        "exchange": (None if comm is None else
                                     ("one kernel over IPC-mapped peer memory (NVLink stores + flags), rank tree"
                                      if L.flgpu_comm_uses_peer_memory(comm) else "ncclAllGather + combine kernel")),
                        "reductions": "partition-independent: fixed chunks on global indices + aligned binary tree over "
                                      "chunks and ranks (include/flgpu_reduce.cuh)",
                        "line_search": LS_MODES[fused], "line_search_policy": POLICIES[args.line_search],
                        "device_resident_search": (f"{args.device_search}: " + (
                            "ON (one cooperative kernel per line search, flgpu_search_fn)" if ds_active else
                            "off at this size (host-driven, one round trip per trial)")),
                        "trials_in_timed_region": trials, "trials_per_iteration": trials / K,
                        "bytes_note": ("fused mode: every trial of the reference is evaluated (same points, same values), but up "
                                       "to four share one 2n-double pass over x0 and p and the first four of a search ride on "
                                       "K3, so the step moves (4m+7+2P)n doubles with P = probe passes per iteration -- NOT "
                                       "(4m+6)n + 2n per trial; charging 2n per trial would put this line above the HBM peak. "
                                       "`roofline.whole_step` and `hbm_GBps_whole_step` count the bytes of the passes actually "
                                       "made; `roofline.traffic` / profiles/r02m_kernels.md hold ncu's DRAM counters") if fused else None,
                        "objective_kernel_launches_per_iteration": (mark["c1"][1] - mark["c0"][1]) / K,
                        "batching": "the trials of a bracketing walk (a, a*Increment, ...) share one pass over x0 and p, four "
                                    "at a time; K3 evaluates the first four of every search (flgpu_fused_multi_fn, "
                                    "flgpu_direction_fn): same trial points, decisions, counts and bits, fewer passes"
                                    if fused else None},
                "per_iteration": per_iteration,
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
                "cpu_baseline_all_cores": cpu_all, "parity": parity, "secondary": secondary,
                "other_line_search_mode": other, "fast_line_search_policy": fast}
        print(json.dumps(line), flush=True)
        if parity is not None and not parity["ok"]:
            print("bench.py: PARITY FAILED -- the sharded run does not reproduce the single-GPU run", file=sys.stderr)
            rc = 3
    if comm is not None:
        L.flgpu_comm_destroy(comm)
        L.flgpu_comm_destroy(comm_nccl)
    if dist is not None:
        dist.destroy_process_group()
    return rc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "cpu-sample"])
    ap.add_argument("--log2n", type=int, default=LOG2_N, help="override the global dimension (debugging only)")
    ap.add_argument("--mem", type=int, default=MEM)
    ap.add_argument("--objective", default="rosenbrock", choices=["rosenbrock", "diag"],
                    help="diag = BASELINE.json configs[3] as the headline (the default run already records it under `secondary`)")
    ap.add_argument("--e2e-steps", type=int, default=100,
                    help="main-loop iterations of the end-to-end optimizer call (its 487-trial prologue is amortised over them)")
    ap.add_argument("--device-search", default="auto", choices=["auto", "on", "off"],
                    help="flgpu_options.device_search; auto = up to 2^18 rows per GPU (same bits either way)")
    ap.add_argument("--line-search", default="reference", choices=["reference", "fast"],
                    help="flgpu_options.line_search for the whole run; the headline is `reference` (the default run "
                         "also records the `fast` rate beside it)")
    ap.add_argument("--plain", action="store_true", help="headline with opaque callbacks (no fused line-search evaluation)")
    ap.add_argument("--cpu-log2n", type=int, default=None, help="size of the bounded single-thread CPU sample (all cores: x4)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the configs[2]/[3] records")
    ap.add_argument("--parity-log2n", type=int, default=20, help="global rows of the N > 1 parity problem")
    args = ap.parse_args()
    if args.cpu_log2n is None:       # bounded CPU sample: ~40 s of one core for the reference arm, ~5 s inside our arm
        args.cpu_log2n = 23 if args.impl == "reference" else 22
    if args.warmup < 0 or args.steps < 1:
        raise SystemExit("need --steps >= 1 and --warmup >= 0")
    if args.impl == "cpu-sample":
        return run_cpu_sample(args)
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
