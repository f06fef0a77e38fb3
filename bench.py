#!/usr/bin/env python
"""bench.py -- the reference's headline metric on its headline config (BASELINE.json):
L-BFGS (m = 10) outer iterations per second on extended Rosenbrock at n = 2^28, fp64, row-sharded over
--gpus N B200s (strong scaling: the global n is fixed), plus the HBM roofline of the dominant kernel.

A "step" is one main-loop L-BFGS iteration = Before + line search + After of the reference
(NonlinearOptimization.f90:514-518), i.e. one accepted step: K1 (ring update + all dots) -> K2 -> K3
(direction + first trial) -> Strong-Wolfe trials (x0 + a p, f_fd callback, f'.p).  Iteration 0 and the
m-1 pre-iterations (f90:442-510) always run first and are never timed.

  python bench.py [--gpus N --steps K --warmup W]         our arm (one JSON line on rank 0)
  python bench.py --impl reference [...]                   the reference's CPU algorithm (oracle port): one thread
                                                           (its own semantics) and, as the line's value, all host cores

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


POLICIES = {"reference": "reference (StrongWolfe_fdwithf f90:1582-1698, statement by statement)",
            "fast": "fast (FLGPU_LS_FAST: first trial satisfying the strong Wolfe conditions is accepted; not a "
                    "reference routine)"}
LS_MODES = {True: "fused (flgpu_fused_fn: objective kernel forms x0+a*p; 2n doubles per trial + 4n per accepted step)",
            False: "plain (opaque f/fd/f_fd device callbacks; 7n doubles per f+g trial)"}
NCU_NAMES = {"k1_update_dots": "k1_update_dots_kernel", "k3_direction": "k3_direction_kernel", "trial_x": "trial_kernel",
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
    reference has no threading on this path) on a bounded sample of the same workload and
    scales the iteration rate linearly in n (a streaming workload)."""
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
    threads = int(O.lib().orc_threads())
    build = ("oracle/liboracle_omp.so (the same source with OpenMP-parallel loops and dots, gcc -O3 -march=native "
             f"-fopenmp, {threads} threads: a GENEROUS baseline, the reference itself is serial)" if O.OMP_VARIANT else
             "oracle/liboracle.so (gcc -O2 -ffp-contract=off, 1 thread = the reference's semantics)")
    return {"value": rate_sample * n_sample / n_target, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": (f"{build} on n=2^{n_sample.bit_length() - 1}: "
                       f"{its} main-loop iterations, {trials} trials, {dt:.2f} s = {rate_sample:.3f} it/s; scaled by "
                       f"n_sample/n (streaming)"),
            "sample_it_per_s": rate_sample, "sample_trials_per_iteration": trials / max(its, 1)}


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
    # ... and on all host cores (OpenMP build of the same port).  The line's value is the faster, all-cores one, so the
    # driver's ours/reference ratio is the conservative reading; the single-thread sample is reported beside it.
    cb = cpu_lbfgs_all_cores(n_sample, args.mem, args.warmup, args.steps, n, args.objective) or single
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / cb["value"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload(n, args.mem, args.objective),
                       "timing": "host perf_counter around oracle iterations"},
            "cpu_baseline": cb, "single_thread": single,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.time() - t0}
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------- our arm
def run_ours(args):
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
    comm = None
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
        fl.lib().flgpu_comm_uses_peer_memory.argtypes = [C.c_void_p]

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    n = 1 << args.log2n
    mem, W, K = args.mem, args.warmup, args.steps
    lo = (n * rank // world) // 2 * 2            # even boundaries: Rosenbrock pairs never straddle shards
    hi = n if rank == world - 1 else (n * (rank + 1) // world) // 2 * 2
    n_local = hi - lo
    diag = args.objective == "diag"
    OBJ, START = (fl.OBJ_DIAGQUAD, fl.START_ZERO) if diag else (fl.OBJ_ROSENBROCK, fl.START_ROSEN_PERT)
    prob = fl.builtin_problem(OBJ)
    first, last = mem + W - 1, mem + W + K - 1   # observer indices bracketing exactly K main-loop iterations

    def timed_run(time_kernels, fused=True, policy=None):
        policy = policy or args.line_search
        x = fl.DeviceVector.start(START, n_local, seed=SEED, offset=lo, n_global=n)
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
        st = fl.LBFGS(prob, x, Memory=mem, Warning=False, MaxIteration=W + K, observer=ob, comm=comm, offset=lo,
                      n_global=n, time_kernels=time_kernels, fused=fused, device_search=DS, line_search=policy)
        if "t1" not in mark:
            raise SystemExit(f"bench.py: optimizer stopped after {st.iterations} iterations (status {st.status}) "
                             f"before {last + 1}; lower --steps")
        ms = ev[0].elapsed_time(ev[1])
        try:
            evs = mark.get("per") or []
            mark["iter_ms"] = sorted(evs[k].elapsed_time(evs[k + 1]) for k in range(len(evs) - 1))
        except Exception:
            mark["iter_ms"] = []
        x.free()
        return ms, mark, st, fl.kernel_times() if time_kernels else None

    # ---- pass 1: the metric (device-resident inputs, no per-kernel events)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    fused = not args.plain
    DS = {"auto": None, "on": True, "off": False}[args.device_search]
    ds_active = fused and (DS is True or (DS is None and n_local <= (1 << 25))) and \
        (comm is None or bool(fl.lib().flgpu_comm_uses_peer_memory(comm)))
    ms, mark, st, _ = timed_run(False, fused)
    clocks = sampler.stop(mark["t0"], mark["t1"]) if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    launches = (mark["c1"][0] - mark["c0"][0]) + (mark["c1"][1] - mark["c0"][1])   # library kernels + objective kernels
    trials = mark["c1"][2] - mark["c0"][2]
    it_ms = mark.get("iter_ms") or []
    per_iteration = ({"median_ms": it_ms[len(it_ms) // 2], "min_ms": it_ms[0], "max_ms": it_ms[-1], "n": len(it_ms),
                      "note": "this rank's CUDA-event time between consecutive accepted steps inside the timed window "
                              "(iterations differ by their trial counts)"} if it_ms else None)

    # ---- pass 2: per-kernel CUDA-event times over the same timed region -> roofline of the dominant kernel
    ms2, mark2, st2, kt = timed_run(True, fused)
    # ---- pass 3: the other line-search mode, for the record (same iterates up to summation order)
    ms3, mark3, _, _ = timed_run(False, not fused)
    t3 = torch.tensor([ms3], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t3, op=dist.ReduceOp.MAX)
    other = {"line_search": LS_MODES[not fused], "value": K / (float(t3.item()) * 1e-3), "unit": UNIT,
             "trials_in_timed_region": mark3["c1"][2] - mark3["c0"][2]}
    # ---- pass 4: the optional FLGPU_LS_FAST policy (NOT the reference's searcher: iterates differ, so this is a
    # separate record and never the headline)
    fast = None
    if args.line_search == "reference" and world == 1:   # 1 GPU only: the policy has not been run on row shards yet
        ms4, mark4, _, _ = timed_run(False, fused, "fast")
        t4 = torch.tensor([ms4], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t4, op=dist.ReduceOp.MAX)
        tr4 = mark4["c1"][2] - mark4["c0"][2]
        fast = {"line_search_policy": POLICIES["fast"], "value": K / (float(t4.item()) * 1e-3), "unit": UNIT,
                "ms_per_step": float(t4.item()) / K, "trials_in_timed_region": tr4, "trials_per_iteration": tr4 / K,
                "note": "same K-iteration window of a run made with line_search=fast; iteration rate, not time to "
                        "solution (DESIGN.md 3.2: with WolfeConst2 = 0.9 this policy needs more iterations on Rosenbrock)"}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback)"
    kernels = {}
    for name, v in kt.items():
        if v["ms"] > 0 and v["bytes"] > 0:
            kernels[name] = {"launches": v["launches"], "ms": round(v["ms"], 3), "avg_ms": v["ms"] / v["launches"],
                             "GBps": v["bytes"] / v["ms"] / 1e6, "frac": v["bytes"] / v["ms"] / 1e6 / peak}
    own = {k: v for k, v in kernels.items() if not k.startswith("callback:")}
    top = max(own, key=lambda k: own[k]["ms"])
    traffic = None
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        traffic = prof.get(NCU_NAMES.get(top, top), {}).get("dram_bytes_per_launch")
    except (OSError, ValueError):
        pass
    total_bytes = sum(v["bytes"] for v in kt.values())
    total_kernel_ms = sum(v["ms"] for v in kt.values())
    roofline = {"bound": "hbm", "kernel": top, "achieved": own[top]["GBps"], "peak": peak, "unit": "GB/s",
                "frac": own[top]["frac"], "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": kt[top]["bytes"] / kt[top]["launches"],
                "share_of_step": kt[top]["ms"] / total_kernel_ms,
                "whole_step": {"algorithmic_GB": total_bytes / 1e9, "kernel_ms": total_kernel_ms,
                               "GBps": total_bytes / total_kernel_ms / 1e6, "frac": total_bytes / total_kernel_ms / 1e6 / peak,
                               "note": "all kernels of the K timed iterations (pass 2)"},
                "kernels": kernels}

    # ---- e2e: the reference-facing call with HOST buffers (pinned): H2D of x, the whole optimisation (iteration 0,
    # the m-1 pre-iterations, Ke main iterations) and D2H of x inside the timed region.  Called twice: cold (work space
    # allocated and returned to the driver inside the call, as the reference does) and warm (flgpu_set_workspace_cache:
    # buffers parked by an untimed warm-up call are reused) -- `value` is the warm call, the cold one is reported beside it.
    Ke = args.e2e_steps
    xh = torch.empty(n_local, dtype=torch.float64).pin_memory()
    L = fl.lib()
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
        ms = max(e0.elapsed_time(e1), (time.time() - te) * 1e3)
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), ste

    cold_ms, cold_st = e2e_call(Ke)
    L.flgpu_set_workspace_cache(1)
    e2e_call(0)                                  # untimed warm-up: parks the work space
    e2e_ms, ste = e2e_call(Ke)
    L.flgpu_set_workspace_cache(0)               # releases the parked buffers
    e2e = {"value": ste.iterations / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 8 * n / ste.iterations,
           "d2h_bytes_per_step": 8 * n / ste.iterations, "iterations": ste.iterations, "ms": e2e_ms,
           "call": ("__nonlinearoptimization_MOD_lbfgs (host x, device-pointer callbacks)" if world == 1
                    else "flgpu_lbfgs (host x shard, row-shard communicator)"),
           "cold_call": {"value": cold_st.iterations / (cold_ms * 1e-3), "ms": cold_ms,
                         "note": "first call: work space cudaMalloc'ed and cudaFree'd inside the call"},
           "note": "one optimizer call incl. H2D of x, iteration 0 + m-1 pre-iterations (never part of `value` above: "
                   f"their first line searches take 100-200 trials) + {Ke} main iterations, D2H of x; work space reused "
                   "from an untimed warm-up call (flgpu_set_workspace_cache); bytes/step = 8n/iterations (x crosses "
                   "PCIe once per call)"}

    cpu = cpu_all = None
    if rank == 0 and world == 1 and not args.no_cpu and not diag:
        cpu = cpu_lbfgs(1 << min(args.log2n, args.cpu_log2n), mem, min(W, 3), min(K, 10), n)
        cpu_all = cpu_lbfgs_all_cores(1 << min(args.log2n, args.cpu_log2n), mem, min(W, 3), min(K, 10), n)

    if rank == 0:
        line = {"metric": METRIC, "value": K / (ms_max * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload(n, mem, args.objective), "n_global": n, "rows_per_gpu": n_local,
                           "parallelism": f"row-shard x{world}" if world > 1 else "1 GPU",
                           "exchange": (None if comm is None else
                                        ("one kernel over IPC-mapped peer memory (NVLink stores + flags), rank-ordered sum"
                                         if fl.lib().flgpu_comm_uses_peer_memory(comm) else "ncclAllGather + combine kernel")),
                           "l2": "inputs (2 GiB/vector) exceed L2; no flush needed",
                           "line_search": LS_MODES[fused],
                           "line_search_policy": POLICIES[args.line_search],
                           "device_resident_search": (f"{args.device_search}: " + (
                               "ON (one cooperative kernel per line search, flgpu_search_fn)" if ds_active else
                               "off at this size (host-driven, one round trip per trial)")),
                           "trials_in_timed_region": trials, "trials_per_iteration": trials / K,
                           "trials_per_sec": trials / (ms_max * 1e-3)},
                "per_iteration": per_iteration,
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
                "cpu_baseline_all_cores": cpu_all,
                "other_line_search_mode": other, "fast_line_search_policy": fast}
        print(json.dumps(line), flush=True)
    if comm is not None:
        fl.lib().flgpu_comm_destroy(comm)
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "cpu-sample"])
    ap.add_argument("--log2n", type=int, default=LOG2_N, help="override the global dimension (debugging only)")
    ap.add_argument("--mem", type=int, default=MEM)
    ap.add_argument("--objective", default="rosenbrock", choices=["rosenbrock", "diag"],
                    help="diag = BASELINE.json configs[3] (with --mem 30 --log2n 31 --gpus 8); not the headline")
    ap.add_argument("--e2e-steps", type=int, default=100,
                    help="main-loop iterations of the end-to-end optimizer call (its 487-trial prologue is amortised over them)")
    ap.add_argument("--device-search", default="auto", choices=["auto", "on", "off"],
                    help="flgpu_options.device_search; auto = up to 2^25 rows per GPU (same bits either way)")
    ap.add_argument("--line-search", default="reference", choices=["reference", "fast"],
                    help="flgpu_options.line_search for the whole run; the headline is `reference` (the default run "
                         "also records the `fast` rate beside it)")
    ap.add_argument("--plain", action="store_true", help="headline with opaque callbacks (no fused line-search evaluation)")
    ap.add_argument("--cpu-log2n", type=int, default=None, help="size of the bounded CPU sample")
    ap.add_argument("--no-cpu", action="store_true")
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
