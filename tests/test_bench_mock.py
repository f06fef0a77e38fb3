"""bench.py's own control flow (window marking, the four passes, the e2e calls, the CPU rows, the JSON line) executed on
a machine without a GPU, against stand-ins for `torch.cuda` and for the package: a Python error in bench.py would
otherwise only show at round end on the GPU box.  Nothing here measures anything; the numbers are fake by design."""
import importlib
import io
import json
import os
import sys
import time
import types
from contextlib import redirect_stdout

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Event:
    def __init__(self, enable_timing=True):
        self.t = None

    def record(self, stream=None):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return max((other.t - self.t) * 1e3, 1e-3)


class _Tensor:
    def __init__(self, v):
        self.v = v

    def item(self):
        return self.v[0]

    def pin_memory(self):
        return self

    def data_ptr(self):
        return 0


def _fake_torch():
    t = types.ModuleType("torch")
    t.float64, t.uint8 = "f64", "u8"
    t.tensor = lambda v, dtype=None, device=None: _Tensor(list(v))
    t.empty = lambda n, dtype=None: _Tensor([0.0])
    cuda = types.SimpleNamespace(is_available=lambda: True, set_device=lambda d: None, synchronize=lambda: None,
                                 Event=_Event, ExternalStream=lambda p: ("stream", p))
    t.cuda = cuda
    return t


class _Stats:
    def __init__(self, iterations):
        self.iterations, self.status = iterations, 2


def _fake_fl(calls):
    fl = types.ModuleType("fortran_library_b200")
    import importlib.util
    spec = importlib.util.spec_from_file_location("_flgpu_capi_mock", os.path.join(ROOT, "fortran_library_b200", "_capi.py"))
    capi = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(capi)
    fl.capi = capi
    for k in ("OBJ_ROSENBROCK", "OBJ_DIAGQUAD", "OBJ_QUARTIC", "START_ROSEN_PERT", "START_ZERO", "START_QUARTIC_U",
              "SPACE_HOST", "SPACE_DEVICE", "LS_FAST", "LS_REFERENCE"):
        setattr(fl, k, getattr(capi, k))
    fl.require_gpu = lambda: None
    fl.builtin_problem = lambda kind: object()

    class DV:
        ptr = 0

        @classmethod
        def start(cls, *a, **k):
            return cls()

        def free(self):
            pass
    fl.DeviceVector = DV

    class Observer:
        def __init__(self, on_iteration=None):
            self.on_iteration = on_iteration
    fl.Observer = Observer

    def LBFGS(prob, x, Memory=10, MaxIteration=0, observer=None, line_search=None, fused=True, **kw):
        calls.append(("LBFGS", line_search, fused, kw.get("time_kernels")))
        trials_per_it = 1 if line_search == "fast" else 8
        total = 1 + (Memory - 1) + MaxIteration
        for it in range(total):
            info = types.SimpleNamespace(iteration=it, stream=0, gpu_launches=3 * it, callbacks=trials_per_it * it,
                                         total_trials=trials_per_it * it)
            if observer is not None and observer.on_iteration(info):
                break
        return _Stats(it + 1)
    fl.LBFGS = LBFGS

    def ConjugateGradient(prob, x, Method=None, MaxIteration=0, observer=None, **kw):
        calls.append(("CG", Method, kw.get("fused"), kw.get("time_kernels")))
        for it in range(MaxIteration):
            info = types.SimpleNamespace(iteration=it, stream=0, gpu_launches=3 * it, callbacks=15 * it, total_trials=15 * it)
            if observer is not None and observer.on_iteration(info):
                break
        return _Stats(it + 1)
    fl.ConjugateGradient = ConjugateGradient
    fl.kernel_times = lambda: {"k1_update_dots": {"ms": 7.0, "launches": 30, "bytes": 30 * 5.0e10},
                               "k3_direction": {"ms": 7.5, "launches": 30, "bytes": 30 * 4.7e10},
                               "callback:fused_probe": {"ms": 5.0, "launches": 240, "bytes": 240 * 4.3e9}}

    class Lib:
        def __getattr__(self, name):
            def fn(*a):
                calls.append((name,))
                if name == "flgpu_last_stats":
                    a[0]._obj.iterations = 110
                return 0
            return fn
    lib = Lib()
    fl.lib = lambda: lib
    return fl


@pytest.mark.timeout(300)
def test_bench_control_flow_with_stand_ins(monkeypatch):
    calls = []
    monkeypatch.setitem(sys.modules, "torch", _fake_torch())
    monkeypatch.setitem(sys.modules, "fortran_library_b200", _fake_fl(calls))
    monkeypatch.setenv("RANK", "0"); monkeypatch.setenv("WORLD_SIZE", "1"); monkeypatch.setenv("LOCAL_RANK", "0")
    sys.path.insert(0, ROOT)
    bench = importlib.import_module("bench")
    monkeypatch.setattr(sys, "argv", ["bench.py", "--steps", "6", "--warmup", "3", "--log2n", "16", "--cpu-log2n", "12",
                                      "--e2e-steps", "5"])
    monkeypatch.setattr(bench.ClockSampler, "start", lambda self: None)
    monkeypatch.setattr(bench.ClockSampler, "stop", lambda self, t0, t1: None)
    out = io.StringIO()
    with redirect_stdout(out):
        assert bench.main() == 0
    lines = [ln for ln in out.getvalue().splitlines() if ln.startswith("{")]
    assert len(lines) == 1, out.getvalue()
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline",
                "cpu_baseline_all_cores", "other_line_search_mode", "fast_line_search_policy", "per_iteration", "run",
                "secondary", "parity", "hbm_GBps_whole_step", "trials_per_sec"):
        assert key in d, key
    assert d["steps"] == 6 and d["run"]["trials_in_timed_region"] == 6 * 8
    assert d["parity"] is None                                  # single GPU: nothing to compare
    assert set(d["secondary"]) == {"cg_dy_quartic", "cg_pr_quartic", "lbfgs_m30_diag_2p28"}
    assert d["secondary"]["cg_dy_quartic"]["trials_per_iteration"] == 15.0
    assert d["config"] == bench.static_config(1 << 16, 10, "rosenbrock")   # the reference arm prints the same object
    assert d["per_iteration"]["n"] == 6 and d["per_iteration"]["min_ms"] <= d["per_iteration"]["median_ms"]
    assert d["fast_line_search_policy"]["trials_per_iteration"] == 1.0
    assert d["roofline"]["kernel"] == "k3_direction" and d["roofline"]["bound"] == "hbm"
    assert d["cpu_baseline"]["cores"] == 1 and d["cpu_baseline"]["kind"] == "port"
    if d["cpu_baseline_all_cores"] is not None:
        assert d["cpu_baseline_all_cores"]["cores"] >= 1 and "GENEROUS" in d["cpu_baseline_all_cores"]["sample"]
    assert d["e2e"]["iterations"] == 110
    # four timed passes (metric, per-kernel events, the other line-search mode, the fast policy), in that order
    passes = [c for c in calls if c[0] == "LBFGS"]
    assert [(p[1], p[2], bool(p[3])) for p in passes[:4]] == [("reference", True, False), ("reference", True, True),
                                                              ("reference", False, False), ("fast", True, False)]
