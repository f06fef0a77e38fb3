"""CPU tests of the product's HOST control flow (fortran_library_b200/csrc/driver.cpp, the same
translation unit libflgpu.so links) run over the test-only host simulator of backend.hpp.

Parity criterion (DESIGN.md "parity"): the reference algorithm is chaotic in its rounding errors --
the oracle differs from itself by far more than 1e-12 after a few iterations when only its summation
order changes (tests/test_oracle.py::test_summation_noise_is_above_1e12).  So trajectories are compared
against the oracle's own summation-order envelope, and exactly (bit for bit) where no summation
exists (dim = 1)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import _cases
import _hostsim as H
import _oracle as O

capi = H.capi
ENV_FACTOR = _cases.ENV_FACTOR
FLOOR = _cases.FLOOR
_check_envelope = _cases.check_envelope


@pytest.mark.parametrize("name,kw", [
    ("rosenR0", dict(Memory=10)), ("rosenR1", dict(Memory=10)), ("rosenR1", dict(Memory=5)),
    ("quartic", dict(Memory=10)), ("diag", dict(Memory=30, MaxIteration=40)),
    ("rosenR1", dict(Memory=1, MaxIteration=30)), ("rosenR1", dict(Memory=10, Strong=False, MaxIteration=30)),
    ("rosenR1", dict(Memory=10, use_ffd=False)),
])
def test_lbfgs_directions_within_oracle_envelope(name, kw):
    n = 2000
    kw = dict(kw)
    use = kw.pop("use_ffd", True)
    kind = _cases.OBJECTIVES[name][0]
    traces, env = _cases.oracle_envelope(name, n, lambda cbs, x, **k: O.lbfgs(cbs, x, use_ffd=use, **k), **kw)
    ob = H.Observer(max_vec_iters=20)
    x, st = H.lbfgs(kind, _cases.start(name, n), observer=ob, use_ffd=use, Warning=False, n_global=n, **kw)
    _check_envelope(traces, ob.p, f"lbfgs {name} {kw}")
    # the steepest-descent step has no history: identical direction, identical first step
    assert np.array_equal(ob.p[0], traces[0].p[0])


@pytest.mark.parametrize("method", ["DY", "PR"])
@pytest.mark.parametrize("name,kw", [("quartic", dict()), ("rosenR1", dict(MaxIteration=60)),
                                     ("diag", dict(MaxIteration=60)), ("quartic", dict(Strong=False, MaxIteration=60)),
                                     ("quartic", dict(use_ffd=False))])
def test_cg_directions_within_oracle_envelope(method, name, kw):
    n = 2000
    kw = dict(kw)
    use = kw.pop("use_ffd", True)
    kind = _cases.OBJECTIVES[name][0]
    traces, env = _cases.oracle_envelope(name, n, lambda cbs, x, **k: O.cg(cbs, x, Method=method, use_ffd=use, **k), **kw)
    ob = H.Observer(max_vec_iters=20)
    x, st = H.cg(kind, _cases.start(name, n), observer=ob, use_ffd=use, Warning=False, n_global=n, Method=method, **kw)
    _check_envelope(traces, ob.p, f"cg {method} {name} {kw}")


@pytest.mark.parametrize("name,kw", [("quartic", dict(MaxIteration=40)), ("rosenR1", dict(MaxIteration=40)),
                                     ("diag", dict(MaxIteration=40, Strong=False)),
                                     ("quartic", dict(MaxIteration=40, use_ffd=False))])
def test_sd_directions_within_oracle_envelope(name, kw):
    """SteepestDescent (f90:55-188) through the product's driver."""
    n = 2000
    kw = dict(kw)
    use = kw.pop("use_ffd", True)
    kind = _cases.OBJECTIVES[name][0]
    traces, env = _cases.oracle_envelope(name, n, lambda cbs, x, **k: O.sd(cbs, x, use_ffd=use, **k), **kw)
    ob = H.Observer(max_vec_iters=20)
    x, st = H.sd(kind, _cases.start(name, n), observer=ob, use_ffd=use, Warning=False, n_global=n, **kw)
    _check_envelope(traces, ob.p, f"sd {name} {kw}")
    assert np.array_equal(ob.p[0], traces[0].p[0])


@pytest.mark.parametrize("name,mem", [("rosenR1", 10), ("rosenR1", 3), ("quartic", 10), ("diag", 30), ("rosenR0", 5),
                                      ("quartic", 1)])
def test_one_step_direction_parity_1e12(name, mem):
    """The strict form of "search directions agree to relative 1e-12 over the first 20 iterations":
    fed the ORACLE's own history (x_k, f'_k) at every iteration, the compact two-loop (K1+K2+K3) must
    reproduce the oracle's next direction to 1e-12 -- no chaotic amplification is involved."""
    _cases.check_one_step(H.History, name, mem, strict=False)


def test_minimisers_and_iteration_counts():
    """north_star: minimisers to 1e-8 relative; iteration counts within 2 % of the oracle's own range under
    one-ULP perturbations (_cases.oracle_iteration_range)."""
    n = 2000
    for name in ("rosenR0", "rosenR1"):
        kind = _cases.OBJECTIVES[name][0]
        counts, statuses, x_ref = _cases.oracle_iteration_range(
            name, n, lambda cbs, x, **k: O.lbfgs(cbs, x, use_ffd=True, **k))
        x, st = H.lbfgs(kind, _cases.start(name, n), Warning=False, n_global=n)
        assert _cases.rel(x, x_ref) < 1e-8
        _cases.check_iteration_count(st.iterations, counts, f"lbfgs {name}")
        assert st.status in statuses


@pytest.mark.parametrize("lbfgs,name,kw", [
    (True, "rosenR1", dict(Memory=10)), (True, "rosenR1", dict(Memory=3, use_ffd=False)),
    (True, "quartic", dict(Memory=5, Strong=False)), (True, "diag", dict(Memory=7, MaxIteration=60)),
    (False, "quartic", dict(Method="DY")), (False, "quartic", dict(Method="PR", use_ffd=False)),
    (False, "rosenR1", dict(Method="DY", Strong=False, MaxIteration=80)),
    ("sd", "quartic", dict(MaxIteration=50)), ("sd", "rosenR1", dict(MaxIteration=50, Strong=False)),
])
def test_fused_line_search_is_the_same_algorithm(lbfgs, name, kw):
    """flgpu_fused_fn changes where trial points live, not what is computed: with the host simulator
    (identical reductions on both paths) the fused and the unfused run must agree bit for bit in every
    direction, step, iterate and evaluation count.  The third run takes the driver's device-resident-search branch
    (SearchCore with eager evaluations, as the cooperative CUDA kernels instantiate it) and must agree as well."""
    n = 3001
    kind = _cases.OBJECTIVES[name][0]
    run = H.sd if lbfgs == "sd" else (H.lbfgs if lbfgs else H.cg)
    out = []
    for fused, dsearch in ((True, False), (False, False), (True, True)):
        ob = H.Observer(max_vec_iters=10**9)
        x, st = run(kind, _cases.start(name, n), observer=ob, Warning=False, n_global=n, fused=fused,
                    device_search=dsearch, **kw)
        out.append((x, st, ob))
    (xa, sa, oa), (xb, sb, ob_) = out[0], out[1]
    (xc, sc, oc) = out[2]
    assert np.array_equal(xa, xc) and oa.rows == oc.rows
    assert all(np.array_equal(u, v) for u, v in zip(oa.p, oc.p))
    for k in ("iterations", "status", "n_f", "n_fd", "n_f_fd", "n_trials", "n_f_only_trials", "n_linesearch"):
        assert getattr(sa, k) == getattr(sc, k), ("device-search", k)
    assert sc.host_syncs < sa.host_syncs
    assert np.array_equal(xa, xb)
    assert oa.rows == ob_.rows
    assert all(np.array_equal(u, v) for u, v in zip(oa.p, ob_.p))
    assert all(np.array_equal(u, v) for u, v in zip(oa.x, ob_.x))
    assert all(np.array_equal(u, v) for u, v in zip(oa.g, ob_.g))
    for k in ("iterations", "status", "n_f", "n_fd", "n_f_fd", "n_trials", "n_f_only_trials", "n_linesearch"):
        assert getattr(sa, k) == getattr(sb, k), k
    # the fused run above batched the trials of every bracketing walk four to a pass (flgpu_fused_multi_fn; here a batch
    # is by definition four separate evaluations).  With batching off: the same everything, more passes and round trips.
    import os
    os.environ["FLGPU_FUSED_MULTI"] = "0"
    try:
        od = H.Observer(max_vec_iters=10**9)
        xd, sd_ = run(kind, _cases.start(name, n), observer=od, Warning=False, n_global=n, fused=True, device_search=False, **kw)
    finally:
        del os.environ["FLGPU_FUSED_MULTI"]
    assert np.array_equal(xa, xd) and oa.rows == od.rows
    assert all(np.array_equal(u, v) for u, v in zip(oa.p, od.p))
    for k in ("iterations", "status", "n_f", "n_fd", "n_f_fd", "n_trials", "n_f_only_trials", "n_linesearch"):
        assert getattr(sa, k) == getattr(sd_, k), ("unbatched", k)
    assert sd_.n_batched_passes == 0 and sa.n_batched_passes > 0
    assert sa.host_syncs < sd_.host_syncs and 2 * sa.n_batched_passes < sa.n_trials
    # L-BFGS: the first four trials of every search after the first ride on K3 (flgpu_problem.direction; here: the direction
    # followed by four separate evaluations) and seed the search's first batch.  Without it: the same, more batches.
    if lbfgs is True:
        os.environ["FLGPU_FUSED_DIRECTION"] = "0"
        try:
            oe = H.Observer(max_vec_iters=10**9)
            xe, se = run(kind, _cases.start(name, n), observer=oe, Warning=False, n_global=n, fused=True, device_search=False, **kw)
        finally:
            del os.environ["FLGPU_FUSED_DIRECTION"]
        assert np.array_equal(xa, xe) and oa.rows == oe.rows
        for k in ("iterations", "status", "n_f", "n_fd", "n_f_fd", "n_trials", "n_f_only_trials", "n_linesearch"):
            assert getattr(sa, k) == getattr(se, k), ("no K3 probe", k)
        assert sa.n_batched_passes <= se.n_batched_passes


def _py_problem(fuse):
    """flgpu_problem over Python scalar functions (host pointers: this is the host simulator)."""
    def f(ctx, fp, xp, n):
        C.cast(fp, C.POINTER(C.c_double))[0] = fuse.f(C.cast(xp, C.POINTER(C.c_double))[0])

    def fd(ctx, gp, xp, n):
        C.cast(gp, C.POINTER(C.c_double))[0] = fuse.g(C.cast(xp, C.POINTER(C.c_double))[0])

    def ffd(ctx, fp, gp, xp, n):
        fv, gv = fuse.fg(C.cast(xp, C.POINTER(C.c_double))[0])
        C.cast(fp, C.POINTER(C.c_double))[0] = fv
        C.cast(gp, C.POINTER(C.c_double))[0] = gv
    keep = (capi.F_FN(f), capi.FD_FN(fd), capi.F_FD_FN(ffd))
    p = capi.Problem()
    p.f, p.fd, p.f_fd = (C.cast(k, C.c_void_p) for k in keep)
    p._keep = keep
    return p


@pytest.mark.parametrize("case", sorted(_cases.TORTURE_1D))
@pytest.mark.parametrize("method", ["DY", "PR"])
def test_torture_1d_bitwise_vs_oracle(case, method):
    """dim = 1: no summation order exists, so every trial point, step and iterate must be identical
    to the oracle's, through every branch of the Strong-Wolfe searcher incl. f90:1511-1512."""
    x0, (f, g) = _cases.TORTURE_1D[case]
    for use in (False, True):
        if use and case in _cases.TORTURE_NO_FFD:
            continue
        fa = _cases.Fuse(f, g)
        cf, cfd, cffd = _cases.make_ref_callbacks(fa.f, fa.g, fa.fg)
        keep = (O.F_T(cf), O.FD_T(cfd), O.FFD_T(cffd))
        tr = O.Trace()
        xa, s = O.cg(tuple(C.cast(k, C.c_void_p) for k in keep), np.array([x0]), Method=method, use_ffd=use,
                     Warning=False, MaxIteration=30, trace=tr)
        fb = _cases.Fuse(f, g)
        prob = _py_problem(fb)
        if not use:
            prob.f_fd = None
        L = H.lib()
        o = capi.Options()
        L.flgpu_hostsim_options_default(C.byref(o), 1)
        capi.apply_options(o, Method=method, Warning=False, MaxIteration=30)
        ob = H.Observer()
        o.observer = C.cast(ob.cb, C.c_void_p)
        x = np.array([x0])
        st = capi.Stats()
        L.flgpu_hostsim_cg(C.byref(prob), C.byref(o), x.ctypes.data_as(C.c_void_p), C.c_int64(1), C.byref(st))
        assert fa.xs == fb.xs, "different trial points"
        assert np.array_equal(x, xa, equal_nan=True)
        assert st.iterations == s.n_iter and st.status == s.status
        assert [r[1:] for r in ob.rows] == [r[1:] for r in tr.rows]


def test_option_clamps_and_basic_variant():
    """f90:431-434 clamps vs ConjugateGradient_basic's unclamped constants (f90:2278)."""
    n = 200          # (the iterates are 1e-6-sized by then: 1e-8 relative to them is 1e-14 of the start)
    kind = O.OBJ_QUARTIC
    x0 = _cases.start("quartic", n)
    # c2 <= c1 is clamped to c1+1e-15 by ConjugateGradient but used as given by _basic
    xa, sa = O.cg(O.builtin_callbacks(kind, 0, n), x0.copy(), WolfeConst1=0.3, WolfeConst2=0.1, Warning=False, MaxIteration=20)
    xb, stb = H.cg(kind, x0, WolfeConst1=0.3, WolfeConst2=0.1, Warning=False, MaxIteration=20, n_global=n)
    assert stb.iterations == sa.n_iter and _cases.rel(xb, xa) < 1e-8
    xc, sc = O.cg_basic(O.builtin_callbacks(kind, 0, n), x0.copy(), WolfeConst1=0.3, WolfeConst2=0.1, Warning=False, MaxIteration=20)
    xd, std = H.cg(kind, x0, WolfeConst1=0.3, WolfeConst2=0.1, Warning=False, MaxIteration=20, n_global=n, use_ffd=False,
                   no_clamp=1)
    assert std.iterations == sc.n_iter and _cases.rel(xd, xc) < 1e-8
    # Memory <= 0 is clamped to 1 (f90:419)
    xe, se = O.lbfgs(O.builtin_callbacks(kind, 0, n), x0.copy(), Memory=0, Warning=False, MaxIteration=10)
    xf, stf = H.lbfgs(kind, x0, Memory=0, Warning=False, MaxIteration=10, n_global=n, use_ffd=False)
    assert stf.iterations == se.n_iter and _cases.rel(xf, xe) < 1e-8


def test_edge_cases():
    # start at the minimiser: |f'|^2 < tol, return at once (f90:443 / 237)
    for fn in (H.lbfgs, H.cg):
        x, st = fn(O.OBJ_ROSENBROCK, np.ones(10), Warning=False, n_global=10)
        assert st.status == capi.INITIAL_CONVERGED and st.iterations == 0 and np.array_equal(x, np.ones(10))
    # odd dimension, dimension 1 and 2, Memory larger than the iteration count
    # (more pairs than dimensions makes the history linearly dependent: the Gram-space recurrences lose
    # digits there, so only a loose agreement is asserted; DESIGN.md "accuracy of the compact form")
    for n in (1, 2, 3, 7):
        x0 = _cases.start("quartic", n)
        xa, sa = O.lbfgs(O.builtin_callbacks(O.OBJ_QUARTIC, 0, n), x0.copy(), Memory=4, Warning=False, MaxIteration=5)
        xb, stb = H.lbfgs(O.OBJ_QUARTIC, x0, Memory=4, Warning=False, MaxIteration=5, n_global=n, use_ffd=False)
        assert stb.iterations == sa.n_iter and _cases.rel(xb, xa) < 1e-6


WORKER = r"""
import ctypes as C, os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, {tests!r})
import _hostsim as H, _oracle as O
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank, n = dist.get_rank(), {n}
lo, hi = (0, {split}) if rank == 0 else ({split}, n)

def allgather(user, local, allp, count):
    t = torch.from_numpy(np.ctypeslib.as_array(local, (count,)).copy())
    outs = [torch.empty(count, dtype=torch.float64) for _ in range(2)]
    dist.all_gather(outs, t)
    np.ctypeslib.as_array(allp, (2 * count,))[:] = torch.cat(outs).numpy()

cb = H.ALLGATHER_FN(allgather)
H.set_comm(cb, rank, 2)
x0 = O.start_vector({start}, n, seed={seed})
ob = H.Observer(max_vec_iters=12)
fn = H.lbfgs if {lbfgs} else H.cg
x, st = fn({kind}, x0[lo:hi], observer=ob, Warning=False, offset=lo, n_global=n, MaxIteration={maxit}, **{kw!r})
np.savez({out!r} + str(rank), x=x, rows=np.array(ob.rows), p=np.array(ob.p), iters=st.iterations, status=st.status)
dist.destroy_process_group()
"""


@pytest.mark.parametrize("lbfgs,name,kw", [(True, "rosenR1", dict(Memory=6)), (True, "diag", dict(Memory=4)),
                                           (False, "quartic", dict(Method="PR")),
                                           (True, "rosenR1", dict(Memory=6, line_search="fast")),
                                           (False, "quartic", dict(Method="DY", line_search="fast"))])
@pytest.mark.parametrize("n,split", [(1000, 400), (4096, 2048)], ids=["ragged", "aligned"])
def test_row_sharded_two_ranks_gloo(tmp_path, lbfgs, name, kw, n, split):
    """N>1 path on CPU: two processes, each owning a row shard, exchanging only the per-rank roots
    (all-gather over gloo) and combining them by the rank tree.  Both ranks must take identical decisions
    (bitwise equal scalars) and the assembled result must match a single-process run -- within tolerance for the
    ragged split (400 + 600 rows: partial chunks), BIT FOR BIT for the aligned one (2 + 2 whole chunks of 1024: each
    rank's root is a node of the single-process tree, include/flgpu_reduce.cuh)."""
    maxit = 25
    kind, start, seed = _cases.OBJECTIVES[name]
    out = str(tmp_path / "r")
    H.lib()                     # build once here: the two workers must not race on `make`
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    src = WORKER.format(tests=os.path.dirname(os.path.abspath(__file__)), port=port, n=n, split=split, start=start,
                        seed=seed, lbfgs=lbfgs, kind=kind, maxit=maxit, kw=kw, out=out)
    procs = [subprocess.Popen([sys.executable, "-c", src, str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in (0, 1)]
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    r0, r1 = np.load(out + "0.npz"), np.load(out + "1.npz")
    assert np.array_equal(r0["rows"], r1["rows"]), "ranks diverged: scalars must be bitwise identical on every rank"
    assert int(r0["iters"]) == int(r1["iters"]) and int(r0["status"]) == int(r1["status"])
    x0 = O.start_vector(start, n, seed=seed)
    ob = H.Observer(max_vec_iters=12)
    fn = H.lbfgs if lbfgs else H.cg
    xs, st = fn(kind, x0, observer=ob, Warning=False, n_global=n, MaxIteration=maxit, **kw)
    x2 = np.concatenate([r0["x"], r1["x"]])
    p2 = np.concatenate([r0["p"], r1["p"]], axis=1)
    assert st.iterations == int(r0["iters"])
    if split * 2 == n and split % 1024 == 0:
        assert np.array_equal(x2, xs) and np.array_equal(np.array(ob.rows), r0["rows"])
        assert all(np.array_equal(p2[k], ob.p[k]) for k in range(min(len(ob.p), len(p2))))
    # relative 1e-7, scaled by |x0| where the minimiser is 0 (quartic: |x| ~ 1e-3 |x0| after 25 iterations)
    assert np.linalg.norm(x2 - xs) / max(np.linalg.norm(xs), np.linalg.norm(x0)) < 1e-7
    for k in range(min(len(ob.p), len(p2), 8)):
        assert _cases.rel(p2[k], ob.p[k]) < 1e-9
