"""GPU parity tests (pytest -m gpu, on the B200 box).  Everything goes through the C-ABI of libflgpu.so
(ctypes); the oracle (oracle/liboracle.so) is only the checker.  Nothing here reads /root/reference.

Tolerances (BASELINE.json north_star, and DESIGN.md "parity" for why they are applied this way):
  * integer/index work and element-wise results: bit-exact;
  * one-step search directions on the oracle's own history: 1e-12 relative;
  * whole trajectories: within the reference's own summation-order noise (x64) or 1e-12;
  * minimisers 1e-8 relative, iteration counts 2 % where the oracle itself is that stable.
"""
import ctypes as C
import json
import math
import os
import subprocess
import sys

import numpy as np
import pytest

import _cases
import _oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def fl():
    import fortran_library_b200 as fl
    fl.require_gpu()          # fails loudly: there is no fallback to test instead
    return fl


def _problem(fl, name, use_ffd=True):
    p = fl.builtin_problem(_cases.OBJECTIVES[name][0])
    if not use_ffd:
        p.f_fd = None
    return p


def _dev_start(fl, name, n):
    kind, st, seed = _cases.OBJECTIVES[name]
    return fl.DeviceVector.start(st, n, seed=seed)


# ----------------------------------------------------------------------------- building blocks
@pytest.mark.parametrize("name", sorted(_cases.OBJECTIVES))
@pytest.mark.parametrize("n", [1, 2, 3, 31, 1000, 4097])
def test_start_vectors_bit_exact(fl, name, n):
    assert np.array_equal(_dev_start(fl, name, n).numpy(), _cases.start(name, n))


@pytest.mark.parametrize("name", ["quartic", "rosenR1", "diag"])
@pytest.mark.parametrize("n", [1, 2, 7, 1000, 65537, 1 << 22, (1 << 22) + 3])   # the last two: full grid, unrolled trips, tail
def test_objective_kernels_vs_oracle(fl, name, n):
    """K6: f' bit-exact (same operation order, no FMA), f to summation-order accuracy."""
    kind = _cases.OBJECTIVES[name][0]
    rng = np.random.default_rng(n)
    x = _cases.start(name, n) + 0.01 * rng.standard_normal(n)
    fo, go = C.c_double(), np.empty(n)
    O.lib().orc_obj_select(kind, 0, n)
    O.lib().orc_obj_f_fd(C.byref(fo), go.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p), C.byref(C.c_int(n)))
    prob = fl.builtin_problem(kind)
    xd, gd, fd_ = fl.DeviceVector.from_numpy(x), fl.DeviceVector(n), fl.DeviceVector(2)
    ctx = fl.capi.EvalCtx(None, None, 0, n, 0, 1, 0)
    zeros = np.zeros(n)                     # kept alive: a temporary would be unmapped before the copy reads it
    for which in ("f_fd", "f", "fd"):
        fl.lib().flgpu_memcpy(gd.ptr, zeros.ctypes.data, n * 8, 1, 0, None)
        if which == "f_fd":
            C.cast(prob.f_fd, fl.capi.F_FD_FN)(C.byref(ctx), fd_.ptr, gd.ptr, xd.ptr, n)
        elif which == "f":
            C.cast(prob.f, fl.capi.F_FN)(C.byref(ctx), fd_.ptr, xd.ptr, n)
        else:
            C.cast(prob.fd, fl.capi.FD_FN)(C.byref(ctx), gd.ptr, xd.ptr, n)
        if which != "f":
            assert np.array_equal(gd.numpy(), go), f"{which}: gradient differs"
        if which != "fd":
            assert abs(fd_.numpy()[0] - fo.value) <= 1e-13 * abs(fo.value) + 1e-300


@pytest.mark.parametrize("name", ["quartic", "rosenR1", "diag"])
@pytest.mark.parametrize("n", [1, 2, 7, 1000, 65537, 1 << 22, (1 << 22) + 3])
def test_fused_evaluation_vs_plain_callbacks(fl, name, n):
    """flgpu_fused_fn of the built-in objectives: the point it forms equals flgpu_vec_trial bit for bit,
    f and f' equal the plain callbacks' bit for bit (same mapping, same order), f'.p to dot accuracy."""
    kind = _cases.OBJECTIVES[name][0]
    rng = np.random.default_rng(n + 17)
    x0 = _cases.start(name, n) + 0.01 * rng.standard_normal(n)
    p = rng.standard_normal(n)
    a = 0.3712345
    prob = fl.builtin_problem(kind)
    fused = C.cast(prob.fused, fl.capi.FUSED_FN)
    x0d, pd = fl.DeviceVector.from_numpy(x0), fl.DeviceVector.from_numpy(p)
    xd, gd, xo, go, sc = (fl.DeviceVector(n), fl.DeviceVector(n), fl.DeviceVector(n), fl.DeviceVector(n),
                          fl.DeviceVector(4))
    ctx = fl.capi.EvalCtx(None, None, 0, n, 0, 1, 0)
    fl.lib().flgpu_vec_trial(xd.ptr, x0d.ptr, pd.ptr, a, n, None)
    C.cast(prob.f_fd, fl.capi.F_FD_FN)(C.byref(ctx), sc.ptr, gd.ptr, xd.ptr, n)
    f_plain, x_plain, g_plain = sc.numpy()[0], xd.numpy(), gd.numpy()
    W = fl.capi
    for flags in (W.WANT_F | W.WANT_GP, W.WANT_F, W.WANT_GP, W.WRITE_X | W.WRITE_G, W.WRITE_G, W.WRITE_X,
                  W.WANT_F | W.WANT_GP | W.WRITE_X | W.WRITE_G):
        nans = np.full(4, np.nan)
        fl.lib().flgpu_memcpy(sc.ptr, nans.ctypes.data, 32, 1, 0, None)
        fused(C.byref(ctx), flags, sc.ptr, sc.ptr + 8, xo.ptr, go.ptr, x0d.ptr, pd.ptr, a, n)
        out = sc.numpy()
        if flags & W.WANT_F:
            assert out[0] == f_plain
        if flags & W.WANT_GP:
            exact = math.fsum(g_plain * p)
            assert abs(out[1] - exact) <= 4e-16 * float(np.sum(np.abs(g_plain * p))) + 1e-300
        if flags & W.WRITE_X:
            assert np.array_equal(xo.numpy(), x_plain)
        if flags & W.WRITE_G:
            assert np.array_equal(go.numpy(), g_plain)


@pytest.mark.parametrize("n", [1, 2, 3, 255, 256, 257, 100003, 1 << 20, 1 << 22, (1 << 22) + 3])
def test_vector_primitives(fl, n):
    rng = np.random.default_rng(n)
    a, b = rng.standard_normal(n), rng.standard_normal(n)
    ad, bd, out = fl.DeviceVector.from_numpy(a), fl.DeviceVector.from_numpy(b), fl.DeviceVector(1)
    fl.lib().flgpu_vec_dot(ad.ptr, bd.ptr, n, out.ptr, None)
    exact = math.fsum(a * b)
    bound = 4e-16 * float(np.sum(np.abs(a * b))) + 1e-300
    assert abs(out.numpy()[0] - exact) <= bound
    xd = fl.DeviceVector(n)
    fl.lib().flgpu_vec_trial(xd.ptr, ad.ptr, bd.ptr, 0.37, n, None)
    assert np.array_equal(xd.numpy(), a + 0.37 * b)          # multiply then add, no FMA (f90:1482)
    # determinism: same bits on repetition
    fl.lib().flgpu_vec_dot(ad.ptr, bd.ptr, n, out.ptr, None)
    first = out.numpy()[0]
    for _ in range(3):
        fl.lib().flgpu_vec_dot(ad.ptr, bd.ptr, n, out.ptr, None)
        assert out.numpy()[0] == first
    # integer-valued data: every partial sum is exact, so ANY summation order must give the exact dot, bit for bit
    ai, bi = rng.integers(-30, 31, n), rng.integers(-30, 31, n)
    aid, bid = fl.DeviceVector.from_numpy(ai.astype(np.float64)), fl.DeviceVector.from_numpy(bi.astype(np.float64))
    fl.lib().flgpu_vec_dot(aid.ptr, bid.ptr, n, out.ptr, None)
    assert out.numpy()[0] == float(int(np.dot(ai, bi)))


@pytest.mark.parametrize("n", [1, 2, 3, 31, 255, 257, 4097, 100003])
def test_no_out_of_bounds_writes(fl, n):
    """compute-sanitizer is closed on this pool, so every kernel that writes a caller-visible vector is run on a
    buffer embedded between canary zones (odd lengths, odd 8-byte alignments) and the canaries are checked."""
    G = 67
    canary = -7.25e300
    rng = np.random.default_rng(n)

    a, b = rng.standard_normal(n), rng.standard_normal(n)
    Ga = G + 1                                                 # vectors must stay 16-byte aligned (double2 accesses)

    def guarded_al(values=None):
        host = np.full(n + 2 * Ga, canary)
        if values is not None:
            host[Ga:Ga + n] = values
        return fl.DeviceVector.from_numpy(host)
    def inner_al(v):
        return v.ptr + 8 * Ga

    def intact_al(v):
        h = v.numpy()
        return bool(np.all(h[:Ga] == canary) and np.all(h[Ga + n:] == canary)), h[Ga:Ga + n]

    ad, bd, xd = guarded_al(a), guarded_al(b), guarded_al()
    fl.lib().flgpu_vec_trial(inner_al(xd), inner_al(ad), inner_al(bd), 0.5, n, None)
    ok, x = intact_al(xd)
    assert ok and np.array_equal(x, a + 0.5 * b)
    for name in ("quartic", "rosenR1", "diag"):
        prob = fl.builtin_problem(_cases.OBJECTIVES[name][0])
        ctx = fl.capi.EvalCtx(None, None, 0, n, 0, 1, 0)
        gd, xo, sc = guarded_al(), guarded_al(), fl.DeviceVector(4)
        C.cast(prob.f_fd, fl.capi.F_FD_FN)(C.byref(ctx), sc.ptr, inner_al(gd), inner_al(ad), n)
        assert intact_al(gd)[0], name
        W = fl.capi
        C.cast(prob.fused, fl.capi.FUSED_FN)(C.byref(ctx), W.WRITE_X | W.WRITE_G, sc.ptr, sc.ptr + 8, inner_al(xo),
                                             inner_al(gd), inner_al(ad), inner_al(bd), 0.25, n)
        assert intact_al(gd)[0] and intact_al(xo)[0], name
    # K1 + K2 + K3 through the history operator: p and the trial point are caller buffers
    L = fl.lib()
    L.flgpu_history_create.restype = C.c_void_p
    L.flgpu_history_create.argtypes = [C.c_int64, C.c_int, C.c_void_p, C.c_void_p]
    L.flgpu_history_push.argtypes = [C.c_void_p] * 5
    L.flgpu_history_direction.argtypes = [C.c_void_p] * 7
    L.flgpu_history_destroy.argtypes = [C.c_void_p]
    h = L.flgpu_history_create(n, 12, None, None)
    d = np.exp(rng.uniform(0, 1, n))
    x0 = rng.standard_normal(n)
    pd, xt = guarded_al(), guarded_al()
    for _ in range(14):                                       # more pushes than memory: every K1 shape and the wrap-around
        x1 = x0 + 0.1 * rng.standard_normal(n)
        vs = [guarded_al(v) for v in (x1, x0, d * x1, d * x0)]
        L.flgpu_history_push(h, *[inner_al(v) for v in vs])
        gp, pp = C.c_double(), C.c_double()
        L.flgpu_history_direction(h, inner_al(vs[2]), inner_al(vs[0]), inner_al(pd), inner_al(xt), C.addressof(gp),
                                  C.addressof(pp))
        ok_p, pv = intact_al(pd)
        ok_x, xv = intact_al(xt)
        assert ok_p and ok_x
        assert all(intact_al(v)[0] for v in vs)               # inputs untouched outside AND inside
        assert np.array_equal(xv, x1 + pv)
        x0 = x1
    L.flgpu_history_destroy(h)


# ----------------------------------------------------------------------------- partition-independent reductions
def _rank_tree(vals):
    """red::rank_tree (include/flgpu_reduce_geom.h): aligned binary tree over the rank index."""
    v = list(vals) + [0.0] * (16 - len(vals))
    s = 1
    while s < 16:
        for j in range(0, 16 - s, 2 * s):
            v[j] = v[j] + v[j + s]
        s *= 2
    return v[0]


@pytest.mark.parametrize("log2n", [13, 20, 22, 26])
def test_reductions_do_not_depend_on_the_partition(fl, log2n):
    """The same vector reduced as 1, 2, 4 and 8 row shards (each shard's root from the kernels, the roots combined by
    the rank tree) gives IDENTICAL bits -- dot products and the objective value of all three built-in objectives.
    2^13: 8 chunks (one per shard at 8 shards); 2^22: 4096 chunks = one tree block; 2^26: 8192 chunks of 8192 elements,
    two tree blocks on one GPU against one block per shard."""
    n = 1 << log2n
    L = fl.lib()
    L.flgpu_vec_dot_sharded.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(log2n)
    a, b = rng.standard_normal(n), rng.standard_normal(n)
    ad, bd, out = fl.DeviceVector.from_numpy(a), fl.DeviceVector.from_numpy(b), fl.DeviceVector(2)
    results = {}
    for G in (1, 2, 4, 8):
        m = n // G
        roots = []
        for r in range(G):
            L.flgpu_vec_dot_sharded(ad.ptr + 8 * r * m, bd.ptr + 8 * r * m, m, n, out.ptr, None)
            roots.append(out.numpy()[0])
        results[G] = _rank_tree(roots)
    assert results[1] == results[2] == results[4] == results[8], results
    assert abs(results[1] - math.fsum(a * b)) <= 4e-16 * float(np.sum(np.abs(a * b)))
    for name in ("quartic", "rosenR1", "diag"):
        kind = _cases.OBJECTIVES[name][0]
        prob = fl.builtin_problem(kind)
        fused = C.cast(prob.fused, fl.capi.FUSED_FN)
        x0 = _cases.start(name, n) + 0.01 * a
        x0d = fl.DeviceVector.from_numpy(x0)
        res = {}
        for G in (1, 2, 4, 8):
            m = n // G
            fs, gps = [], []
            for r in range(G):
                ctx = fl.capi.EvalCtx(None, None, r * m, n, r, G, 0)
                fused(C.byref(ctx), fl.capi.WANT_F | fl.capi.WANT_GP, out.ptr, out.ptr + 8, None, None, x0d.ptr + 8 * r * m,
                      bd.ptr + 8 * r * m, 0.125, m)
                o = out.numpy()
                fs.append(o[0]); gps.append(o[1])
            res[G] = (_rank_tree(fs), _rank_tree(gps))
        assert res[1] == res[2] == res[4] == res[8], (name, res)


# ----------------------------------------------------------------------------- parity: strict tier
@pytest.mark.parametrize("name,mem", [("rosenR1", 10), ("rosenR1", 3), ("quartic", 10), ("diag", 30), ("rosenR0", 5),
                                      ("quartic", 1), ("rosenR1", 17), ("rosenR0", 10), ("quartic1", 10)])
def test_one_step_direction_parity_1e12(fl, name, mem):
    """K1+K2+K3 on the oracle's own history reproduce its next direction to 1e-12 (20 iterations)."""
    _cases.check_one_step(fl.History, name, mem, n=10_000)


@pytest.mark.parametrize("mem", [1, 2, 5, 10, 30, 33, 64])
def test_two_loop_operator_all_kernel_shapes(fl, mem):
    """Every (columns-per-group, groups) instantiation of K1 and the two-slot-per-lane path of K2
    against the extended-precision two-loop on random (well-conditioned) pairs."""
    n = 5003
    rng = np.random.default_rng(mem)
    d = np.exp(rng.uniform(0, 2, n))                     # SPD diagonal Hessian: y = d * s keeps s.y > 0
    h = fl.History(n, mem)
    x0, pairs = rng.standard_normal(n), []
    g0 = d * x0
    for k in range(mem + 3):
        x1 = x0 + 0.1 * rng.standard_normal(n)
        g1 = d * x1
        h.push(x1, x0, g1, g0)
        pairs = (pairs + [(x1 - x0, g1 - g0)])[-mem:]
        p, xt, gp, pp = h.direction(g1, x1)
        exact = _cases.two_loop_extended(pairs, g1)
        assert _cases.rel(p, exact) < 1e-12, (mem, k)
        assert np.array_equal(xt, x1 + p)
        x0, g0 = x1, g1
    h.close()


# ----------------------------------------------------------------------------- parity: trajectories
@pytest.mark.parametrize("name,kw", [
    ("rosenR0", dict(Memory=10)), ("rosenR1", dict(Memory=10)), ("rosenR1", dict(Memory=5)),
    ("quartic", dict(Memory=10)), ("diag", dict(Memory=30, MaxIteration=40)),
    ("rosenR1", dict(Memory=1, MaxIteration=30)), ("rosenR1", dict(Memory=10, Strong=False, MaxIteration=30)),
    ("rosenR1", dict(Memory=10, use_ffd=False)), ("quartic1", dict(Memory=10)),
])
@pytest.mark.parametrize("fused", [True, False], ids=["fused", "plain"])
def test_lbfgs_trajectory_within_oracle_envelope(fl, name, kw, fused):
    n = 10_000                                             # BASELINE.json configs[0]
    kw = dict(kw, fused=fused)
    use = kw.pop("use_ffd", True)
    okw = {k: v for k, v in kw.items() if k != "fused"}
    traces, _ = _cases.oracle_envelope(name, n, lambda cbs, x, **k: O.lbfgs(cbs, x, use_ffd=use, **k), **okw)
    ob = fl.Observer(keep_vectors=True, max_vec_iters=20)
    x = _dev_start(fl, name, n)
    st = fl.LBFGS(_problem(fl, name, use), x, observer=ob, Warning=False, **kw)
    _cases.check_envelope(traces, ob.p, f"lbfgs {name} {kw}")
    assert np.array_equal(ob.p[0], traces[0].p[0])         # steepest-descent step: bit-exact direction
    assert st.gpu_launches > 0


@pytest.mark.parametrize("method", ["DY", "PR"])
@pytest.mark.parametrize("name,kw", [("quartic", dict()), ("rosenR1", dict(MaxIteration=60)),
                                     ("diag", dict(MaxIteration=60)), ("quartic", dict(Strong=False, MaxIteration=60)),
                                     ("quartic", dict(use_ffd=False)), ("quartic1", dict())])
@pytest.mark.parametrize("fused", [True, False], ids=["fused", "plain"])
def test_cg_trajectory_within_oracle_envelope(fl, method, name, kw, fused):
    n = 10_000
    kw = dict(kw)
    use = kw.pop("use_ffd", True)
    traces, _ = _cases.oracle_envelope(name, n, lambda cbs, x, **k: O.cg(cbs, x, Method=method, use_ffd=use, **k), **kw)
    kw["fused"] = fused
    ob = fl.Observer(keep_vectors=True, max_vec_iters=20)
    x = _dev_start(fl, name, n)
    fl.ConjugateGradient(_problem(fl, name, use), x, Method=method, observer=ob, Warning=False, **kw)
    _cases.check_envelope(traces, ob.p, f"cg {method} {name} {kw}")


@pytest.mark.parametrize("fused", [True, False], ids=["fused", "plain"])
@pytest.mark.parametrize("name,kw", [("quartic", dict(MaxIteration=40)), ("rosenR1", dict(MaxIteration=40)),
                                     ("diag", dict(MaxIteration=40, Strong=False)),
                                     ("quartic", dict(MaxIteration=40, use_ffd=False))])
def test_sd_trajectory_within_oracle_envelope(fl, name, kw, fused):
    """SteepestDescent (f90:55-188) through the same kernels."""
    n = 10_000
    kw = dict(kw)
    use = kw.pop("use_ffd", True)
    traces, _ = _cases.oracle_envelope(name, n, lambda cbs, x, **k: O.sd(cbs, x, use_ffd=use, **k), **kw)
    ob = fl.Observer(keep_vectors=True, max_vec_iters=20)
    x = _dev_start(fl, name, n)
    st = fl.SteepestDescent(_problem(fl, name, use), x, observer=ob, Warning=False, fused=fused, **kw)
    _cases.check_envelope(traces, ob.p, f"sd {name} {kw}")
    assert np.array_equal(ob.p[0], traces[0].p[0])
    assert st.iterations == len(traces[0].rows)


def test_minimisers_and_iteration_counts(fl):
    """north_star: minimisers to 1e-8 relative; iteration counts within 2 % of the oracle's own range under
    one-ULP perturbations (_cases.oracle_iteration_range explains why a single count is not a target)."""
    n = 10_000
    for name in ("rosenR0", "rosenR1"):
        counts, statuses, x_ref = _cases.oracle_iteration_range(
            name, n, lambda cbs, x, **k: O.lbfgs(cbs, x, use_ffd=True, **k))
        for fused in (True, False):
            x = _dev_start(fl, name, n)
            st = fl.LBFGS(_problem(fl, name), x, Warning=False, fused=fused)
            assert _cases.rel(x.numpy(), x_ref) < 1e-8              # minimiser, relative 1e-8
            _cases.check_iteration_count(st.iterations, counts, f"lbfgs {name} fused={fused}")
            assert st.status in statuses
    # CG where the minimiser is well defined: sum (x-1)^4 + (x-1)^2, x* = 1, linear convergence.  north_star as
    # written: minimiser to relative 1e-8, iteration count within 2 % (17 and 7 iterations: equal), same exit.
    # L-BFGS on it likewise.
    x0 = _cases.start("quartic1", n)
    cbs = lambda: O.builtin_callbacks(O.OBJ_QUARTIC_SHIFTED, 0, n)     # noqa: E731
    for M in ("DY", "PR"):
        xr, sr = O.cg(cbs(), x0.copy(), Method=M, use_ffd=True, Warning=False)
        assert sr.status == 0 and np.abs(xr - 1.0).max() < 1e-12
        for fused in (True, False):
            x = _dev_start(fl, "quartic1", n)
            st = fl.ConjugateGradient(_problem(fl, "quartic1"), x, Method=M, Warning=False, fused=fused)
            assert _cases.rel(x.numpy(), xr) < 1e-8, (M, fused, _cases.rel(x.numpy(), xr))
            assert abs(st.iterations - sr.n_iter) <= 0.02 * sr.n_iter, (M, fused, st.iterations, sr.n_iter)
            assert st.status == sr.status
    xr, sr = O.lbfgs(cbs(), x0.copy(), use_ffd=True, Warning=False)
    x = _dev_start(fl, "quartic1", n)
    st = fl.LBFGS(_problem(fl, "quartic1"), x, Warning=False)
    assert _cases.rel(x.numpy(), xr) < 1e-8 and abs(st.iterations - sr.n_iter) <= 0.02 * sr.n_iter and st.status == sr.status
    # CG on the reference's own test objective sum x^4 (config 3; test.f90:350-373 expects "norm2(x) close to 0"): the
    # minimiser is 0 and convergence is sub-linear, so a RELATIVE distance between two runs has no scale (the oracle's
    # own summation orders differ by 15 % of |x| there).  Asserted: the reference's criterion -- both runs end as
    # close to 0 -- the distance in units of |x0|, and the iteration count within 2 % (+-1).
    x0 = _cases.start("quartic", n)
    for M in ("DY", "PR"):
        xr, sr = O.cg(O.builtin_callbacks(O.OBJ_QUARTIC, 0, n), x0.copy(), Method=M, use_ffd=True, Warning=False)
        x = _dev_start(fl, "quartic", n)
        st = fl.ConjugateGradient(_problem(fl, "quartic"), x, Method=M, Warning=False)
        nx, nr, n0 = np.linalg.norm(x.numpy()), np.linalg.norm(xr), np.linalg.norm(x0)
        assert nx / n0 < 1e-5 and 0.5 < nx / nr < 2.0
        assert np.linalg.norm(x.numpy() - xr) / n0 < 1e-6
        assert abs(st.iterations - sr.n_iter) <= max(1, 0.02 * sr.n_iter) and st.status == sr.status


# ----------------------------------------------------------------------------- Fortran ABI (drop-in)
def _ref_call_cg(fl, cf, cfd, cffd, x, method, use, maxit):
    L = fl.lib()
    keep = (fl.capi.REF_F_FN(cf), fl.capi.REF_FD_FN(cfd), fl.capi.REF_F_FD_FN(cffd))
    dim = C.c_int(x.size)
    m = method.encode()
    L.flgpu_set_callback_space(fl.SPACE_HOST)
    try:
        L.__getattr__("__nonlinearoptimization_MOD_conjugategradient")(
            keep[0], keep[1], x.ctypes.data_as(C.c_void_p), C.byref(dim), m, keep[2] if use else None,
            None, C.byref(C.c_int32(0)), C.byref(C.c_int(maxit)), None, None, None, None, None, C.c_int(len(m)))
    finally:
        L.flgpu_set_callback_space(-1)             # back to the automatic choice
    st = fl.capi.Stats()
    L.flgpu_last_stats(C.byref(st))
    return st


@pytest.mark.parametrize("case", sorted(_cases.TORTURE_1D))
@pytest.mark.parametrize("method", ["DY", "PR"])
def test_fortran_abi_torture_1d_bitwise(fl, case, method):
    """The reference's own symbol (__nonlinearoptimization_MOD_conjugategradient), absent optionals as
    NULL, host callbacks staged through pinned memory.  dim = 1 has no summation order, so every trial
    point and the result must equal the oracle's bit for bit -- through every Strong-Wolfe branch."""
    x0, (f, g) = _cases.TORTURE_1D[case]
    for use in (False, True):
        if use and case in _cases.TORTURE_NO_FFD:
            continue
        fa = _cases.Fuse(f, g)
        cf, cfd, cffd = _cases.make_ref_callbacks(fa.f, fa.g, fa.fg)
        keep = (O.F_T(cf), O.FD_T(cfd), O.FFD_T(cffd))
        xa, s = O.cg(tuple(C.cast(k, C.c_void_p) for k in keep), np.array([x0]), Method=method, use_ffd=use,
                     Warning=False, MaxIteration=30)
        fb = _cases.Fuse(f, g)
        cf2, cfd2, cffd2 = _cases.make_ref_callbacks(fb.f, fb.g, fb.fg)
        x = np.array([x0])
        st = _ref_call_cg(fl, cf2, cfd2, cffd2, x, method, use, 30)
        assert fa.xs == fb.xs, "different trial points"
        assert np.array_equal(x, xa, equal_nan=True)
        assert st.iterations == s.n_iter and st.status == s.status


def test_fortran_abi_device_callbacks_lbfgs(fl):
    """__nonlinearoptimization_MOD_lbfgs with the built-in CUDA objective in reference-ABI form: host x
    in/out, device pointers in the callbacks (the default spaces)."""
    n = 10_000
    L = fl.lib()
    f, fd, ffd = fl.capi.REF_F_FN(), fl.capi.REF_FD_FN(), fl.capi.REF_F_FD_FN()
    L.flgpu_builtin_ref_callbacks(fl.OBJ_ROSENBROCK, C.byref(f), C.byref(fd), C.byref(ffd))
    x = _cases.start("rosenR0", n)
    xr, sr = O.lbfgs(O.builtin_callbacks(O.OBJ_ROSENBROCK, 0, n), x.copy(), use_ffd=True, Warning=False)
    L.__getattr__("__nonlinearoptimization_MOD_lbfgs")(
        f, fd, x.ctypes.data_as(C.c_void_p), C.byref(C.c_int(n)), C.byref(C.c_int(10)), ffd, None,
        C.byref(C.c_int32(0)), None, None, None, None, None, None)
    st = fl.capi.Stats()
    L.flgpu_last_stats(C.byref(st))
    assert st.iterations == sr.n_iter and _cases.rel(x, xr) < 1e-8
    # ifort spelling of the same entry point
    x2 = _cases.start("rosenR0", n)
    L.nonlinearoptimization_mp_lbfgs_(f, fd, x2.ctypes.data_as(C.c_void_p), C.byref(C.c_int(n)), None, ffd, None,
                                      C.byref(C.c_int32(0)), None, None, None, None, None, None)
    assert np.array_equal(x2, x)


def test_fortran_abi_steepest_descent(fl):
    """__nonlinearoptimization_MOD_steepestdescent (hpp:279-291) with host callbacks staged by the library:
    dim = 1, so the result must equal the oracle's bit for bit."""
    x0, (f, g) = _cases.TORTURE_1D["quartic1"]
    L = fl.lib()
    for use in (False, True):
        fa = _cases.Fuse(f, g)
        cf, cfd, cffd = _cases.make_ref_callbacks(fa.f, fa.g, fa.fg)
        keep = (O.F_T(cf), O.FD_T(cfd), O.FFD_T(cffd))
        xa, s = O.sd(tuple(C.cast(k, C.c_void_p) for k in keep), np.array([x0]), use_ffd=use, Warning=False,
                     MaxIteration=25)
        fb = _cases.Fuse(f, g)
        cf2, cfd2, cffd2 = _cases.make_ref_callbacks(fb.f, fb.g, fb.fg)
        keep2 = (fl.capi.REF_F_FN(cf2), fl.capi.REF_FD_FN(cfd2), fl.capi.REF_F_FD_FN(cffd2))
        x = np.array([x0])
        L.flgpu_set_callback_space(fl.SPACE_HOST)
        try:
            L.__getattr__("__nonlinearoptimization_MOD_steepestdescent")(
                keep2[0], keep2[1], x.ctypes.data_as(C.c_void_p), C.byref(C.c_int(1)), keep2[2] if use else None,
                None, C.byref(C.c_int32(0)), C.byref(C.c_int(25)), None, None, None, None, None)
        finally:
            L.flgpu_set_callback_space(-1)             # back to the automatic choice
        st = fl.capi.Stats()
        L.flgpu_last_stats(C.byref(st))
        assert fa.xs == fb.xs, "different trial points"
        assert np.array_equal(x, xa, equal_nan=True) and st.iterations == s.n_iter and st.status == s.status


# ----------------------------------------------------------------------------- device-resident line search
@pytest.mark.parametrize("algo,name,kw", [
    ("lbfgs", "rosenR1", dict(Memory=10, MaxIteration=60)), ("lbfgs", "rosenR1", dict(Memory=5, use_ffd=False, MaxIteration=40)),
    ("lbfgs", "quartic", dict(Memory=7)), ("lbfgs", "diag", dict(Memory=30, MaxIteration=40)),
    ("lbfgs", "rosenR1", dict(Memory=4, Strong=False, MaxIteration=40)),
    ("cg", "quartic", dict(Method="DY")), ("cg", "quartic", dict(Method="PR", use_ffd=False)),
    ("cg", "rosenR1", dict(Method="DY", Strong=False, MaxIteration=60)), ("sd", "rosenR1", dict(MaxIteration=40)),
    ("sd", "quartic", dict(MaxIteration=40, Strong=False)),
])
@pytest.mark.parametrize("n", [10_000, 4097])
def test_device_resident_search_is_the_same_algorithm(fl, algo, name, kw, n):
    """flgpu_search_fn: the whole line search in one cooperative kernel, driven by the same SearchCore source as the
    host.  Same launch geometry -> same bits for f and f'.p -> same decisions: every step, f, trial count, iterate
    and evaluation counter must be IDENTICAL to the host-driven fused search."""
    kw = dict(kw)
    use = kw.pop("use_ffd", True)
    run = {"lbfgs": fl.LBFGS, "cg": fl.ConjugateGradient, "sd": fl.SteepestDescent}[algo]
    out = []
    for dev in (False, True):
        x = _dev_start(fl, name, n)
        ob = fl.Observer(keep_vectors=True, max_vec_iters=12)
        st = run(_problem(fl, name, use), x, observer=ob, Warning=False, device_search=dev, **kw)
        out.append((x.numpy(), st, ob))
    (xa, sa, oa), (xb, sb, ob_) = out
    assert oa.rows == ob_.rows
    assert np.array_equal(xa, xb)
    assert all(np.array_equal(u, v) for u, v in zip(oa.p, ob_.p))
    for k in ("iterations", "status", "n_f", "n_fd", "n_f_fd", "n_trials", "n_f_only_trials", "n_linesearch"):
        assert getattr(sa, k) == getattr(sb, k), k
    assert sb.host_syncs < sa.host_syncs            # the point of it: fewer host round trips


def test_device_resident_search_terminates_on_nan():
    """A NaN objective makes the reference's zoom spin forever (f90:1684,1695) -- acceptable on a CPU, not inside a GPU
    kernel: the device-resident search gives up after its evaluation budget and says so."""
    code = ("import sys; sys.path.insert(0, %r)\nimport numpy as np, fortran_library_b200 as fl\n"
            "x = np.full(64, np.nan)\n"
            "st = fl.LBFGS(fl.builtin_problem(fl.OBJ_QUARTIC), x, Memory=3, MaxIteration=0, device_search=True)\n"
            "print('returned', st.iterations)\n" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=240)
    assert r.returncode == 0 and "returned" in r.stdout, r.stdout + r.stderr
    assert "gave up after" in r.stdout


@pytest.mark.parametrize("name,kw", [("rosenR1", dict(Memory=10, MaxIteration=40)), ("quartic", dict(Memory=3)),
                                     ("diag", dict(Memory=30, MaxIteration=45)), ("rosenR1", dict(Memory=12, use_ffd=False, MaxIteration=30))])
@pytest.mark.parametrize("n", [10_001, 1 << 20])
def test_fused_update_is_the_same_algorithm(name, kw, n):
    """flgpu_problem.update: K1 forms x1 = x0 + a*p and f'(x1) in registers and stores them, instead of the line search
    storing the accepted point and K1 reading it back.  Same roundings, same chunk sums: every iterate, step and counter
    must be IDENTICAL with the callback switched off (FLGPU_FUSED_UPDATE=0), host-driven and device-resident search."""
    kw = dict(kw)
    use = kw.pop("use_ffd", True)
    code = ("import sys, json; sys.path.insert(0, %r); sys.path.insert(0, %r)\nimport numpy as np, fortran_library_b200 as fl\n"
            "kind, start, seed = %r\n"
            "out = {}\n"
            "for ds in (False, True):\n"
            "    x = fl.DeviceVector.start(start, %d, seed=seed)\n"
            "    p = fl.builtin_problem(kind)\n"
            "    if not %r: p.f_fd = None\n"
            "    ob = fl.Observer()\n"
            "    st = fl.LBFGS(p, x, Warning=False, observer=ob, device_search=ds, **%r)\n"
            "    out[str(ds)] = dict(rows=ob.rows, x=x.numpy().tobytes().hex()[:4096], xs=float(np.sum(x.numpy())), n_f_fd=st.n_f_fd, n_f=st.n_f,\n"
            "                        n_fd=st.n_fd, trials=st.n_trials, it=st.iterations, status=st.status, launches=st.gpu_launches)\n"
            "print(json.dumps(out))\n" % (ROOT, os.path.join(ROOT, "tests"), _cases.OBJECTIVES[name], n, use, kw))
    res = []
    for env in ({}, {"FLGPU_FUSED_UPDATE": "0"}):
        r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env), capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        res.append(json.loads(r.stdout.strip().splitlines()[-1]))
    on, off = res
    for ds in ("False", "True"):
        a, b = on[ds], off[ds]
        assert a["rows"] == b["rows"] and a["x"] == b["x"] and a["xs"] == b["xs"]
        for k in ("n_f_fd", "n_f", "n_fd", "trials", "it", "status"):
            assert a[k] == b[k], k
    assert on["False"]["rows"] == on["True"]["rows"] and on["False"]["x"] == on["True"]["x"]


@pytest.mark.parametrize("name,kw", [("rosenR1", dict(Memory=10, MaxIteration=40)), ("quartic", dict(Memory=3)),
                                     ("diag", dict(Memory=30, MaxIteration=45)), ("rosenR1", dict(Memory=12, use_ffd=False, MaxIteration=30)),
                                     ("rosenR1", dict(Memory=10, MaxIteration=30, line_search="fast")),
                                     ("quartic1", dict(Memory=5, Strong=False))])
@pytest.mark.parametrize("n", [10_001, (1 << 22) + 3])
def test_fused_direction_is_the_same_algorithm(name, kw, n):
    """flgpu_problem.direction: K3 evaluates the a = 1 trial of the next line search while it writes p (x1 + p formed in
    registers; f and f'.p reduced in the chunk order of the fused evaluation).  Every iterate, step and counter must be
    IDENTICAL with the callback switched off (FLGPU_FUSED_DIRECTION=0: a separate probe launch after K3);
    n = 2^22+3 runs the full 296-block grid with an odd tail."""
    kw = dict(kw)
    use = kw.pop("use_ffd", True)
    code = ("import sys, json; sys.path.insert(0, %r); sys.path.insert(0, %r)\nimport numpy as np, fortran_library_b200 as fl\n"
            "kind, start, seed = %r\n"
            "x = fl.DeviceVector.start(start, %d, seed=seed)\n"
            "p = fl.builtin_problem(kind)\n"
            "if not %r: p.f_fd = None\n"
            "ob = fl.Observer()\n"
            "st = fl.LBFGS(p, x, Warning=False, observer=ob, device_search=False, **%r)\n"
            "xn = x.numpy()\n"
            "import hashlib\n"
            "print(json.dumps(dict(rows=ob.rows, x=hashlib.sha256(xn.tobytes()).hexdigest(), xs=float(np.sum(xn)), n_f_fd=st.n_f_fd, n_f=st.n_f,\n"
            "                 n_fd=st.n_fd, trials=st.n_trials, it=st.iterations, status=st.status, f=st.f, g2=st.gnorm2)))\n"
            % (ROOT, os.path.join(ROOT, "tests"), _cases.OBJECTIVES[name], n, use, kw))
    res = []
    for env in ({}, {"FLGPU_FUSED_DIRECTION": "0"}):
        r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env), capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        res.append(json.loads(r.stdout.strip().splitlines()[-1]))
    on, off = res
    assert on["rows"] == off["rows"] and on["x"] == off["x"] and on["xs"] == off["xs"]
    for k in ("n_f_fd", "n_f", "n_fd", "trials", "it", "status", "f", "g2"):
        assert on[k] == off[k], k
    assert on["it"] > 3


@pytest.mark.parametrize("kind", ["quartic", "rosenR1", "diag", "quartic1"])
@pytest.mark.parametrize("n", [1, 2, 777, 10_001, 1 << 22, (1 << 22) + 3, (1 << 25) + 1])
def test_fused_multi_kernel_bitwise(fl, kind, n):
    """flgpu_fused_multi_fn: f and f'.p at four steps from ONE pass over x0 and p carry exactly the bits of four separate
    flgpu_fused_fn evaluations (same per-element roundings, same chunk sums, same tree).  Up to 4096 chunks (n <= 2^22)
    the register kernel with its in-kernel finish runs; above, the shared-memory-ring kernel: n = 2^22+3 (1024-element
    chunks: half-filled stages, odd tail) and 2^25+1 (8192-element chunks: four stages per chunk, grid-stride wrap, odd
    tail); also with fewer than four steps."""
    k, start, seed = _cases.OBJECTIVES[kind]
    prob = fl.builtin_problem(k)
    assert prob.fused_multi
    MULTI_FN = C.CFUNCTYPE(None, C.POINTER(fl.capi.EvalCtx), C.c_int, C.POINTER(C.c_double), C.c_void_p, C.c_void_p,
                           C.c_void_p, C.c_int64)
    x = fl.DeviceVector.start(start, n, seed=seed)
    rng = np.random.default_rng(n)
    p = fl.DeviceVector.from_numpy(rng.standard_normal(n) * 0.1)
    ctx = fl.capi.EvalCtx(C.c_void_p(prob.user), None, 0, n, 0, 1, 0)
    W = fl.capi
    for steps in ([1.05, 1.05 * 1.05, 1.05 * 1.05 * 1.05, 1.2155062500000001], [0.3, -0.7], [1e-3, 0.0, 2.5]):
        out = fl.DeviceVector(8)
        arr = (C.c_double * len(steps))(*steps)
        C.cast(prob.fused_multi, MULTI_FN)(C.byref(ctx), len(steps), arr, out.ptr, x.ptr, p.ptr, n)
        got = out.numpy()
        for j, a in enumerate(steps):
            sc = fl.DeviceVector(2)
            C.cast(prob.fused, W.FUSED_FN)(C.byref(ctx), W.WANT_F | W.WANT_GP, sc.ptr, sc.ptr + 8, None, None, x.ptr, p.ptr,
                                           a, n)
            one = sc.numpy()
            assert got[2 * j] == one[0] and got[2 * j + 1] == one[1], (kind, n, j, got[2 * j:2 * j + 2], one)


@pytest.mark.parametrize("algo,name,kw", [
    ("lbfgs", "rosenR1", dict(Memory=10, MaxIteration=40)), ("lbfgs", "rosenR1", dict(Memory=10, MaxIteration=30, use_ffd=False)),
    ("lbfgs", "rosenR1", dict(Memory=5, MaxIteration=30, Strong=False)), ("lbfgs", "diag", dict(Memory=30, MaxIteration=45)),
    ("lbfgs", "quartic", dict(Memory=3)), ("cg", "quartic", dict(Method="DY")), ("cg", "quartic1", dict(Method="PR")),
    ("cg", "rosenR1", dict(Method="DY", MaxIteration=40, Strong=False)), ("sd", "rosenR1", dict(MaxIteration=25)),
    ("lbfgs", "rosenR1", dict(Memory=10, MaxIteration=30, Increment=2.0)),
])
@pytest.mark.parametrize("n", [10_001, (1 << 22) + 3])
def test_fused_multi_is_the_same_algorithm(algo, name, kw, n):
    """Batched evaluation of a bracketing walk (flgpu_problem.fused_multi): the next four steps a, a*Increment, ... (or
    a/Increment, ...) are evaluated in the pass that evaluates the first.  Every iterate, step, trial count and
    evaluation count must be IDENTICAL with batching off (FLGPU_FUSED_MULTI=0: one probe launch per trial) -- only the
    number of passes and of host round trips drops."""
    kw = dict(kw)
    use = kw.pop("use_ffd", True)
    fn = {"lbfgs": "LBFGS", "cg": "ConjugateGradient", "sd": "SteepestDescent"}[algo]
    code = ("import sys, json, hashlib; sys.path.insert(0, %r); sys.path.insert(0, %r)\nimport numpy as np, fortran_library_b200 as fl\n"
            "kind, start, seed = %r\n"
            "x = fl.DeviceVector.start(start, %d, seed=seed)\n"
            "p = fl.builtin_problem(kind)\n"
            "if not %r: p.f_fd = None\n"
            "ob = fl.Observer()\n"
            "st = fl.%s(p, x, Warning=False, observer=ob, device_search=False, **%r)\n"
            "xn = x.numpy()\n"
            "print(json.dumps(dict(rows=ob.rows, x=hashlib.sha256(xn.tobytes()).hexdigest(), n_f_fd=st.n_f_fd, n_f=st.n_f,\n"
            "                 n_fd=st.n_fd, trials=st.n_trials, fonly=st.n_f_only_trials, it=st.iterations, status=st.status, f=st.f,\n"
            "                 g2=st.gnorm2, syncs=st.host_syncs, batched=st.n_batched_passes)))\n"
            % (ROOT, os.path.join(ROOT, "tests"), _cases.OBJECTIVES[name], n, use, fn, kw))
    res = []
    for env in ({}, {"FLGPU_FUSED_MULTI": "0"}):
        r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env), capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        res.append(json.loads(r.stdout.strip().splitlines()[-1]))
    on, off = res
    assert on["rows"] == off["rows"] and on["x"] == off["x"]
    for k in ("n_f_fd", "n_f", "n_fd", "trials", "fonly", "it", "status", "f", "g2"):
        assert on[k] == off[k], k
    assert off["batched"] == 0 and on["batched"] > 0 and on["syncs"] < off["syncs"], (on["batched"], on["syncs"], off["syncs"])


# ----------------------------------------------------------------------------- GPU == scalar C++ statement, bit for bit
@pytest.mark.parametrize("algo,name,kw", [
    ("lbfgs", "rosenR1", dict(Memory=10, MaxIteration=40)), ("lbfgs", "rosenR1", dict(Memory=10, MaxIteration=25, fused=False)),
    ("lbfgs", "diag", dict(Memory=30, MaxIteration=20)), ("lbfgs", "quartic", dict(Memory=7, use_ffd=False)),
    ("lbfgs", "quartic1", dict(Memory=10)), ("lbfgs", "rosenR1", dict(Memory=4, Strong=False, MaxIteration=30)),
    ("cg", "quartic", dict(Method="DY")), ("cg", "quartic", dict(Method="PR", fused=False)), ("cg", "quartic1", dict(Method="PR")),
    ("cg", "rosenR1", dict(Method="DY", MaxIteration=60)), ("sd", "rosenR1", dict(MaxIteration=30)),
    ("lbfgs", "rosenR1", dict(Memory=10, MaxIteration=30, line_search="fast")),
])
@pytest.mark.parametrize("n", [10_001, (1 << 17) + 2])
def test_gpu_equals_host_simulator_bit_for_bit(fl, algo, name, kw, n):
    """tests/hostsim is the product's driver.cpp over a scalar C++ backend whose element-wise arithmetic AND reductions
    restate the CUDA kernels (which thread adds which unit, FMA accumulation, lane butterfly, aligned binary tree:
    namespace model in backend_host.cpp) -- written independently of the kernels and checked against the oracle on the
    CPU.  A whole optimisation on the GPU must therefore reproduce it BIT FOR BIT: every step length, objective value,
    trial count and the final iterate; 10 001 rows = 10 chunks with an odd tail, 2^17 + 2 rows = 129 chunks (a ragged
    tree), K1 in every thread shape, fused and plain line searches."""
    import _hostsim as H
    kw = dict(kw)
    use = kw.pop("use_ffd", True)
    kind = _cases.OBJECTIVES[name][0]
    x0 = _cases.start(name, n)
    obh = H.Observer(keep_vectors=False)
    hrun = {"lbfgs": H.lbfgs, "cg": H.cg, "sd": H.sd}[algo]
    xh, sth = hrun(kind, x0, observer=obh, use_ffd=use, Warning=False, n_global=n, **kw)
    grun = {"lbfgs": fl.LBFGS, "cg": fl.ConjugateGradient, "sd": fl.SteepestDescent}[algo]
    x = _dev_start(fl, name, n)
    ob = fl.Observer()
    st = grun(_problem(fl, name, use), x, observer=ob, Warning=False, **kw)
    assert ob.rows == obh.rows, next((k, a, b) for k, (a, b) in enumerate(zip(ob.rows, obh.rows)) if a != b)
    assert np.array_equal(x.numpy(), xh)
    for k in ("iterations", "status", "n_f", "n_fd", "n_f_fd", "n_trials", "n_f_only_trials", "n_linesearch"):
        assert getattr(st, k) == getattr(sth, k), k


# ----------------------------------------------------------------------------- user objectives (flgpu_objective.cuh)
def _np_user_objective(which, n):
    """NumPy statement of tests/link/user_objective.cu (same operation order, no FMA)."""
    if which == 0:
        w = 1.0 + (np.arange(n) % 7).astype(np.float64)

        def fg(x):
            t = x - 1.0
            t2 = t * t
            return float(np.cumsum(w * (t2 * t2) + t2)[-1]), (4.0 * w) * (t2 * t) + 2.0 * t
    else:
        def fg(x):
            a, b = x[0:n - n % 2:2], x[1::2]
            t1, t2 = a - 2.0, b - a * a
            terms = t1 * t1 + (5.0 * t2) * t2
            g = np.empty(n)
            g[0:n - n % 2:2] = 2.0 * t1 - (20.0 * a) * t2
            g[1::2] = 10.0 * t2
            f = float(np.cumsum(terms)[-1]) if terms.size else 0.0
            if n % 2:
                t = x[-1] - 2.0
                f += float(t * t)
                g[-1] = 2.0 * t
            return f, g
    return fg


@pytest.fixture(scope="module")
def user_objective_lib(tmp_path_factory):
    out = tmp_path_factory.mktemp("user_objective") / "libuser_objective.so"
    libdir = os.path.join(ROOT, "fortran_library_b200")
    import shutil
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-fmad=false", "-std=c++17", "-Xcompiler",
                    "-fPIC", "-shared", "-o", str(out), os.path.join(ROOT, "tests", "link", "user_objective.cu"),
                    "-L" + libdir, "-lflgpu", "-Xlinker", "-rpath," + libdir], check=True)
    return C.CDLL(str(out))


@pytest.mark.parametrize("which,n", [(0, 1000), (0, 1001), (1, 1000), (1, 777)])
def test_user_objective_header(fl, user_objective_lib, which, n):
    """include/flgpu_objective.cuh: one device functor -> f, fd, f_fd and the fused evaluation.  Gradients bit-exact
    against the NumPy statement of the same formulas, and L-BFGS / CG on it (fused and plain) land where the oracle
    lands with the same objective given as host callbacks."""
    prob = fl.capi.Problem()
    user_objective_lib.user_problem(which, C.byref(prob))
    fg = _np_user_objective(which, n)
    rng = np.random.default_rng(n + which)
    x0 = rng.uniform(-0.5, 1.5, n)
    # kernels vs NumPy
    xd, gd, sc = fl.DeviceVector.from_numpy(x0), fl.DeviceVector(n), fl.DeviceVector(4)
    ctx = fl.capi.EvalCtx(C.c_void_p(prob.user), None, 0, n, 0, 1, 0)
    C.cast(prob.f_fd, fl.capi.F_FD_FN)(C.byref(ctx), sc.ptr, gd.ptr, xd.ptr, n)
    f_ref, g_ref = fg(x0)
    assert np.array_equal(gd.numpy(), g_ref)
    assert abs(sc.numpy()[0] - f_ref) <= 1e-13 * abs(f_ref)
    p = rng.standard_normal(n)
    pd, xo, go = fl.DeviceVector.from_numpy(p), fl.DeviceVector(n), fl.DeviceVector(n)
    W = fl.capi
    C.cast(prob.fused, fl.capi.FUSED_FN)(C.byref(ctx), W.WANT_F | W.WANT_GP | W.WRITE_X | W.WRITE_G, sc.ptr, sc.ptr + 8,
                                         xo.ptr, go.ptr, xd.ptr, pd.ptr, 0.125, n)
    xt = x0 + 0.125 * p
    ft, gt = fg(xt)
    out = sc.numpy()
    assert np.array_equal(xo.numpy(), xt) and np.array_equal(go.numpy(), gt)
    assert abs(out[0] - ft) <= 1e-13 * abs(ft) and abs(out[1] - float(np.dot(gt, p))) <= 1e-12 * float(np.abs(gt * p).sum())
    # optimizers vs the oracle on the same objective through host callbacks
    def cf(fx, xp, dim):
        fx[0] = fg(np.ctypeslib.as_array(C.cast(xp, C.POINTER(C.c_double)), (dim[0],)))[0]

    def cfd(gp, xp, dim):
        np.ctypeslib.as_array(C.cast(gp, C.POINTER(C.c_double)), (dim[0],))[:] = \
            fg(np.ctypeslib.as_array(C.cast(xp, C.POINTER(C.c_double)), (dim[0],)))[1]

    def cffd(fx, gp, xp, dim):
        fv, gv = fg(np.ctypeslib.as_array(C.cast(xp, C.POINTER(C.c_double)), (dim[0],)))
        fx[0] = fv
        np.ctypeslib.as_array(C.cast(gp, C.POINTER(C.c_double)), (dim[0],))[:] = gv
        return 0
    keep = (O.F_T(cf), O.FD_T(cfd), O.FFD_T(cffd))
    cbs = tuple(C.cast(k, C.c_void_p) for k in keep)
    tr = O.Trace(keep_vectors=False)
    xr, sr = O.lbfgs(cbs, x0.copy(), use_ffd=True, Warning=False, MaxIteration=300, trace=tr)
    for fused in (True, False):
        x = x0.copy()
        ob = fl.Observer()
        st = fl.LBFGS(prob, x, Warning=False, MaxIteration=300, observer=ob, fused=fused)
        assert _cases.rel(x, xr) < 1e-8 and st.status == sr.status
        for k in range(5):                                   # same searches at the start: trial counts and f
            assert ob.rows[k][4] == tr.rows[k][4] and abs(ob.rows[k][2] - tr.rows[k][2]) <= 1e-10 * abs(tr.rows[k][2])
    xr, sr = O.cg(cbs, x0.copy(), Method="PR", use_ffd=True, Warning=False, MaxIteration=300)
    x = x0.copy()
    st = fl.ConjugateGradient(prob, x, Method="PR", Warning=False, MaxIteration=300)
    assert _cases.rel(x, xr) < 1e-8
    # the header's device-resident search (one cooperative kernel per line search) == the host-driven fused search
    res = []
    for dev in (False, True):
        x = x0.copy()
        ob = fl.Observer()
        st = fl.LBFGS(prob, x, Memory=6, Warning=False, MaxIteration=80, observer=ob, device_search=dev)
        res.append((x, ob.rows, st.n_trials, st.n_f_fd, st.n_f, st.n_fd, st.host_syncs))
    assert np.array_equal(res[0][0], res[1][0]) and res[0][1] == res[1][1] and res[0][2:6] == res[1][2:6]
    assert res[1][6] < res[0][6]


# ----------------------------------------------------------------------------- AugmentedLagrangian
@pytest.mark.parametrize("solver,kw", [("LBFGS", dict()), ("LBFGS", dict(Memory=5, miu0=4.0, lambda0=[0.3], Increment=1.3)),
                                       ("ConjugateGradient", dict()), ("ConjugateGradient", dict(Method="PR"))])
@pytest.mark.parametrize("n", [10, 4097])
def test_augmented_lagrangian_vs_oracle(fl, solver, kw, n):
    """f = sum x^4 on the unit sphere (the reference's own smoke case, test.f90:466-478, at dim 10 and larger).
    Precision 1e-6: the oracle's outer-iteration count does not depend on its summation order there, so it is
    matched exactly (+-1 for PR, which itself moves 4..6) together with the multiplier growth and the minimiser.
    Precision 1e-10: the count is chaotic in the oracle too (15 / 51 / 27 outer iterations under its three summation
    orders), so only the answer is compared: on the sphere, same point."""
    x0 = _cases.start("quartic", n)
    prob, con = fl.builtin_problem(fl.OBJ_QUARTIC), fl.builtin_constraints()
    xr, sr = O.al(O.builtin_callbacks(O.OBJ_QUARTIC, 0, n), O.sphere_constraint(), x0.copy(), UnconstrainedSolver=solver,
                  use_ffd=True, Warning=False, MaxIteration=60, Precision=1e-6, **kw)
    x = x0.copy()
    st = fl.AugmentedLagrangian(prob, con, x, UnconstrainedSolver=solver, Warning=False, MaxIteration=60, Precision=1e-6,
                                **kw)
    assert st.status == 0 and sr.status == 0 and st.gpu_launches > 0
    slack = 1 if kw.get("Method") == "PR" else 0
    assert abs(st.outer_iterations - sr.outer_iterations) <= slack
    if slack == 0:
        assert st.miu == sr.miu
    # the inner solves stop at |L'| < 1e-6 and the Hessian of L at the solution is 12 x^2 ~ 12/n: the minimiser is
    # determined to ~1e-6 n / 12 only (3e-4 at n = 4097) -- for the oracle's own summation orders as well
    assert abs(np.linalg.norm(x) - 1.0) < 1e-6 and _cases.rel(x, xr) < max(1e-6, 1e-6 * n)
    xr, sr = O.al(O.builtin_callbacks(O.OBJ_QUARTIC, 0, n), O.sphere_constraint(), x0.copy(), UnconstrainedSolver=solver,
                  use_ffd=True, Warning=False, MaxIteration=60, Precision=1e-10, **kw)
    x = x0.copy()
    st = fl.AugmentedLagrangian(prob, con, x, UnconstrainedSolver=solver, Warning=False, MaxIteration=60, Precision=1e-10,
                                **kw)
    assert st.status == 0 and abs(np.linalg.norm(x) - 1.0) < 1e-9      # "norm2(x)-1 should print close to 0"
    assert _cases.rel(x, xr) < max(1e-7, 1e-10 * n)


@pytest.mark.parametrize("solver,kw", [("LBFGS", dict()), ("LBFGS", dict(Memory=5, miu0=4.0, lambda0=[0.3], Increment=1.3)),
                                       ("ConjugateGradient", dict()), ("ConjugateGradient", dict(Method="PR")),
                                       ("LBFGS", dict(line_search="fast"))])
@pytest.mark.parametrize("n", [10, 4097, (1 << 20) + 1])
def test_augmented_lagrangian_fused_probe(fl, solver, kw, n):
    """flgpu_constraints.fused: the inner solves' line searches ask for L(x0+a p) and L'(x0+a p).p as scalars (objective
    probe + constraint probe, nothing stored) instead of materialising the point, c, the Jacobian and L'.  L is the
    same sum; L'.p = f'.p + sum_j (miu c_j - lambda_j)(cd_j.p) is associated differently from dot_product(L', p), so the
    two compositions agree to rounding, not bit for bit.  Small n, Precision 1e-6 (where the unfused composition is
    compared with the oracle, see above): same outer iterations (+-1 where the unfused run itself is on an edge: PR,
    fast policy), same multiplier schedule, same point to the inner tolerance, and the oracle's answer.  n = 2^20+1 (grid
    wraps, odd tail), Precision 1e-10 so that the minimiser is determined (to ~1e-10 n / 12): same point, same f."""
    x0 = _cases.start("quartic", n)
    prob, con = fl.builtin_problem(fl.OBJ_QUARTIC), fl.builtin_constraints()
    assert con.fused
    big = n > 4097
    run_kw = dict(kw, Increment=2.0) if big else dict(kw)
    prec, maxit = (1e-10, 200) if big else (1e-6, 60)
    res = []
    for fused in (True, False):
        x = fl.DeviceVector.from_numpy(x0)
        st = fl.AugmentedLagrangian(prob, con, x, UnconstrainedSolver=solver, Warning=False, MaxIteration=maxit,
                                    Precision=prec, fused=fused, **run_kw)
        res.append((x.numpy(), st))
    (xf, sf), (xp, sp) = res
    msg = (f"outer {sf.outer_iterations}/{sp.outer_iterations} inner {sf.inner_iterations}/{sp.inner_iterations} trials "
           f"{sf.trials}/{sp.trials} f {sf.f!r}/{sp.f!r} |x|-1 {np.linalg.norm(xf) - 1.0:.3e}/{np.linalg.norm(xp) - 1.0:.3e} "
           f"rel {_cases.rel(xf, xp):.3e} status {sf.status}/{sp.status}")
    print(msg)
    assert sf.status == 0 and sp.status == 0, msg
    if big:
        assert abs(np.linalg.norm(xf) - 1.0) < 1e-9 and abs(np.linalg.norm(xp) - 1.0) < 1e-9, msg
        assert _cases.rel(xf, xp) < 1e-3, msg
        assert abs(sf.f - sp.f) <= 1e-6 * abs(sp.f), msg
        return
    slack = 1 if kw.get("Method") == "PR" or kw.get("line_search") == "fast" else 0
    assert abs(sf.outer_iterations - sp.outer_iterations) <= slack, msg
    if sf.outer_iterations == sp.outer_iterations:
        assert sf.miu == sp.miu
    assert abs(np.linalg.norm(xf) - 1.0) < 1e-6
    assert _cases.rel(xf, xp) < max(1e-6, 1e-6 * n), msg
    if "line_search" not in kw:                               # and against the oracle, like the unfused composition
        xr, sr = O.al(O.builtin_callbacks(O.OBJ_QUARTIC, 0, n), O.sphere_constraint(), x0.copy(), UnconstrainedSolver=solver,
                      use_ffd=True, Warning=False, MaxIteration=60, Precision=1e-6, **kw)
        assert abs(sf.outer_iterations - sr.outer_iterations) <= slack
        assert _cases.rel(xf, xr) < max(1e-6, 1e-6 * n)


def test_fortran_abi_augmented_lagrangian(fl):
    """__nonlinearoptimization_MOD_augmentedlagrangian (hpp:369-392) with the built-in CUDA objective and constraint
    in reference-ABI form (device pointers, host scalars), host x in/out; and with plain host callbacks."""
    n = 10
    L = fl.lib()
    f, fd, ffd = fl.capi.REF_F_FN(), fl.capi.REF_FD_FN(), fl.capi.REF_F_FD_FN()
    L.flgpu_builtin_ref_callbacks(fl.OBJ_QUARTIC, C.byref(f), C.byref(fd), C.byref(ffd))
    c, cd = fl.capi.REF_C_FN(), fl.capi.REF_CD_FN()
    L.flgpu_builtin_ref_constraints(fl.CON_SPHERE, C.byref(c), C.byref(cd))
    x0 = _cases.start("quartic", n)
    for solver in (b"LBFGS", b"ConjugateGradient"):
        xr, sr = O.al(O.builtin_callbacks(O.OBJ_QUARTIC, 0, n), O.sphere_constraint(), x0.copy(),
                      UnconstrainedSolver=solver.decode(), use_ffd=True, Warning=False, MaxIteration=60, Precision=1e-6)
        x = x0.copy()
        L.__getattr__("__nonlinearoptimization_MOD_augmentedlagrangian")(
            f, fd, c, cd, x.ctypes.data_as(C.c_void_p), C.byref(C.c_int(n)), C.byref(C.c_int(1)), solver, None, None,
            None, None, None, None, None, ffd, None, C.byref(C.c_int32(0)), C.byref(C.c_int(60)),
            C.byref(C.c_double(1e-6)), None, None, None, None, C.c_int(len(solver)), C.c_int(0))
        st = fl.capi.ALStats()
        L.flgpu_last_al_stats(C.byref(st))
        assert st.status == 0 and st.outer_iterations == sr.outer_iterations
        assert abs(np.linalg.norm(x) - 1.0) < 1e-6 and _cases.rel(x, xr) < 1e-6
    # host callbacks (the reference's own test functions, test.f90:630-705) staged by the library
    def hf(fx, xp, dim):
        v = np.ctypeslib.as_array(C.cast(xp, C.POINTER(C.c_double)), (dim[0],))
        fx[0] = float(np.sum(v ** 4))

    def hfd(gp, xp, dim):
        v = np.ctypeslib.as_array(C.cast(xp, C.POINTER(C.c_double)), (dim[0],))
        np.ctypeslib.as_array(C.cast(gp, C.POINTER(C.c_double)), (dim[0],))[:] = 4.0 * v ** 3

    def hc(cx, xp, M, N):
        v = np.ctypeslib.as_array(C.cast(xp, C.POINTER(C.c_double)), (N[0],))
        cx[0] = float(np.dot(v, v)) - 1.0

    def hcd(cdp, xp, M, N):
        v = np.ctypeslib.as_array(C.cast(xp, C.POINTER(C.c_double)), (N[0],))
        np.ctypeslib.as_array(C.cast(cdp, C.POINTER(C.c_double)), (N[0],))[:] = 2.0 * v
    keep = (fl.capi.REF_F_FN(hf), fl.capi.REF_FD_FN(hfd), fl.capi.REF_C_FN(hc), fl.capi.REF_CD_FN(hcd))
    x = x0.copy()
    L.flgpu_set_callback_space(fl.SPACE_HOST)
    try:
        L.__getattr__("__nonlinearoptimization_MOD_augmentedlagrangian")(
            keep[0], keep[1], keep[2], keep[3], x.ctypes.data_as(C.c_void_p), C.byref(C.c_int(n)), C.byref(C.c_int(1)),
            b"LBFGS", None, None, None, None, None, None, None, None, None, C.byref(C.c_int32(0)),
            C.byref(C.c_int(60)), C.byref(C.c_double(1e-10)), None, None, None, None, C.c_int(5), C.c_int(0))
    finally:
        L.flgpu_set_callback_space(-1)             # back to the automatic choice
    assert abs(np.linalg.norm(x) - 1.0) < 1e-9


def test_cpp_dropin_program_runs(fl, tmp_path):
    """tests/link/cpp_dropin.cpp = the reference's test.cpp optimizer section with host callbacks."""
    exe = tmp_path / "cpp_dropin"
    libdir = os.path.join(ROOT, "fortran_library_b200")
    subprocess.run(["g++", "-std=c++11", os.path.join(ROOT, "tests", "link", "cpp_dropin.cpp"), "-o", str(exe),
                    "-L" + libdir, "-lflgpu", "-Wl,-rpath," + libdir], check=True)
    env = {k: v for k, v in os.environ.items() if k != "FLGPU_CALLBACK_SPACE"}   # host callbacks are the default
    r = subprocess.run([str(exe)], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr


def test_row_sharded_nccl(fl, user_objective_lib):
    """Row-sharded over every visible GPU (one process per GPU, NCCL exchange) against the 1-GPU run.
    Needs >= 2 GPUs; the round-end single-GPU run skips it (the CPU suite covers the same host logic with
    gloo at world_size 2: tests/test_hostsim.py::test_row_sharded_two_ranks_gloo)."""
    ngpu = fl.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={min(ngpu, 8)}",
                        "--master-addr", "127.0.0.1", "--master-port", "29512",
                        os.path.join(ROOT, "tests", "gpu_multi.py"), "20", user_objective_lib._name],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    if "peer-memory exchange: yes" in r.stdout:
        assert "user functor 1" in r.stdout, r.stdout[-3000:]


def test_fortran_abi_device_x_and_stream_sync_mode(fl):
    """Two switches of the drop-in layer: x handed over as a DEVICE pointer (flgpu_set_x_space) gives the same bits
    as host x; FLGPU_SYNC=stream (cudaStreamSynchronize per round trip instead of the polled flag) gives the same bits
    as the default -- both only change how results travel, never what is computed."""
    n = 4097
    L = fl.lib()
    f, fd, ffd = fl.capi.REF_F_FN(), fl.capi.REF_FD_FN(), fl.capi.REF_F_FD_FN()
    L.flgpu_builtin_ref_callbacks(fl.OBJ_ROSENBROCK, C.byref(f), C.byref(fd), C.byref(ffd))
    sym = L.__getattr__("__nonlinearoptimization_MOD_lbfgs")

    def call(xptr):
        sym(f, fd, C.c_void_p(xptr), C.byref(C.c_int(n)), C.byref(C.c_int(7)), ffd, None, C.byref(C.c_int32(0)),
            C.byref(C.c_int(25)), None, None, None, None, None)
    xh = _cases.start("rosenR1", n)
    call(xh.ctypes.data)
    xd = _dev_start(fl, "rosenR1", n)
    L.flgpu_set_x_space(fl.SPACE_DEVICE)
    try:
        call(xd.ptr)
    finally:
        L.flgpu_set_x_space(fl.SPACE_HOST)
    assert np.array_equal(xd.numpy(), xh)
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\nimport numpy as np, fortran_library_b200 as fl\n"
            "x = fl.DeviceVector.start(fl.START_ROSEN_PERT, %d, seed=7)\n"
            "fl.LBFGS(fl.builtin_problem(fl.OBJ_ROSENBROCK), x, Memory=7, Warning=False, MaxIteration=25)\n"
            "np.save(sys.argv[1], x.numpy())\n" % (ROOT, os.path.join(ROOT, "tests"), n))
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, "x.npy")
        r = subprocess.run([sys.executable, "-c", code, out], env=dict(os.environ, FLGPU_SYNC="stream"),
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        assert np.array_equal(np.load(out), xh)


def test_reference_header_program_runs(fl):
    """tests/link/ref_header_prog: user code + the UNMODIFIED reference header (compiled by __graft_entry__.build()
    where the reference tree is mounted), linked against libflgpu.so, run here with host callbacks."""
    exe = os.path.join(ROOT, "tests", "link", "ref_header_prog")
    if not os.path.exists(exe):
        pytest.skip("tests/link/ref_header_prog was not built (reference header not mounted at build time)")
    env = {k: v for k, v in os.environ.items() if k != "FLGPU_CALLBACK_SPACE"}   # host callbacks are the default
    r = subprocess.run([exe], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr


# ----------------------------------------------------------------------------- edge cases
def test_edge_cases(fl):
    for fn in (fl.LBFGS, fl.ConjugateGradient):                 # start at the minimiser (f90:443 / 237)
        x = fl.DeviceVector.from_numpy(np.ones(10))
        st = fn(_problem(fl, "rosenR0"), x, Warning=False)
        assert st.status == fl.INITIAL_CONVERGED and st.iterations == 0 and np.array_equal(x.numpy(), np.ones(10))
    for fn in (fl.LBFGS, fl.ConjugateGradient, fl.SteepestDescent):   # dim = 0: nothing to do, nothing to touch
        st = fn(_problem(fl, "quartic"), np.zeros(0), Warning=False)
        assert st.status == fl.INITIAL_CONVERGED and st.iterations == 0
    for n in (1, 2, 3, 7, 33, 1001):                            # tiny, odd, not a multiple of the vector width
        x0 = _cases.start("quartic", n)
        xs = [O.lbfgs(O.builtin_callbacks(O.OBJ_QUARTIC, 0, n), x0.copy(), Memory=4, Warning=False, MaxIteration=5,
                      sum_mode=mode) for mode in (0, 1, 2)]
        xr, sr = xs[0]
        # Memory = 4 > n makes the pairs linearly dependent: at n = 2 the oracle itself moves by 8e-9 when its dots are
        # summed in long double.  Bound: 1e-10, or 4x the oracle's own spread over its summation orders.
        spread = max(_cases.rel(xs[0][0], xs[1][0]), _cases.rel(xs[2][0], xs[1][0]))
        x = fl.DeviceVector.from_numpy(x0)
        st = fl.LBFGS(_problem(fl, "quartic", False), x, Memory=4, Warning=False, MaxIteration=5)
        err = _cases.rel(x.numpy(), xs[1][0])
        assert st.iterations == sr.n_iter and err <= max(1e-10, _cases.ENV_FACTOR * spread), (n, err, spread)
    # host x (numpy, updated in place) equals device x
    n = 4097
    xh = _cases.start("rosenR1", n)
    fl.LBFGS(_problem(fl, "rosenR1"), xh, Warning=False, MaxIteration=20)
    xd = _dev_start(fl, "rosenR1", n)
    fl.LBFGS(_problem(fl, "rosenR1"), xd, Warning=False, MaxIteration=20)
    assert np.array_equal(xh, xd.numpy())                       # deterministic: bitwise repeatable
    # MaxIteration = 0 runs the steepest-descent step and the Memory-1 pre-iterations only (f90:448-510)
    x = _dev_start(fl, "rosenR1", 100)
    st = fl.LBFGS(_problem(fl, "rosenR1"), x, Memory=4, Warning=False, MaxIteration=0)
    assert st.iterations == 4 and st.status == fl.MAX_ITERATION and st.n_f_fd == 1


def test_memory_above_limit_is_refused_not_fatal(fl):
    """The reference allocates s(dim,0:mem) for any Memory (f90:419-420); the kernels hold FLGPU_MAX_MEMORY = 64 pairs.
    flgpu_lbfgs refuses more with an error code (x untouched, the process lives); the Fortran-ABI symbol, which has no
    error channel, warns and runs with 64."""
    x0 = _cases.start("rosenR1", 100)
    x = fl.DeviceVector.from_numpy(x0)
    with pytest.raises(fl.FlgpuError, match="FLGPU_MAX_MEMORY"):
        fl.LBFGS(fl.builtin_problem(fl.OBJ_ROSENBROCK), x, Memory=65, Warning=False)
    assert np.array_equal(x.numpy(), x0)
    st = fl.LBFGS(fl.builtin_problem(fl.OBJ_ROSENBROCK), x, Memory=64, Warning=False, MaxIteration=5)
    assert st.iterations == 69
    L = fl.lib()
    f, fd, ffd = fl.capi.REF_F_FN(), fl.capi.REF_FD_FN(), fl.capi.REF_F_FD_FN()
    L.flgpu_builtin_ref_callbacks(fl.OBJ_ROSENBROCK, C.byref(f), C.byref(fd), C.byref(ffd))
    xa, xb = x0.copy(), x0.copy()
    sym = L.__getattr__("__nonlinearoptimization_MOD_lbfgs")
    for xv, mem in ((xa, 100), (xb, 64)):
        sym(f, fd, xv.ctypes.data_as(C.c_void_p), C.byref(C.c_int(100)), C.byref(C.c_int(mem)), ffd, None,
            C.byref(C.c_int32(0)), C.byref(C.c_int(5)), None, None, None, None, None)
    assert np.array_equal(xa, xb)


# ----------------------------------------------------------------------------- full size (BASELINE configs[1], [2])
def test_full_size_properties(fl):
    """n = 2^28: size-independent properties.  (a) exact dots of exactly representable data;
    (b) x0 + 0*p == x0; (c) every accepted L-BFGS step satisfies the strong Wolfe conditions it was
    searched for and f decreases monotonically; (d) two runs give identical bits."""
    n = 1 << 28
    L = fl.lib()
    ones = fl.DeviceVector.start(fl.START_ROSEN_STD, n)          # (-1.2, 1, -1.2, 1, ...)
    out = fl.DeviceVector(1)
    L.flgpu_vec_dot(ones.ptr, ones.ptr, n, out.ptr, None)
    assert abs(out.numpy()[0] - (n // 2) * (1.44 + 1.0)) <= 1e-9 * n
    xz, z = fl.DeviceVector(n), fl.DeviceVector(n)
    L.flgpu_memcpy(z.ptr, ones.ptr, n * 8, 1, 1, None)
    L.flgpu_vec_trial(xz.ptr, ones.ptr, z.ptr, 0.0, n, None)
    L.flgpu_vec_dot(xz.ptr, ones.ptr, n, out.ptr, None)
    first = out.numpy()[0]
    L.flgpu_vec_dot(ones.ptr, ones.ptr, n, out.ptr, None)
    assert out.numpy()[0] == first
    for v in (ones, xz, z):
        v.free()

    results = []
    for rep in range(2):
        rows = []

        def on_iter(i, rows=rows):
            gp = fl.DeviceVector(1)
            L.flgpu_vec_dot(i.g_dev, i.p_dev, n, gp.ptr, i.stream)
            L.flgpu_memcpy(None, None, 0, 1, 1, i.stream)
            rows.append((i.f, i.step, i.phid0, gp.numpy()[0]))
            return False
        x = fl.DeviceVector.start(fl.START_ROSEN_PERT, n, seed=7)
        ob = fl.Observer(on_iteration=on_iter)
        st = fl.LBFGS(fl.builtin_problem(fl.OBJ_ROSENBROCK), x, Memory=10, Warning=False, MaxIteration=6, observer=ob)
        assert st.iterations == 16
        L.flgpu_vec_dot(x.ptr, x.ptr, n, out.ptr, None)
        results.append((out.numpy()[0], st.f, st.n_trials, [r[:3] for r in rows]))
        f_prev = None
        for k, (f, a, phid0, gp) in enumerate(rows):
            assert phid0 < 0.0
            if f_prev is not None:
                assert f <= f_prev + 1e-4 * a * phid0 + 1e-9 * abs(f_prev), k     # sufficient decrease, c1 = 1e-4
                assert abs(gp) <= 0.9 * abs(phid0) * (1 + 1e-9), k                # curvature, c2 = 0.9
            f_prev = f
        x.free()
    assert results[0] == results[1], "two identical runs must give identical bits"


def test_beyond_int32_rows(fl):
    """n = 2^31 + 4098 rows on one GPU (the reference's `dim` is a 32-bit integer; the flgpu_* entry points
    take int64): 64-bit indexing in the start, objective, K1, K3 and line-search kernels.  Memory = 1 keeps the work
    space at 7 vectors (112 GiB)."""
    n = (1 << 31) + 4098
    L = fl.lib()
    x = fl.DeviceVector.start(fl.START_QUARTIC_U, n, seed=12345)
    # the start kernel indexes with 64 bits: the tail equals a small vector generated at the same global offset
    tail = fl.copy_to_numpy(x.ptr + 8 * (n - 1000), 1000)
    assert np.array_equal(tail, O.start_vector(O.START_QUARTIC_U, 1000, seed=12345, offset=n - 1000, n_global=n))
    out = fl.DeviceVector(1)
    L.flgpu_vec_dot(x.ptr, x.ptr, n, out.ptr, None)
    xx = out.numpy()[0]
    assert abs(xx / n - 1.0 / 3.0) < 1e-4                      # u ~ U[0,1): E[u^2] = 1/3 over ALL 2^31+ rows
    ob = fl.Observer()
    st = fl.LBFGS(fl.builtin_problem(fl.OBJ_QUARTIC), x, Memory=1, Warning=False, MaxIteration=3, observer=ob)
    assert st.iterations == 4 and st.status == fl.MAX_ITERATION
    fs = [r[2] for r in ob.rows]
    assert all(b < a for a, b in zip(fs, fs[1:]))               # f decreases at every accepted step
    assert abs(fs[0]) < n / 5.0 * 1.01                          # f0 = sum u^4 ~ n/5: every row contributed once
    tail2 = fl.copy_to_numpy(x.ptr + 8 * (n - 1000), 1000)
    assert np.all(np.abs(tail2) < np.abs(tail) + 1e-300) and not np.array_equal(tail2, tail)   # the tail was optimised too
    x.free()


def test_full_size_cg_quartic(fl):
    """BASELINE configs[2]: CG DY and PR+ on the separable quartic, n = 2^28: monotone decrease and the
    closed-form symmetry of the problem (f = sum x^4 decreases; x stays in [0,1))."""
    n = 1 << 28
    for M in ("DY", "PR"):
        x = fl.DeviceVector.start(fl.START_QUARTIC_U, n, seed=12345)
        ob = fl.Observer()
        st = fl.ConjugateGradient(fl.builtin_problem(fl.OBJ_QUARTIC), x, Method=M, Warning=False, MaxIteration=5,
                                  observer=ob)
        fs = [r[2] for r in ob.rows]
        assert st.iterations == 5 and all(b < a for a, b in zip(fs, fs[1:]))
        x.free()
