"""FLGPU_LS_FAST (flgpu_options.line_search = "fast"): the optional accept-at-first-Wolfe-point
searcher.  It is NOT a reference routine -- the reference-exact searchers stay the default and every parity test
elsewhere runs them -- so what is checked here is
  (a) the product's published algorithm (include/flgpu.h, flgpu_search_core.hpp SearchCore::fast) against two
      independent restatements (oracle.c, oracle_np.py): identical bits in 1-D and on small n-D problems;
  (b) the properties that define it: every accepted step satisfies the Wolfe conditions it was searched for, f
      decreases monotonically, the minimiser is the reference policy's, and it needs far fewer objective passes;
  (c) that where the trial point lives (fused / plain / device-resident search) does not change a single bit.
The GPU half (-m gpu) runs the same checks through libflgpu.so."""
import ctypes as C

import numpy as np
import pytest

import _cases
import _hostsim as H
import _oracle as O
import oracle_np as N

capi = H.capi


@pytest.fixture
def np_fast():
    N.LINE_SEARCH_POLICY = 1
    yield
    N.LINE_SEARCH_POLICY = 0


def _np_objective(name, n):
    if name.startswith("rosen"):
        return N.rosenbrock()
    if name == "quartic":
        return N.quartic()
    return N.diagquad(np.array([O.lib().orc_diag_coeff(i, n) for i in range(n)]))


def _same_history(hist, tr):
    assert len(hist) == len(tr.rows)
    for k, (h, r) in enumerate(zip(hist, tr.rows)):
        assert np.array_equal(np.array(h[3:7], dtype=float), np.array(r[1:5], dtype=float), equal_nan=True), \
            f"scalars differ at iteration {k}"
        assert np.array_equal(h[0], tr.p[k], equal_nan=True) and np.array_equal(h[1], tr.x[k], equal_nan=True), \
            f"vectors differ at iteration {k}"


def _same_points(a, b):
    """Trial points as lists of floats; NaN equals NaN (f9_quirk is unbounded below for x < 0: once a step lands there
    every implementation runs off to -inf and NaN in the same way)."""
    return len(a) == len(b) and np.array_equal(np.array(a), np.array(b), equal_nan=True)


def _same_rows(a, b):
    return len(a) == len(b) and all(np.array_equal(np.array(u[1:], dtype=float), np.array(v[1:], dtype=float),
                                                   equal_nan=True) for u, v in zip(a, b))


# ----------------------------------------------------------------------------- (a) the two restatements agree
ND_CASES = [
    ("lbfgs", "rosenR1", 64, dict(Memory=5)), ("lbfgs", "rosenR1", 100, dict(use_ffd=True)),
    ("lbfgs", "rosenR0", 100, dict(use_ffd=True, Memory=3)), ("lbfgs", "quartic", 10, dict()),
    ("lbfgs", "rosenR1", 100, dict(Strong=False, use_ffd=True, MaxIteration=80)),
    ("lbfgs", "diag", 200, dict(Memory=30, use_ffd=True, MaxIteration=80)),
    ("lbfgs", "rosenR1", 100, dict(WolfeConst2=0.1, use_ffd=True)),
    ("cg", "quartic", 10, dict(Method="DY")), ("cg", "quartic", 300, dict(Method="PR", use_ffd=True)),
    ("cg", "rosenR1", 100, dict(Method="DY", use_ffd=True, MaxIteration=150)),
    ("cg", "diag", 200, dict(Method="PR", MaxIteration=100)),
    ("cg", "quartic", 50, dict(Method="DY", Strong=False, use_ffd=True)),
    ("sd", "quartic", 20, dict(MaxIteration=60)), ("sd", "rosenR1", 64, dict(MaxIteration=60, use_ffd=True)),
]


@pytest.mark.parametrize("algo,name,n,kw", ND_CASES)
def test_fast_c_equals_numpy_bitwise(np_fast, algo, name, n, kw):
    kw = dict(kw)
    use = kw.pop("use_ffd", False)
    kind = _cases.OBJECTIVES[name][0]
    x0 = _cases.start(name, n)
    tr = O.Trace()
    crun = {"lbfgs": O.lbfgs, "cg": O.cg, "sd": O.sd}[algo]
    nrun = {"lbfgs": N.lbfgs, "cg": N.conjugate_gradient, "sd": N.steepest_descent}[algo]
    with O.fast_line_search():
        xa, s = crun(O.builtin_callbacks(kind, 0, n), x0.copy(), use_ffd=use, Warning=False, trace=tr, **kw)
    f, fd, ffd = _np_objective(name, n)
    with np.errstate(all="ignore"):
        xb, c = nrun(f, fd, x0.copy(), f_fd=ffd if use else None, Warning=False, **kw)
    _same_history(c.history, tr)
    assert np.array_equal(xa, xb)
    assert c.status == s.status and c.trials == s.n_trials
    assert (c.n_f, c.n_fd, c.n_ffd) == (s.n_f, s.n_fd, s.n_ffd)


# The 1-D torture functions of the reference-policy tests, except the two odd-power walls: those are unbounded below
# for x < 0, a searcher that extrapolates lands there and every implementation then runs off to -inf / NaN, where
# C and Python disagree on fmax(NaN, x).  Their even-power twins (bounded below) take their place.
TORTURE = {k: v for k, v in _cases.TORTURE_1D.items() if k not in ("f9_quirk", "wall")}
TORTURE["f9_even"] = (0.0, _cases._poly_wall(1.05 / 1.01748, 42))
TORTURE["wall_even"] = (0.2, _cases._poly_wall(2.0, 42))
TORTURE["far_start"] = (40.0, (lambda x: 0.5 * (x - 1.0) ** 2 + 0.1 * np.cos(3.0 * x).item(),
                               lambda x: (x - 1.0) - 0.3 * np.sin(3.0 * x).item()))
TORTURE["tiny_first_step"] = (5.0, (lambda x: 1e6 * _cases._pw(x, 2) + 1.0, lambda x: 2e6 * x))


def _run_1d_oracle(case, method, use):
    x0, (f, g) = TORTURE[case]
    fa = _cases.Fuse(f, g)
    cf, cfd, cffd = _cases.make_ref_callbacks(fa.f, fa.g, fa.fg)
    keep = (O.F_T(cf), O.FD_T(cfd), O.FFD_T(cffd))
    tr = O.Trace()
    with O.fast_line_search():
        xa, s = O.cg(tuple(C.cast(k, C.c_void_p) for k in keep), np.array([x0]), Method=method, use_ffd=use,
                     Warning=False, MaxIteration=30, trace=tr)
    return fa, xa, s, tr


@pytest.mark.parametrize("case", sorted(TORTURE))
@pytest.mark.parametrize("method", ["DY", "PR"])
def test_fast_torture_1d_c_equals_numpy(np_fast, case, method):
    """The 1-D functions that steer a searcher through bracketing, both zoom orientations, bisection and NaN regions."""
    x0, (f, g) = TORTURE[case]
    for use in (False, True):
        fa, xa, s, tr = _run_1d_oracle(case, method, use)
        fb = _cases.Fuse(f, g)
        with np.errstate(all="ignore"):
            xb, c = N.conjugate_gradient(lambda x: fb.f(float(x[0])), lambda x: np.array([fb.g(float(x[0]))]),
                                         np.array([x0]), Method=method, Warning=False, MaxIteration=30,
                                         f_fd=(lambda x: (lambda r: (r[0], np.array([r[1]])))(fb.fg(float(x[0]))))
                                         if use else None)
        assert _same_points(fa.xs, fb.xs), "the two restatements evaluated different trial points"
        _same_history(c.history, tr)
        assert np.array_equal(xa, xb, equal_nan=True)


# ----------------------------------------------------------------------------- (a) product host control flow
def _py_problem(fuse):
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


@pytest.mark.parametrize("case", sorted(TORTURE))
@pytest.mark.parametrize("method", ["DY", "PR"])
def test_fast_torture_1d_driver_bitwise_vs_oracle(case, method):
    """driver.cpp + SearchCore::fast over the host simulator: dim = 1 has no summation order, so every trial point,
    step and iterate must equal the oracle's."""
    x0, (f, g) = TORTURE[case]
    for use in (False, True):
        fa, xa, s, tr = _run_1d_oracle(case, method, use)
        fb = _cases.Fuse(f, g)
        prob = _py_problem(fb)
        if not use:
            prob.f_fd = None
        L = H.lib()
        o = capi.Options()
        L.flgpu_hostsim_options_default(C.byref(o), 1)
        capi.apply_options(o, Method=method, Warning=False, MaxIteration=30, line_search="fast")
        ob = H.Observer()
        o.observer = C.cast(ob.cb, C.c_void_p)
        x = np.array([x0])
        st = capi.Stats()
        L.flgpu_hostsim_cg(C.byref(prob), C.byref(o), x.ctypes.data_as(C.c_void_p), C.c_int64(1), C.byref(st))
        assert _same_points(fa.xs, fb.xs), "different trial points"
        assert np.array_equal(x, xa, equal_nan=True)
        assert st.iterations == s.n_iter and st.status == s.status
        assert _same_rows(ob.rows, tr.rows)
        assert (st.n_f, st.n_fd, st.n_f_fd, st.n_trials) == (s.n_f, s.n_fd, s.n_ffd, s.n_trials)


RUNS = {"lbfgs": (H.lbfgs, O.lbfgs), "cg": (H.cg, O.cg), "sd": (H.sd, O.sd)}
TRAJ_CASES = [
    ("lbfgs", "rosenR1", dict(Memory=10)), ("lbfgs", "rosenR1", dict(Memory=3, use_ffd=False)),
    ("lbfgs", "quartic", dict(Memory=5, Strong=False)), ("lbfgs", "diag", dict(Memory=7, MaxIteration=60)),
    ("cg", "quartic", dict(Method="DY")), ("cg", "quartic", dict(Method="PR", use_ffd=False)),
    ("cg", "rosenR1", dict(Method="DY", Strong=False, MaxIteration=80)),
    ("sd", "quartic", dict(MaxIteration=50)), ("sd", "rosenR1", dict(MaxIteration=50, Strong=False)),
]


@pytest.mark.parametrize("algo,name,kw", TRAJ_CASES)
def test_fast_trajectory_within_oracle_envelope(algo, name, kw):
    """n-D: the driver's first 20 directions against the oracle's fast-policy run, inside the oracle's own
    summation-order envelope (the same criterion the reference-exact policy is held to)."""
    n = 2000
    kw = dict(kw)
    use = kw.pop("use_ffd", True)
    kind = _cases.OBJECTIVES[name][0]
    hrun, orun = RUNS[algo]
    with O.fast_line_search():
        traces, _ = _cases.oracle_envelope(name, n, lambda cbs, x, **k: orun(cbs, x, use_ffd=use, **k), **kw)
    ob = H.Observer(max_vec_iters=20)
    # plain mode: the oracle has no fused evaluation, so evaluation counters are comparable as well
    x, st = hrun(kind, _cases.start(name, n), observer=ob, use_ffd=use, Warning=False, n_global=n, fused=False,
                 line_search="fast", **kw)
    _cases.check_envelope(traces, ob.p, f"fast {algo} {name} {kw}")
    assert [r[4] for r in ob.rows[:5]] == [r[4] for r in traces[0].rows[:5]]     # trial counts of the first searches


# ----------------------------------------------------------------------------- (c) fused = plain = device-resident
@pytest.mark.parametrize("algo,name,kw", TRAJ_CASES)
def test_fast_fused_plain_and_device_search_identical(algo, name, kw):
    n = 3001
    kw = dict(kw)
    use = kw.pop("use_ffd", True)
    kind = _cases.OBJECTIVES[name][0]
    out = []
    for fused, dsearch in ((True, False), (False, False), (True, True)):
        ob = H.Observer(max_vec_iters=10**9)
        x, st = RUNS[algo][0](kind, _cases.start(name, n), observer=ob, Warning=False, n_global=n, fused=fused,
                              device_search=dsearch, use_ffd=use, line_search="fast", **kw)
        out.append((x, st, ob))
    (xa, sa, oa) = out[0]
    for tag, (xb, sb, ob_) in (("plain", out[1]), ("device-search", out[2])):
        assert np.array_equal(xa, xb), tag
        assert oa.rows == ob_.rows, tag
        assert all(np.array_equal(u, v) for u, v in zip(oa.p, ob_.p)), tag
        assert all(np.array_equal(u, v) for u, v in zip(oa.x, ob_.x)), tag
        assert all(np.array_equal(u, v) for u, v in zip(oa.g, ob_.g)), tag
        assert (sa.iterations, sa.status, sa.n_trials, sa.n_linesearch) == \
               (sb.iterations, sb.status, sb.n_trials, sb.n_linesearch), tag
        # one evaluation of f and f' per trial on every path; a fused probe counts as f_fd, separate callbacks as f + fd
        assert sa.n_f + sa.n_f_fd == sb.n_f + sb.n_f_fd and sa.n_fd + sa.n_f_fd == sb.n_fd + sb.n_f_fd, tag
    if algo == "lbfgs":
        # the first trial of every search rides on the speculative K1->K2->K3 chain and is usually accepted: about one
        # host round trip per iteration, which is why device_search = auto stays off under this policy
        assert sa.host_syncs < 1.6 * sa.iterations + 20


# ----------------------------------------------------------------------------- (b) defining properties
def check_wolfe_rows(rows, ps, gs, f_start, c1, c2, strong, what, rtol=1e-9):
    """rows[k] = (iteration, step, f, phid0, trials), ps/gs = direction searched / gradient accepted.  Every accepted
    step must satisfy sufficient decrease and the curvature condition for (c1, c2); f must not increase."""
    f_prev, checked = f_start, 0
    for k, (row, p, g) in enumerate(zip(rows, ps, gs)):
        _, a, f, phid0, trials = row
        if f_prev <= 1e-16 * abs(f_start):
            break            # the tail where f is rounding noise (x = x* to 1e-13): searches end on their safeguards
        checked += 1
        assert phid0 < 0.0
        slack = rtol * (abs(f_prev) + abs(f)) + 1e-300
        assert f <= f_prev + c1 * a * phid0 + slack, f"{what}: step {k} violates sufficient decrease"
        gp = float(np.dot(g, p))
        cslack = rtol * float(np.linalg.norm(g) * np.linalg.norm(p)) + 1e-300
        if strong:
            assert abs(gp) <= c2 * abs(phid0) + cslack, f"{what}: step {k} violates strong curvature ({gp} vs {phid0})"
        else:
            assert gp >= -c2 * abs(phid0) - cslack, f"{what}: step {k} violates weak curvature"
        f_prev = f
    return checked


def _objective_value(kind, x):
    fo = C.c_double()
    n = x.size
    O.lib().orc_obj_select(kind, 0, n)
    O.lib().orc_obj_f(C.byref(fo), x.ctypes.data_as(C.c_void_p), C.byref(C.c_int(n)))
    return fo.value


@pytest.mark.parametrize("algo,name,kw,c2", [
    ("lbfgs", "rosenR1", dict(Memory=10), 0.9), ("lbfgs", "rosenR1", dict(Memory=5, WolfeConst2=0.1), 0.1),
    ("lbfgs", "quartic", dict(Memory=5, Strong=False), 0.9), ("lbfgs", "diag", dict(Memory=30, MaxIteration=300), 0.9),
    ("cg", "quartic", dict(Method="DY"), 0.45), ("cg", "rosenR1", dict(Method="PR", MaxIteration=400), 0.45),
    ("cg", "quartic", dict(Method="DY", Strong=False), 0.45), ("sd", "quartic", dict(MaxIteration=80), 0.9),
])
def test_fast_steps_satisfy_the_wolfe_conditions(algo, name, kw, c2):
    n = 2000
    kind = _cases.OBJECTIVES[name][0]
    x0 = _cases.start(name, n)
    ob = H.Observer()
    x, st = RUNS[algo][0](kind, x0, observer=ob, Warning=False, n_global=n, line_search="fast", **kw)
    strong = kw.get("Strong", True) or kw.get("Method") == "PR"
    rows = ob.rows
    # the run may end on a collapsed bracket (step-length convergence): the last step is then exempt
    if st.status == capi.STEP_CONVERGED:
        rows = rows[:-1]
    assert check_wolfe_rows(rows, ob.p, ob.g, _objective_value(kind, x0), 1e-4, c2, strong, f"{algo} {name}") > 3


@pytest.mark.parametrize("name,mem", [("rosenR0", 10), ("rosenR1", 10), ("rosenR1", 5)])
def test_fast_reaches_the_reference_minimiser_with_fewer_passes(name, mem):
    """Same minimiser as the reference-exact policy (north_star's 1e-8), at a fraction of the objective passes."""
    n = 2000
    kind = _cases.OBJECTIVES[name][0]
    xr, sr = H.lbfgs(kind, _cases.start(name, n), Warning=False, n_global=n, Memory=mem)
    xf, sf = H.lbfgs(kind, _cases.start(name, n), Warning=False, n_global=n, Memory=mem, line_search="fast")
    assert sf.status in (capi.CONVERGED, capi.STEP_CONVERGED)
    assert _cases.rel(xf, xr) < 1e-8 and np.abs(xf - 1.0).max() < 1e-8
    # one f and f' per trial: trials + the initial evaluation
    assert sf.n_f_fd == sf.n_trials + 1 and sf.n_f == 0 and sf.n_fd == 0
    assert sf.n_trials / sf.iterations < 2.0 < sr.n_trials / sr.iterations
    assert sf.n_trials < 0.5 * sr.n_trials, (sf.n_trials, sr.n_trials)


def test_fast_default_is_reference_and_increment_is_ignored():
    n = 500
    kind = _cases.OBJECTIVES["rosenR1"][0]
    xa, sa = H.lbfgs(kind, _cases.start("rosenR1", n), Warning=False, n_global=n, MaxIteration=30)
    xb, sb = H.lbfgs(kind, _cases.start("rosenR1", n), Warning=False, n_global=n, MaxIteration=30, line_search="reference")
    assert np.array_equal(xa, xb) and sa.n_trials == sb.n_trials
    xc, sc = H.lbfgs(kind, _cases.start("rosenR1", n), Warning=False, n_global=n, MaxIteration=30, line_search="fast")
    xd, sd_ = H.lbfgs(kind, _cases.start("rosenR1", n), Warning=False, n_global=n, MaxIteration=30, line_search="fast",
                      Increment=3.0)
    assert np.array_equal(xc, xd) and sc.n_trials == sd_.n_trials
    assert not np.array_equal(xa, xc)


def test_fast_nan_objective_terminates():
    """A NaN objective value counts as insufficient decrease: the bracket shrinks away from it and the search ends."""
    fa = _cases.Fuse(*_cases.TORTURE_1D["nan_region"][1], limit=10**9)
    prob = _py_problem(fa)
    L = H.lib()
    o = capi.Options()
    L.flgpu_hostsim_options_default(C.byref(o), 1)
    capi.apply_options(o, Warning=False, MaxIteration=50, line_search="fast")
    x = np.array([_cases.TORTURE_1D["nan_region"][0]])
    st = capi.Stats()
    L.flgpu_hostsim_cg(C.byref(prob), C.byref(o), x.ctypes.data_as(C.c_void_p), C.c_int64(1), C.byref(st))
    assert abs(x[0]) < 1e-7 and st.n_trials < 400


def test_fast_edge_cases():
    """Start at the minimiser, tiny dimensions, Memory = 1 and more pairs than dimensions: same exits as the reference
    policy (f90:443 / 237 initial test) and agreement with the oracle's fast-policy run."""
    for fn in (H.lbfgs, H.cg, H.sd):
        x, st = fn(O.OBJ_ROSENBROCK, np.ones(10), Warning=False, n_global=10, line_search="fast")
        assert st.status == capi.INITIAL_CONVERGED and st.iterations == 0 and np.array_equal(x, np.ones(10))
    for n in (1, 2, 3, 7):
        x0 = _cases.start("quartic", n)
        with O.fast_line_search():
            xa, sa = O.lbfgs(O.builtin_callbacks(O.OBJ_QUARTIC, 0, n), x0.copy(), Memory=4, Warning=False, MaxIteration=5)
        xb, stb = H.lbfgs(O.OBJ_QUARTIC, x0, Memory=4, Warning=False, MaxIteration=5, n_global=n, use_ffd=False,
                          line_search="fast")
        assert stb.iterations == sa.n_iter and _cases.rel(xb, xa) < 1e-6
    n = 400
    x0 = _cases.start("rosenR1", n)
    with O.fast_line_search():
        xa, sa = O.lbfgs(O.builtin_callbacks(O.OBJ_ROSENBROCK, 0, n), x0.copy(), Memory=1, use_ffd=True, Warning=False,
                         MaxIteration=25)
    xb, stb = H.lbfgs(O.OBJ_ROSENBROCK, x0, Memory=1, Warning=False, MaxIteration=25, n_global=n, line_search="fast")
    assert stb.iterations == sa.n_iter and stb.n_trials == sa.n_trials and _cases.rel(xb, xa) < 1e-9


# ----------------------------------------------------------------------------- known-answer vectors (CPU)
import glob            # noqa: E402
import json            # noqa: E402
import os              # noqa: E402

GOLDEN_FAST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_fast")
GOLDEN_FAST_NAMES = sorted(os.path.splitext(os.path.basename(q))[0] for q in glob.glob(os.path.join(GOLDEN_FAST, "*.json")))
# runs that end in a long tail of rounding-noise steps: only the early iterations and the exit status are compared
GOLDEN_FAST_TAIL = {"fast_lbfgs_rosenR1_64_m5_c2_01", "fast_lbfgs_diag_60_m30", "fast_cg_dy_rosenR1_64", "fast_sd_quartic10",
                    # CG creeps into the quartic's flat bottom (|x| ~ 1e-6 at exit): the last iterates move by 1e-6 under
                    # summation-order noise, 1e-7 of |x0|
                    "fast_cg_pr_quartic_200"}


def _load_golden_fast(name):
    with open(os.path.join(GOLDEN_FAST, name + ".json")) as fh:
        d = json.load(fh)
    d["x0"] = np.array([float.fromhex(v) for v in d["x0"]])
    d["x_final"] = np.array([float.fromhex(v) for v in d["x_final"]])
    d["rows"] = [(r[0], float.fromhex(r[1]), float.fromhex(r[2]), float.fromhex(r[3]), r[4]) for r in d["rows"]]
    d["p_first"] = [np.array([float.fromhex(v) for v in q]) for q in d["p_first"]]
    return d


def test_golden_fast_vectors_exist():
    assert len(GOLDEN_FAST_NAMES) >= 9


@pytest.mark.parametrize("name", GOLDEN_FAST_NAMES)
def test_fast_oracle_reproduces_golden_bitwise(name):
    """tests/golden_fast/*.json (made by make_golden_fast.py): the oracle's fast-policy restatement must reproduce every
    stored bit -- regression protection for the algorithm itself."""
    d = _load_golden_fast(name)
    kw = dict(d["options"])
    use = kw.pop("use_ffd", False)
    kind = _cases.OBJECTIVES[d["objective"]][0]
    assert np.array_equal(_cases.start(d["objective"], d["n"]), d["x0"])
    tr = O.Trace()
    with O.fast_line_search():
        x, st = RUNS[d["algorithm"]][1](O.builtin_callbacks(kind, 0, d["n"]), d["x0"].copy(), use_ffd=use, Warning=False,
                                        trace=tr, **kw)
    assert np.array_equal(x, d["x_final"])
    assert (st.n_iter, st.status, st.n_f, st.n_fd, st.n_ffd, st.n_trials) == \
        (d["iterations"], d["status"], d["n_f"], d["n_fd"], d["n_ffd"], d["n_trials"])
    assert [tuple(r) for r in tr.rows] == d["rows"]
    for q, gq in zip(tr.p, d["p_first"]):
        assert np.array_equal(q, gq)


@pytest.mark.parametrize("fused", [True, False], ids=["fused", "plain"])
@pytest.mark.parametrize("name", GOLDEN_FAST_NAMES)
def test_fast_host_control_flow_lands_on_golden(name, fused):
    """driver.cpp + SearchCore::fast over the host simulator from the stored starts: same trial counts and (to 1e-9) the
    same steps over the first 8 iterations, same exit status; minimiser to 1e-8 and iteration count to 2 % where the
    stored run does not end in a rounding-noise tail (the criteria tests/test_golden.py applies to the reference policy)."""
    d = _load_golden_fast(name)
    kw = dict(d["options"])
    use = kw.pop("use_ffd", False)
    kind = _cases.OBJECTIVES[d["objective"]][0]
    ob = H.Observer(keep_vectors=False)
    x, st = RUNS[d["algorithm"]][0](kind, d["x0"], observer=ob, use_ffd=use, Warning=False, n_global=d["n"], fused=fused,
                                    line_search="fast", **kw)
    assert st.status == d["status"]
    for k in range(min(8, len(ob.rows), len(d["rows"]))):
        (_, a, f, phid0, trials), (_, ga, gf, gphid0, gtrials) = ob.rows[k], d["rows"][k]
        assert trials == gtrials, f"{name}: iteration {k} took {trials} trials, stored {gtrials}"
        assert abs(a - ga) <= 1e-9 * abs(ga) and abs(f - gf) <= 1e-9 * abs(gf) + 1e-300
        assert abs(phid0 - gphid0) <= 1e-8 * abs(gphid0)
    if name in GOLDEN_FAST_TAIL:
        return
    scale = max(np.linalg.norm(d["x_final"]), np.linalg.norm(d["x0"]))
    assert np.linalg.norm(x - d["x_final"]) <= 1e-8 * scale
    assert abs(st.iterations - d["iterations"]) <= max(1, 0.02 * d["iterations"])


# ============================================================================= GPU half (libflgpu.so, -m gpu)
@pytest.fixture(scope="module")
def fl():
    import fortran_library_b200 as fl
    fl.require_gpu()          # fails loudly: there is no fallback to test instead
    return fl


def _gpu_problem(fl, name, use_ffd=True):
    p = fl.builtin_problem(_cases.OBJECTIVES[name][0])
    if not use_ffd:
        p.f_fd = None
    return p


def _gpu_start(fl, name, n):
    kind, st, seed = _cases.OBJECTIVES[name]
    return fl.DeviceVector.start(st, n, seed=seed)


def _gpu_run(fl, algo):
    return {"lbfgs": fl.LBFGS, "cg": fl.ConjugateGradient, "sd": fl.SteepestDescent}[algo]


@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(TORTURE))
@pytest.mark.parametrize("method", ["DY", "PR"])
def test_gpu_fast_fortran_abi_torture_1d_bitwise(fl, case, method):
    """The reference's own symbol with the policy set through flgpu_set_line_search (the Fortran signature has no room
    for it), host callbacks staged by the library: every trial point and the result equal the oracle's fast-policy
    run bit for bit."""
    x0, (f, g) = TORTURE[case]
    L = fl.lib()
    for use in (False, True):
        fa, xa, s, tr = _run_1d_oracle(case, method, use)
        fb = _cases.Fuse(f, g)
        cf, cfd, cffd = _cases.make_ref_callbacks(fb.f, fb.g, fb.fg)
        keep = (fl.capi.REF_F_FN(cf), fl.capi.REF_FD_FN(cfd), fl.capi.REF_F_FD_FN(cffd))
        x = np.array([x0])
        m = method.encode()
        L.flgpu_set_callback_space(fl.SPACE_HOST)
        L.flgpu_set_line_search(fl.LS_FAST)
        try:
            L.__getattr__("__nonlinearoptimization_MOD_conjugategradient")(
                keep[0], keep[1], x.ctypes.data_as(C.c_void_p), C.byref(C.c_int(1)), m, keep[2] if use else None,
                None, C.byref(C.c_int32(0)), C.byref(C.c_int(30)), None, None, None, None, None, C.c_int(len(m)))
        finally:
            L.flgpu_set_callback_space(-1)             # back to the automatic choice
            L.flgpu_set_line_search(-1)
        st = fl.capi.Stats()
        L.flgpu_last_stats(C.byref(st))
        assert _same_points(fa.xs, fb.xs), "different trial points"
        assert np.array_equal(x, xa, equal_nan=True)
        assert st.iterations == s.n_iter and st.status == s.status
        assert (st.n_f, st.n_fd, st.n_f_fd, st.n_trials) == (s.n_f, s.n_fd, s.n_ffd, s.n_trials)


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [True, False], ids=["fused", "plain"])
@pytest.mark.parametrize("algo,name,kw", TRAJ_CASES)
def test_gpu_fast_trajectory_within_oracle_envelope(fl, algo, name, kw, fused):
    n = 10_000
    kw = dict(kw)
    use = kw.pop("use_ffd", True)
    orun = RUNS[algo][1]
    with O.fast_line_search():
        traces, _ = _cases.oracle_envelope(name, n, lambda cbs, x, **k: orun(cbs, x, use_ffd=use, **k), **kw)
    ob = fl.Observer(keep_vectors=True, max_vec_iters=20)
    x = _gpu_start(fl, name, n)
    st = _gpu_run(fl, algo)(_gpu_problem(fl, name, use), x, observer=ob, Warning=False, fused=fused, line_search="fast", **kw)
    _cases.check_envelope(traces, ob.p, f"gpu fast {algo} {name} {kw} fused={fused}")
    assert ob.rows[0][4] == traces[0].rows[0][4]          # first search: same trial count as the oracle's
    assert st.gpu_launches > 0


@pytest.mark.gpu
@pytest.mark.parametrize("n", [10_000, 4097])
@pytest.mark.parametrize("algo,name,kw", TRAJ_CASES)
def test_gpu_fast_device_resident_search_identical(fl, algo, name, kw, n):
    """search_kernel<KIND, FAST = true>: SearchCore::fast run by every thread of one cooperative kernel gives the bits
    of the host-driven fused search (same launch geometry, same reductions)."""
    kw = dict(kw)
    use = kw.pop("use_ffd", True)
    out = []
    for dev in (False, True):
        x = _gpu_start(fl, name, n)
        ob = fl.Observer(keep_vectors=True, max_vec_iters=12)
        st = _gpu_run(fl, algo)(_gpu_problem(fl, name, use), x, observer=ob, Warning=False, device_search=dev,
                                line_search="fast", **kw)
        out.append((x.numpy(), st, ob))
    (xa, sa, oa), (xb, sb, ob_) = out
    assert oa.rows == ob_.rows
    assert np.array_equal(xa, xb)
    assert all(np.array_equal(u, v) for u, v in zip(oa.p, ob_.p))
    for k in ("iterations", "status", "n_f", "n_fd", "n_f_fd", "n_trials", "n_linesearch"):
        assert getattr(sa, k) == getattr(sb, k), k


@pytest.mark.gpu
@pytest.mark.parametrize("algo,name,kw,c2", [
    ("lbfgs", "rosenR1", dict(Memory=10), 0.9), ("lbfgs", "rosenR1", dict(Memory=5, WolfeConst2=0.1), 0.1),
    ("lbfgs", "quartic", dict(Memory=5, Strong=False), 0.9), ("cg", "quartic", dict(Method="DY"), 0.45),
    ("cg", "rosenR1", dict(Method="PR", MaxIteration=400), 0.45), ("sd", "quartic", dict(MaxIteration=80), 0.9),
])
def test_gpu_fast_steps_satisfy_the_wolfe_conditions(fl, algo, name, kw, c2):
    n = 10_000
    kind = _cases.OBJECTIVES[name][0]
    ob = fl.Observer(keep_vectors=True)
    x = _gpu_start(fl, name, n)
    st = _gpu_run(fl, algo)(_gpu_problem(fl, name), x, observer=ob, Warning=False, line_search="fast", **kw)
    strong = kw.get("Strong", True) or kw.get("Method") == "PR"
    rows = ob.rows[:-1] if st.status == fl.STEP_CONVERGED else ob.rows
    assert check_wolfe_rows(rows, ob.p, ob.g, _objective_value(kind, _cases.start(name, n)), 1e-4, c2, strong,
                            f"gpu {algo} {name}") > 3


@pytest.mark.gpu
def test_gpu_fast_reaches_the_reference_minimiser_with_fewer_passes(fl):
    n = 10_000
    for name in ("rosenR0", "rosenR1"):
        xr = _gpu_start(fl, name, n)
        sr = fl.LBFGS(_gpu_problem(fl, name), xr, Warning=False)
        xf = _gpu_start(fl, name, n)
        sf = fl.LBFGS(_gpu_problem(fl, name), xf, Warning=False, line_search="fast")
        assert _cases.rel(xf.numpy(), xr.numpy()) < 1e-8 and np.abs(xf.numpy() - 1.0).max() < 1e-8
        assert sf.n_f_fd == sf.n_trials + 1 and sf.n_f == 0 and sf.n_fd == 0
        assert sf.n_trials / sf.iterations < 2.0 < sr.n_trials / sr.iterations
        # the first trial rides on the speculative K1 -> K2 -> K3 chain: about one host round trip per iteration
        assert sf.host_syncs < 1.6 * sf.iterations + 20


@pytest.mark.gpu
@pytest.mark.parametrize("solver,kw", [("LBFGS", dict()), ("ConjugateGradient", dict())])
@pytest.mark.parametrize("n", [10, 4097])
def test_gpu_fast_augmented_lagrangian(fl, solver, kw, n):
    """AugmentedLagrangian (f90:2005-2241) hands `inner.line_search` to every inner solve.  sum x^4 on the unit sphere has
    2^n equivalent minimisers (|x_i| = n^-1/2), so the comparison with the oracle's fast-policy run is on what they share:
    the constraint, the objective value 1/n and the outer-iteration count."""
    x0 = _cases.start("quartic", n)
    prob, con = fl.builtin_problem(fl.OBJ_QUARTIC), fl.builtin_constraints()
    with O.fast_line_search():
        xr, sr = O.al(O.builtin_callbacks(O.OBJ_QUARTIC, 0, n), O.sphere_constraint(), x0.copy(),
                      UnconstrainedSolver=solver, use_ffd=True, Warning=False, MaxIteration=60, Precision=1e-6, **kw)
    x = x0.copy()
    st = fl.AugmentedLagrangian(prob, con, x, UnconstrainedSolver=solver, Warning=False, MaxIteration=60, Precision=1e-6,
                                line_search="fast", **kw)
    assert st.status == 0 and sr.status == 0 and st.gpu_launches > 0
    assert abs(np.linalg.norm(x) - 1.0) < 1e-6
    assert abs(float(np.sum(x ** 4)) - float(np.sum(xr ** 4))) < 1e-5 / n
    assert abs(st.outer_iterations - sr.outer_iterations) <= 1
    # far fewer objective passes than the reference policy needs for the same job
    xs = x0.copy()
    ss = fl.AugmentedLagrangian(prob, con, xs, UnconstrainedSolver=solver, Warning=False, MaxIteration=60, Precision=1e-6,
                                **kw)
    assert st.trials < 0.5 * ss.trials


# ============================================================================= property test (CPU)
from hypothesis import HealthCheck, given, settings      # noqa: E402
from hypothesis import strategies as hst                 # noqa: E402
from test_property_1d import _objective                  # noqa: E402  (the objective families of the reference-policy property test)

FAST_COUNTS = {"compared": 0, "skipped": 0}


@pytest.mark.timeout(300)
@settings(max_examples=300, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.too_slow])
@given(algo=hst.sampled_from(["cg", "sd"]), kind=hst.sampled_from(["quartic", "steep", "cosh", "well"]),
       p=hst.floats(0.1, 10.0), q=hst.floats(-2.0, 2.0), r=hst.floats(0.01, 3.0), x0=hst.floats(-3.0, 3.0),
       method=hst.sampled_from(["DY", "PR"]), strong=hst.booleans(), use=hst.booleans(),
       c1=hst.floats(1e-6, 0.3), c2frac=hst.floats(0.05, 0.95), fused=hst.booleans())
def test_fast_property_1d_driver_bitwise_vs_oracle(algo, kind, p, q, r, x0, method, strong, use, c1, c2frac, fused):
    """Random objective families, starts and tunables in ONE dimension (no summation order): driver.cpp + SearchCore::fast
    over the host simulator must reproduce the oracle's fast-policy run bit for bit -- every trial point, accepted step,
    evaluation counter and the result -- and every accepted step must satisfy the Wolfe conditions it was searched for."""
    f, g = _objective(kind, p, q, r)
    c2 = c1 + c2frac * (0.99 - c1)
    opts = dict(Strong=strong, Warning=False, MaxIteration=12, WolfeConst1=c1, WolfeConst2=c2)
    fa = _cases.Fuse(f, g, limit=10**9)
    cf, cfd, cffd = _cases.make_ref_callbacks(fa.f, fa.g, fa.fg)
    keep = (O.F_T(cf), O.FD_T(cfd), O.FFD_T(cffd))
    cbs = tuple(C.cast(k, C.c_void_p) for k in keep)
    tr = O.Trace()
    with np.errstate(all="ignore"), O.fast_line_search(), O.eval_budget(20000):
        if algo == "cg":
            xa, s = O.cg(cbs, np.array([x0]), Method=method, use_ffd=use, trace=tr, **opts)
        else:
            xa, s = O.sd(cbs, np.array([x0]), use_ffd=use, trace=tr, **opts)
    assert s.status != 9, "the fast searcher is bounded: 100 trials per search at most"
    if any(not np.isfinite(v) for v in fa.xs):
        FAST_COUNTS["skipped"] += 1
        return                    # a step ran off to inf / NaN (the guarded families saturate at 1e300)
    FAST_COUNTS["compared"] += 1
    fb = _cases.Fuse(f, g, limit=10**9)
    prob = _py_problem(fb)
    if not use:
        prob.f_fd = None
    L = H.lib()
    o = capi.Options()
    L.flgpu_hostsim_options_default(C.byref(o), int(algo == "cg"))
    capi.apply_options(o, Method=method if algo == "cg" else None, line_search="fast", **opts)
    o.no_fused = int(not fused)     # _py_problem supplies no fused callback: both settings must take the plain path
    ob = H.Observer()
    o.observer = C.cast(ob.cb, C.c_void_p)
    x = np.array([x0])
    stt = capi.Stats()
    fn = L.flgpu_hostsim_cg if algo == "cg" else L.flgpu_hostsim_sd
    fn(C.byref(prob), C.byref(o), x.ctypes.data_as(C.c_void_p), C.c_int64(1), C.byref(stt))
    assert _same_points(fa.xs, fb.xs), "different trial points"
    assert np.array_equal(x, xa, equal_nan=True)
    assert stt.iterations == s.n_iter and stt.status == s.status
    assert _same_rows(ob.rows, tr.rows)
    assert (stt.n_f, stt.n_fd, stt.n_f_fd, stt.n_trials) == (s.n_f, s.n_fd, s.n_ffd, s.n_trials)
    assert max(r_[4] for r_ in ob.rows) <= 101 if ob.rows else True
    # Wolfe conditions of every accepted step, from the observer's exact scalars (1-D: phi'(a) = g(x) * p)
    is_strong = strong or (algo == "cg" and method == "PR")
    f_prev = f(x0)
    for k, (row, pk, gk) in enumerate(zip(ob.rows, ob.p, ob.g)):
        _, a, fk, phid0, trials = row
        if trials >= 40 or not (phid0 < 0.0) or abs(a * phid0) <= 1e-13 * abs(f_prev):
            break                 # a safeguard exit (growth / zoom cap, collapsed bracket) or the rounding floor
        assert fk <= f_prev + c1 * a * phid0
        gp = float(gk[0] * pk[0])
        assert (abs(gp) <= c2 * abs(phid0)) if is_strong else (gp >= -c2 * abs(phid0))
        f_prev = fk


def test_fast_property_cases_are_not_vacuous():
    total = FAST_COUNTS["compared"] + FAST_COUNTS["skipped"]
    assert total == 0 or FAST_COUNTS["compared"] >= 0.7 * total, FAST_COUNTS
