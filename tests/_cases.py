"""Shared test cases: the synthetic objectives of SURVEY.md 8(d) and 1-D line-search torture functions."""
import ctypes as C

import numpy as np

import _oracle as O

# (objective kind, start kind, seed, name)
OBJECTIVES = {
    "rosenR0": (O.OBJ_ROSENBROCK, O.START_ROSEN_STD, 0),
    "rosenR1": (O.OBJ_ROSENBROCK, O.START_ROSEN_PERT, 7),
    "quartic": (O.OBJ_QUARTIC, O.START_QUARTIC_U, 12345),
    "diag": (O.OBJ_DIAGQUAD, O.START_ZERO, 0),
    # sum (x-1)^4 + (x-1)^2 from u in [0,1): x* = 1, Hessian 2 at the minimiser -- the CG case with a well-defined minimiser
    "quartic1": (O.OBJ_QUARTIC_SHIFTED, O.START_QUARTIC_U, 12345),
}


def start(name, n):
    kind, st, seed = OBJECTIVES[name]
    return O.start_vector(st, n, seed=seed)


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(np.asarray(b)))


ENV_FACTOR = 4.0     # allowed multiple of the oracle's own summation-order noise (measured on B200 over every case of
                     # the GPU suite, tests/gpu_calibrate.py: worst ratio 1.66, typical 0.0-0.5)
FLOOR = 1e-12        # north_star tolerance, asserted wherever that noise is below it


def check_envelope(traces, p_got, what):
    """traces = oracle runs under (sequential, long double, pairwise) sums.  Our deviation from the
    long-double run must stay within ENV_FACTOR x the running maximum of the sequential / pairwise
    runs' own deviation from it (or within 1e-12)."""
    ref, ld = 0.0, traces[1].p
    for k in range(min(len(p_got), 20, *(len(t.p) for t in traces))):
        ref = max(ref, rel(traces[0].p[k], ld[k]), rel(traces[2].p[k], ld[k]))
        err = rel(p_got[k], ld[k])
        assert err <= max(FLOOR, ENV_FACTOR * ref), f"{what}: direction {k} off by {err:.2e} (oracle noise {ref:.2e})"


def two_loop_extended(pairs, g):
    """The reference's two-loop recursion (f90:589-607) in 80-bit extended precision.
    pairs: [(s, y)] oldest -> newest.  Used as the 'exact' one-step direction."""
    L = np.longdouble
    S = [np.asarray(s, dtype=L) for s, _ in pairs]
    Y = [np.asarray(y, dtype=L) for _, y in pairs]
    rho = [L(1) / np.dot(y, s) for s, y in zip(S, Y)]
    q = np.asarray(g, dtype=L).copy()
    alpha = [None] * len(S)
    for i in range(len(S) - 1, -1, -1):
        alpha[i] = rho[i] * np.dot(S[i], q)
        q = q - alpha[i] * Y[i]
    q = q / rho[-1] / np.dot(Y[-1], Y[-1])
    for i in range(len(S)):
        beta = rho[i] * np.dot(Y[i], q)
        q = q + (alpha[i] - beta) * S[i]
    return np.asarray(-q, dtype=np.float64)


def check_one_step(history_cls, name, mem, n=2000, steps=20, strict=True):
    """One-step direction parity (strict tier): fed the ORACLE's own accepted points and gradients,
    the compact two-loop (K1+K2+K3) must reproduce the exact (extended-precision) two-loop direction on the
    same history to 1e-12, the north_star tolerance, on every objective."""
    import ctypes as C
    kind = OBJECTIVES[name][0]
    x0 = start(name, n)
    tr = O.Trace(max_vec_iters=steps + 2)
    O.lbfgs(O.builtin_callbacks(kind, 0, n), x0.copy(), Memory=mem, use_ffd=True, Warning=False, MaxIteration=steps + 2,
            trace=tr)
    g0 = np.empty(n)
    O.lib().orc_obj_select(kind, 0, n)
    O.lib().orc_obj_fd(g0.ctypes.data_as(C.c_void_p), x0.ctypes.data_as(C.c_void_p), C.byref(C.c_int(n)))
    xs, gs = [x0] + tr.x, [g0] + tr.g
    h = history_cls(n, mem)
    pairs, worst = [], 0.0
    for k in range(min(steps, len(tr.p) - 1)):
        h.push(xs[k + 1], xs[k], gs[k + 1], gs[k])
        p, xt, gp, pp = h.direction(gs[k + 1], xs[k + 1])
        pairs = (pairs + [(xs[k + 1] - xs[k], gs[k + 1] - gs[k])])[-mem:]
        exact = two_loop_extended(pairs, gs[k + 1])
        noise = rel(tr.p[k + 1], exact)            # the reference's own rounding error on this step
        err = rel(p, exact)
        worst = max(worst, err)
        # north_star: search directions to relative 1e-12.  Measured (tests/gpu_calibrate.py): <= 1.3e-15 on the
        # benchmark objectives, 9.5e-14 on the degenerate standard-start Rosenbrock (n/2 identical 2-D problems: the
        # Gram matrix of the pairs has rank 2), where the oracle's own sequential sums are 3.2e-12 from the exact value
        # (strict=False: the CPU host simulator, whose blocked sequential sums are only as accurate as the oracle's:
        # within 4x the oracle's own distance from the exact direction)
        bound = 1e-12 if strict else max(1e-12, 4.0 * noise)
        assert err <= bound, f"direction after step {k}: {err:.2e} (reference's own: {noise:.2e})"
        assert np.array_equal(xt, xs[k + 1] + p)
        assert abs(gp - float(np.dot(gs[k + 1], p))) <= 1e-10 * abs(gp)
        assert abs(pp - float(np.dot(p, p))) <= 1e-10 * abs(pp)
    h.close()
    return worst


def oracle_iteration_range(name, n, run, **kw):
    """Iteration counts of the oracle under perturbations no implementation can be distinguished from:
    its three summation orders, and the start vector moved by ONE ULP in a single entry (six variants).
    On Rosenbrock from the perturbed start the count moves by 40 % under such a nudge (156 ... 223 at
    n = 10^4), so "iteration counts within 2 %" (north_star) is asserted against this range: the run under
    test must land within [0.98 min, 1.02 max].  Returns (counts, statuses, x of the unperturbed run)."""
    kind, _, _ = OBJECTIVES[name]
    x0 = start(name, n)
    counts, statuses, x_ref = [], [], None
    for mode in (0, 1, 2):
        x, st = run(O.builtin_callbacks(kind, 0, n), x0.copy(), sum_mode=mode, Warning=False, **kw)
        counts.append(st.n_iter); statuses.append(st.status)
        if mode == 0:
            x_ref = x
    for k in range(6):
        xp = x0.copy()
        j = (k * 37) % n
        xp[j] = np.nextafter(xp[j], 2.0)
        x, st = run(O.builtin_callbacks(kind, 0, n), xp, Warning=False, **kw)
        counts.append(st.n_iter); statuses.append(st.status)
    return counts, statuses, x_ref


def check_iteration_count(got, counts, what=""):
    lo, hi = min(counts), max(counts)
    assert 0.98 * lo - 1 <= got <= 1.02 * hi + 1, f"{what}: {got} iterations, oracle range {lo}..{hi}"


def oracle_envelope(name, n, run, iters=20, **kw):
    """Directions of the oracle under its three summation orders.  Returns (seq_trace, env) where
    env[k] = max relative deviation of the long-double / pairwise runs from the sequential run at
    iteration k: the reference algorithm's own sensitivity to summation order."""
    kind, _, _ = OBJECTIVES[name]
    x0 = start(name, n)
    traces = []
    for mode in (0, 1, 2):
        tr = O.Trace(max_vec_iters=iters)
        run(O.builtin_callbacks(kind, 0, n), x0.copy(), trace=tr, sum_mode=mode, Warning=False, **kw)
        traces.append(tr)
    env = []
    for k in range(min(len(t.p) for t in traces)):
        env.append(max(rel(traces[1].p[k], traces[0].p[k]), rel(traces[2].p[k], traces[0].p[k])))
    return traces, env


# ---- 1-D objectives that steer the line searchers into specific branches.  In one dimension every
# dot product is a single multiply, so the oracle, the NumPy transcription, the host simulator and
# the GPU must agree BIT FOR BIT on the whole trajectory (CG; and LBFGS' steepest-descent step).
def _pw(b, k):
    try:
        return b ** k
    except OverflowError:      # Python raises where IEEE arithmetic returns inf
        return float("inf") if (b > 0 or k % 2 == 0) else float("-inf")


def _poly_wall(c, k):
    """f(x) = -x + c (x/c)^k / k : slope jumps from <0 to >> c2|phi'(0)| within one 1.05 step near x=c."""
    def f(x):
        return -x + c * _pw(x / c, k) / k

    def g(x):
        return -1.0 + _pw(x / c, k - 1)
    return f, g


TORTURE_1D = {
    # start, (f, g) built from plain Python floats (IEEE double), description
    "f9_quirk": (0.0, _poly_wall(1.05 / 1.01748, 41)),          # StrongWolfe falls through f90:1511-1512
    "quartic1": (0.7, (lambda x: _pw(x, 4), lambda x: 4.0 * _pw(x, 3))),
    "steep": (3.0, (lambda x: 50.0 * _pw(x - 1.0, 2), lambda x: 100.0 * (x - 1.0))),   # Armijo fails first (branch D)
    "flat": (2.0, (lambda x: 1e-3 * _pw(x - 1.0, 2) + 1.0, lambda x: 2e-3 * (x - 1.0))),  # long grow loop (branch C)
    "cosh": (1.5, (lambda x: float(np.cosh(min(x, 700.0))), lambda x: float(np.sinh(min(x, 700.0))))),
    "wall": (0.2, _poly_wall(2.0, 41)),
    # f is NaN beyond x = 2: the first trial lands there and `fx<=fx0+c1*a*phid0` (f90:1483) must read it as a
    # violation, not as a pass (the `>` form of the test, f90:1502, is false for NaN too)
    "nan_region": (-3.0, (lambda x: x * x if x < 2.0 else float("nan"), lambda x: 2.0 * x if x < 2.0 else float("nan"))),
}


class Fuse:
    """Wraps (f, g): after `limit` evaluations f becomes +inf and g becomes 0, which makes every
    searcher and driver terminate (bisection collapses the bracket, then |f'| = 0 < tol).  Needed
    because the reference's fall-through at f90:1511-1512 can zoom with inconsistent data for a very
    long time; all implementations see the same call sequence, so trajectories stay comparable."""

    def __init__(self, f, g, limit=400):
        self.f0, self.g0, self.limit, self.calls = f, g, limit, 0
        self.xs = []

    def f(self, x):
        self.calls += 1
        self.xs.append(x)
        return self.f0(x) if self.calls <= self.limit else float("inf")

    def g(self, x):
        self.calls += 1
        self.xs.append(x)
        return self.g0(x) if self.calls <= self.limit else 0.0

    def fg(self, x):
        self.calls += 1
        self.xs.append(x)
        return (self.f0(x), self.g0(x)) if self.calls <= self.limit else (float("inf"), 0.0)


def make_ref_callbacks(f, g, fg=None):
    """Reference-ABI (f90:33-38) ctypes callbacks over Python scalar functions; dim must be 1."""
    if fg is None:
        def fg(x):
            return f(x), g(x)
    def cf(fx, x, dim):
        xs = C.cast(x, C.POINTER(C.c_double))
        fx[0] = f(xs[0])

    def cfd(fdx, x, dim):
        xs = C.cast(x, C.POINTER(C.c_double))
        C.cast(fdx, C.POINTER(C.c_double))[0] = g(xs[0])

    def cffd(fx, fdx, x, dim):
        xs = C.cast(x, C.POINTER(C.c_double))
        fv, gv = fg(xs[0])
        fx[0] = fv
        C.cast(fdx, C.POINTER(C.c_double))[0] = gv
        return 0
    return cf, cfd, cffd

# With f_fd present these two produce a NaN step inside the reference's zoom, which then never
# terminates (NaN defeats both the bracket test and the collapse test, f90:1684,1695): the reference
# itself would hang, so there is nothing to compare.
TORTURE_NO_FFD = ("f9_quirk", "wall")
