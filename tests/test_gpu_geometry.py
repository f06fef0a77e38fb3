"""GPU parity at PRODUCTION LAUNCH GEOMETRY (pytest -m gpu): n = 2^22 and 2^22 + 3.

At these sizes every library kernel launches its full grid (148 SMs x resident CTAs), every grid-stride loop wraps at
least twice, the 4x-unrolled branches of the streaming kernels are taken and the odd tail element exists -- none of
which the n <= 65537 cases of test_gpu.py reach.  Three kinds of check, all through the C-ABI:

  (a) K1 + K2 + K3 (flgpu_history_*) on integer-valued pairs: every dot product is then exact in any summation
      order, so the direction must equal a NumPy statement of K2 / K3 BIT FOR BIT (any stride, tail or ring-slot
      error changes a dot or an element); and on the ORACLE's own accepted points at this size: 1e-12 against the
      extended-precision two-loop recursion (north_star: search directions to relative 1e-12);
  (b) whole trajectories -- LBFGS m = 10 on Rosenbrock R1, CG Dai-Yuan and Polak-Ribiere+ on the quartic
      (BASELINE.json configs 1-3 shapes) -- against oracle.c: first 20 directions within the oracle's own
      summation-order envelope, first direction bit-exact, same trial counts while the searches are not yet chaotic;
  (c) the element-wise and objective kernels at this size live in test_gpu.py (same tests, larger n).

The oracle needs ~20 s of one host core per run at this size; the nine runs are made concurrently in worker
processes (tests/_oracle_traj.py) while the checks that need no oracle run.
"""
import os
import shutil

import numpy as np
import pytest

import _cases
import _oracle as O
from _oracle_traj import Job

pytestmark = pytest.mark.gpu

N_FULL = 1 << 22
N_ODD = (1 << 22) + 3
ITERS = 20


@pytest.fixture(scope="module")
def fl():
    import fortran_library_b200 as fl
    fl.require_gpu()
    return fl


@pytest.fixture(scope="module", autouse=True)
def oracle_jobs(tmp_path_factory):
    """Starts the oracle trajectories (3 cases x 3 summation orders) as soon as the module is entered."""
    try:
        import fortran_library_b200 as fl_
        have_gpu = fl_.device_count() > 0
    except Exception:
        have_gpu = False
    if not have_gpu:
        yield {}
        return
    root = str(tmp_path_factory.mktemp("oracle_traj"))
    jobs = {}
    for mode in (0, 1, 2):
        jobs["lbfgs", mode] = Job(root, "lbfgs", "rosenR1", N_ODD, mode, ITERS, keep_xg=(mode == 0), Memory=10,
                                  use_ffd=True, MaxIteration=10)
        for M in ("DY", "PR"):
            jobs[M, mode] = Job(root, "cg", "quartic", N_ODD, mode, ITERS, Method=M, use_ffd=True, MaxIteration=ITERS)
    yield jobs
    for j in jobs.values():
        if j.proc.poll() is None:
            j.proc.kill()
    shutil.rmtree(root, ignore_errors=True)


# ----------------------------------------------------------------------------- GPU == scalar C++ statement at full grid
def test_gpu_equals_host_simulator_at_full_grid(fl):
    """The host simulator (tests/hostsim: the product's driver over a scalar C++ statement of the kernels' arithmetic,
    reductions included) against the GPU at n = 2^22 + 3 -- 4097 chunks, so every kernel's chunk loop wraps its grid
    several times and the tree has two blocks: L-BFGS m = 3 through the 181-trial first line search and four main-loop
    iterations, bit for bit.  (Runs while the oracle workers of this module compute.)"""
    import _hostsim as H
    n = N_ODD
    x0 = _cases.start("rosenR1", n)
    obh = H.Observer(keep_vectors=False)
    xh, sth = H.lbfgs(O.OBJ_ROSENBROCK, x0, observer=obh, use_ffd=True, Warning=False, n_global=n, Memory=3, MaxIteration=4)
    x = fl.DeviceVector.start(fl.START_ROSEN_PERT, n, seed=7)
    ob = fl.Observer()
    st = fl.LBFGS(fl.builtin_problem(fl.OBJ_ROSENBROCK), x, observer=ob, Warning=False, Memory=3, MaxIteration=4)
    assert ob.rows == obh.rows, next((k, a, b) for k, (a, b) in enumerate(zip(ob.rows, obh.rows)) if a != b)
    assert np.array_equal(x.numpy(), xh) and st.n_trials == sth.n_trials and st.iterations == sth.iterations == 7


# ----------------------------------------------------------------------------- (a) K1 + K2 + K3, bit-exact
def _gram_solve(m, k, recent, D, SY, YY):
    """lbfgs_gram_solve() of include/flgpu_lbfgs_gram.hpp (the scalar statement of K2) in Python floats: same operations in
    the same order, every multiply and add rounded separately."""
    r = recent
    slot = lambda t: (recent - t + m) % m          # noqa: E731
    sq, yq, al = [0.0] * m, [0.0] * m, [0.0] * m
    for t in range(k):
        j = slot(t)
        SY[j][r] = D["SYN"][j]
        YY[j][r] = D["YYN"][j]
        YY[r][j] = D["YYN"][j]
        sq[j], yq[j] = D["A"][j], D["B"][j]
    C = [0.0] * (2 * m + 1)
    for t in range(k):
        i = slot(t)
        rho = 1.0 / SY[i][i]
        alpha = rho * sq[i]
        al[i] = alpha
        for u in range(k):
            j = slot(u)
            if u > t:
                sq[j] = sq[j] - alpha * SY[j][i]
            yq[j] = yq[j] - alpha * YY[j][i]
    rho_r = 1.0 / SY[r][r]
    gamma = 1.0 / rho_r / YY[r][r]
    for u in range(k):
        j = slot(u)
        yq[j] = gamma * yq[j]
    for t in range(k - 1, -1, -1):
        i = slot(t)
        rho = 1.0 / SY[i][i]
        beta = rho * yq[i]
        e = al[i] - beta
        C[1 + i] = al[i]
        C[1 + m + i] = e
        for u in range(t):
            j = slot(u)
            yq[j] = yq[j] + e * SY[i][j]
    C[0] = gamma
    return C


def _k3_numpy(g, S, Y, C, m, k, recent):
    """K3 element by element in its own order (kernels.cuh k3_direction_kernel), separate roundings."""
    v = g.copy()
    for t in range(k):                              # y newest -> oldest, coefficient -alpha
        j = (recent - t + m) % m
        v = v + (-C[1 + j]) * Y[j]
    v = C[0] * v
    for t in range(k - 1, -1, -1):                  # s oldest -> newest, coefficient e
        j = (recent - t + m) % m
        v = v + C[1 + m + j] * S[j]
    return -v


@pytest.mark.parametrize("n", [N_FULL, N_ODD])
@pytest.mark.parametrize("mem", [5, 10, 30])
def test_two_loop_operator_bit_exact_on_integer_data(fl, mem, n):
    """Integer-valued x and f' (|values| small): s.g, y.g, s.y, y.y are sums of integers far below 2^53, hence exact
    under ANY summation order, and K2 / K3 round every multiply and add separately -- so the GPU direction, the trial
    point and the two reduced scalars' inputs must reproduce the NumPy statement bit for bit at the full launch
    geometry (m = 30 runs K1 in three passes and K3 in eight chunks)."""
    rng = np.random.default_rng(1000 * mem + (n & 7))
    d = rng.integers(1, 5, n).astype(np.float64)            # SPD diagonal Hessian with integer entries
    x0 = rng.integers(-8, 9, n).astype(np.float64)
    g0 = d * x0
    h = fl.History(n, mem)
    S, Y = [None] * mem, [None] * mem
    SY = [[0.0] * mem for _ in range(mem)]
    YY = [[0.0] * mem for _ in range(mem)]
    recent, k = -1, 0
    steps = mem + 3
    check_at = set(range(steps)) if mem <= 10 else {0, 1, 2, mem - 2, mem - 1, mem, mem + 2}
    for it in range(steps):
        x1 = x0 + rng.integers(-2, 3, n).astype(np.float64)
        x1[it] = x0[it] + 1.0                                # never an all-zero step
        g1 = d * x1
        h.push(x1, x0, g1, g0)
        if k < mem:
            recent, k = recent + 1, k + 1
        else:
            recent = (recent + 1) % mem
        S[recent], Y[recent] = x1 - x0, g1 - g0
        D = {"A": [0.0] * mem, "B": [0.0] * mem, "SYN": [0.0] * mem, "YYN": [0.0] * mem}
        for t in range(k):
            j = (recent - t + mem) % mem
            D["A"][j] = float(np.dot(S[j], g1)); D["B"][j] = float(np.dot(Y[j], g1))
            D["SYN"][j] = float(np.dot(S[j], Y[recent])); D["YYN"][j] = float(np.dot(Y[j], Y[recent]))
        C = _gram_solve(mem, k, recent, D, SY, YY)           # persistent Gram blocks updated every step, as on the device
        if it in check_at:
            p, xt, gp, pp = h.direction(g1, x1)
            want = _k3_numpy(g1, S, Y, C, mem, k, recent)
            bad = np.flatnonzero(p != want)
            assert bad.size == 0, (f"m={mem} n={n} step {it}: {bad.size} elements differ, first at {bad[:5]} "
                                   f"(gpu {p[bad[:3]]}, numpy {want[bad[:3]]})")
            assert np.array_equal(xt, x1 + p)
            assert abs(gp - float(np.dot(g1, p))) <= 1e-12 * float(np.dot(np.abs(g1), np.abs(p)))
            assert abs(pp - float(np.dot(p, p))) <= 1e-12 * pp
        x0, g0 = x1, g1
    h.close()


def test_one_step_direction_on_oracle_history_at_full_grid(fl, oracle_jobs):
    """K1+K2+K3 fed the oracle's own accepted points and gradients at n = 2^22 + 3 (LBFGS m = 10, Rosenbrock R1):
    the next direction to 1e-12 of the extended-precision two-loop recursion (north_star's tolerance)."""
    n, mem = N_ODD, 10
    tr = oracle_jobs["lbfgs", 0].result()
    x0 = _cases.start("rosenR1", n)
    g0 = np.empty(n)
    import ctypes as C
    O.lib().orc_obj_select(O.OBJ_ROSENBROCK, 0, n)
    O.lib().orc_obj_fd(g0.ctypes.data_as(C.c_void_p), x0.ctypes.data_as(C.c_void_p), C.byref(C.c_int(n)))
    xs = [x0] + [np.asarray(v) for v in tr["x"]]
    gs = [g0] + [np.asarray(v) for v in tr["g"]]
    h = fl.History(n, mem)
    pairs, worst = [], 0.0
    for k in range(len(tr["p"]) - 1):
        h.push(xs[k + 1], xs[k], gs[k + 1], gs[k])
        p, xt, gp, pp = h.direction(gs[k + 1], xs[k + 1])
        pairs = (pairs + [(xs[k + 1] - xs[k], gs[k + 1] - gs[k])])[-mem:]
        exact = _cases.two_loop_extended(pairs, gs[k + 1])
        noise = _cases.rel(tr["p"][k + 1], exact)
        err = _cases.rel(p, exact)
        worst = max(worst, err)
        assert err <= 1e-12, f"direction after step {k}: {err:.2e} (oracle's own: {noise:.2e})"
        assert np.array_equal(xt, xs[k + 1] + p)
    h.close()
    print(f"worst one-step error at n=2^22+3: {worst:.2e}")


# ----------------------------------------------------------------------------- (b) trajectories vs oracle.c
def _check_trajectory(jobs, key, ob, what):
    tr = [jobs[key, mode].result() for mode in (0, 1, 2)]
    ld = tr[1]["p"]
    assert np.array_equal(ob.p[0], tr[0]["p"][0]), f"{what}: steepest-descent direction must be bit-exact"
    ref = 0.0
    nk = min(len(ob.p), ITERS, *(len(t["p"]) for t in tr))
    assert nk == ITERS
    for k in range(nk):
        ref = max(ref, _cases.rel(tr[0]["p"][k], ld[k]), _cases.rel(tr[2]["p"][k], ld[k]))
        err = _cases.rel(ob.p[k], ld[k])
        assert err <= max(_cases.FLOOR, _cases.ENV_FACTOR * ref), \
            f"{what}: direction {k} off by {err:.2e} (oracle's own summation-order noise {ref:.2e})"
    # the searches: same trial counts as the oracle, steps and f within the oracle's own summation-order spread, for
    # as long as the three oracle runs agree among themselves on the trial counts (once they part ways there is no
    # single reference sequence)
    spread_a = spread_f = 0.0
    agreed = 0
    for k in range(nk):
        rows = [t["rows"][k] for t in tr]
        if not all(r[4] == rows[0][4] for r in rows):
            break
        agreed = k + 1
        a_ld, f_ld = rows[1][1], rows[1][2]
        spread_a = max(spread_a, abs(rows[0][1] - a_ld) / abs(a_ld), abs(rows[2][1] - a_ld) / abs(a_ld))
        spread_f = max(spread_f, abs(rows[0][2] - f_ld) / abs(f_ld), abs(rows[2][2] - f_ld) / abs(f_ld))
        assert ob.rows[k][4] == rows[0][4], f"{what}: iteration {k}: {ob.rows[k][4]} trials, oracle {rows[0][4]}"
        da, df = abs(ob.rows[k][1] - a_ld) / abs(a_ld), abs(ob.rows[k][2] - f_ld) / abs(f_ld)
        assert da <= max(_cases.FLOOR, _cases.ENV_FACTOR * spread_a), \
            f"{what}: step {k} off by {da:.2e} (oracle's own spread {spread_a:.2e})"
        assert df <= max(_cases.FLOOR, _cases.ENV_FACTOR * spread_f), \
            f"{what}: f at {k} off by {df:.2e} (oracle's own spread {spread_f:.2e})"
    assert agreed >= 4, f"{what}: the oracle's own runs disagree on trial counts from iteration {agreed}"


@pytest.mark.parametrize("fused", [True, False], ids=["fused", "plain"])
def test_lbfgs_trajectory_at_full_grid(fl, oracle_jobs, fused):
    """BASELINE.json configs[1] shape (LBFGS m = 10, extended Rosenbrock, start R1) at n = 2^22 + 3."""
    x = fl.DeviceVector.start(fl.START_ROSEN_PERT, N_ODD, seed=7)
    ob = fl.Observer(keep_vectors=True, max_vec_iters=ITERS)
    st = fl.LBFGS(fl.builtin_problem(fl.OBJ_ROSENBROCK), x, Memory=10, MaxIteration=10, Warning=False, observer=ob,
                  fused=fused)
    assert st.iterations == ITERS
    _check_trajectory(oracle_jobs, "lbfgs", ob, f"lbfgs rosenR1 n=2^22+3 fused={fused}")


@pytest.mark.parametrize("method", ["DY", "PR"])
def test_cg_trajectory_at_full_grid(fl, oracle_jobs, method):
    """BASELINE.json configs[2] shape (CG Dai-Yuan / Polak-Ribiere+, separable quartic) at n = 2^22 + 3."""
    x = fl.DeviceVector.start(fl.START_QUARTIC_U, N_ODD, seed=12345)
    ob = fl.Observer(keep_vectors=True, max_vec_iters=ITERS)
    st = fl.ConjugateGradient(fl.builtin_problem(fl.OBJ_QUARTIC), x, Method=method, MaxIteration=ITERS, Warning=False,
                              observer=ob)
    assert st.iterations == ITERS
    _check_trajectory(oracle_jobs, method, ob, f"cg {method} quartic n=2^22+3")
