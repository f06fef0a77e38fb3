"""oracle/oracle_np.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  PARITY UNPINNED.

Second, independent transcription of the reference hot path, written from
/root/reference/source/NonlinearOptimization.f90 (not from oracle.c) in NumPy, with
Fortran host association mirrored by closures.  Its only job is to catch
transcription errors in oracle/oracle.c: tests require both to produce IDENTICAL
bits (sequential sums: np.cumsum is a strict left-to-right recurrence; products are
rounded before the add, i.e. no FMA), since the reference pins no numbers itself
(SURVEY.md F8).

"f90:" = NonlinearOptimization.f90.  Callbacks: f(x)->float, fd(x)->ndarray,
f_fd(x)->(float, ndarray) or None.  Optional arguments: None = absent.
"""
import math

import numpy as np


def div(a, b):
    """IEEE-754 division (Python floats raise on /0; Fortran and C do not)."""
    with np.errstate(all="ignore"):
        return float(np.float64(a) / np.float64(b))


def dot(a, b):
    """dot_product(a,b) as gfortran -O3 without fast-math evaluates it."""
    if a.size == 0:
        return 0.0
    return float(np.cumsum(a * b)[-1])


class Counters:
    def __init__(self):
        self.trials = 0
        self.n_f = self.n_fd = self.n_ffd = 0
        self.iters = 0
        self.status = -1
        self.history = []  # (p, x, g, a, fx, phid0, trials) per outer iteration


# ----------------------------------------------------------------- line searchers
def wolfe(c1, c2, f, fd, x, a, p, fx, phid0, Increment, cnt):
    """f90:1286-1371 (Wolfe) == f90:1373-1459 (Wolfe_fdwithf).  Returns x, a, fx, fdx."""
    incrmt = max(1.0 + 1e-15, Increment) if Increment is not None else 1.05
    x0 = x.copy()
    fx0 = fx
    c2_m_abs_phid0 = c2 * abs(phid0)
    st = {"a": a, "x": x, "fx": fx, "fdx": None}

    def setx(aa):
        st["x"] = x0 + aa * p
        cnt.trials += 1

    def cf():
        st["fx"] = f(st["x"]); cnt.n_f += 1

    def cfd():
        st["fdx"] = fd(st["x"]); cnt.n_fd += 1

    def zoom(low, up, flow, fup, phidlow):  # f90:1347-1370
        phidlow_m_a = phidlow * st["a"]
        while True:
            st["a"] = div(phidlow_m_a * st["a"] / 2.0, flow + phidlow_m_a - fup)
            if not (st["a"] > low and st["a"] < up):
                st["a"] = (low + up) / 2.0
            setx(st["a"]); cf()
            if st["fx"] > fx0 + c1 * st["a"] * phid0:
                up = st["a"]
                if up - low < 1e-15 or (up - low) / max(abs(low), abs(up)) < 1e-15:
                    cfd(); return
                fup = st["fx"]
            else:
                cfd(); phidnew = dot(st["fdx"], p)
                if phidnew > c2_m_abs_phid0:
                    return
                low = st["a"]
                if up - low < 1e-15 or (up - low) / max(abs(low), abs(up)) < 1e-15:
                    return
                flow = st["fx"]; phidlow = phidnew; phidlow_m_a = phidlow * st["a"]

    setx(st["a"]); cf()
    if st["fx"] <= fx0 + c1 * st["a"] * phid0:
        while True:
            aold = st["a"]; fold = st["fx"]
            st["a"] = aold * incrmt; setx(st["a"]); cf()
            if st["fx"] > fx0 + c1 * st["a"] * phid0:
                setx(aold)
                cfd()
                phidx = dot(st["fdx"], p)
                if phidx > c2_m_abs_phid0:
                    st["a"] = aold; st["fx"] = fold
                else:
                    zoom(aold, st["a"], fold, st["fx"], phidx)
                break
    else:
        while True:
            aold = st["a"]; fold = st["fx"]
            st["a"] = aold / incrmt; setx(st["a"]); cf()
            if st["fx"] <= fx0 + c1 * st["a"] * phid0:
                cfd()
                phidx = dot(st["fdx"], p)
                if phidx < c2_m_abs_phid0:
                    zoom(st["a"], aold, st["fx"], fold, phidx)
                break
            if st["a"] < 1e-15:
                cfd(); break
    return st["x"], st["a"], st["fx"], st["fdx"]


def strongwolfe(c1, c2, f, fd, f_fd, x, a, p, fx, phid0, Increment, cnt, fdwithf):
    """f90:1462-1580 (fdwithf=False) / f90:1582-1698 (fdwithf=True)."""
    incrmt = max(1.0 + 1e-15, Increment) if Increment is not None else 1.05
    x0 = x.copy()
    fx0 = fx
    c2_m_abs_phid0 = c2 * abs(phid0)
    # Fortran locals that zoom's by-reference dummies may alias live in this dict
    v = {"a": a, "x": x, "fx": fx, "fdx": None, "aold": 0.0, "fold": 0.0, "atemp": 0.0,
         "ftemp": 0.0, "phidnew": 0.0, "phidold": 0.0}

    def setx(aa):
        v["x"] = x0 + aa * p
        cnt.trials += 1

    def cf():
        v["fx"] = f(v["x"]); cnt.n_f += 1

    def cfd():
        v["fdx"] = fd(v["x"]); cnt.n_fd += 1

    def both():
        if fdwithf:
            v["fx"], v["fdx"] = f_fd(v["x"]); cnt.n_ffd += 1
        else:
            cf(); cfd()

    def zoom(low, up, flow, fup, phidlow, phidup):
        """f90:1557-1579.  Arguments are NAMES of entries of v (by-reference aliasing)."""
        while True:
            d1 = v[phidlow] + v[phidup] - div(3.0 * (v[flow] - v[fup]), v[low] - v[up])
            d2 = v[up] - v[low]
            disc = d1 * d1 - v[phidlow] * v[phidup]
            root = math.sqrt(disc) if disc >= 0.0 else float("nan")
            d2 = root if d2 > 0.0 else -root
            den = v[phidup] - v[phidlow] + 2.0 * d2
            num = (v[up] - v[low]) * (v[phidup] + d2 - d1)
            v["a"] = v[up] - div(num, den)
            if not (v["a"] > min(v[low], v[up]) and v["a"] < max(v[low], v[up])):
                v["a"] = (v[low] + v[up]) / 2.0
            setx(v["a"]); both(); phidnew_local = dot(v["fdx"], p)
            if v["fx"] > fx0 + c1 * v["a"] * phid0 or v["fx"] >= v[flow]:
                v[up] = v["a"]; v[fup] = v["fx"]; v[phidup] = phidnew_local
            else:
                if abs(phidnew_local) <= c2_m_abs_phid0:
                    return
                if phidnew_local * (v[up] - v[low]) >= 0.0:
                    v[up] = v[low]; v[fup] = v[flow]; v[phidup] = v[phidlow]
                v[low] = v["a"]; v[flow] = v["fx"]; v[phidlow] = phidnew_local
            gap = abs(v[up] - v[low])
            if gap < 1e-15 or gap / max(abs(v[low]), abs(v[up])) < 1e-15:
                return

    def finish():
        return v["x"], v["a"], v["fx"], v["fdx"]

    setx(v["a"])
    if fdwithf:
        v["fx"], v["fdx"] = f_fd(v["x"]); cnt.n_ffd += 1
    else:
        cf()
    if v["fx"] <= fx0 + c1 * v["a"] * phid0:
        if not fdwithf:
            cfd()
        v["phidnew"] = dot(v["fdx"], p)
        if v["phidnew"] > 0.0:
            if abs(v["phidnew"]) <= c2_m_abs_phid0:
                return finish()
            while True:
                v["aold"] = v["a"]; v["fold"] = v["fx"]; v["phidold"] = v["phidnew"]
                v["a"] = v["aold"] / incrmt; setx(v["a"]); both(); v["phidnew"] = dot(v["fdx"], p)
                if v["fx"] >= v["fold"] or v["phidnew"] <= 0.0:
                    v["atemp"] = v["a"]; v["ftemp"] = v["fx"]
                    zoom("aold", "atemp", "fold", "ftemp", "phidold", "phidnew")
                    return finish()
                if v["a"] < 1e-15:
                    return finish()
        else:
            while True:
                v["aold"] = v["a"]; v["fold"] = v["fx"]; v["phidold"] = v["phidnew"]
                v["a"] = v["aold"] * incrmt; setx(v["a"]); both(); v["phidnew"] = dot(v["fdx"], p)
                if v["fx"] > fx0 + c1 * v["a"] * phid0 or v["fx"] >= v["fold"]:
                    v["atemp"] = v["a"]; v["ftemp"] = v["fx"]
                    zoom("aold", "atemp", "fold", "ftemp", "phidold", "phidnew")
                    return finish()
                if v["phidnew"] > 0.0:
                    if abs(v["phidnew"]) <= c2_m_abs_phid0:
                        return finish()
                    v["atemp"] = v["a"]; v["ftemp"] = v["fx"]
                    zoom("atemp", "aold", "ftemp", "fold", "phidnew", "phidold")
                    if fdwithf:
                        return finish()          # f90:1632
                    v["fx"] = fx0                # f90:1512, falls back into the loop
    else:
        while True:
            v["aold"] = v["a"]; v["fold"] = v["fx"]
            v["a"] = v["aold"] / incrmt; setx(v["a"]); cf()
            if v["fx"] <= fx0 + c1 * v["a"] * phid0:
                cfd(); v["phidnew"] = dot(v["fdx"], p)
                if abs(v["phidnew"]) <= c2_m_abs_phid0:
                    return finish()
                if v["phidnew"] < 0.0:
                    setx(v["aold"]); cfd(); v["phidold"] = dot(v["fdx"], p)
                    v["atemp"] = v["a"]; v["ftemp"] = v["fx"]
                    zoom("atemp", "aold", "ftemp", "fold", "phidnew", "phidold")
                    return finish()
                else:
                    while True:
                        v["aold"] = v["a"]; v["fold"] = v["fx"]; v["phidold"] = v["phidnew"]
                        v["a"] = v["aold"] / incrmt; setx(v["a"]); both(); v["phidnew"] = dot(v["fdx"], p)
                        if v["fx"] >= v["fold"] or v["phidnew"] <= 0.0:
                            v["atemp"] = v["a"]; v["ftemp"] = v["fx"]
                            zoom("aold", "atemp", "fold", "ftemp", "phidold", "phidnew")
                            return finish()
                        if v["a"] < 1e-15:
                            return finish()
            if v["a"] < 1e-15:
                cfd(); return finish()


# ----------------------------------------------------------------- FLGPU_LS_FAST (NOT a reference routine)
LINE_SEARCH_POLICY = 0   # 1 = the product's optional accept-at-first-Wolfe-point searcher (flgpu_options.line_search)


def _cubic_minimiser(u, v, fu, fv, gu, gv):
    """Minimiser of the cubic through (u, fu, gu), (v, fv, gv); NaN when there is none."""
    d1 = gu + gv - div(3.0 * (fu - fv), u - v)
    disc = d1 * d1 - gu * gv
    root = math.sqrt(disc) if disc >= 0.0 else float("nan")
    d2 = root if v - u > 0.0 else -root
    return v - div((v - u) * (gv + d2 - d1), gv - gu + 2.0 * d2)


def fast_search(strong, c1, c2, f, fd, f_fd, x, a, p, fx, phid0, cnt):
    """Written from the algorithm's description in include/flgpu.h / flgpu_search_core.hpp (SearchCore::fast): Nocedal &
    Wright Alg. 3.5/3.6; f and f' at every trial; accept the first trial with sufficient decrease and (strong or weak)
    curvature; grow to the cubic minimiser clipped to [a+1.1(a-a_prev), a+4(a-a_prev)] (40 trials at most); zoom with
    the cubic minimiser kept 5 % off both ends (60 trials at most), NaN counting as insufficient decrease; on a
    collapsed bracket return the best sufficient-decrease point.  The reference has no such routine."""
    x0 = x.copy(); fx0 = fx
    c2abs = c2 * abs(phid0)
    st = {"x": x, "fx": fx, "g": None}

    def evaluate(aa):
        st["x"] = x0 + aa * p; cnt.trials += 1
        if f_fd is not None:
            st["fx"], st["g"] = f_fd(st["x"]); cnt.n_ffd += 1
        else:
            st["fx"] = f(st["x"]); cnt.n_f += 1
            st["g"] = fd(st["x"]); cnt.n_fd += 1
        return st["fx"], dot(st["g"], p)

    def decrease(aa, fa):
        return fa <= fx0 + c1 * aa * phid0           # False for NaN

    def curvature(ga):
        return abs(ga) <= c2abs if strong else ga >= -c2abs

    def zoom(lo, hi, flo, fhi, glo, ghi):
        it = 0
        while True:
            w = hi - lo
            t = div(_cubic_minimiser(lo, hi, flo, fhi, glo, ghi) - lo, w)
            if not (0.0 < t < 1.0):
                t = 0.5
            else:
                t = min(max(t, 0.05), 0.95)
            aa = lo + t * w
            fa, ga = evaluate(aa)
            if (not decrease(aa, fa)) or fa >= flo:
                hi, fhi, ghi = aa, fa, ga
            else:
                if curvature(ga):
                    return aa
                if ga * (hi - lo) >= 0.0:
                    hi, fhi, ghi = lo, flo, glo
                lo, flo, glo = aa, fa, ga
            gap = abs(hi - lo)
            if gap < 1e-15 or div(gap, max(abs(lo), abs(hi))) < 1e-15 or it >= 59:
                if aa != lo:
                    evaluate(lo)
                return lo
            it += 1

    prev = (0.0, fx0, phid0)
    fa, ga = evaluate(a)
    for grow in range(40):
        if (not decrease(a, fa)) or (grow > 0 and fa >= prev[1]):
            a = zoom(prev[0], a, prev[1], fa, prev[2], ga); break
        if curvature(ga):
            break
        if ga >= 0.0:
            a = zoom(a, prev[0], fa, prev[1], ga, prev[2]); break
        if grow == 39:
            break
        nxt = _cubic_minimiser(prev[0], a, prev[1], fa, prev[2], ga)
        lo_b = a + 1.1 * (a - prev[0]); hi_b = a + 4.0 * (a - prev[0])
        if not (nxt <= hi_b):
            nxt = hi_b
        if nxt < lo_b:
            nxt = lo_b
        prev = (a, fa, ga)
        a = nxt
        fa, ga = evaluate(a)
    return st["x"], a, st["fx"], st["g"]


def _search(strong, fdwithf, c1, c2, f, fd, f_fd, x, a, p, fx, phid0, Increment, cnt):
    if LINE_SEARCH_POLICY == 1:
        return fast_search(strong, c1, c2, f, fd, f_fd, x, a, p, fx, phid0, cnt)
    if strong:
        return strongwolfe(c1, c2, f, fd, f_fd, x, a, p, fx, phid0, Increment, cnt, fdwithf)
    return wolfe(c1, c2, f, fd, x, a, p, fx, phid0, Increment, cnt)


def _defaults(Strong, Warning, MaxIteration, Precision, MinStepLength, WolfeConst1, WolfeConst2, c2def):
    sw = True if Strong is None else bool(Strong)
    warn = True if Warning is None else bool(Warning)
    maxit = 1000 if MaxIteration is None else MaxIteration
    tol = 1e-30 if Precision is None else Precision * Precision
    minstep = 1e-30 if MinStepLength is None else MinStepLength * MinStepLength
    c1 = 1e-4 if WolfeConst1 is None else max(1e-15, WolfeConst1)
    c2 = c2def if WolfeConst2 is None else min(1.0 - 1e-15, max(c1 + 1e-15, WolfeConst2))
    return sw, warn, maxit, tol, minstep, c1, c2


# ----------------------------------------------------------------- LBFGS f90:398-625
def lbfgs(f, fd, x, Memory=None, f_fd=None, Strong=None, Warning=None, MaxIteration=None,
          Precision=None, MinStepLength=None, WolfeConst1=None, WolfeConst2=None, Increment=None):
    cnt = Counters()
    x = np.array(x, dtype=np.float64)
    dim = x.size
    mem = 10 if Memory is None else max(1, Memory)
    sw, warn, maxit, tol, minstep, c1, c2 = _defaults(Strong, Warning, MaxIteration, Precision,
                                                      MinStepLength, WolfeConst1, WolfeConst2, 0.9)
    rho = np.zeros(mem + 1); alpha = np.zeros(mem + 1)
    s = np.zeros((mem + 1, dim)); y = np.zeros((mem + 1, dim))
    if f_fd is not None:
        fnew, fdnew = f_fd(x); cnt.n_ffd += 1
    else:
        fnew = f(x); cnt.n_f += 1; fdnew = fd(x); cnt.n_fd += 1
    p = -fdnew; phidnew = -dot(fdnew, fdnew)
    if -phidnew < tol:
        cnt.status = 3; return x, cnt
    a = 1.0 if fnew == 0.0 else div(abs(fnew), math.sqrt(-phidnew))
    xold = x.copy(); fdold = fdnew.copy()

    def search(fdwithf):
        nonlocal x, a, fnew, fdnew
        t0 = cnt.trials; ph0 = phidnew
        x, a, fnew, fdnew = _search(sw, fdwithf, c1, c2, f, fd, f_fd, x, a, p, fnew, phidnew,
                                    Increment, cnt)
        cnt.history.append((p.copy(), x.copy(), fdnew.copy(), a, fnew, ph0, cnt.trials - t0))
        cnt.iters += 1

    def converged():
        nonlocal phidnew
        phidnew = dot(fdnew, fdnew)
        if phidnew < tol:
            cnt.status = 0; return True
        if dot(p, p) * a * a < minstep:
            cnt.status = 1; return True
        return False

    search(False)                                   # f90:448-460
    if converged():
        return x, cnt
    recent = 0
    s[0] = x - xold; y[0] = fdnew - fdold; rho[0] = div(1.0, dot(y[0], s[0]))
    for _ in range(1, mem):                         # f90:472-510
        xold = x.copy(); fdold = fdnew.copy()
        p = fdnew.copy()
        for i in range(recent, -1, -1):
            alpha[i] = rho[i] * dot(s[i], p)
            p = p - alpha[i] * y[i]
        p = p / rho[recent] / dot(y[recent], y[recent])
        for i in range(0, recent + 1):
            phidnew = rho[i] * dot(y[i], p)
            p = p + (alpha[i] - phidnew) * s[i]
        p = -p; phidnew = dot(fdnew, p); a = 1.0
        search(False)
        if converged():
            return x, cnt
        recent = recent + 1
        s[recent] = x - xold; y[recent] = fdnew - fdold; rho[recent] = div(1.0, dot(y[recent], s[recent]))
    for _ in range(maxit):                          # f90:511-579
        xold = x.copy(); fdold = fdnew.copy()       # Before() f90:586-608
        p = fdnew.copy()
        for i in list(range(recent, -1, -1)) + list(range(mem - 1, recent, -1)):
            alpha[i] = rho[i] * dot(s[i], p)
            p = p - alpha[i] * y[i]
        p = p / rho[recent] / dot(y[recent], y[recent])
        for i in list(range(recent + 1, mem)) + list(range(0, recent + 1)):
            phidnew = rho[i] * dot(y[i], p)
            p = p + (alpha[i] - phidnew) * s[i]
        p = -p; phidnew = dot(fdnew, p); a = 1.0
        search(f_fd is not None)
        if converged():                             # After() f90:609-624
            return x, cnt
        recent = (recent + 1) % mem
        s[recent] = x - xold; y[recent] = fdnew - fdold; rho[recent] = div(1.0, dot(y[recent], s[recent]))
    cnt.status = 2
    return x, cnt


# ----------------------------------------------------------------- CG f90:193-394
def conjugate_gradient(f, fd, x, Method=None, f_fd=None, Strong=None, Warning=None, MaxIteration=None,
                       Precision=None, MinStepLength=None, WolfeConst1=None, WolfeConst2=None,
                       Increment=None):
    cnt = Counters()
    x = np.array(x, dtype=np.float64)
    typ = "DY" if Method is None else (Method + "  ")[:2]
    sw, warn, maxit, tol, minstep, c1, c2 = _defaults(Strong, Warning, MaxIteration, Precision,
                                                      MinStepLength, WolfeConst1, WolfeConst2, 0.45)
    if f_fd is not None:
        fnew, fdnew = f_fd(x); cnt.n_ffd += 1
    else:
        fnew = f(x); cnt.n_f += 1; fdnew = fd(x); cnt.n_fd += 1
    p = -fdnew; phidnew = -dot(fdnew, fdnew)
    if -phidnew < tol:
        cnt.status = 3; return x, cnt
    a = 1.0 if fnew == 0.0 else div(abs(fnew), math.sqrt(-phidnew))
    if typ not in ("DY", "PR"):
        raise SystemExit("Program abort: unsupported conjugate gradient method " + typ)
    strong = True if typ == "PR" else sw
    for _ in range(maxit):
        fdold = fdnew.copy(); phidold = phidnew
        t0 = cnt.trials
        x, a, fnew, fdnew = _search(strong, f_fd is not None, c1, c2, f, fd, f_fd, x, a, p, fnew,
                                    phidnew, Increment, cnt)
        cnt.history.append((p.copy(), x.copy(), fdnew.copy(), a, fnew, phidold, cnt.trials - t0))
        cnt.iters += 1
        phidnew = dot(fdnew, fdnew)                 # DY()/PR() f90:352-393
        if phidnew < tol:
            cnt.status = 0; return x, cnt
        if dot(p, p) * a * a < minstep:
            cnt.status = 1; return x, cnt
        if typ == "DY":
            p = -fdnew + div(dot(fdnew, fdnew), dot(fdnew - fdold, p)) * p
        else:
            p = -fdnew + div(dot(fdnew, fdnew - fdold), dot(fdold, fdold)) * p
        phidnew = dot(fdnew, p)
        if phidnew > 0.0:
            p = -fdnew; phidnew = -dot(fdnew, fdnew)
        a = div(a * phidold, phidnew)
    cnt.status = 2
    return x, cnt


# ----------------------------------------------------------------- SteepestDescent f90:55-188
def steepest_descent(f, fd, x, f_fd=None, Strong=None, Warning=None, MaxIteration=None, Precision=None,
                     MinStepLength=None, WolfeConst1=None, WolfeConst2=None, Increment=None):
    cnt = Counters()
    x = np.array(x, dtype=np.float64)
    sw, warn, maxit, tol, minstep, c1, c2 = _defaults(Strong, Warning, MaxIteration, Precision,
                                                      MinStepLength, WolfeConst1, WolfeConst2, 0.9)
    if f_fd is not None:                            # f90:86-90
        fnew, fdnew = f_fd(x); cnt.n_ffd += 1
    else:
        fnew = f(x); cnt.n_f += 1; fdnew = fd(x); cnt.n_fd += 1
    p = -fdnew; phidnew = -dot(fdnew, fdnew)        # f90:92
    if -phidnew < tol:
        cnt.status = 3; return x, cnt
    a = 1.0 if fnew == 0.0 else div(abs(fnew), math.sqrt(-phidnew))
    for _ in range(maxit):                          # f90:101-167: one of 8 copies, same body
        phidold = phidnew
        t0 = cnt.trials
        x, a, fnew, fdnew = _search(sw, f_fd is not None, c1, c2, f, fd, f_fd, x, a, p, fnew, phidnew,
                                    Increment, cnt)
        cnt.history.append((p.copy(), x.copy(), fdnew.copy(), a, fnew, phidold, cnt.trials - t0))
        cnt.iters += 1
        phidnew = dot(fdnew, fdnew)                 # After() f90:172-187
        if phidnew < tol:
            cnt.status = 0; return x, cnt
        if dot(p, p) * a * a < minstep:
            cnt.status = 1; return x, cnt
        p = -fdnew; phidnew = -dot(fdnew, fdnew)
        a = div(a * phidold, phidnew)
    cnt.status = 2
    return x, cnt


# ----------------------------------------------------------------- AugmentedLagrangian f90:2005-2241
def augmented_lagrangian(f, fd, c, cd, x, M, UnconstrainedSolver=None, lambda0=None, miu0=None, Memory=None,
                         Method=None, f_fd=None, Strong=None, Warning=None, MaxIteration=None, Precision=None,
                         MinStepLength=None, WolfeConst1=None, WolfeConst2=None, Increment=None):
    """Branches 'LBFGS' (f90:2150-2167) and 'ConjugateGradient' (f90:2168-2185) only.  c(x) -> ndarray (M,),
    cd(x) -> ndarray (N, M).  Returns x, dict(outer, inner, trials, cnorm2, miu, status)."""
    x = np.array(x, dtype=np.float64)
    solver = "BFGS" if UnconstrainedSolver is None else UnconstrainedSolver
    lam = np.zeros(M) if lambda0 is None else np.array(lambda0, dtype=np.float64)
    miu = 1.0 if miu0 is None else max(1.0, miu0)
    sw = True if Strong is None else bool(Strong)
    warn = True if Warning is None else bool(Warning)
    maxit = 1000 if MaxIteration is None else MaxIteration
    tol = 1e-15 if Precision is None else Precision
    minstep = 1e-15 if MinStepLength is None else MinStepLength
    c1 = 1e-4 if WolfeConst1 is None else max(1e-15, WolfeConst1)
    if WolfeConst2 is not None:
        c2 = min(1.0 - 1e-15, max(c1 + 1e-15, WolfeConst2))
    else:
        c2 = 0.45 if solver == "ConjugateGradient" else 0.9
    incrmt = 1.05 if Increment is None else Increment
    mem = 10 if Memory is None else max(1, Memory)
    typ = "DY" if Method is None else Method
    tolsq = tol * tol
    if solver not in ("LBFGS", "ConjugateGradient"):
        raise SystemExit("oracle_np: AugmentedLagrangian is transcribed for LBFGS / ConjugateGradient only")
    state = {"lam": lam, "miu": miu}

    def terms(Lx, cx):                              # Lx - dot(lambda,cx) + miu/2*dot(cx,cx)
        return Lx - dot(state["lam"], cx) + state["miu"] / 2.0 * dot(cx, cx)

    def grad(Ldx, cx, cdx):                         # Ldx + matmul(cdx, miu*cx-lambda), ascending constraint index
        w = state["miu"] * cx - state["lam"]
        r = np.zeros_like(Ldx)
        for j in range(M):
            r = r + cdx[:, j] * w[j]
        return Ldx + r

    def L(xx):                                      # f90:2193-2199
        return terms(f(xx), c(xx))

    def Ld(xx):                                     # f90:2200-2206
        g = fd(xx); cx = c(xx); cdx = cd(xx)
        return grad(g, cx, cdx)

    def L_Ld(xx):                                   # f90:2207-2217
        Lx = f(xx); cx = c(xx); Lx = terms(Lx, cx)
        g = fd(xx); cdx = cd(xx)
        return Lx, grad(g, cx, cdx)

    def L_Ld_fdwithf(xx):                           # f90:2218-2228
        Lx, g = f_fd(xx); cx = c(xx); Lx = terms(Lx, cx)
        return Lx, grad(g, cx, cd(xx))

    out = {"outer": 0, "inner": 0, "trials": 0, "cnorm2": 0.0, "status": 2}
    inner_ffd = L_Ld_fdwithf if f_fd is not None else L_Ld
    for it in range(1, maxit + 1):
        if solver == "LBFGS":
            x, cnt = lbfgs(L, Ld, x, Memory=mem, f_fd=inner_ffd, Strong=sw, Warning=warn, MaxIteration=maxit,
                           Precision=tol, MinStepLength=minstep, WolfeConst1=c1, WolfeConst2=c2, Increment=incrmt)
        else:
            x, cnt = conjugate_gradient(L, Ld, x, Method=typ, f_fd=inner_ffd, Strong=sw, Warning=warn,
                                        MaxIteration=maxit, Precision=tol, MinStepLength=minstep, WolfeConst1=c1,
                                        WolfeConst2=c2, Increment=incrmt)
        out["outer"] = it; out["inner"] += cnt.iters; out["trials"] += cnt.trials
        cx = c(x)
        out["cnorm2"] = dot(cx, cx)
        if out["cnorm2"] < tolsq:
            out["status"] = 0
            break
        state["lam"] = state["lam"] - state["miu"] * cx
        state["miu"] = state["miu"] * incrmt
    out["miu"] = state["miu"]
    return x, out


def sphere_constraint():
    """test.f90:692-705: cx(1) = dot_product(x,x) - 1, cdx(:,1) = 2 x."""
    def c(x):
        return np.array([dot(x, x) - 1.0])

    def cd(x):
        return (2.0 * x).reshape(-1, 1)
    return c, cd


# ----------------------------------------------------------------- objectives (same op order as objectives.c)
def quartic():
    def f(x):
        x2 = x * x
        return float(np.cumsum(x2 * x2)[-1])

    def fd(x):
        return 4.0 * ((x * x) * x)

    def f_fd(x):
        return f(x), fd(x)
    return f, fd, f_fd


def quartic_shifted():
    """f = sum (x-1)^4 + (x-1)^2 (objectives.c ORC_OBJ_QUARTIC_SHIFTED): x* = 1, Hessian 2 at the minimiser."""
    def f(x):
        t = x - 1.0
        t2 = t * t
        return float(np.cumsum(t2 * t2 + t2)[-1])

    def fd(x):
        t = x - 1.0
        return 4.0 * ((t * t) * t) + 2.0 * t

    def f_fd(x):
        return f(x), fd(x)
    return f, fd, f_fd


def rosenbrock():
    def parts(x):
        a = x[0::2]; b = x[1::2]
        t1 = b - a * a; t2 = 1.0 - a
        return a, t1, t2

    def f(x):
        a, t1, t2 = parts(x)
        return float(np.cumsum((100.0 * t1) * t1 + t2 * t2)[-1])

    def fd(x):
        a, t1, t2 = parts(x)
        g = np.empty_like(x)
        g[0::2] = (-400.0 * a) * t1 - 2.0 * t2
        g[1::2] = 200.0 * t1
        return g

    def f_fd(x):
        return f(x), fd(x)
    return f, fd, f_fd


def diagquad(d):
    def f(x):
        t = x - 1.0
        return float(np.cumsum(((0.5 * d) * t) * t)[-1])

    def fd(x):
        return d * (x - 1.0)

    def f_fd(x):
        return f(x), fd(x)
    return f, fd, f_fd
