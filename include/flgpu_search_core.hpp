// flgpu_search_core.hpp -- the reference's line searchers as ONE piece of source for the host driver and for device
// code (libflgpu's own kernels and include/flgpu_objective.cuh).
//
// Wolfe / Wolfe_fdwithf (f90:1286-1459, quadratic zoom f90:1347-1370) and StrongWolfe / StrongWolfe_fdwithf
// (f90:1462-1698, cubic zoom f90:1557-1579) of NonlinearOptimization.f90, statement by statement.  The control flow
// lives here once; WHERE a trial point is formed and evaluated is supplied by the derived class (CRTP):
//     form(step)            x = x0 + step*p                         (f90:1482 and every later trial)
//     call_f / call_fd / call_ffd                                   the user callbacks
//     slope()               dot_product(fdx, p) at the last gradient
//     fx() / set_fx(v)      Fortran `fx` (may be fetched lazily)
//     adopt_pre()           the first trial was already formed/evaluated by the caller's chain (pre != 0)
//     count_f_only()        statistics hook (branch D's f-only probes, f90:1518-1520)
//     aborted()             checked at the top of every loop; the host version always answers false (the reference has
//                           no such exit: a NaN objective makes its zoom spin forever, f90:1684,1695), the device
//                           version answers true after an evaluation budget so a kernel can never hang the GPU
// driver.cpp derives the host version (asynchronous kernels + one host round trip per evaluation); a device kernel can
// derive a version whose evaluations are grid-wide cooperative reductions.  Arithmetic that the host compiles without
// FMA contraction (-ffp-contract=off) is written with nf_mul / nf_add so device code rounds identically.
#pragma once
#include <cmath>

#ifdef __CUDACC__
#define FLGPU_SC_HD __host__ __device__
#else
#define FLGPU_SC_HD
#endif

namespace flgpu {

FLGPU_SC_HD inline double nf_mul(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
FLGPU_SC_HD inline double nf_add(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}

template <class Derived>
struct SearchCore {
    double c1 = 0.0, c2abs = 0.0, fx0 = 0.0, phid0 = 0.0, incr = 0.0;
    bool fdwithf = false;
    double a = 0.0;        // Fortran `a`
    // the first trial may already have been formed/evaluated by the caller's chain:
    // 0 nothing, 2 x formed and f known, 3 x formed, f, f' and f'.p known
    int pre = 0;
    double pre_f = 0.0, pre_gp = 0.0;

    FLGPU_SC_HD Derived &self() { return *static_cast<Derived *>(this); }
    FLGPU_SC_HD void both() { if (fdwithf) self().call_ffd(); else { self().call_f(); self().call_fd(); } }
    // The reference writes the sufficient-decrease test in two forms, `fx>fx0+c1*a*phid0` (f90:1311,1356,1502,1568,...)
    // and `fx<=fx0+c1*a*phid0` (f90:1307,1328,1483,1521,...).  They are complements except for a NaN objective value,
    // where both are false -- so both forms exist here and each site uses the one the reference uses.
    FLGPU_SC_HD bool armijo_violated() { return self().fx() > nf_add(fx0, nf_mul(nf_mul(c1, a), phid0)); }
    FLGPU_SC_HD bool armijo_ok() { return self().fx() <= nf_add(fx0, nf_mul(nf_mul(c1, a), phid0)); }
    FLGPU_SC_HD static bool collapsed(double low, double up) {
        return fabs(up - low) < 1e-15 || fabs(up - low) / fmax(fabs(low), fabs(up)) < 1e-15;
    }

    // ---------------- Wolfe / Wolfe_fdwithf, f90:1286-1459 (quadratic zoom f90:1347-1370)
    FLGPU_SC_HD void wolfe_zoom(double &low, double &up, double &flow, double &fup, double &phidlow) {
        double phidlow_m_a = nf_mul(phidlow, a);
        for (;;) {
            if (self().aborted()) return;
            a = nf_mul(phidlow_m_a, a) / 2.0 / nf_add(nf_add(flow, phidlow_m_a), -fup);
            if (!(a > low && a < up)) a = (low + up) / 2.0;
            self().form(a); self().call_f();
            if (armijo_violated()) {
                up = a;
                if (up - low < 1e-15 || (up - low) / fmax(fabs(low), fabs(up)) < 1e-15) {
                    self().call_fd(); return;
                }
                fup = self().fx();
            } else {
                self().call_fd();
                const double phidnew = self().slope();
                if (phidnew > c2abs) return;
                low = a;
                if (up - low < 1e-15 || (up - low) / fmax(fabs(low), fabs(up)) < 1e-15) return;
                flow = self().fx(); phidlow = phidnew; phidlow_m_a = nf_mul(phidlow, a);
            }
        }
    }
    FLGPU_SC_HD void wolfe() {
        double aold, fold, atemp, ftemp, phidx;
        if (pre == 0) { self().form(a); self().call_f(); } else { self().adopt_pre(); }          // f90:1306
        if (armijo_ok()) {                                                           // f90:1307
            for (;;) {
            if (self().aborted()) return;
                aold = a; fold = self().fx();
                a = nf_mul(aold, incr); self().form(a); self().call_f();
                if (armijo_violated()) {
                    self().form(aold);                                               // f90:1312
                    self().call_fd();
                    phidx = self().slope();
                    if (phidx > c2abs) {
                        a = aold; self().set_fx(fold);
                    } else {
                        atemp = a; ftemp = self().fx();
                        wolfe_zoom(aold, atemp, fold, ftemp, phidx);
                    }
                    return;
                }
            }
        } else {
            for (;;) {
            if (self().aborted()) return;
                aold = a; fold = self().fx();
                a = aold / incr; self().form(a); self().call_f();
                if (armijo_ok()) {                                                   // f90:1328
                    self().call_fd();
                    phidx = self().slope();
                    if (phidx < c2abs) {
                        atemp = a; ftemp = self().fx();
                        wolfe_zoom(atemp, aold, ftemp, fold, phidx);
                    }
                    return;
                }
                if (a < 1e-15) { self().call_fd(); return; }
            }
        }
    }

    // ---------------- StrongWolfe / StrongWolfe_fdwithf, f90:1462-1698 (cubic zoom f90:1557-1579)
    // The six zoom arguments alias the caller's locals exactly as the Fortran by-reference dummies do.
    FLGPU_SC_HD void strong_zoom(double &low, double &up, double &flow, double &fup, double &phidlow, double &phidup) {
        for (;;) {
            if (self().aborted()) return;
            double d1 = nf_add(nf_add(phidlow, phidup), -(nf_mul(3.0, flow - fup) / (low - up)));
            double d2 = up - low;
            const double disc = nf_add(nf_mul(d1, d1), -nf_mul(phidlow, phidup));
            if (d2 > 0.0) d2 = sqrt(disc);
            else d2 = -sqrt(disc);
            a = nf_add(up, -(nf_mul(up - low, nf_add(nf_add(phidup, d2), -d1)) /
                             nf_add(nf_add(phidup, -phidlow), nf_mul(2.0, d2))));
            if (!(a > fmin(low, up) && a < fmax(low, up))) a = (low + up) / 2.0;
            self().form(a); both();
            const double phidnew = self().slope();
            if (armijo_violated() || self().fx() >= flow) {
                up = a; fup = self().fx(); phidup = phidnew;
            } else {
                if (fabs(phidnew) <= c2abs) return;
                if (nf_mul(phidnew, up - low) >= 0.0) { up = low; fup = flow; phidup = phidlow; }
                low = a; flow = self().fx(); phidlow = phidnew;
            }
            if (collapsed(low, up)) return;
        }
    }
    FLGPU_SC_HD void strongwolfe() {
        double aold = 0, fold = 0, atemp, ftemp, phidnew, phidold = 0;
        if (pre == 0) {                                                  // f90:1482 / 1604
            self().form(a);
            if (fdwithf) self().call_ffd(); else self().call_f();
        } else {
            self().adopt_pre();
        }
        if (armijo_ok()) {                                               // f90:1483 / 1605
            if (!fdwithf) self().call_fd();
            phidnew = (pre == 3) ? pre_gp : self().slope();
            if (phidnew > 0.0) {
                if (fabs(phidnew) <= c2abs) return;
                for (;;) {
            if (self().aborted()) return;                                               // f90:1488-1497
                    aold = a; fold = self().fx(); phidold = phidnew;
                    a = aold / incr; self().form(a); both(); phidnew = self().slope();
                    if (self().fx() >= fold || phidnew <= 0.0) {
                        atemp = a; ftemp = self().fx();
                        strong_zoom(aold, atemp, fold, ftemp, phidold, phidnew);
                        return;
                    }
                    if (a < 1e-15) return;
                }
            } else {
                for (;;) {
            if (self().aborted()) return;                                               // f90:1499-1515
                    aold = a; fold = self().fx(); phidold = phidnew;
                    a = nf_mul(aold, incr); self().form(a); both(); phidnew = self().slope();
                    if (armijo_violated() || self().fx() >= fold) {
                        atemp = a; ftemp = self().fx();
                        strong_zoom(aold, atemp, fold, ftemp, phidold, phidnew);
                        return;
                    }
                    if (phidnew > 0.0) {
                        if (fabs(phidnew) <= c2abs) return;
                        atemp = a; ftemp = self().fx();
                        strong_zoom(atemp, aold, ftemp, fold, phidnew, phidold);
                        if (fdwithf) return;                             // f90:1632
                        self().set_fx(fx0);                              // f90:1512 (no return there)
                    }
                }
            }
        } else {                                                         // f90:1517-1546
            for (;;) {
            if (self().aborted()) return;
                aold = a; fold = self().fx();
                a = aold / incr; self().form(a); self().call_f();
                self().count_f_only();
                if (armijo_ok()) {                                       // f90:1521 / 1640
                    self().call_fd();
                    phidnew = self().slope();
                    if (fabs(phidnew) <= c2abs) return;
                    if (phidnew < 0.0) {
                        self().form(aold); self().call_fd(); phidold = self().slope();   // f90:1526
                        atemp = a; ftemp = self().fx();
                        strong_zoom(atemp, aold, ftemp, fold, phidnew, phidold);
                        return;
                    } else {
                        for (;;) {
            if (self().aborted()) return;
                            aold = a; fold = self().fx(); phidold = phidnew;
                            a = aold / incr; self().form(a); both(); phidnew = self().slope();
                            if (self().fx() >= fold || phidnew <= 0.0) {
                                atemp = a; ftemp = self().fx();
                                strong_zoom(aold, atemp, fold, ftemp, phidold, phidnew);
                                return;
                            }
                            if (a < 1e-15) return;
                        }
                    }
                }
                if (a < 1e-15) { self().call_fd(); return; }
            }
        }
    }

    // ---------------- FLGPU_LS_FAST: NOT a reference routine (flgpu_options.line_search).
    // The reference's StrongWolfe never returns from a trial whose slope is still negative, even when that trial
    // already satisfies both Wolfe conditions: it keeps multiplying the step by Increment (1.05) until f rises or the
    // slope turns, then zooms (f90:1498-1515) -- 8 to 30 objective passes per accepted step.  This searcher
    // is the textbook bracketing/zoom scheme (Nocedal & Wright, Numerical Optimization, Alg. 3.5/3.6) with
    // More'-Thuente-style safeguards: every trial evaluates f and f' together and is ACCEPTED AS SOON AS it satisfies
    //     f(a) <= f(0) + c1 a phi'(0)   and   |phi'(a)| <= c2 |phi'(0)|   (weak form: phi'(a) >= -c2 |phi'(0)|);
    // while no bracket exists the step grows to the minimiser of the cubic through the last two trials, kept inside
    // [a + 1.1 (a - a_prev), a + 4 (a - a_prev)]; inside a bracket the cubic minimiser is kept 5 % away from both
    // ends.  A NaN or +inf objective counts as an Armijo violation (the bracket shrinks away from it).  If the
    // bracket collapses, or after 40 growth / 60 zoom steps, the best Armijo point found so far is returned.
    // The accepted step always satisfies the conditions above unless one of those exits fired.
    FLGPU_SC_HD bool fast_curvature_ok(bool strong, double phid) const {
        return strong ? fabs(phid) <= c2abs : phid >= -c2abs;
    }
    // minimiser of the cubic interpolating (u, fu, gu) and (v, fv, gv); NaN when it has none
    FLGPU_SC_HD static double cubic_minimiser(double u, double v, double fu, double fv, double gu, double gv) {
        const double d1 = nf_add(nf_add(gu, gv), -(nf_mul(3.0, fu - fv) / (u - v)));
        const double disc = nf_add(nf_mul(d1, d1), -nf_mul(gu, gv));
        const double d2 = (v - u > 0.0) ? sqrt(disc) : -sqrt(disc);
        return nf_add(v, -(nf_mul(v - u, nf_add(nf_add(gv, d2), -d1)) / nf_add(nf_add(gv, -gu), nf_mul(2.0, d2))));
    }
    // lo: best point with sufficient decrease so far (slope glo points towards hi); hi: the other end
    FLGPU_SC_HD void fast_zoom(bool strong, double lo, double hi, double flo, double fhi, double glo, double ghi) {
        for (int it = 0;; it++) {
            if (self().aborted()) return;
            const double w = hi - lo;
            double t = (cubic_minimiser(lo, hi, flo, fhi, glo, ghi) - lo) / w;
            if (!(t > 0.0 && t < 1.0)) t = 0.5;
            else if (t < 0.05) t = 0.05;
            else if (t > 0.95) t = 0.95;
            a = nf_add(lo, nf_mul(t, w));
            self().form(a); both();
            const double g = self().slope();
            const double f = self().fx();
            if (!armijo_ok() || f >= flo) {
                hi = a; fhi = f; ghi = g;
            } else {
                if (fast_curvature_ok(strong, g)) return;
                if (nf_mul(g, hi - lo) >= 0.0) { hi = lo; fhi = flo; ghi = glo; }
                lo = a; flo = f; glo = g;
            }
            if (collapsed(lo, hi) || it >= 59) {
                if (a != lo) { a = lo; self().form(a); both(); (void)self().slope(); }
                return;
            }
        }
    }
    // pre must be 0 or 3 (the caller's chain evaluated f, f' and f'.p at the first trial)
    FLGPU_SC_HD void fast(bool strong) {
        double a_lo = 0.0, f_lo = fx0, g_lo = phid0;
        if (pre == 0) { self().form(a); both(); } else { self().adopt_pre(); }
        double g = (pre == 3) ? pre_gp : self().slope();
        for (int grow = 0;; grow++) {
            if (self().aborted()) return;
            const double f = self().fx();
            if (!armijo_ok() || (grow > 0 && f >= f_lo)) { fast_zoom(strong, a_lo, a, f_lo, f, g_lo, g); return; }
            if (fast_curvature_ok(strong, g)) return;
            if (g >= 0.0) { fast_zoom(strong, a, a_lo, f, f_lo, g, g_lo); return; }
            if (grow >= 39) return;
            double an = cubic_minimiser(a_lo, a, f_lo, f, g_lo, g);
            const double lo_b = nf_add(a, nf_mul(1.1, a - a_lo)), hi_b = nf_add(a, nf_mul(4.0, a - a_lo));
            if (!(an <= hi_b)) an = hi_b;
            if (an < lo_b) an = lo_b;
            a_lo = a; f_lo = f; g_lo = g;
            a = an;
            self().form(a); both(); g = self().slope();
        }
    }
};

}  // namespace flgpu
