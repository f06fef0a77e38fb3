/*
 * oracle/oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  PARITY UNPINNED (see oracle.h).
 *
 * Statement-by-statement C restatement of the reference drivers and line searchers.
 * Each block cites the f90 lines it follows ("f90:" = /root/reference/source/
 * NonlinearOptimization.f90).  Arithmetic mimics gfortran -O3 on baseline x86-64:
 * strict left-to-right sums in dot_product, no FMA (build with -ffp-contract=off),
 * array expressions evaluated element by element in source operator order.
 *
 * The reference duplicates each main loop 8 times by presence of Increment /
 * presence of f_fd / value of Strong (f90:511-579, 241-344); the copies differ only
 * in which line searcher is called, so they are folded into run-time selection here.
 */
#include "oracle.h"

#include <math.h>
#include <setjmp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ OpenMP build (liboracle_omp.so ONLY)
 * The reference runs this path on ONE thread (no OpenMP / MPI / BLAS: SURVEY.md F4) and so does liboracle.so, the
 * checker: without -fopenmp the macros below are empty and this file is the strict sequential restatement, bit for
 * bit what it was.  `make liboracle_omp.so` (-O3 -march=native -fopenmp) turns the element-wise loops and the dot
 * products into parallel loops: a GENEROUS CPU baseline for bench.py, labelled as such.
 * Its sums are reassociated, so it is never used as a checker. */
#ifdef _OPENMP
#include <omp.h>
#define OMP_FOR _Pragma("omp parallel for schedule(static)")
#define OMP_FOR_SUM_S _Pragma("omp parallel for schedule(static) reduction(+:s)")
#else
#define OMP_FOR
#define OMP_FOR_SUM_S
#endif

/* threads the element-wise loops and dots run on: 1 for liboracle.so (the checker, = the reference) */
int orc_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------ state */
static orc_trace_t g_trace = NULL;
static void *g_trace_user = NULL;
static int g_sum_mode = 0;
static orc_stats_t g_st;

/* Evaluation budget (a test-harness guard, not reference behaviour).  The reference's searchers never terminate once
 * the step is NaN (f90:1518-1546: `a<1d-15` and both Armijo forms are false for NaN), and nothing a callback returns
 * can end such a loop.  With a budget set, the driver gives up after that many callback invocations: status 9, x left
 * as it was at that moment (the search's private copy of x0 is not freed).  0 = unlimited = the reference. */
static long g_eval_budget = 0, g_evals = 0;
static int g_armed = 0;
static jmp_buf g_abort;
void orc_set_eval_budget(long n) { g_eval_budget = n; }
static void count_eval(void) {
    if (g_armed && ++g_evals > g_eval_budget) { g_armed = 0; longjmp(g_abort, 1); }
}
#define ORC_ARM_BUDGET(label)                                                          \
    do {                                                                               \
        g_evals = 0; g_armed = 0;                                                      \
        if (g_eval_budget > 0) {                                                       \
            if (setjmp(g_abort)) { g_st.status = 9; goto label; }                      \
            g_armed = 1;                                                               \
        }                                                                              \
    } while (0)

void orc_set_trace(orc_trace_t cb, void *user) { g_trace = cb; g_trace_user = user; }
void orc_set_sum_mode(int mode) { g_sum_mode = mode; orc_obj_set_sum_mode(mode); }
void orc_get_stats(orc_stats_t *out) { *out = g_st; }

/* ------------------------------------------------------------------ primitives (a8) */
static double pairwise(const double *a, const double *b, const double *c, long n) {
    /* sum (a-b)*c or a*c if b==NULL, pairwise; only for noise bounding */
    if (n <= 32) {
        double s = 0.0;
        for (long i = 0; i < n; i++) s += (b ? (a[i] - b[i]) : a[i]) * c[i];
        return s;
    }
    long h = n / 2;
    return pairwise(a, b ? b : NULL, c, h) + pairwise(a + h, b ? b + h : NULL, c + h, n - h);
}

/* dot_product(a,b): gfortran inlines a sequential multiply-add loop */
static double dot(const double *a, const double *b, int n) {
    if (g_sum_mode == 1) {
        long double s = 0.0L;
        for (int i = 0; i < n; i++) s += (long double)a[i] * (long double)b[i];
        return (double)s;
    }
    if (g_sum_mode == 2) return pairwise(a, NULL, b, n);
    double s = 0.0;
    OMP_FOR_SUM_S
    for (int i = 0; i < n; i++) s += a[i] * b[i];
    return s;
}
/* dot_product(a-b,c): the temporary a-b is rounded to double before the multiply */
static double dot_diff_l(const double *a, const double *b, const double *c, int n) {
    if (g_sum_mode == 1) {
        long double s = 0.0L;
        for (int i = 0; i < n; i++) s += (long double)(a[i] - b[i]) * (long double)c[i];
        return (double)s;
    }
    if (g_sum_mode == 2) return pairwise(a, b, c, n);
    double s = 0.0;
    OMP_FOR_SUM_S
    for (int i = 0; i < n; i++) s += (a[i] - b[i]) * c[i];
    return s;
}
/* dot_product(c,a-b) */
static double dot_diff_r(const double *c, const double *a, const double *b, int n) {
    if (g_sum_mode == 1) {
        long double s = 0.0L;
        for (int i = 0; i < n; i++) s += (long double)c[i] * (long double)(a[i] - b[i]);
        return (double)s;
    }
    if (g_sum_mode == 2) return pairwise(a, b, c, n);
    double s = 0.0;
    OMP_FOR_SUM_S
    for (int i = 0; i < n; i++) s += c[i] * (a[i] - b[i]);
    return s;
}
/* array assignment a = b */
static void vcopy(double *dst, const double *src, int n) {
#ifdef _OPENMP
    OMP_FOR
    for (int i = 0; i < n; i++) dst[i] = src[i];
#else
    memcpy(dst, src, sizeof(double) * (size_t)n);
#endif
}

static double *valloc(long n) {
    double *v = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    if (!v) { fprintf(stderr, "oracle: out of memory\n"); abort(); }
    return v;
}

/* ------------------------------------------------------------------ line searchers */
typedef struct {
    double c1, c2_m_abs_phid0, fx0, phid0;
    orc_f_t f; orc_fd_t fd; orc_ffd_t f_fd;
    double *x; const double *x0; const double *p; double *fdx;
    double *a; double *fx;
    int dim;
} ls_t;

/* x=x0+a*p (e.g. f90:1482): element-wise, multiply then add, no FMA */
static void trial_x(ls_t *L) {
    const double a = *L->a;
    OMP_FOR
    for (int i = 0; i < L->dim; i++) L->x[i] = L->x0[i] + a * L->p[i];
    g_st.n_trials++;
}
static void call_f(ls_t *L) { count_eval(); L->f(L->fx, L->x, &L->dim); g_st.n_f++; }
static void call_fd(ls_t *L) { count_eval(); L->fd(L->fdx, L->x, &L->dim); g_st.n_fd++; }
static void call_ffd(ls_t *L) { count_eval(); (void)L->f_fd(L->fx, L->fdx, L->x, &L->dim); g_st.n_ffd++; }
static double slope(ls_t *L) { return dot(L->fdx, L->p, L->dim); }
static int armijo_ok(ls_t *L) { /* fx<=fx0+c1*a*phid0 (f90:1307,1328,1483,1521): NOT the complement of the next for NaN */
    return *L->fx <= L->fx0 + L->c1 * (*L->a) * L->phid0;
}
static int armijo_violated(ls_t *L) { /* fx>fx0+c1*a*phid0 */
    return *L->fx > L->fx0 + L->c1 * (*L->a) * L->phid0;
}
static int collapsed(double low, double up) { /* f90:1358,1577 */
    return fabs(up - low) < 1e-15 || fabs(up - low) / fmax(fabs(low), fabs(up)) < 1e-15;
}

/* zoom of Wolfe / Wolfe_fdwithf, f90:1347-1370 == 1435-1458 (quadratic interpolation).
 * low < up there, so up-low needs no abs in the reference; kept literal. */
static void wolfe_zoom(ls_t *L, double *low, double *up, double *flow, double *fup, double *phidlow) {
    double phidnew, phidlow_m_a;
    phidlow_m_a = *phidlow * (*L->a);                                       /* f90:1350 */
    for (;;) {
        *L->a = phidlow_m_a * (*L->a) / 2.0 / (*flow + phidlow_m_a - *fup); /* f90:1353 */
        if (!(*L->a > *low && *L->a < *up)) *L->a = (*low + *up) / 2.0;     /* f90:1354 */
        trial_x(L); call_f(L);                                              /* f90:1355 */
        if (armijo_violated(L)) {                                           /* f90:1356 */
            *up = *L->a;
            if (*up - *low < 1e-15 || (*up - *low) / fmax(fabs(*low), fabs(*up)) < 1e-15) {
                call_fd(L); return;                                         /* f90:1358-1360 */
            }
            *fup = *L->fx;
        } else {
            call_fd(L); phidnew = slope(L);                                 /* f90:1363 */
            if (phidnew > L->c2_m_abs_phid0) return;                        /* f90:1364 */
            *low = *L->a;
            if (*up - *low < 1e-15 || (*up - *low) / fmax(fabs(*low), fabs(*up)) < 1e-15) return;
            *flow = *L->fx; *phidlow = phidnew; phidlow_m_a = *phidlow * (*L->a); /* f90:1367 */
        }
    }
}

/* Wolfe f90:1286-1371 and Wolfe_fdwithf f90:1373-1459: the two bodies are textually
 * identical apart from the unused f_fd dummy ("CURRENTLY NO BETTER THAN Wolfe"). */
static void wolfe_impl(double c1, double c2, orc_f_t f, orc_fd_t fd, double *x, double *a,
                       const double *p, double *fx, double phid0, double *fdx, int dim,
                       const double *Increment) {
    double incrmt, ftemp, atemp, aold, fold, phidx;
    double *x0 = valloc(dim);
    ls_t L;
    g_st.n_linesearch++;
    if (Increment) incrmt = fmax(1.0 + 1e-15, *Increment); else incrmt = 1.05; /* f90:1302-1303 */
    vcopy(x0, x, dim);                               /* f90:1304 */
    L.c1 = c1; L.c2_m_abs_phid0 = c2 * fabs(phid0); L.fx0 = *fx; L.phid0 = phid0;
    L.f = f; L.fd = fd; L.f_fd = NULL; L.x = x; L.x0 = x0; L.p = p; L.fdx = fdx; L.a = a; L.fx = fx;
    L.dim = dim;
    trial_x(&L); call_f(&L);                                                   /* f90:1306 */
    if (armijo_ok(&L)) {                                                /* f90:1307 */
        for (;;) {
            aold = *a; fold = *fx;
            *a = aold * incrmt; trial_x(&L); call_f(&L);                       /* f90:1310 */
            if (armijo_violated(&L)) {                                         /* f90:1311 */
                double save = *a;
                *a = aold; trial_x(&L); *a = save;                             /* x=x0+aold*p f90:1312 */
                call_fd(&L);
                phidx = slope(&L);
                if (phidx > L.c2_m_abs_phid0) {
                    *a = aold; *fx = fold;                                     /* f90:1316 */
                } else {
                    atemp = *a; ftemp = *fx;
                    wolfe_zoom(&L, &aold, &atemp, &fold, &ftemp, &phidx);      /* f90:1319 */
                }
                free(x0); return;
            }
        }
    } else {                                                                   /* f90:1324 */
        for (;;) {
            aold = *a; fold = *fx;
            *a = aold / incrmt; trial_x(&L); call_f(&L);                       /* f90:1327 */
            if (armijo_ok(&L)) {
                call_fd(&L);
                phidx = slope(&L);
                if (phidx < L.c2_m_abs_phid0) {                                /* f90:1331 */
                    atemp = *a; ftemp = *fx;
                    wolfe_zoom(&L, &atemp, &aold, &ftemp, &fold, &phidx);
                }
                free(x0); return;
            }
            if (*a < 1e-15) { call_fd(&L); free(x0); return; }                 /* f90:1337-1339 */
        }
    }
}

void orc_wolfe(const double *c1, const double *c2, orc_f_t f, orc_fd_t fd, double *x, double *a,
               const double *p, double *fx, const double *phid0, double *fdx, const int *dim,
               const double *Increment) {
    wolfe_impl(*c1, *c2, f, fd, x, a, p, fx, *phid0, fdx, *dim, Increment);
}
void orc_wolfe_fdwithf(const double *c1, const double *c2, orc_f_t f, orc_fd_t fd, orc_ffd_t f_fd,
                       double *x, double *a, const double *p, double *fx, const double *phid0,
                       double *fdx, const int *dim, const double *Increment) {
    (void)f_fd; /* never called by the reference either, f90:1373-1459 */
    wolfe_impl(*c1, *c2, f, fd, x, a, p, fx, *phid0, fdx, *dim, Increment);
}

/* one f and one f' at the current trial: "call f; call fd" (StrongWolfe) or "f_fd" (_fdwithf) */
static void eval_both(ls_t *L, int fdwithf) {
    if (fdwithf) call_ffd(L); else { call_f(L); call_fd(L); }
}

/* zoom of StrongWolfe f90:1557-1579 == StrongWolfe_fdwithf f90:1675-1697 (cubic).
 * All six arguments are by reference and ALIAS the caller's locals, which matters
 * for the fall-through at f90:1511-1512. */
static void strong_zoom(ls_t *L, int fdwithf, double *low, double *up, double *flow, double *fup,
                        double *phidlow, double *phidup) {
    double phidnew, d1, d2;
    for (;;) {
        d1 = *phidlow + *phidup - 3.0 * (*flow - *fup) / (*low - *up); d2 = *up - *low; /* f90:1562 */
        if (d2 > 0.0) d2 = sqrt(d1 * d1 - *phidlow * *phidup);
        else d2 = -sqrt(d1 * d1 - *phidlow * *phidup);                                  /* f90:1563-1564 */
        *L->a = *up - (*up - *low) * (*phidup + d2 - d1) / (*phidup - *phidlow + 2.0 * d2); /* f90:1565 */
        if (!(*L->a > fmin(*low, *up) && *L->a < fmax(*low, *up))) *L->a = (*low + *up) / 2.0;
        trial_x(L); eval_both(L, fdwithf); phidnew = slope(L);                          /* f90:1567 */
        if (armijo_violated(L) || *L->fx >= *flow) {                                    /* f90:1568 */
            *up = *L->a; *fup = *L->fx; *phidup = phidnew;
        } else {
            if (fabs(phidnew) <= L->c2_m_abs_phid0) return;                             /* f90:1571 */
            if (phidnew * (*up - *low) >= 0.0) {
                *up = *low; *fup = *flow; *phidup = *phidlow;
            }
            *low = *L->a; *flow = *L->fx; *phidlow = phidnew;
        }
        if (collapsed(*low, *up)) return;                                               /* f90:1577 */
    }
}

/* StrongWolfe f90:1462-1580 (fdwithf=0) and StrongWolfe_fdwithf f90:1582-1698 (fdwithf=1).
 * Differences kept: (1) which callbacks evaluate a trial, (2) the missing return after
 * the zoom at f90:1511-1512 exists only in StrongWolfe; _fdwithf returns (f90:1631-1632),
 * (3) branch D's first shrink loop uses separate f then fd in BOTH (f90:1520-1526,1639-1644). */
static void strongwolfe_impl(int fdwithf, double c1, double c2, orc_f_t f, orc_fd_t fd,
                             orc_ffd_t f_fd, double *x, double *a, const double *p, double *fx,
                             double phid0, double *fdx, int dim, const double *Increment) {
    double incrmt, ftemp, atemp, aold, fold, phidnew, phidold;
    double *x0 = valloc(dim);
    ls_t L;
    g_st.n_linesearch++;
    if (Increment) incrmt = fmax(1.0 + 1e-15, *Increment); else incrmt = 1.05;  /* f90:1478-1479 */
    vcopy(x0, x, dim);                                /* f90:1480 */
    L.c1 = c1; L.c2_m_abs_phid0 = c2 * fabs(phid0); L.fx0 = *fx; L.phid0 = phid0;
    L.f = f; L.fd = fd; L.f_fd = f_fd; L.x = x; L.x0 = x0; L.p = p; L.fdx = fdx; L.a = a; L.fx = fx;
    L.dim = dim;
    phidold = 0.0; aold = 0.0; fold = 0.0;
    /* f90:1482 / f90:1604 */
    trial_x(&L);
    if (fdwithf) call_ffd(&L); else call_f(&L);
    if (armijo_ok(&L)) {                                                 /* f90:1483 */
        if (!fdwithf) call_fd(&L);                                              /* f90:1484 */
        phidnew = slope(&L);
        if (phidnew > 0.0) {                                                    /* f90:1486 */
            if (fabs(phidnew) <= L.c2_m_abs_phid0) { free(x0); return; }
            for (;;) {                                                          /* f90:1488-1497 */
                aold = *a; fold = *fx; phidold = phidnew;
                *a = aold / incrmt; trial_x(&L); eval_both(&L, fdwithf); phidnew = slope(&L);
                if (*fx >= fold || phidnew <= 0.0) {
                    atemp = *a; ftemp = *fx;
                    strong_zoom(&L, fdwithf, &aold, &atemp, &fold, &ftemp, &phidold, &phidnew);
                    free(x0); return;
                }
                if (*a < 1e-15) { free(x0); return; }
            }
        } else {                                                                /* f90:1498-1515 */
            for (;;) {
                aold = *a; fold = *fx; phidold = phidnew;
                *a = aold * incrmt; trial_x(&L); eval_both(&L, fdwithf); phidnew = slope(&L);
                if (armijo_violated(&L) || *fx >= fold) {                       /* f90:1502 */
                    atemp = *a; ftemp = *fx;
                    strong_zoom(&L, fdwithf, &aold, &atemp, &fold, &ftemp, &phidold, &phidnew);
                    free(x0); return;
                }
                if (phidnew > 0.0) {                                            /* f90:1507 */
                    if (fabs(phidnew) <= L.c2_m_abs_phid0) { free(x0); return; }
                    atemp = *a; ftemp = *fx;
                    strong_zoom(&L, fdwithf, &atemp, &aold, &ftemp, &fold, &phidnew, &phidold);
                    if (fdwithf) { free(x0); return; }                          /* f90:1632 */
                    *fx = L.fx0;                                                /* f90:1512: no return */
                    g_st.n_quirk_f9++;
                }
            }
        }
    } else {                                                                    /* f90:1517-1546 */
        for (;;) {
            aold = *a; fold = *fx;
            *a = aold / incrmt; trial_x(&L); call_f(&L);                        /* f90:1520 */
            if (armijo_ok(&L)) {                                         /* f90:1521 */
                call_fd(&L);
                phidnew = slope(&L);
                if (fabs(phidnew) <= L.c2_m_abs_phid0) { free(x0); return; }    /* f90:1524 */
                if (phidnew < 0.0) {                                            /* f90:1525-1529 */
                    double save = *a;
                    *a = aold; trial_x(&L); *a = save;                          /* x=x0+aold*p */
                    call_fd(&L); phidold = slope(&L);
                    atemp = *a; ftemp = *fx;
                    strong_zoom(&L, fdwithf, &atemp, &aold, &ftemp, &fold, &phidnew, &phidold);
                    free(x0); return;
                } else {                                                        /* f90:1530-1540 */
                    for (;;) {
                        aold = *a; fold = *fx; phidold = phidnew;
                        *a = aold / incrmt; trial_x(&L); eval_both(&L, fdwithf); phidnew = slope(&L);
                        if (*fx >= fold || phidnew <= 0.0) {
                            atemp = *a; ftemp = *fx;
                            strong_zoom(&L, fdwithf, &aold, &atemp, &fold, &ftemp, &phidold, &phidnew);
                            free(x0); return;
                        }
                        if (*a < 1e-15) { free(x0); return; }
                    }
                }
            }
            if (*a < 1e-15) { call_fd(&L); free(x0); return; }                  /* f90:1543-1545 */
        }
    }
}

void orc_strongwolfe(const double *c1, const double *c2, orc_f_t f, orc_fd_t fd, double *x, double *a,
                     const double *p, double *fx, const double *phid0, double *fdx, const int *dim,
                     const double *Increment) {
    strongwolfe_impl(0, *c1, *c2, f, fd, NULL, x, a, p, fx, *phid0, fdx, *dim, Increment);
}
void orc_strongwolfe_fdwithf(const double *c1, const double *c2, orc_f_t f, orc_fd_t fd,
                             orc_ffd_t f_fd, double *x, double *a, const double *p, double *fx,
                             const double *phid0, double *fdx, const int *dim,
                             const double *Increment) {
    strongwolfe_impl(1, *c1, *c2, f, fd, f_fd, x, a, p, fx, *phid0, fdx, *dim, Increment);
}

/* ------------------------------------------------------------------ FLGPU_LS_FAST (NOT a reference routine)
 * Restates the product's optional accept-at-first-Wolfe-point searcher (include/flgpu_search_core.hpp,
 * SearchCore::fast; flgpu_options.line_search = FLGPU_LS_FAST) so that the host-driven and the
 * device-resident implementations can be checked bit for bit.  The reference has no counterpart: with this policy the
 * oracle pins the PRODUCT's published algorithm (Nocedal & Wright Alg. 3.5/3.6 + the safeguards below), nothing else.
 * Selected by orc_set_line_search(1); the drivers below are unchanged (they all search through line_search()). */
static int g_ls_policy = 0;
void orc_set_line_search(int policy) { g_ls_policy = policy; }

static double cubic_minimiser(double u, double v, double fu, double fv, double gu, double gv) {
    const double d1 = gu + gv - 3.0 * (fu - fv) / (u - v);
    const double disc = d1 * d1 - gu * gv;
    const double d2 = (v - u > 0.0) ? sqrt(disc) : -sqrt(disc);
    return v - (v - u) * (gv + d2 - d1) / (gv - gu + 2.0 * d2);
}
static int fast_curvature_ok(const ls_t *L, int strong, double phid) {
    return strong ? fabs(phid) <= L->c2_m_abs_phid0 : phid >= -L->c2_m_abs_phid0;
}
static void fast_zoom(ls_t *L, int strong, int ffd, double lo, double hi, double flo, double fhi, double glo,
                      double ghi) {
    int it;
    for (it = 0;; it++) {
        const double w = hi - lo;
        double t = (cubic_minimiser(lo, hi, flo, fhi, glo, ghi) - lo) / w, g, f;
        if (!(t > 0.0 && t < 1.0)) t = 0.5;             /* no interior minimiser (or NaN): bisect */
        else if (t < 0.05) t = 0.05;
        else if (t > 0.95) t = 0.95;
        *L->a = lo + t * w;
        trial_x(L); eval_both(L, ffd); g = slope(L); f = *L->fx;
        if (!armijo_ok(L) || f >= flo) {
            hi = *L->a; fhi = f; ghi = g;
        } else {
            if (fast_curvature_ok(L, strong, g)) return;
            if (g * (hi - lo) >= 0.0) { hi = lo; fhi = flo; ghi = glo; }
            lo = *L->a; flo = f; glo = g;
        }
        if (collapsed(lo, hi) || it >= 59) {            /* give back the best sufficient-decrease point */
            if (*L->a != lo) { *L->a = lo; trial_x(L); eval_both(L, ffd); }
            return;
        }
    }
}
static void fast_impl(int strong, double c1, double c2, orc_f_t f, orc_fd_t fd, orc_ffd_t f_fd, double *x, double *a,
                      const double *p, double *fx, double phid0, double *fdx, int dim) {
    const int ffd = f_fd != NULL;
    double a_lo = 0.0, f_lo = *fx, g_lo = phid0, g;
    double *x0 = valloc(dim);
    int grow;
    ls_t L;
    g_st.n_linesearch++;
    vcopy(x0, x, dim);
    L.c1 = c1; L.c2_m_abs_phid0 = c2 * fabs(phid0); L.fx0 = *fx; L.phid0 = phid0;
    L.f = f; L.fd = fd; L.f_fd = f_fd; L.x = x; L.x0 = x0; L.p = p; L.fdx = fdx; L.a = a; L.fx = fx;
    L.dim = dim;
    trial_x(&L); eval_both(&L, ffd); g = slope(&L);
    for (grow = 0;; grow++) {
        const double fa = *fx;
        double an, lo_b, hi_b;
        if (!armijo_ok(&L) || (grow > 0 && fa >= f_lo)) { fast_zoom(&L, strong, ffd, a_lo, *a, f_lo, fa, g_lo, g); break; }
        if (fast_curvature_ok(&L, strong, g)) break;
        if (g >= 0.0) { fast_zoom(&L, strong, ffd, *a, a_lo, fa, f_lo, g, g_lo); break; }
        if (grow >= 39) break;
        an = cubic_minimiser(a_lo, *a, f_lo, fa, g_lo, g);
        lo_b = *a + 1.1 * (*a - a_lo); hi_b = *a + 4.0 * (*a - a_lo);
        if (!(an <= hi_b)) an = hi_b;
        if (an < lo_b) an = lo_b;
        a_lo = *a; f_lo = fa; g_lo = g;
        *a = an;
        trial_x(&L); eval_both(&L, ffd); g = slope(&L);
    }
    free(x0);
}

/* Dispatch that every driver's 8-way (or 4-way) textual copy reduces to. */
static void line_search(int strong, int use_ffd, double c1, double c2, orc_f_t f, orc_fd_t fd,
                        orc_ffd_t f_fd, double *x, double *a, const double *p, double *fx,
                        double phid0, double *fdx, int dim, const double *Increment) {
    if (g_ls_policy == 1) { fast_impl(strong, c1, c2, f, fd, f_fd, x, a, p, fx, phid0, fdx, dim); return; }
    if (strong) strongwolfe_impl(use_ffd, c1, c2, f, fd, f_fd, x, a, p, fx, phid0, fdx, dim, Increment);
    else wolfe_impl(c1, c2, f, fd, x, a, p, fx, phid0, fdx, dim, Increment);
}

static void trace(int iter, int dim, const double *p, const double *x, const double *g, double a,
                  double fx, double phid0, long trials_before) {
    g_st.n_iter = iter + 1;
    if (g_trace) g_trace(g_trace_user, iter, dim, p, x, g, a, fx, phid0, g_st.n_trials - trials_before);
}

/* ------------------------------------------------------------------ LBFGS f90:398-625 */
void orc_lbfgs(orc_f_t f, orc_fd_t fd, double *x, const int *dim_, const int *Memory, orc_ffd_t f_fd,
               const int *Strong, const int *Warning, const int *MaxIteration, const double *Precision,
               const double *MinStepLength, const double *WolfeConst1, const double *WolfeConst2,
               const double *Increment) {
    const int dim = *dim_;
    int sw, warn, mem, maxit, iIteration, i, recent, outer = 0;
    double tol, minstep, c1, c2, a, fnew, phidnew, phid0;
    double *p, *fdnew, *xold, *fdold, *rho, *alpha, *s, *y;
    long tb;
    memset(&g_st, 0, sizeof g_st);
    /* f90:419-434 */
    if (Memory) mem = (*Memory > 1 ? *Memory : 1); else mem = 10;
    if (Strong) sw = (*Strong != 0); else sw = 1;
    if (Warning) warn = (*Warning != 0); else warn = 1;
    if (MaxIteration) maxit = *MaxIteration; else maxit = 1000;
    if (Precision) tol = *Precision * *Precision; else tol = 1e-30;
    if (MinStepLength) minstep = *MinStepLength * *MinStepLength; else minstep = 1e-30;
    if (WolfeConst1) c1 = fmax(1e-15, *WolfeConst1); else c1 = 1e-4;
    if (WolfeConst2) c2 = fmin(1.0 - 1e-15, fmax(c1 + 1e-15, *WolfeConst2)); else c2 = 0.9;
    p = valloc(dim); fdnew = valloc(dim); xold = valloc(dim); fdold = valloc(dim);
    /* f90:435 allocates (dim,0:mem); column mem is never touched, so mem columns suffice */
    rho = valloc(mem + 1); alpha = valloc(mem + 1);
    s = valloc((long)dim * mem); y = valloc((long)dim * mem);
#define S(i) (s + (long)(i) * dim)
#define Y(i) (y + (long)(i) * dim)
    ORC_ARM_BUDGET(done);
    /* f90:436-440 */
    if (f_fd) { (void)f_fd(&fnew, fdnew, x, &dim); g_st.n_ffd++; }
    else { f(&fnew, x, &dim); g_st.n_f++; fd(fdnew, x, &dim); g_st.n_fd++; }
    /* f90:442-446 */
    OMP_FOR for (i = 0; i < dim; i++) p[i] = -fdnew[i];
    phidnew = -dot(fdnew, fdnew, dim);
    if (-phidnew < tol) { g_st.status = 3; goto done; }
    if (fnew == 0.0) a = 1.0; else a = fabs(fnew) / sqrt(-phidnew);
    vcopy(xold, x, dim); vcopy(fdold, fdnew, dim);
    /* f90:448-460: never the _fdwithf variant here */
    tb = g_st.n_trials; phid0 = phidnew;
    line_search(sw, 0, c1, c2, f, fd, f_fd, x, &a, p, &fnew, phidnew, fdnew, dim, Increment);
    trace(outer++, dim, p, x, fdnew, a, fnew, phid0, tb);
    /* f90:461-469 */
    phidnew = dot(fdnew, fdnew, dim);
    if (phidnew < tol) { g_st.status = 0; goto done; }
    if (dot(p, p, dim) * a * a < minstep) {
        if (warn) {
            printf(" BFGS warning: step length has converged, but gradient norm has not met accuracy goal\n");
            printf(" Euclidean norm of gradient = %.17g\n", sqrt(phidnew));
        }
        g_st.status = 1; goto done;
    }
    /* f90:470-471 */
    recent = 0;
    OMP_FOR for (i = 0; i < dim; i++) { S(0)[i] = x[i] - xold[i]; Y(0)[i] = fdnew[i] - fdold[i]; }
    rho[0] = 1.0 / dot(Y(0), S(0), dim);
    /* f90:472-510 pre-iterations */
    for (iIteration = 1; iIteration <= mem - 1; iIteration++) {
        int k;
        vcopy(xold, x, dim); vcopy(fdold, fdnew, dim);
        vcopy(p, fdnew, dim);                               /* f90:475 */
        for (i = recent; i >= 0; i--) {                                               /* f90:476-479 */
            alpha[i] = rho[i] * dot(S(i), p, dim);
            OMP_FOR for (k = 0; k < dim; k++) p[k] = p[k] - alpha[i] * Y(i)[k];
        }
        {   /* f90:480 p=p/rho(recent)/dot_product(y,y) */
            const double r = rho[recent], yy = dot(Y(recent), Y(recent), dim);
            OMP_FOR for (k = 0; k < dim; k++) p[k] = p[k] / r / yy;
        }
        for (i = 0; i <= recent; i++) {                                               /* f90:481-484 */
            phidnew = rho[i] * dot(Y(i), p, dim);
            { const double c = alpha[i] - phidnew; OMP_FOR for (k = 0; k < dim; k++) p[k] = p[k] + c * S(i)[k]; }
        }
        OMP_FOR for (k = 0; k < dim; k++) p[k] = -p[k];
        phidnew = dot(fdnew, p, dim); a = 1.0;                                        /* f90:485 */
        tb = g_st.n_trials; phid0 = phidnew;
        line_search(sw, 0, c1, c2, f, fd, f_fd, x, &a, p, &fnew, phidnew, fdnew, dim, Increment);
        trace(outer++, dim, p, x, fdnew, a, fnew, phid0, tb);
        phidnew = dot(fdnew, fdnew, dim);                                             /* f90:499-507 */
        if (phidnew < tol) { g_st.status = 0; goto done; }
        if (dot(p, p, dim) * a * a < minstep) {
            if (warn) {
                printf(" BFGS warning: step length has converged, but gradient norm has not met accuracy goal\n");
                printf(" Euclidean norm of gradient = %.17g\n", sqrt(phidnew));
            }
            g_st.status = 1; goto done;
        }
        recent = recent + 1;                                                          /* f90:508-509 */
        OMP_FOR for (k = 0; k < dim; k++) { S(recent)[k] = x[k] - xold[k]; Y(recent)[k] = fdnew[k] - fdold[k]; }
        rho[recent] = 1.0 / dot(Y(recent), S(recent), dim);
    }
    /* f90:511-579 main loop */
    for (iIteration = 1; iIteration <= maxit; iIteration++) {
        int k;
        /* Before() f90:586-608 */
        vcopy(xold, x, dim); vcopy(fdold, fdnew, dim);
        vcopy(p, fdnew, dim);
        for (i = recent; i >= 0; i--) {
            alpha[i] = rho[i] * dot(S(i), p, dim);
            OMP_FOR for (k = 0; k < dim; k++) p[k] = p[k] - alpha[i] * Y(i)[k];
        }
        for (i = mem - 1; i >= recent + 1; i--) {
            alpha[i] = rho[i] * dot(S(i), p, dim);
            OMP_FOR for (k = 0; k < dim; k++) p[k] = p[k] - alpha[i] * Y(i)[k];
        }
        {
            const double r = rho[recent], yy = dot(Y(recent), Y(recent), dim);        /* f90:598 */
            OMP_FOR for (k = 0; k < dim; k++) p[k] = p[k] / r / yy;
        }
        for (i = recent + 1; i <= mem - 1; i++) {
            phidnew = rho[i] * dot(Y(i), p, dim);
            { const double c = alpha[i] - phidnew; OMP_FOR for (k = 0; k < dim; k++) p[k] = p[k] + c * S(i)[k]; }
        }
        for (i = 0; i <= recent; i++) {
            phidnew = rho[i] * dot(Y(i), p, dim);
            { const double c = alpha[i] - phidnew; OMP_FOR for (k = 0; k < dim; k++) p[k] = p[k] + c * S(i)[k]; }
        }
        OMP_FOR for (k = 0; k < dim; k++) p[k] = -p[k];
        phidnew = dot(fdnew, p, dim); a = 1.0;                                        /* f90:607 */
        /* line search: _fdwithf iff f_fd present */
        tb = g_st.n_trials; phid0 = phidnew;
        line_search(sw, f_fd != NULL, c1, c2, f, fd, f_fd, x, &a, p, &fnew, phidnew, fdnew, dim, Increment);
        trace(outer++, dim, p, x, fdnew, a, fnew, phid0, tb);
        /* After() f90:609-624 */
        phidnew = dot(fdnew, fdnew, dim);
        if (phidnew < tol) { g_st.status = 0; goto done; }
        if (dot(p, p, dim) * a * a < minstep) {
            if (warn) {
                printf(" L-BFGS warning: step length has converged, but gradient norm has not met accuracy goal\n");
                printf(" Euclidean norm of gradient = %.17g\n", sqrt(phidnew));
            }
            g_st.status = 1; goto done;
        }
        recent = (recent + 1) % mem;
        OMP_FOR for (k = 0; k < dim; k++) { S(recent)[k] = x[k] - xold[k]; Y(recent)[k] = fdnew[k] - fdold[k]; }
        rho[recent] = 1.0 / dot(Y(recent), S(recent), dim);
    }
    g_st.status = 2;
    if (warn) {                                                                       /* f90:580-583 */
        printf(" Failed L-BFGS: max iteration exceeded!\n");
        printf(" Euclidean norm of gradient = %.17g\n", sqrt(dot(fdnew, fdnew, dim)));
    }
done:
    g_armed = 0;
    free(p); free(fdnew); free(xold); free(fdold); free(rho); free(alpha); free(s); free(y);
#undef S
#undef Y
}

/* ------------------------------------------------------------------ CG f90:193-394, 2249-2346 */
/* DY() f90:352-372 / PR() f90:373-393; returns 1 to terminate */
static int cg_after(int is_pr, int warn, double tol, double minstep, int dim, double *p,
                    const double *fdnew, const double *fdold, double *a, double *phidnew,
                    double phidold) {
    int k;
    double beta;
    *phidnew = dot(fdnew, fdnew, dim);
    if (*phidnew < tol) { g_st.status = 0; return 1; }
    if (dot(p, p, dim) * *a * *a < minstep) {
        if (warn) {
            if (is_pr) printf(" Polak-Ribiere+ conjugate gradient warning: step length has converged, but gradient norm has not met accuracy goal\n");
            else printf(" Dai-Yuan conjugate gradient warning: step length has converged, but gradient norm has not met accuracy goal\n");
            printf(" Euclidean norm of gradient = %.17g\n", sqrt(*phidnew));
        }
        g_st.status = 1; return 1;
    }
    if (is_pr) beta = dot_diff_r(fdnew, fdnew, fdold, dim) / dot(fdold, fdold, dim);  /* f90:387 */
    else beta = dot(fdnew, fdnew, dim) / dot_diff_l(fdnew, fdold, p, dim);            /* f90:366 */
    OMP_FOR for (k = 0; k < dim; k++) p[k] = -fdnew[k] + beta * p[k];
    *phidnew = dot(fdnew, p, dim);
    if (*phidnew > 0.0) {                                                             /* f90:368-370 */
        OMP_FOR for (k = 0; k < dim; k++) p[k] = -fdnew[k];
        *phidnew = -dot(fdnew, fdnew, dim);
    }
    *a = *a * phidold / *phidnew;                                                     /* f90:371 */
    return 0;
}

static void cg_core(orc_f_t f, orc_fd_t fd, orc_ffd_t f_fd, double *x, int dim, const char *type,
                    int sw, int warn, int maxit, double tol, double minstep, double c1, double c2,
                    const double *Increment) {
    int iIteration, i, is_pr, outer = 0;
    double a, fnew, fold, phidnew, phidold;
    double *p = valloc(dim), *fdnew = valloc(dim), *fdold = valloc(dim);
    long tb;
    (void)fold;
    memset(&g_st, 0, sizeof g_st);
    ORC_ARM_BUDGET(done);
    if (f_fd) { (void)f_fd(&fnew, fdnew, x, &dim); g_st.n_ffd++; }                    /* f90:230-234 */
    else { f(&fnew, x, &dim); g_st.n_f++; fd(fdnew, x, &dim); g_st.n_fd++; }
    OMP_FOR for (i = 0; i < dim; i++) p[i] = -fdnew[i];                                       /* f90:236 */
    phidnew = -dot(fdnew, fdnew, dim);
    if (-phidnew < tol) { g_st.status = 3; goto done; }
    if (fnew == 0.0) a = 1.0; else a = fabs(fnew) / sqrt(-phidnew);
    if (type[0] == 'D' && type[1] == 'Y') is_pr = 0;
    else if (type[0] == 'P' && type[1] == 'R') is_pr = 1;
    else {                                                                            /* f90:345 */
        printf(" Program abort: unsupported conjugate gradient method %.2s\n", type);
        exit(1);
    }
    for (iIteration = 1; iIteration <= maxit; iIteration++) {
        const int strong = is_pr ? 1 : sw; /* PR always strong Wolfe, f90:311-344 */
        double phid0;
        fold = fnew; vcopy(fdold, fdnew, dim); phidold = phidnew;
        tb = g_st.n_trials; phid0 = phidnew;
        line_search(strong, f_fd != NULL, c1, c2, f, fd, f_fd, x, &a, p, &fnew, phidnew, fdnew, dim, Increment);
        trace(outer++, dim, p, x, fdnew, a, fnew, phid0, tb);
        if (cg_after(is_pr, warn, tol, minstep, dim, p, fdnew, fdold, &a, &phidnew, phidold)) goto done;
    }
    g_st.status = 2;
    if (warn) {                                                                       /* f90:347-350 */
        printf(" Failed conjugate gradient: max iteration exceeded!\n");
        printf(" Euclidean norm of gradient = %.17g\n", sqrt(dot(fdnew, fdnew, dim)));
    }
done:
    g_armed = 0;
    free(p); free(fdnew); free(fdold);
}

void orc_conjugategradient(orc_f_t f, orc_fd_t fd, double *x, const int *dim, const char *Method,
                           orc_ffd_t f_fd, const int *Strong, const int *Warning,
                           const int *MaxIteration, const double *Precision,
                           const double *MinStepLength, const double *WolfeConst1,
                           const double *WolfeConst2, const double *Increment, int len_Method) {
    char type[2] = {'D', 'Y'};
    int sw, warn, maxit;
    double tol, minstep, c1, c2;
    if (Method) { /* character*2 :: type = Method, blank padded, f90:207,214 */
        type[0] = len_Method > 0 ? Method[0] : ' ';
        type[1] = len_Method > 1 ? Method[1] : ' ';
    }
    if (Strong) sw = (*Strong != 0); else sw = 1;
    if (Warning) warn = (*Warning != 0); else warn = 1;
    if (MaxIteration) maxit = *MaxIteration; else maxit = 1000;
    if (Precision) tol = *Precision * *Precision; else tol = 1e-30;
    if (MinStepLength) minstep = *MinStepLength * *MinStepLength; else minstep = 1e-30;
    if (WolfeConst1) c1 = fmax(1e-15, *WolfeConst1); else c1 = 1e-4;
    if (WolfeConst2) c2 = fmin(1.0 - 1e-15, fmax(c1 + 1e-15, *WolfeConst2)); else c2 = 0.45;
    cg_core(f, fd, f_fd, x, *dim, type, sw, warn, maxit, tol, minstep, c1, c2, Increment);
}

void orc_conjugategradient_basic(orc_f_t f, orc_fd_t fd, double *x, const int *dim, const char *Method,
                                 const int *Strong, const int *Warning, const int *MaxIteration,
                                 const double *Precision, const double *MinStepLength,
                                 const double *WolfeConst1, const double *WolfeConst2,
                                 const double *Increment, int len_Method) {
    /* select case(Method) compares the whole string blank-padded, f90:2273 */
    char type[2];
    int j;
    type[0] = len_Method > 0 ? Method[0] : ' ';
    type[1] = len_Method > 1 ? Method[1] : ' ';
    for (j = 2; j < len_Method; j++) if (Method[j] != ' ') type[0] = '?';
    /* no clamps on c1/c2 (f90:2278); Increment still clamped inside the searcher */
    cg_core(f, fd, NULL, x, *dim, type, *Strong != 0, *Warning != 0, *MaxIteration,
            *Precision * *Precision, *MinStepLength * *MinStepLength, *WolfeConst1, *WolfeConst2,
            Increment);
}

/* ------------------------------------------------------------------ SteepestDescent f90:55-188 */
/* The 8 textual copies of the main loop (f90:101-167) differ only in the line searcher called:
 * Strong -> StrongWolfe(_fdwithf when f_fd is present), else Wolfe(_fdwithf == Wolfe, f90:1373). */
void orc_steepestdescent(orc_f_t f, orc_fd_t fd, double *x, const int *dim_, orc_ffd_t f_fd,
                         const int *Strong, const int *Warning, const int *MaxIteration,
                         const double *Precision, const double *MinStepLength, const double *WolfeConst1,
                         const double *WolfeConst2, const double *Increment) {
    int dim = *dim_, sw, warn, maxit, iIteration, i, outer = 0;
    double tol, minstep, c1, c2, a, fnew, fold, phidnew, phidold;
    double *p = valloc(dim), *fdnew = valloc(dim), *fdold = valloc(dim);
    long tb;
    (void)fold;
    memset(&g_st, 0, sizeof g_st);
    if (Strong) sw = (*Strong != 0); else sw = 1;                                     /* f90:72-85 */
    if (Warning) warn = (*Warning != 0); else warn = 1;
    if (MaxIteration) maxit = *MaxIteration; else maxit = 1000;
    if (Precision) tol = *Precision * *Precision; else tol = 1e-30;
    if (MinStepLength) minstep = *MinStepLength * *MinStepLength; else minstep = 1e-30;
    if (WolfeConst1) c1 = fmax(1e-15, *WolfeConst1); else c1 = 1e-4;
    if (WolfeConst2) c2 = fmin(1.0 - 1e-15, fmax(c1 + 1e-15, *WolfeConst2)); else c2 = 0.9;
    ORC_ARM_BUDGET(done);
    if (f_fd) { (void)f_fd(&fnew, fdnew, x, &dim); g_st.n_ffd++; }                    /* f90:86-90 */
    else { f(&fnew, x, &dim); g_st.n_f++; fd(fdnew, x, &dim); g_st.n_fd++; }
    OMP_FOR for (i = 0; i < dim; i++) p[i] = -fdnew[i];                                       /* f90:92 */
    phidnew = -dot(fdnew, fdnew, dim);
    if (-phidnew < tol) { g_st.status = 3; goto done; }
    if (fnew == 0.0) a = 1.0; else a = fabs(fnew) / sqrt(-phidnew);                   /* f90:94-95 */
    for (iIteration = 1; iIteration <= maxit; iIteration++) {
        double phid0;
        fold = fnew; vcopy(fdold, fdnew, dim); phidold = phidnew;
        tb = g_st.n_trials; phid0 = phidnew;
        line_search(sw, f_fd != NULL, c1, c2, f, fd, f_fd, x, &a, p, &fnew, phidnew, fdnew, dim, Increment);
        trace(outer++, dim, p, x, fdnew, a, fnew, phid0, tb);
        /* After() f90:172-187 */
        phidnew = dot(fdnew, fdnew, dim);
        if (phidnew < tol) { g_st.status = 0; goto done; }
        if (dot(p, p, dim) * a * a < minstep) {
            if (warn) {
                printf(" Steepest descent warning: step length has converged, but gradient norm has not met accuracy goal\n");
                printf(" Euclidean norm of gradient = %.17g\n", sqrt(phidnew));
            }
            g_st.status = 1; goto done;
        }
        OMP_FOR for (i = 0; i < dim; i++) p[i] = -fdnew[i];
        phidnew = -dot(fdnew, fdnew, dim);
        a = a * phidold / phidnew;
    }
    g_st.status = 2;
    if (warn) {                                                                       /* f90:168-171 */
        printf(" Failed steepest descent: max iteration exceeded!\n");
        printf(" Euclidean norm of gradient = %.17g\n", sqrt(dot(fdnew, fdnew, dim)));
    }
done:
    g_armed = 0;
    free(p); free(fdnew); free(fdold);
}

/* ------------------------------------------------------------------ AugmentedLagrangian f90:2005-2241 */
/* Only the branches that call the hot path are restated: UnconstrainedSolver = 'LBFGS' (f90:2150-2167) and
 * 'ConjugateGradient' (f90:2168-2185).  The internal procedures L, Ld, L_Ld, L_Ld_fdwithf (f90:2193-2228) reach
 * the host variables lambda, miu, cx, cdx by host association; C has no closures, so they live in this struct. */
static struct {
    orc_f_t f; orc_fd_t fd; orc_ffd_t f_fd; orc_c_t c; orc_cd_t cd;
    int M; double miu; double *lambda, *cx, *cdx;
} AL;
static orc_al_stats_t g_al;

static void al_terms(double *Lx, int N) {                /* Lx = Lx - dot(lambda,cx) + miu/2*dot(cx,cx) */
    double d1 = 0.0, d2 = 0.0;
    int j;
    (void)N;
    for (j = 0; j < AL.M; j++) d1 += AL.lambda[j] * AL.cx[j];
    for (j = 0; j < AL.M; j++) d2 += AL.cx[j] * AL.cx[j];
    *Lx = *Lx - d1 + AL.miu / 2.0 * d2;
}
static void al_grad(double *Ldx, int N) {                /* Ldx = Ldx + matmul(cdx, miu*cx-lambda), cdx(N,M) */
    int i, j;
    for (i = 0; i < N; i++) {
        double r = 0.0;
        for (j = 0; j < AL.M; j++) r += AL.cdx[(size_t)j * (size_t)N + i] * (AL.miu * AL.cx[j] - AL.lambda[j]);
        Ldx[i] = Ldx[i] + r;
    }
}
static void al_L(double *Lx, const double *x, const int *N) {                  /* f90:2193-2199 */
    AL.f(Lx, x, N); AL.c(AL.cx, x, &AL.M, N);
    al_terms(Lx, *N);
}
static void al_Ld(double *Ldx, const double *x, const int *N) {                /* f90:2200-2206 */
    AL.fd(Ldx, x, N); AL.c(AL.cx, x, &AL.M, N); AL.cd(AL.cdx, x, &AL.M, N);
    al_grad(Ldx, *N);
}
static int al_L_Ld(double *Lx, double *Ldx, const double *x, const int *N) {   /* f90:2207-2217 */
    AL.f(Lx, x, N); AL.c(AL.cx, x, &AL.M, N);
    al_terms(Lx, *N);
    AL.fd(Ldx, x, N); AL.cd(AL.cdx, x, &AL.M, N);
    al_grad(Ldx, *N);
    return 0;
}
static int al_L_Ld_fdwithf(double *Lx, double *Ldx, const double *x, const int *N) { /* f90:2218-2228 */
    (void)AL.f_fd(Lx, Ldx, x, N); AL.c(AL.cx, x, &AL.M, N);
    al_terms(Lx, *N);
    AL.cd(AL.cdx, x, &AL.M, N);
    al_grad(Ldx, *N);
    return 0;
}
void orc_get_al_stats(orc_al_stats_t *out) { *out = g_al; }

static int str_is(const char *s, int len, const char *lit) { /* Fortran == on blank-padded strings */
    int n = (int)strlen(lit), i;
    if (!s) return 0;
    for (i = 0; i < n; i++) if (i >= len || s[i] != lit[i]) return 0;
    for (i = n; i < len; i++) if (s[i] != ' ') return 0;
    return 1;
}

void orc_augmentedlagrangian(orc_f_t f, orc_fd_t fd, orc_c_t c, orc_cd_t cd, double *x, const int *N_, const int *M_,
                             const char *UnconstrainedSolver, const double *lambda0, const double *miu0,
                             const void *fdd, const void *cdd, const int *ExactStep, const int *Memory,
                             const char *Method, orc_ffd_t f_fd, const int *Strong, const int *Warning,
                             const int *MaxIteration, const double *Precision, const double *MinStepLength,
                             const double *WolfeConst1, const double *WolfeConst2, const double *Increment,
                             int len_solver, int len_Method) {
    const int N = *N_, M = *M_;
    int sw, warn, maxit, mem, iIteration, j, is_cg;
    double tol, minstep, c1, c2, incrmt, tolsq, cc = 0.0;
    char type[32];
    (void)fdd; (void)cdd; (void)ExactStep;
    memset(&g_al, 0, sizeof g_al);
    AL.f = f; AL.fd = fd; AL.f_fd = f_fd; AL.c = c; AL.cd = cd; AL.M = M;
    AL.lambda = valloc(M); AL.cx = valloc(M); AL.cdx = valloc((long)N * M);
    for (j = 0; j < M; j++) AL.lambda[j] = lambda0 ? lambda0[j] : 0.0;                 /* f90:2037-2038 */
    if (miu0) AL.miu = fmax(1.0, *miu0); else AL.miu = 1.0;                           /* f90:2039-2040 */
    if (Strong) sw = (*Strong != 0); else sw = 1;                                     /* f90:2042-2075 */
    if (Warning) warn = (*Warning != 0); else warn = 1;
    if (MaxIteration) maxit = *MaxIteration; else maxit = 1000;
    if (Precision) tol = *Precision; else tol = 1e-15;
    if (MinStepLength) minstep = *MinStepLength; else minstep = 1e-15;
    if (WolfeConst1) c1 = fmax(1e-15, *WolfeConst1); else c1 = 1e-4;
    is_cg = str_is(UnconstrainedSolver, len_solver, "ConjugateGradient");
    if (WolfeConst2) c2 = fmin(1.0 - 1e-15, fmax(c1 + 1e-15, *WolfeConst2));
    else c2 = is_cg ? 0.45 : 0.9;
    if (Increment) incrmt = *Increment; else incrmt = 1.05;
    if (Memory) mem = *Memory > 1 ? *Memory : 1; else mem = 10;
    memset(type, ' ', sizeof type);
    if (Method) memcpy(type, Method, (size_t)(len_Method < 32 ? len_Method : 32)); else { type[0] = 'D'; type[1] = 'Y'; }
    tolsq = tol * tol;
    if (!is_cg && !str_is(UnconstrainedSolver, len_solver, "LBFGS")) {
        printf(" oracle: AugmentedLagrangian is restated for UnconstrainedSolver = LBFGS / ConjugateGradient only\n");
        exit(2);
    }
    for (iIteration = 1; iIteration <= maxit; iIteration++) {                         /* f90:2150-2185 */
        orc_ffd_t inner_ffd = f_fd ? al_L_Ld_fdwithf : al_L_Ld;
        if (is_cg)
            orc_conjugategradient(al_L, al_Ld, x, &N, type, inner_ffd, &sw, &warn, &maxit, &tol, &minstep, &c1, &c2,
                                  &incrmt, 32);
        else
            orc_lbfgs(al_L, al_Ld, x, &N, &mem, inner_ffd, &sw, &warn, &maxit, &tol, &minstep, &c1, &c2, &incrmt);
        g_al.inner_iterations += g_st.n_iter;
        g_al.trials += g_st.n_trials;
        g_al.outer_iterations = iIteration;
        c(AL.cx, x, &M, &N);
        cc = 0.0;
        for (j = 0; j < M; j++) cc += AL.cx[j] * AL.cx[j];
        g_al.cnorm2 = cc;
        if (cc < tolsq) break;
        for (j = 0; j < M; j++) AL.lambda[j] = AL.lambda[j] - AL.miu * AL.cx[j];
        AL.miu = AL.miu * incrmt;
    }
    g_al.miu = AL.miu;
    g_al.status = iIteration > maxit ? 2 : 0;
    if (iIteration > maxit && warn) {                                                 /* f90:2187-2190 */
        printf(" Failed augmented Lagrangian: max iteration exceeded!\n");
        printf(" Euclidean norm of constraint violation = %.17g\n", sqrt(cc));
    }
    free(AL.lambda); free(AL.cx); free(AL.cdx);
}
