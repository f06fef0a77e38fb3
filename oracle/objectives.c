/*
 * oracle/objectives.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * The synthetic objectives of SURVEY.md 8(d) as reference-ABI callbacks
 * (f90:33-38), plus the quartic shifted to a non-zero, well-conditioned minimiser
 * (f = sum (x-1)^4 + (x-1)^2, x* = 1: "minimisers to relative 1e-8" needs a scale and
 * linear convergence, which sum x^4 with x* = 0 does not offer).
 * Only the quartic exists in the reference (test/test.f90:630-663:
 * f = sum x**4, f' = 4 x**3, evaluated as gfortran expands integer powers:
 * x**4 = (x*x)*(x*x), x**3 = (x*x)*x).  Extended Rosenbrock and the diagonal
 * quadratic are defined here; the CUDA objective kernels use the same operation
 * order without FMA so that gradients agree bit for bit given the same x.
 *
 * Build with -ffp-contract=off.
 */
#include "oracle.h"

#include <math.h>
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static int g_kind = ORC_OBJ_QUARTIC;
static long long g_offset = 0, g_nglobal = 0;
static int g_fsum_mode = 0;

/* Summation order of the objective value, to bound its share of the summation-order noise:
 * 0 = sequential double (what the reference's own test callback does, test.f90:630-640),
 * 1 = long double, 2 = pairwise (binary-counter cascade).  Gradients do not depend on it. */
void orc_obj_set_sum_mode(int mode) { g_fsum_mode = mode; }

typedef struct {
    int mode, top;
    unsigned long count;
    double s;
    long double sl;
    double stack[64];
} facc_t;
static void facc_init(facc_t *A) { A->mode = g_fsum_mode; A->top = 0; A->count = 0; A->s = 0.0; A->sl = 0.0L; }
static void facc_add(facc_t *A, double v) {
    if (A->mode == 0) { A->s = A->s + v; return; }
    if (A->mode == 1) { A->sl += (long double)v; return; }
    unsigned long c = ++A->count;
    while ((c & 1ul) == 0) { v = A->stack[--A->top] + v; c >>= 1; }
    A->stack[A->top++] = v;
}
static double facc_value(facc_t *A) {
    if (A->mode == 0) return A->s;
    if (A->mode == 1) return (double)A->sl;
    double s = 0.0;
    while (A->top > 0) s = A->stack[--A->top] + s;
    return s;
}
static double T0[256], T1[256], T2[256];
static int g_tables = 0;

static void build_tables(void) {
    for (int k = 0; k < 256; k++) {
        T0[k] = pow(10.0, 6.0 * (double)k / 16777216.0);
        T1[k] = pow(10.0, 6.0 * (double)k / 65536.0);
        T2[k] = pow(10.0, 6.0 * (double)k / 256.0);
    }
    g_tables = 1;
}

/* d_i = 10^(6 q / 2^24), q = trunc(i * (2^24/(n-1))) evaluated in IEEE double (one divide, one
 * multiply, one truncation: identical on host and device): log-uniform in [1, 1e6], built from
 * three table factors with exact IEEE multiplies so host and device agree bitwise. */
double orc_diag_coeff(long long i, long long n_global) {
    if (!g_tables) build_tables();
    if (n_global <= 1) return 1.0;
    const double scale = 16777216.0 / (double)(n_global - 1);
    uint64_t q = (uint64_t)((double)i * scale);
    if (q >> 24) return 1.0e6;
    return T2[(q >> 16) & 255] * T1[(q >> 8) & 255] * T0[q & 255];
}

void orc_obj_select(int kind, long long offset, long long n_global) {
    g_kind = kind; g_offset = offset; g_nglobal = n_global;
}

static double splitmix_u(uint64_t z) {
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

void orc_obj_start(int start_kind, unsigned long long seed, double *x, long long offset, long long n,
                   long long n_global) {
    (void)n_global;
    for (long long k = 0; k < n; k++) {
        const long long i = offset + k;
        const double u = splitmix_u((uint64_t)i + seed);
        switch (start_kind) {
        case ORC_START_QUARTIC_U: x[k] = u; break;
        case ORC_START_ROSEN_STD: x[k] = (i & 1) ? 1.0 : -1.2; break;
        case ORC_START_ROSEN_PERT: x[k] = ((i & 1) ? 1.0 : -1.2) + 0.1 * (u - 0.5); break;
        default: x[k] = 0.0; break;
        }
    }
}

/* rows [lo, hi) of the current objective; lo is even (Rosenbrock pairs), hi == n closes an odd-length problem.
 * The sequential build evaluates [0, n) in one call: the loops below are the whole of the reference-order arithmetic. */
static double eval_range(int want_f, double *g, const double *x, long lo, long hi, long n) {
    facc_t f;
    facc_init(&f);
    if (g_kind == ORC_OBJ_QUARTIC) {
        for (long i = lo; i < hi; i++) {
            const double x2 = x[i] * x[i];
            if (want_f) facc_add(&f, x2 * x2);
            if (g) g[i] = 4.0 * (x2 * x[i]);
        }
    } else if (g_kind == ORC_OBJ_ROSENBROCK) {
        /* pairs (2j, 2j+1) in GLOBAL indexing; shards start on even offsets */
        long i = lo;
        for (; i + 1 < hi; i += 2) {
            const double a = x[i], b = x[i + 1];
            const double t1 = b - a * a, t2 = 1.0 - a;
            if (want_f) facc_add(&f, (100.0 * t1) * t1 + t2 * t2);
            if (g) { g[i] = (-400.0 * a) * t1 - 2.0 * t2; g[i + 1] = 200.0 * t1; }
        }
        if (i < hi && hi == n) { /* unpaired last element of an odd-length problem */
            const double t2 = 1.0 - x[i];
            if (want_f) facc_add(&f, t2 * t2);
            if (g) g[i] = -2.0 * t2;
        }
    } else if (g_kind == ORC_OBJ_QUARTIC_SHIFTED) {
        for (long i = lo; i < hi; i++) {
            const double t = x[i] - 1.0;
            const double t2 = t * t;
            if (want_f) facc_add(&f, t2 * t2 + t2);
            if (g) g[i] = 4.0 * (t2 * t) + 2.0 * t;
        }
    } else {
        for (long i = lo; i < hi; i++) {
            const double d = orc_diag_coeff(g_offset + i, g_nglobal);
            const double t = x[i] - 1.0;
            if (want_f) facc_add(&f, ((0.5 * d) * t) * t);
            if (g) g[i] = d * t;
        }
    }
    return want_f ? facc_value(&f) : 0.0;
}

static int eval(double *fx, double *g, const double *x, int n) {
#ifdef _OPENMP
    /* liboracle_omp.so only (the generous CPU baseline of bench.py; never a checker): one even-aligned block per thread,
     * partial sums added in thread order */
    double part[256];
    int nt_used = 1;
    if (!g_tables) build_tables();
#pragma omp parallel
    {
        int nt = omp_get_num_threads(), t = omp_get_thread_num();
        if (nt > 256) nt = 256;
        if (t < nt) {
            const long chunk = ((long)n / nt) & ~1L;
            const long lo = (long)t * chunk, hi = (t == nt - 1) ? (long)n : lo + chunk;
            part[t] = eval_range(fx != 0, g, x, lo, hi, n);
        }
#pragma omp single
        nt_used = nt;
    }
    if (fx) { double s = 0.0; for (int t = 0; t < nt_used; t++) s += part[t]; *fx = s; }
#else
    const double f = eval_range(fx != 0, g, x, 0, n, n);
    if (fx) *fx = f;
#endif
    return 0;
}

/* f-term of every element (Rosenbrock: the pair's term at the even index, +0.0 at the odd one; the unpaired last
 * element of an odd-length problem carries its own term), in the arithmetic of eval_range().  For the host simulator's
 * bit-level model of the CUDA reductions (tests/hostsim): sum of terms[] in any order = f. */
void orc_obj_terms(double *terms, const double *x, const int *dim) {
    const long n = *dim;
    if (g_kind == ORC_OBJ_QUARTIC) {
        for (long i = 0; i < n; i++) { const double x2 = x[i] * x[i]; terms[i] = x2 * x2; }
    } else if (g_kind == ORC_OBJ_ROSENBROCK) {
        long i = 0;
        for (; i + 1 < n; i += 2) {
            const double a = x[i], b = x[i + 1];
            const double t1 = b - a * a, t2 = 1.0 - a;
            terms[i] = (100.0 * t1) * t1 + t2 * t2;
            terms[i + 1] = 0.0;
        }
        if (i < n) { const double t2 = 1.0 - x[i]; terms[i] = t2 * t2; }
    } else if (g_kind == ORC_OBJ_QUARTIC_SHIFTED) {
        for (long i = 0; i < n; i++) { const double t = x[i] - 1.0; const double t2 = t * t; terms[i] = t2 * t2 + t2; }
    } else {
        for (long i = 0; i < n; i++) {
            const double d = orc_diag_coeff(g_offset + i, g_nglobal);
            const double t = x[i] - 1.0;
            terms[i] = ((0.5 * d) * t) * t;
        }
    }
}

void orc_obj_f(double *fx, const double *x, const int *dim) { eval(fx, 0, x, *dim); }
void orc_obj_fd(double *fdx, const double *x, const int *dim) { eval(0, fdx, x, *dim); }
int orc_obj_f_fd(double *fx, double *fdx, const double *x, const int *dim) {
    return eval(fx, fdx, x, *dim);
}

/* test.f90:692-705: the unit-sphere equality constraint used by the reference's AugmentedLagrangian smoke test */
void orc_con_sphere_c(double *cx, const double *x, const int *M, const int *N) {
    double s = 0.0;
    (void)M;
    for (int i = 0; i < *N; i++) s = s + x[i] * x[i];
    cx[0] = s - 1.0;
}
void orc_con_sphere_cd(double *cdx, const double *x, const int *M, const int *N) {
    (void)M;
    for (int i = 0; i < *N; i++) cdx[i] = 2.0 * x[i];
}
