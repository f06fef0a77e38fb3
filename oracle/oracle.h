/*
 * oracle/oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, strict IEEE: -O2 -ffp-contract=off, sequential sums)
 * of the reference's large-dimension unconstrained-optimizer path in
 * /root/reference/source/NonlinearOptimization.f90:
 *     SteepestDescent          f90:55-188
 *     ConjugateGradient        f90:193-394
 *     LBFGS                    f90:398-625
 *     Wolfe / Wolfe_fdwithf    f90:1286-1459
 *     StrongWolfe / _fdwithf   f90:1462-1698
 *     ConjugateGradient_basic  f90:2249-2346
 *     AugmentedLagrangian      f90:2005-2241 (LBFGS / ConjugateGradient branches)
 *
 * PARITY UNPINNED: the reference cannot be compiled here (no Fortran compiler,
 * no MKL: SURVEY.md F1/F2) and its tests hold no golden vectors for this path
 * (F8).  This restatement is pinned only by (i) an independent NumPy
 * transcription (oracle/oracle_np.py) that must agree with it, (ii) closed-form
 * minimisers and scipy, (iii) structural checks.  See DESIGN.md.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  The product (libflgpu.so) never does.
 *
 * Calling convention mirrors gfortran's for the reference routines: every
 * argument by reference, an absent OPTIONAL is a NULL pointer, LOGICAL is a
 * 4-byte int (non-zero = true), CHARACTER(*) is char* + trailing hidden length.
 */
#ifndef FLGPU_ORACLE_H
#define FLGPU_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* user callbacks, f90:33-38; C shapes as cpp/NonlinearOptimization.hpp:281-284 */
typedef void (*orc_f_t)(double *fx, const double *x, const int *dim);
typedef void (*orc_fd_t)(double *fdx, const double *x, const int *dim);
typedef int (*orc_ffd_t)(double *fx, double *fdx, const double *x, const int *dim);

/* Called after every line search of an outer iteration (iteration 0 = the
 * steepest-descent step of LBFGS / first CG step).  p = direction searched along,
 * x, g = point and gradient accepted, a = accepted step, phid0 = phi'(0). */
typedef void (*orc_trace_t)(void *user, int iter, int dim, const double *p, const double *x,
                            const double *g, double a, double fx, double phid0, long trials);

typedef struct {
    long n_f, n_fd, n_ffd;  /* callback invocations */
    long n_trials;          /* x = x0 + a*p formations */
    long n_linesearch;      /* line searches run */
    long n_iter;            /* outer iterations completed (accepted steps) */
    long n_quirk_f9;        /* times StrongWolfe fell through f90:1511-1512 */
    int status;             /* 0 converged(g), 1 step converged, 2 max iteration, 3 initial g small */
} orc_stats_t;

/* 1 for liboracle.so (sequential, as the reference); the host's thread count for liboracle_omp.so, the separately
 * built generous CPU baseline of bench.py, which is never used as a checker */
int orc_threads(void);
void orc_set_trace(orc_trace_t cb, void *user);
/* 0 = sequential double (reference semantics), 1 = long double accumulate,
 * 2 = pairwise double.  Modes 1/2 exist only to bound summation noise. */
void orc_set_sum_mode(int mode);
void orc_get_stats(orc_stats_t *out);
/* Test-harness guard: give up (status 9) after n callback invocations inside line searches; 0 = unlimited.  The
 * reference itself never terminates once a step is NaN (f90:1518-1546). */
void orc_set_eval_budget(long n);
/* 0 (default) = the reference's searchers.  1 = the product's optional FLGPU_LS_FAST searcher (NOT in the reference;
 * restated here only so its host and device implementations can be compared bit for bit -- see oracle.c). */
void orc_set_line_search(int policy);

/* f90:398-625 */
void orc_lbfgs(orc_f_t f, orc_fd_t fd, double *x, const int *dim, const int *Memory, orc_ffd_t f_fd,
               const int *Strong, const int *Warning, const int *MaxIteration, const double *Precision,
               const double *MinStepLength, const double *WolfeConst1, const double *WolfeConst2,
               const double *Increment);
/* f90:193-394 */
void orc_conjugategradient(orc_f_t f, orc_fd_t fd, double *x, const int *dim, const char *Method,
                           orc_ffd_t f_fd, const int *Strong, const int *Warning,
                           const int *MaxIteration, const double *Precision,
                           const double *MinStepLength, const double *WolfeConst1,
                           const double *WolfeConst2, const double *Increment, int len_Method);
/* f90:2249-2346 (all arguments mandatory, no clamps) */
void orc_conjugategradient_basic(orc_f_t f, orc_fd_t fd, double *x, const int *dim, const char *Method,
                                 const int *Strong, const int *Warning, const int *MaxIteration,
                                 const double *Precision, const double *MinStepLength,
                                 const double *WolfeConst1, const double *WolfeConst2,
                                 const double *Increment, int len_Method);

/* f90:55-188 */
void orc_steepestdescent(orc_f_t f, orc_fd_t fd, double *x, const int *dim, orc_ffd_t f_fd,
                         const int *Strong, const int *Warning, const int *MaxIteration,
                         const double *Precision, const double *MinStepLength, const double *WolfeConst1,
                         const double *WolfeConst2, const double *Increment);

/* constraint callbacks of AugmentedLagrangian (f90:2012, test.f90:692-705; hpp:372-373): c(cx,x,M,N), cd(cdx,x,M,N)
 * with cdx(N,M) column-major */
typedef void (*orc_c_t)(double *cx, const double *x, const int *M, const int *N);
typedef void (*orc_cd_t)(double *cdx, const double *x, const int *M, const int *N);
typedef struct {
    long outer_iterations, inner_iterations, trials;
    int status;        /* 0 constraint converged, 2 max iteration */
    double cnorm2, miu;
} orc_al_stats_t;
/* f90:2005-2241, branches UnconstrainedSolver = 'LBFGS' / 'ConjugateGradient' only (the others are dense
 * Newton / BFGS solvers outside the hot path).  Argument order and hidden lengths as hpp:369-392. */
void orc_augmentedlagrangian(orc_f_t f, orc_fd_t fd, orc_c_t c, orc_cd_t cd, double *x, const int *N, const int *M,
                             const char *UnconstrainedSolver, const double *lambda0, const double *miu0,
                             const void *fdd, const void *cdd, const int *ExactStep, const int *Memory,
                             const char *Method, orc_ffd_t f_fd, const int *Strong, const int *Warning,
                             const int *MaxIteration, const double *Precision, const double *MinStepLength,
                             const double *WolfeConst1, const double *WolfeConst2, const double *Increment,
                             int len_solver, int len_Method);
void orc_get_al_stats(orc_al_stats_t *out);
/* the reference's test constraint (test.f90:692-705): unit sphere, cx(1) = dot_product(x,x) - 1, cdx(:,1) = 2 x */
void orc_con_sphere_c(double *cx, const double *x, const int *M, const int *N);
void orc_con_sphere_cd(double *cdx, const double *x, const int *M, const int *N);

/* line searchers, f90:1286,1373,1462,1582 (exported like the reference's module procedures) */
void orc_wolfe(const double *c1, const double *c2, orc_f_t f, orc_fd_t fd, double *x, double *a,
               const double *p, double *fx, const double *phid0, double *fdx, const int *dim,
               const double *Increment);
void orc_wolfe_fdwithf(const double *c1, const double *c2, orc_f_t f, orc_fd_t fd, orc_ffd_t f_fd,
                       double *x, double *a, const double *p, double *fx, const double *phid0,
                       double *fdx, const int *dim, const double *Increment);
void orc_strongwolfe(const double *c1, const double *c2, orc_f_t f, orc_fd_t fd, double *x, double *a,
                     const double *p, double *fx, const double *phid0, double *fdx, const int *dim,
                     const double *Increment);
void orc_strongwolfe_fdwithf(const double *c1, const double *c2, orc_f_t f, orc_fd_t fd,
                             orc_ffd_t f_fd, double *x, double *a, const double *p, double *fx,
                             const double *phid0, double *fdx, const int *dim,
                             const double *Increment);

/* ---- synthetic objectives (oracle/objectives.c), SURVEY.md 8(d) ---- */
enum { ORC_OBJ_QUARTIC = 0, ORC_OBJ_ROSENBROCK = 1, ORC_OBJ_DIAGQUAD = 2, ORC_OBJ_QUARTIC_SHIFTED = 3 };
enum { ORC_START_QUARTIC_U = 0, ORC_START_ROSEN_STD = 1, ORC_START_ROSEN_PERT = 2, ORC_START_ZERO = 3 };
/* Select the objective the three callbacks below evaluate.  offset/n_global let a
 * row shard be evaluated (index-dependent objectives). */
void orc_obj_select(int kind, long long offset, long long n_global);
/* summation order of the objective VALUE (0 sequential = reference test callback, 1 long double,
 * 2 pairwise); orc_set_sum_mode() sets it too, so the noise envelope covers every reduction. */
void orc_obj_set_sum_mode(int mode);
void orc_obj_terms(double *terms, const double *x, const int *dim);
void orc_obj_f(double *fx, const double *x, const int *dim);
void orc_obj_fd(double *fdx, const double *x, const int *dim);
int orc_obj_f_fd(double *fx, double *fdx, const double *x, const int *dim);
void orc_obj_start(int start_kind, unsigned long long seed, double *x, long long offset, long long n,
                   long long n_global);
double orc_diag_coeff(long long i, long long n_global);

#ifdef __cplusplus
}
#endif
#endif
