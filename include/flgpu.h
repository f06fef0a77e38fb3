/*
 * flgpu.h -- C-ABI of libflgpu.so: the B200 (sm_100a, fp64 CUDA) implementation of
 * Fortran-Library's large-dimension unconstrained-optimizer hot path.
 *
 * Reference interfaces replaced ("f90:" = source/NonlinearOptimization.f90,
 * "hpp:" = cpp/NonlinearOptimization.hpp of YifanShenSZ/Fortran-Library):
 *   LBFGS                   f90:398-400   (no hpp declaration exists; added by analogy)
 *   SteepestDescent         f90:55-56     hpp:279-291 (gnu) / hpp:11-23 (intel)   [SURVEY 8f row N1]
 *   AugmentedLagrangian     f90:2005-2008 hpp:369-392 (gnu), LBFGS / ConjugateGradient branches [SURVEY 8f row N2]
 *   ConjugateGradient       f90:193-195   hpp:310-324 (gnu) / hpp:42-56 (intel)
 *   ConjugateGradient_basic f90:2249-2251 hpp:292-306 (gnu) / hpp:24-38 (intel)
 *   Strong-Wolfe / Wolfe line searchers f90:1286,1373,1462,1582 (internal to the above)
 *   callbacks f, fd, f_fd   f90:33-38     hpp:281-284
 *
 * Two layers are exported:
 *   1. the reference's own compiled symbol names (section "Fortran ABI"), same argument
 *      order, by-reference passing, NULL = absent OPTIONAL, 4-byte LOGICAL, trailing
 *      hidden CHARACTER length -- a binary drop-in for libFL.so on this path;
 *   2. flgpu_* entry points with 64-bit dimensions, device-resident x, an explicit
 *      stream and a row-shard communicator for multi-GPU runs.
 *
 * No torch types, no C++ types.  There is no CPU fallback: every entry point aborts
 * with a message if no CUDA device is usable.
 */
#ifndef FLGPU_H
#define FLGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ callbacks */

/* Reference callback ABI (f90:33-38, hpp:281-284).  Under the Fortran-ABI entry points x and fdx are HOST pointers
 * by default, DEVICE pointers for callbacks known or declared to be device code (flgpu_set_callback_space); fx is a
 * host scalar that must be valid when the callback returns.  Device callbacks enqueue on flgpu_current_stream(). */
typedef void (*flgpu_ref_f_fn)(double *fx, const double *x, const int *dim);
typedef void (*flgpu_ref_fd_fn)(double *fdx, const double *x, const int *dim);
typedef int (*flgpu_ref_f_fd_fn)(double *fx, double *fdx, const double *x, const int *dim);

typedef struct flgpu_comm flgpu_comm; /* row-shard communicator (peer-memory mailboxes + NCCL plumbing) */

/* Context handed to the 64-bit device callbacks. */
typedef struct flgpu_eval_ctx {
    void *user;       /* flgpu_problem.user */
    void *stream;     /* cudaStream_t the callback must enqueue on */
    int64_t offset;   /* global index of local row 0 (row-sharded runs) */
    int64_t n_global; /* global dimension */
    int rank, nranks; /* position in the communicator (0,1 when single GPU) */
    int device;       /* CUDA device ordinal */
} flgpu_eval_ctx;

/* Device callbacks: enqueue kernels on ctx->stream, do NOT synchronise.
 * f_dev is a device scalar receiving this rank's partial sum of f (plain store). */
typedef void (*flgpu_f_fn)(const flgpu_eval_ctx *ctx, double *f_dev, const double *x_dev, int64_t n_local);
typedef void (*flgpu_fd_fn)(const flgpu_eval_ctx *ctx, double *g_dev, const double *x_dev, int64_t n_local);
typedef void (*flgpu_f_fd_fn)(const flgpu_eval_ctx *ctx, double *f_dev, double *g_dev, const double *x_dev,
                              int64_t n_local);

/* Optional FUSED line-search evaluation (an extension; the reference has no counterpart).
 * Every line-search trial of the reference is `x=x0+a*p; call f / fd / f_fd; dot_product(fdx,p)`
 * (f90:1482-1485, 1490, 1501, 1567): 7n doubles of traffic through opaque callbacks.  An objective
 * whose kernel can form the trial point itself (element-local or block-local objectives) may supply
 * this callback instead; the library then never materialises rejected trial points:
 *     x  = x0 + a*p          element-wise, multiply THEN add (two roundings, as the reference)
 *     FLGPU_WANT_F   : *f_dev  = this rank's partial sum of f(x)
 *     FLGPU_WANT_GP  : *gp_dev = this rank's partial sum of f'(x).p
 *     FLGPU_WRITE_X  : x_out[i] = x[i]          (accepted point only)
 *     FLGPU_WRITE_G  : g_out[i] = f'(x)[i]      (accepted point only)
 * Enqueue on ctx->stream, do not synchronise.  Which of f / f' are requested follows the
 * reference's own call sequence (f-only probes in branch D f90:1518-1520, f_fd elsewhere), so
 * evaluation counts and every host decision are the same as on the unfused path. */
enum { FLGPU_WANT_F = 1, FLGPU_WANT_GP = 2, FLGPU_WRITE_X = 4, FLGPU_WRITE_G = 8 };
typedef void (*flgpu_fused_fn)(const flgpu_eval_ctx *ctx, int flags, double *f_dev, double *gp_dev,
                               double *x_out, double *g_out, const double *x0_dev, const double *p_dev,
                               double a, int64_t n_local);

/* Optional DEVICE-RESIDENT line search (an extension beyond flgpu_fused_fn).  One cooperative kernel runs the whole
 * Wolfe / Strong-Wolfe search of the reference (f90:1286-1698) on the device: every thread executes the same state
 * machine (include/flgpu_search_core.hpp, the source the host driver uses too), every evaluation is a grid-wide reduction with
 * the library's fixed summation order, and the accepted point and gradient are stored by the same kernel -- no host
 * round trip per trial.  result_dev receives FLGPU_SEARCH_RESULT_DOUBLES doubles:
 *   [0] accepted step a  [1] f at it  [2] trial points formed  [3] f calls  [4] fd calls  [5] f_fd calls
 *   [6] f-only trials (branch D)  [7] rank exchanges made inside the kernel, or -1 if the search gave up after its
 *   evaluation budget (100000: a kernel must terminate even if the objective returns NaN, where the reference spins)
 * The decisions and the values are those of the host-driven fused search, bit for bit.  On row-sharded runs block 0
 * trades the partial sums with the other ranks inside the kernel (peer-memory mailboxes; needs the peer-memory
 * exchange, not the ncclAllGather fallback). */
#define FLGPU_SEARCH_RESULT_DOUBLES 8
#define FLGPU_SEARCH_ROW_SHARDS 1
typedef struct flgpu_search_args {
    const double *x0_dev, *p_dev;   /* start point and direction */
    double *x_out, *g_out;          /* accepted point and its gradient */
    double c1, c2abs;               /* WolfeConst1, WolfeConst2 * |phi'(0)| */
    double fx0, phid0, incr, a;     /* f(x0), phi'(0), Increment, first step */
    int strong, fdwithf;            /* which of the four searchers */
    double *result_dev;
    flgpu_comm *comm;               /* row-sharded runs: the communicator whose search mailboxes carry the exchanges */
    int policy;                     /* FLGPU_LS_REFERENCE: the reference's searchers; FLGPU_LS_FAST: SearchCore::fast(strong) */
    int no_store;                   /* 1: leave x_out / g_out alone (the caller's fused K1 forms the accepted point itself) */
} flgpu_search_args;
typedef void (*flgpu_search_fn)(const flgpu_eval_ctx *ctx, const flgpu_search_args *args, int64_t n_local);

/* Optional FUSED ACCEPTED-POINT UPDATE (an extension; L-BFGS with a fused line search only).  After a line search the
 * reference has x and f'(x) in memory and forms s = x - xold, y = f' - f'old and the dot products of the next
 * two-loop recursion (f90:609-624, 590-606).  With this callback the accepted point is never stored by the search:
 * the library's K1 kernel, instantiated by the objective with a source that forms x1 = x0 + a*p and f'(x1) in
 * registers (include/flgpu_k1.cuh), writes x1, f'(x1), s and y and accumulates every dot in ONE pass -- 7n instead of
 * 10n doubles per iteration, the same bits.  The callback launches exactly that pass on ctx->stream:
 *     flgpu::k::launch_k1_pass(*(const flgpu::k::K1Launch *)args->k1, MySource{...});
 * (libflgpu's built-in objectives and include/flgpu_objective.cuh provide it). */
typedef struct flgpu_update_args {
    const void *k1;     /* flgpu::k::K1Launch prepared by the library: vectors, step, ring buffers, shape, geometry */
    size_t k1_bytes;    /* sizeof(flgpu::k::K1Launch) the library was built with (callbacks check it) */
} flgpu_update_args;
typedef void (*flgpu_update_fn)(const flgpu_eval_ctx *ctx, const flgpu_update_args *args, int64_t n_local);

/* Optional FUSED FIRST TRIALS (an extension; L-BFGS with a fused line search only).  Every line search of the L-BFGS main
 * loop starts at a = 1 (f90:607), i.e. at x + p with the direction p that Before() has just formed (f90:589-607), and while
 * the slope stays negative it walks on to Increment, Increment^2, ... (f90:1499-1501).  With this callback the library's K3
 * kernel, instantiated by the objective with a probe (include/flgpu_k3.cuh), evaluates the first FOUR steps of that walk
 * while it writes p: it reads x as well, forms x + a*p in registers and reduces f and f'.p at each step with the chunk order
 * of the fused evaluation -- the same bits as separate flgpu_fused_fn calls, for n doubles of extra traffic in all and no
 * launch.  The search consumes the values in the reference's order with the reference's counts; what it does not reach is
 * dropped.  The callback launches exactly that pass on ctx->stream:
 *     flgpu::k::launch_k3_probe(*(const flgpu::k::K3Launch *)args->k3, MyProbe{...});
 * (libflgpu's built-in objectives and include/flgpu_objective.cuh provide it). */
typedef struct flgpu_direction_args {
    const void *k3;     /* flgpu::k::K3Launch prepared by the library: vectors, coefficients, ring buffers, geometry */
    size_t k3_bytes;    /* sizeof(flgpu::k::K3Launch) the library was built with (callbacks check it) */
    int flags;          /* what the reference evaluates at the first trial (FLGPU_WANT_F [| FLGPU_WANT_GP]); informational:
                           the probe always reduces both sums at every step */
} flgpu_direction_args;
typedef void (*flgpu_direction_fn)(const flgpu_eval_ctx *ctx, const flgpu_direction_args *args, int64_t n_local);

/* Optional BATCHED fused evaluation (an extension; needs `fused`).  While the reference's searchers bracket, the next
 * trial steps are known in advance: a, a*Increment, a*Increment^2, ... (f90:1499-1501, 1308-1310) or a/Increment, ...
 * (f90:1488-1490, 1518, 1325).  One pass over x0 and p can evaluate several of them -- the traffic of ONE trial, a few more
 * flops per element -- and the host then takes its decisions from values it already holds, in the reference's order, with
 * the reference's counts; evaluations the search never reaches are dropped.  For j < count <= FLGPU_MULTI_MAX:
 *     out_dev[2j]   = this rank's partial sum of f (x0 + steps[j]*p)
 *     out_dev[2j+1] = this rank's partial sum of f'(x0 + steps[j]*p) . p
 * each with exactly the bits a flgpu_fused_fn call with a = steps[j] delivers (same per-element roundings, same
 * summation order).  `steps` is a HOST array (pass it to the kernel by value).  Enqueue on ctx->stream, do not synchronise. */
#define FLGPU_MULTI_MAX 4
typedef void (*flgpu_fused_multi_fn)(const flgpu_eval_ctx *ctx, int count, const double *steps, double *out_dev,
                                     const double *x0_dev, const double *p_dev, int64_t n_local);

typedef struct flgpu_problem {
    flgpu_f_fn f;       /* required */
    flgpu_fd_fn fd;     /* required */
    flgpu_f_fd_fn f_fd; /* optional (NULL = absent, f90:42-43) */
    void *user;
    flgpu_fused_fn fused; /* optional (NULL = trial points are materialised and f / fd / f_fd are called) */
    flgpu_search_fn search; /* optional (NULL = the host drives the search, one round trip per evaluation) */
    int search_caps;        /* FLGPU_SEARCH_ROW_SHARDS if `search` handles flgpu_search_args.comm != NULL; else 0 */
    flgpu_update_fn update; /* optional, needs `fused` (NULL = the search stores the accepted point, K1 reads it back) */
    flgpu_direction_fn direction; /* optional, needs `fused` (NULL = the first trial of a search is a flgpu_fused_fn call) */
    flgpu_fused_multi_fn fused_multi; /* optional, needs `fused` (NULL = one flgpu_fused_fn call per trial) */
} flgpu_problem;

/* ------------------------------------------------------------------ options / results */

enum { FLGPU_CG_DY = 0, FLGPU_CG_PR = 1 };
/* flgpu_options.line_search.  REFERENCE (default): Wolfe / StrongWolfe(_fdwithf) of f90:1286-1698, statement by
 * statement -- same trial points, same evaluation counts as the reference.  FAST (not a reference
 * routine, iterates differ): the bracketing / cubic-zoom scheme of Nocedal & Wright Alg. 3.5/3.6 with
 * More'-Thuente-style safeguards (include/flgpu_search_core.hpp, SearchCore::fast): f and f' are evaluated together
 * at every trial and the first trial that satisfies the (strong, or with Strong = 0 the weak) Wolfe conditions for
 * the given WolfeConst1/2 is accepted, where the reference keeps growing the step by Increment while the slope is
 * negative (f90:1498-1515).  Increment is not used by FAST.  f_fd (or the fused evaluation) is used from the first
 * line search on, not only in the main loop. */
enum { FLGPU_LS_REFERENCE = 0, FLGPU_LS_FAST = 1 };
enum { FLGPU_SPACE_HOST = 0, FLGPU_SPACE_DEVICE = 1 };
/* flgpu_stats.status */
enum {
    FLGPU_CONVERGED = 0,        /* |f'|^2 < Precision^2 (f90:612) */
    FLGPU_STEP_CONVERGED = 1,   /* |p|^2 a^2 < MinStepLength^2 (f90:615) */
    FLGPU_MAX_ITERATION = 2,    /* f90:580 */
    FLGPU_INITIAL_CONVERGED = 3,/* f90:443 */
    FLGPU_STOPPED_BY_OBSERVER = 4,
    FLGPU_INVALID_ARGUMENT = 5  /* the call was refused (see the return code); x is unchanged */
};
/* Return codes of the flgpu_* entry points.  CUDA / NCCL failures still abort with a message: like the reference,
 * the optimizers have no channel to report them and no way to continue. */
enum { FLGPU_OK = 0, FLGPU_ERR_MEMORY_LIMIT = 2 };
/* Largest LBFGS Memory the kernels hold (K2: two ring slots per lane of one warp; K3: coefficient tables in shared
 * memory).  The reference allocates s(dim,0:mem) for any mem (f90:419-420, 435) and recommends [3, 30] (f90:397).
 * flgpu_lbfgs refuses a larger Memory with FLGPU_ERR_MEMORY_LIMIT; the Fortran-ABI symbol, which cannot return an
 * error, prints a warning and runs with Memory = FLGPU_MAX_MEMORY. */
#define FLGPU_MAX_MEMORY 64


/* Per outer iteration, called on the host after the line search accepted a step.
 * Device pointers stay valid until the observer returns.  Return non-zero to stop. */
typedef struct flgpu_iter_info {
    int64_t iteration;  /* 0-based count of accepted steps, including LBFGS' pre-iterations */
    int64_t n_local;
    double step;        /* accepted a */
    double f;           /* f at the accepted point */
    double phid0;       /* phi'(0) of this search */
    int64_t trials;     /* x0 + a p formations in this search */
    const double *p_dev, *x_dev, *g_dev;
    void *stream;
    int64_t gpu_launches; /* library kernels enqueued so far in this call */
    int64_t callbacks;    /* objective callbacks invoked so far in this call (f, fd, f_fd, fused, search: one kernel
                             launch each for the built-in objectives) */
    int64_t total_trials; /* trial points formed so far in this call */
} flgpu_iter_info;
typedef int (*flgpu_observer_fn)(void *user, const flgpu_iter_info *info);

/* Tunables: names, defaults and fail-safe clamps of f90:417-434 (LBFGS), f90:212-229
 * (CG) and f90:1478-1479 (Increment).  Fill with flgpu_options_default(). */
typedef struct flgpu_options {
    int memory;             /* LBFGS Memory, default 10, clamped max(1,.); at most FLGPU_MAX_MEMORY */
    int method;             /* CG: FLGPU_CG_DY (default) / FLGPU_CG_PR */
    int strong;             /* default 1 */
    int warning;            /* default 1 */
    int max_iteration;      /* default 1000 */
    double precision;       /* default 1e-15 (compared squared) */
    double min_step_length; /* default 1e-15 (compared squared) */
    double wolfe_c1;        /* default 1e-4 */
    double wolfe_c2;        /* default 0.9 (LBFGS) / 0.45 (CG) */
    double increment;       /* default 1.05 */
    int no_clamp;           /* 1 = ConjugateGradient_basic semantics: c1/c2 used as given (f90:2278) */
    /* execution */
    void *stream;           /* cudaStream_t; NULL = library-owned non-blocking stream */
    flgpu_comm *comm;       /* NULL = single GPU */
    int64_t offset;         /* global index of local row 0 (with comm) */
    int64_t n_global;       /* global dimension (0 = n_local) */
    flgpu_observer_fn observer;
    void *observer_user;
    int time_kernels;       /* 1 = bracket every library kernel with CUDA events (flgpu_kernel_times) */
    int no_fused;           /* 1 = ignore flgpu_problem.fused (always materialise trial points) */
    int device_search;      /* flgpu_problem.search (single GPU, fused mode): 0 never, 1 always, 2 = auto (default):
                               used up to 2^18 rows per GPU, where the per-trial host round trip dominates; same bits either way */
    int line_search;        /* FLGPU_LS_REFERENCE (default) / FLGPU_LS_FAST */
} flgpu_options;

typedef struct flgpu_stats {
    int64_t iterations;   /* accepted steps */
    int status;
    int64_t n_f, n_fd, n_f_fd;      /* callback invocations */
    int64_t n_trials, n_f_only_trials, n_linesearch;
    int64_t gpu_launches;           /* library kernels launched (callback kernels not counted) */
    int64_t host_syncs;
    double f;             /* final objective */
    double gnorm2;        /* final |f'|^2 */
    int64_t n_batched_passes; /* flgpu_fused_multi_fn launches: passes that evaluated up to FLGPU_MULTI_MAX trials at once */
} flgpu_stats;

void flgpu_options_default(flgpu_options *o, int for_cg);

/* x: n_local doubles, host (x_space = FLGPU_SPACE_HOST) or device memory; in: initial
 * guess, out: minimiser (f90:52).  Returns FLGPU_OK, or FLGPU_ERR_MEMORY_LIMIT (nothing done); aborts on CUDA/NCCL
 * failure (the reference has no error channel either). */
int flgpu_lbfgs(const flgpu_problem *prob, const flgpu_options *opt, double *x, int64_t n_local,
                int x_space, flgpu_stats *stats);
int flgpu_conjugate_gradient(const flgpu_problem *prob, const flgpu_options *opt, double *x,
                             int64_t n_local, int x_space, flgpu_stats *stats);
/* SteepestDescent (f90:55-188): the same line searchers and kernels with p = -f'(x) */
int flgpu_steepest_descent(const flgpu_problem *prob, const flgpu_options *opt, double *x,
                           int64_t n_local, int x_space, flgpu_stats *stats);

/* ------------------------------------------------------------------ AugmentedLagrangian over LBFGS / CG */
/* Reference: AugmentedLagrangian (f90:2005-2241), branches UnconstrainedSolver = 'LBFGS' (f90:2150-2167) and
 * 'ConjugateGradient' (f90:2168-2185) -- the only in-library caller of the hot path.  Minimises
 * f(x) subject to c(x) = 0 (M equality constraints) by repeated unconstrained solves of
 * L(x) = f - lambda.c + miu/2 c.c (f90:2193-2228) followed by lambda -= miu c, miu *= Increment.
 * Device callbacks: c stores this rank's PARTIAL values of the M constraints (constants such as the "-1" of a
 * sphere constraint belong to rank 0 only); cd stores the local rows of the N x M Jacobian, column-major with
 * leading dimension ld (cdx(N,M) of the reference).  M <= 64. */
typedef void (*flgpu_c_fn)(const flgpu_eval_ctx *ctx, double *c_dev, const double *x_dev, int m, int64_t n_local);
typedef void (*flgpu_cd_fn)(const flgpu_eval_ctx *ctx, double *cd_dev, const double *x_dev, int m, int64_t n_local,
                            int64_t ld);
/* Optional FUSED constraint evaluation (an extension; used when flgpu_problem.fused exists too).  The reference composes
 * L = f - lambda.c + miu/2 c.c and L' = f' + cd (miu c - lambda) (f90:2193-2228) and its line search evaluates them at
 * every trial point x = x0 + a*p: x stored, c and the N x M Jacobian stored, L' stored, then dot_product(L', p) --
 * 13n doubles per trial with one constraint.  A line search only needs the scalars L(x) and L'(x).p =
 * f'.p + sum_j (miu c_j - lambda_j) (cd_j . p); with this callback the library asks for them without storing anything:
 *     x = x0 + a*p                         element-wise, multiply THEN add, never stored
 *     FLGPU_WANT_F  : c_dev[j]   = this rank's partial of c_j(x)
 *     FLGPU_WANT_GP : cdp_dev[j] = this rank's partial of (d c_j / d x)(x) . p
 * 2n per trial for the objective plus what the constraints read (2n for an element-local constraint).  The accepted
 * point, c, the Jacobian and L' are formed and stored once per search, as before.  The slope L'.p is then summed in a
 * different association than dot_product(L', p): iterates agree with the unfused composition to rounding, not bit for bit
 * (unlike flgpu_fused_fn for a plain objective). */
typedef void (*flgpu_c_fused_fn)(const flgpu_eval_ctx *ctx, int flags, double *c_dev, double *cdp_dev,
                                 const double *x0_dev, const double *p_dev, double a, int m, int64_t n_local);
typedef struct flgpu_constraints {
    flgpu_c_fn c;
    flgpu_cd_fn cd;
    int m;
    flgpu_c_fused_fn fused; /* optional (NULL = trial points of the inner solves are materialised) */
} flgpu_constraints;
enum { FLGPU_AL_LBFGS = 0, FLGPU_AL_CG = 1 };
typedef struct flgpu_al_options {
    int solver;             /* FLGPU_AL_LBFGS / FLGPU_AL_CG */
    const double *lambda0;  /* host, m values; NULL = 0 (f90:2037-2038) */
    double miu0;            /* default 1, clamped max(1,.) (f90:2039-2040) */
    flgpu_options inner;    /* Memory, Method, Strong, Warning, MaxIteration (also the outer limit), Precision (also the
                               |c| tolerance), MinStepLength, WolfeConst1/2, Increment (also the miu growth), stream,
                               comm, offset, n_global: passed to every inner solve as the reference does */
} flgpu_al_options;
typedef struct flgpu_al_stats {
    int64_t outer_iterations, inner_iterations, trials;
    int status;             /* 0 = |c|^2 < Precision^2, FLGPU_MAX_ITERATION otherwise */
    double cnorm2, miu, f;
    int64_t gpu_launches;
} flgpu_al_stats;
void flgpu_al_options_default(flgpu_al_options *o, int solver);   /* WolfeConst2 0.45 for CG, 0.9 otherwise (f90:2060-2067) */
int flgpu_augmented_lagrangian(const flgpu_problem *prob, const flgpu_constraints *con, const flgpu_al_options *opt,
                               double *x, int64_t n_local, int x_space, flgpu_al_stats *stats);
void flgpu_last_al_stats(flgpu_al_stats *out);
/* the reference's test constraint (test.f90:692-705): unit sphere, c = x.x - 1, cd = 2x, as CUDA kernels */
enum { FLGPU_CON_SPHERE = 0 };
int flgpu_builtin_constraints(int kind, flgpu_constraints *out);
/* reference-ABI constraint callbacks (hpp:372-373): c(cx, x, M, N), cd(cdx, x, M, N); with device callback space x
 * and cdx are device pointers and cx is a HOST array the callback must have filled when it returns */
typedef void (*flgpu_ref_c_fn)(double *cx, const double *x, const int *M, const int *N);
typedef void (*flgpu_ref_cd_fn)(double *cdx, const double *x, const int *M, const int *N);
int flgpu_builtin_ref_constraints(int kind, flgpu_ref_c_fn *c, flgpu_ref_cd_fn *cd);

/* ------------------------------------------------------------------ multi-GPU (row shards) */
/* One process per GPU.  Rank 0 obtains a 128-byte id, the caller distributes it
 * (MPI / torch.distributed / file), every rank creates the communicator. */
int flgpu_comm_unique_id(void *id128);
/* For a device-resident line search written outside the library (include/flgpu_objective.cuh generates one): fills
 * `out` (a flgpu::k::SearchExchange, include/flgpu_exchange.cuh; out_bytes = its size, checked) with what a kernel needs to
 * trade partial sums with the other ranks INSIDE the kernel over the communicator's search mailboxes.  Returns 0; 1 if
 * the communicator has no peer-memory exchange (NCCL fallback: use the host-driven search); 2 on a size mismatch. */
int flgpu_comm_search_exchange(const flgpu_comm *c, void *stream, void *out, size_t out_bytes);
flgpu_comm *flgpu_comm_create(const void *id128, int rank, int nranks);
void flgpu_comm_destroy(flgpu_comm *c);
/* 1 when the per-reduction exchange runs as one kernel over IPC-mapped peer memory (NVLink/NVSwitch
 * stores + flags), 0 when it fell back to ncclAllGather (FLGPU_EXCHANGE=nccl forces the fallback). */
int flgpu_comm_uses_peer_memory(const flgpu_comm *c);

/* ------------------------------------------------------------------ Fortran ABI (drop-in symbols) */
/* Where x and the callbacks' vectors live for the entry points below.  Defaults: x in HOST memory and HOST
 * callbacks, exactly what a program compiled against the reference passes: the library stages x / f' through pinned
 * host buffers around every callback, so unmodified code (e.g. the reference's test/test.cpp) runs as it is -- correct,
 * and PCIe-bound.  Objectives the library knows to be device callbacks -- its built-in ones
 * (flgpu_builtin_ref_callbacks) and anything registered with flgpu_register_fused -- get DEVICE pointers without any
 * setting.  A user callback that expects device pointers and is not registered opts in with
 * flgpu_set_callback_space(FLGPU_SPACE_DEVICE) or FLGPU_CALLBACK_SPACE=device.  -1 restores the automatic choice.
 * Environment overrides: FLGPU_X_SPACE, FLGPU_CALLBACK_SPACE = host|device. */
void flgpu_set_x_space(int space);
void flgpu_set_callback_space(int space);
/* Stream / shard of the optimizer call currently executing on this thread. */
void *flgpu_current_stream(void);
int flgpu_current_device(void);
/* Statistics of the last Fortran-ABI call on this thread. */
void flgpu_last_stats(flgpu_stats *out);
/* Observer applied to Fortran-ABI calls on this thread (NULL to clear). */
void flgpu_set_observer(flgpu_observer_fn fn, void *user);
/* Associates a fused line-search evaluation (see flgpu_fused_fn) with a reference-ABI objective:
 * a later Fortran-ABI call whose `f` argument equals `f` uses it.  fused = NULL removes the entry.
 * `user` is handed to the fused callback as ctx->user.  The built-in objectives are pre-registered. */
void flgpu_register_fused(flgpu_ref_f_fn f, flgpu_fused_fn fused, void *user);
/* Line-search policy (FLGPU_LS_*) of later Fortran-ABI calls on this thread; the reference signatures have no room
 * for it.  -1 = unset: the environment variable FLGPU_LINE_SEARCH = reference|fast decides (default reference). */
void flgpu_set_line_search(int policy);

/* gfortran names (hpp:278-393 "#elif __GNUC__") */
void __nonlinearoptimization_MOD_lbfgs(
    flgpu_ref_f_fn f, flgpu_ref_fd_fn fd, double *x, const int *dim, const int *Memory,
    flgpu_ref_f_fd_fn f_fd, const int32_t *Strong, const int32_t *Warning, const int *MaxIteration,
    const double *Precision, const double *MinStepLength, const double *WolfeConst1,
    const double *WolfeConst2, const double *Increment);
void __nonlinearoptimization_MOD_steepestdescent( /* f90:55-56, hpp:279-291 */
    flgpu_ref_f_fn f, flgpu_ref_fd_fn fd, double *x, const int *dim, flgpu_ref_f_fd_fn f_fd,
    const int32_t *Strong, const int32_t *Warning, const int *MaxIteration, const double *Precision,
    const double *MinStepLength, const double *WolfeConst1, const double *WolfeConst2,
    const double *Increment);
void __nonlinearoptimization_MOD_conjugategradient(
    flgpu_ref_f_fn f, flgpu_ref_fd_fn fd, double *x, const int *dim, const char *Method,
    flgpu_ref_f_fd_fn f_fd, const int32_t *Strong, const int32_t *Warning, const int *MaxIteration,
    const double *Precision, const double *MinStepLength, const double *WolfeConst1,
    const double *WolfeConst2, const double *Increment, int len_Method);
void __nonlinearoptimization_MOD_conjugategradient_basic(
    flgpu_ref_f_fn f, flgpu_ref_fd_fn fd, double *x, const int *dim, const char *Method,
    const int32_t *Strong, const int32_t *Warning, const int *MaxIteration, const double *Precision,
    const double *MinStepLength, const double *WolfeConst1, const double *WolfeConst2,
    const double *Increment, int len_Method);
/* f90:2005-2008, hpp:369-392.  UnconstrainedSolver = 'LBFGS' / 'ConjugateGradient' run here; the dense-Hessian
 * solvers ('BFGS' -- the reference default --, 'NewtonRaphson') are forwarded to the next definition of this symbol
 * (libFL linked after libflgpu) or, if there is none, print a message and exit.  fdd, cdd, ExactStep serve those
 * solvers only. */
void __nonlinearoptimization_MOD_augmentedlagrangian(
    flgpu_ref_f_fn f, flgpu_ref_fd_fn fd, flgpu_ref_c_fn c, flgpu_ref_cd_fn cd, double *x, const int *N, const int *M,
    const char *UnconstrainedSolver, const double *lambda0, const double *miu0, void *fdd, void *cdd,
    const int *ExactStep, const int *Memory, const char *Method, flgpu_ref_f_fd_fn f_fd, const int32_t *Strong,
    const int32_t *Warning, const int *MaxIteration, const double *Precision, const double *MinStepLength,
    const double *WolfeConst1, const double *WolfeConst2, const double *Increment, int len_UnconstrainedSolver,
    int len_Method);
/* ifort names (hpp:9-276 "#ifdef __INTEL_COMPILER") */
void nonlinearoptimization_mp_augmentedlagrangian_(
    flgpu_ref_f_fn f, flgpu_ref_fd_fn fd, flgpu_ref_c_fn c, flgpu_ref_cd_fn cd, double *x, const int *N, const int *M,
    const char *UnconstrainedSolver, const double *lambda0, const double *miu0, void *fdd, void *cdd,
    const int *ExactStep, const int *Memory, const char *Method, flgpu_ref_f_fd_fn f_fd, const int32_t *Strong,
    const int32_t *Warning, const int *MaxIteration, const double *Precision, const double *MinStepLength,
    const double *WolfeConst1, const double *WolfeConst2, const double *Increment, int len_UnconstrainedSolver,
    int len_Method);
void nonlinearoptimization_mp_lbfgs_(
    flgpu_ref_f_fn f, flgpu_ref_fd_fn fd, double *x, const int *dim, const int *Memory,
    flgpu_ref_f_fd_fn f_fd, const int32_t *Strong, const int32_t *Warning, const int *MaxIteration,
    const double *Precision, const double *MinStepLength, const double *WolfeConst1,
    const double *WolfeConst2, const double *Increment);
void nonlinearoptimization_mp_steepestdescent_( /* hpp:11-23 */
    flgpu_ref_f_fn f, flgpu_ref_fd_fn fd, double *x, const int *dim, flgpu_ref_f_fd_fn f_fd,
    const int32_t *Strong, const int32_t *Warning, const int *MaxIteration, const double *Precision,
    const double *MinStepLength, const double *WolfeConst1, const double *WolfeConst2,
    const double *Increment);
void nonlinearoptimization_mp_conjugategradient_(
    flgpu_ref_f_fn f, flgpu_ref_fd_fn fd, double *x, const int *dim, const char *Method,
    flgpu_ref_f_fd_fn f_fd, const int32_t *Strong, const int32_t *Warning, const int *MaxIteration,
    const double *Precision, const double *MinStepLength, const double *WolfeConst1,
    const double *WolfeConst2, const double *Increment, int len_Method);
void nonlinearoptimization_mp_conjugategradient_basic_(
    flgpu_ref_f_fn f, flgpu_ref_fd_fn fd, double *x, const int *dim, const char *Method,
    const int32_t *Strong, const int32_t *Warning, const int *MaxIteration, const double *Precision,
    const double *MinStepLength, const double *WolfeConst1, const double *WolfeConst2,
    const double *Increment, int len_Method);

/* ------------------------------------------------------------------ built-in objectives (CUDA) */
/* The synthetic objectives of the benchmark configs (BASELINE.json), as CUDA kernels. */
/* QUARTIC_SHIFTED: f = sum (x-1)^4 + (x-1)^2, x* = 1 -- the quartic with a non-zero, well-conditioned minimiser */
enum { FLGPU_OBJ_QUARTIC = 0, FLGPU_OBJ_ROSENBROCK = 1, FLGPU_OBJ_DIAGQUAD = 2, FLGPU_OBJ_QUARTIC_SHIFTED = 3 };
enum { FLGPU_START_QUARTIC_U = 0, FLGPU_START_ROSEN_STD = 1, FLGPU_START_ROSEN_PERT = 2, FLGPU_START_ZERO = 3 };
/* 64-bit device-callback form */
int flgpu_builtin_problem(int kind, flgpu_problem *out);
/* reference-ABI form (device pointers in, host fx out; use with the Fortran-ABI symbols) */
int flgpu_builtin_ref_callbacks(int kind, flgpu_ref_f_fn *f, flgpu_ref_fd_fn *fd, flgpu_ref_f_fd_fn *f_fd);
/* x_dev[k] = start(offset + k), k < n_local; enqueued on stream */
int flgpu_fill_start(int start_kind, uint64_t seed, double *x_dev, int64_t offset, int64_t n_local,
                     int64_t n_global, void *stream);

/* Device vectors handed to the primitives, the history operator and the built-in objectives must be 16-byte
 * aligned (128-bit accesses; cudaMalloc / flgpu_malloc pointers are).  x passed to the optimizers may have any
 * alignment: it is copied into library-owned work space. */
/* ------------------------------------------------------------------ vector primitives (a8) */
/* Deterministic device primitives the optimizers are built from; exported for tests and
 * for users writing device callbacks.  All enqueue on `stream`; *_dev outputs are device
 * scalars.  Replaces the dot_product / array-expression sites of f90:591-606,1482,1485. */
int flgpu_vec_dot(const double *a_dev, const double *b_dev, int64_t n, double *out_dev, void *stream);
/* the same over one shard (n local rows) of a vector of n_global rows: this rank's root of the tree above */
int flgpu_vec_dot_sharded(const double *a_dev, const double *b_dev, int64_t n, int64_t n_global, double *out_dev,
                          void *stream);
int flgpu_vec_trial(double *x_dev, const double *x0_dev, const double *p_dev, double a, int64_t n,
                    void *stream); /* x = x0 + a*p, multiply then add (no FMA), f90:1482 */

/* ------------------------------------------------------------------ two-loop recursion as an operator */
/* LBFGS::Before (f90:586-608) + the ring-buffer update of After (f90:622-623) on their own, over the
 * same kernels the optimizer uses: push() appends (or overwrites the oldest of `memory`) the pair
 * s = x1-x0, y = g1-g0 and prepares the coefficients for g1; direction() forms p = -H g1 and
 * xt = x1 + p and returns g1.p and p.p.  g1 must be the vector given to the last push(). */
typedef struct flgpu_history flgpu_history;
flgpu_history *flgpu_history_create(int64_t n_local, int memory, void *stream, flgpu_comm *comm);
int flgpu_history_push(flgpu_history *h, const double *x1_dev, const double *x0_dev, const double *g1_dev,
                       const double *g0_dev);
int flgpu_history_direction(flgpu_history *h, const double *g1_dev, const double *x1_dev, double *p_dev,
                            double *xt_dev, double *gp, double *pp);
void flgpu_history_destroy(flgpu_history *h);

/* ------------------------------------------------------------------ reductions (include/flgpu_reduce.cuh) */
/* Every sum the library forms is deterministic and PARTITION-INDEPENDENT: the vector is cut into chunks of
 * flgpu_chunk_elems(n_global) consecutive elements, a thread block sums one chunk in a fixed order, and the chunk
 * sums are combined by the aligned binary tree over the chunk index, then over the rank index.  Shards that hold the
 * same power-of-two number of whole chunks (any power-of-two dimension on 1, 2, 4, 8 GPUs) therefore give the bits
 * of the single-GPU run; other partitions give a deterministic, rank-identical sum.
 * Kernels written with flgpu_reduce.cuh (flgpu_objective.cuh does) store chunk sums into the work space of their
 * stream -- 8 rows, row r at partials + r * *stride, each with room for at least `nchunks` values -- and call
 * flgpu_reduce_tree(), which enqueues the tree kernel: root of row r -> out_dev[r] (device pointers, NULL = skip).
 * The work space is library-owned and valid until the stream's scratch is released; users must be stream-ordered. */
int64_t flgpu_chunk_elems(int64_t n_global);
int flgpu_reduction_workspace(void *stream, int64_t nchunks, double **partials, int64_t *stride);
int flgpu_reduce_tree(void *stream, int64_t nchunks, int nrows, double *const *out_dev);

/* ------------------------------------------------------------------ device memory helpers */
/* Thin wrappers (cudaMalloc / cudaFree / cudaMemcpyAsync + stream sync) so that C, Fortran
 * (iso_c_binding) and ctypes callers need no CUDA runtime binding of their own. */
void *flgpu_malloc(size_t bytes);
void flgpu_free(void *dev_ptr);
int flgpu_memcpy(void *dst, const void *src, size_t bytes, int dst_space, int src_space, void *stream);
int flgpu_device_count(void); /* 0 when no CUDA device is usable; never aborts */

/* Work-space cache.  By default every call allocates its work space ((2m+5) n doubles) and returns it to the
 * driver before it returns, like the reference (f90:435, 584).  With the cache on (or FLGPU_WORKSPACE_CACHE=1) the
 * buffers are parked for the next call on the same device -- repeated solves (e.g. an augmented-Lagrangian outer
 * loop, f90:2150-2185) then skip cudaMalloc/cudaFree, which cost 0.3-0.6 s for 50 GiB.  flgpu_release_workspace()
 * returns parked buffers to the driver; switching the cache off does so too. */
void flgpu_set_workspace_cache(int on);
void flgpu_release_workspace(void);

/* ------------------------------------------------------------------ introspection */
const char *flgpu_version(void);
/* Per-kernel accumulated CUDA-event time of the last call run with time_kernels=1.
 * names/ms/launches/bytes: arrays of capacity cap; returns the number of kernels. */
int flgpu_kernel_times(const char **names, double *ms, int64_t *launches, double *bytes, int cap);
/* Tuning aid: force K1's (columns per group, groups) shape; (0,0) restores the built-in table. */
void flgpu_debug_set_k1_shape(int columns_per_group, int groups);
/* From inside an observer: zero the accumulators of the call in progress (to time a window). */
void flgpu_reset_kernel_times(void);

#ifdef __cplusplus
}
#endif
#endif /* FLGPU_H */
