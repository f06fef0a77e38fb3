// al.cu -- AugmentedLagrangian over the hot path (reference f90:2005-2241, branches 'LBFGS' f90:2150-2167 and
// 'ConjugateGradient' f90:2168-2185).  The reference builds L, Ld, L_Ld, L_Ld_fdwithf as internal
// procedures over host-associated lambda, miu, cx, cdx (f90:2193-2228) and hands them to LBFGS / ConjugateGradient;
// here they are device callbacks composed from the user's f, fd, f_fd, c, cd plus two small kernels, and x, the
// Jacobian and the multipliers stay in HBM across the outer iterations.  "f90:" = NonlinearOptimization.f90.
#include <dlfcn.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "api_internal.hpp"
#include "backend_cuda.cuh"

using namespace flgpu;
using namespace flgpu_api;

namespace {

constexpr int kMaxConstraints = 64;

// Lx = Lx - dot_product(lambda,cx) + miu/2*dot_product(cx,cx)   (f90:2197, 2213, 2224), applied to the partial of
// ONE rank (apply != 0 on rank 0 only) so that the rank-ordered sum of the partials is L.
__global__ void al_terms_kernel(double *f_dev, const double *c, const double *lambda, double miu, int m, int apply) {
    if (threadIdx.x != 0 || !apply) return;
    double d1 = 0.0, d2 = 0.0;
    for (int j = 0; j < m; j++) d1 = __dadd_rn(d1, __dmul_rn(lambda[j], c[j]));
    for (int j = 0; j < m; j++) d2 = __dadd_rn(d2, __dmul_rn(c[j], c[j]));
    *f_dev = __dadd_rn(__dadd_rn(*f_dev, -d1), __dmul_rn(miu / 2.0, d2));
}

// Ldx = Ldx + matmul(cdx, miu*cx - lambda)   (f90:2205, 2215, 2226): ascending constraint index, separate roundings
__global__ void __launch_bounds__(k::kThreads) al_grad_kernel(double *g, const double *cd, int64_t ld, const double *c,
                                                              const double *lambda, double miu, int m, int64_t n) {
    __shared__ double w[kMaxConstraints];
    for (int j = threadIdx.x; j < m; j += k::kThreads) w[j] = __dadd_rn(__dmul_rn(miu, c[j]), -lambda[j]);
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * k::kThreads;
    for (int64_t i = (int64_t)blockIdx.x * k::kThreads + threadIdx.x; i < n; i += stride) {
        double r = 0.0;
        for (int j = 0; j < m; j++) r = __dadd_rn(r, __dmul_rn(cd[i + (int64_t)j * ld], w[j]));
        g[i] = __dadd_rn(g[i], r);
    }
}

// Fused probe (flgpu_c_fused_fn): L and L'.p from the scalars f, f'.p, c_j, cd_j.p at the trial point.
//   L    = f - lambda.c + miu/2 c.c            on ONE rank's partial (apply != 0 on rank 0 only), as al_terms_kernel
//   L'.p = f'.p + sum_j (miu c_j - lambda_j) (cd_j.p)    on every rank's partial (c is already summed over the ranks)
__global__ void al_probe_terms_kernel(double *f_dev, double *gp_dev, const double *c, const double *cdp, const double *lambda,
                                      double miu, int m, int apply_f, int want_gp) {
    if (threadIdx.x != 0) return;
    if (apply_f) {
        double d1 = 0.0, d2 = 0.0;
        for (int j = 0; j < m; j++) d1 = __dadd_rn(d1, __dmul_rn(lambda[j], c[j]));
        for (int j = 0; j < m; j++) d2 = __dadd_rn(d2, __dmul_rn(c[j], c[j]));
        *f_dev = __dadd_rn(__dadd_rn(*f_dev, -d1), __dmul_rn(miu / 2.0, d2));
    }
    if (want_gp) {
        double r = 0.0;
        for (int j = 0; j < m; j++) r = __dadd_rn(r, __dmul_rn(cdp[j], __dadd_rn(__dmul_rn(miu, c[j]), -lambda[j])));
        *gp_dev = __dadd_rn(*gp_dev, r);
    }
}

struct ALState {
    flgpu_problem user;
    flgpu_constraints con;
    flgpu_comm *comm = nullptr;
    int m = 0;
    int64_t ld = 0;
    double miu = 1.0;
    double *lambda_dev = nullptr, *cpart = nullptr, *cglob = nullptr, *cd_dev = nullptr, *gather = nullptr;
    double *cdp = nullptr;      // fused probe: this rank's partials of cd_j . p
    double *xtmp = nullptr;     // fused store of f' alone (never requested by the reference's searchers): the point
    int64_t n = 0;
    int grid = 1;
    int64_t fused_probes = 0;
};

flgpu_eval_ctx user_ctx(const flgpu_eval_ctx *c, const ALState *S) {
    flgpu_eval_ctx u = *c;
    u.user = S->user.user;
    return u;
}
// c(x): partial values -> values summed over the ranks (rank order) in S->cglob
void eval_c(const flgpu_eval_ctx *u, ALState *S, const double *x, int64_t n) {
    cudaStream_t s = (cudaStream_t)u->stream;
    S->con.c(u, S->cpart, x, S->m, n);
    if (S->comm && S->comm->nranks > 1) rank_sum(S->comm, s, S->cpart, S->m, S->cglob, S->gather, nullptr, 0);
    else FLGPU_CUDA_CHECK(cudaMemcpyAsync(S->cglob, S->cpart, sizeof(double) * S->m, cudaMemcpyDeviceToDevice, s));
}
void add_terms(const flgpu_eval_ctx *u, ALState *S, double *f_dev) {
    al_terms_kernel<<<1, 32, 0, (cudaStream_t)u->stream>>>(f_dev, S->cglob, S->lambda_dev, S->miu, S->m, u->rank == 0);
}
void add_grad(const flgpu_eval_ctx *u, ALState *S, double *g, const double *x, int64_t n) {
    S->con.cd(u, S->cd_dev, x, S->m, n, S->ld);
    al_grad_kernel<<<S->grid, k::kThreads, 0, (cudaStream_t)u->stream>>>(g, S->cd_dev, S->ld, S->cglob, S->lambda_dev,
                                                                         S->miu, S->m, n);
}
// L (f90:2193-2199)
void al_f(const flgpu_eval_ctx *c, double *f_dev, const double *x, int64_t n) {
    ALState *S = (ALState *)c->user;
    flgpu_eval_ctx u = user_ctx(c, S);
    S->user.f(&u, f_dev, x, n);
    eval_c(&u, S, x, n);
    add_terms(&u, S, f_dev);
}
// Ld (f90:2200-2206)
void al_fd(const flgpu_eval_ctx *c, double *g, const double *x, int64_t n) {
    ALState *S = (ALState *)c->user;
    flgpu_eval_ctx u = user_ctx(c, S);
    S->user.fd(&u, g, x, n);
    eval_c(&u, S, x, n);
    add_grad(&u, S, g, x, n);
}
// L_Ld (f90:2207-2217) / L_Ld_fdwithf (f90:2218-2228) when the user supplied f_fd
void al_ffd(const flgpu_eval_ctx *c, double *f_dev, double *g, const double *x, int64_t n) {
    ALState *S = (ALState *)c->user;
    flgpu_eval_ctx u = user_ctx(c, S);
    if (S->user.f_fd) S->user.f_fd(&u, f_dev, g, x, n);
    else { S->user.f(&u, f_dev, x, n); S->user.fd(&u, g, x, n); }
    eval_c(&u, S, x, n);
    add_terms(&u, S, f_dev);
    add_grad(&u, S, g, x, n);
}

// L / L'.p at x0 + a*p without storing the point (flgpu_fused_fn of the composed problem); the accepted point is
// stored by the user's fused evaluation and L' is completed in place exactly as al_ffd does.
void al_fused(const flgpu_eval_ctx *c, int flags, double *f_dev, double *gp_dev, double *x_out, double *g_out,
              const double *x0, const double *p, double a, int64_t n) {
    ALState *S = (ALState *)c->user;
    flgpu_eval_ctx u = user_ctx(c, S);
    cudaStream_t s = (cudaStream_t)u.stream;
    const int want = flags & (FLGPU_WANT_F | FLGPU_WANT_GP), write = flags & (FLGPU_WRITE_X | FLGPU_WRITE_G);
    if (want) {
        S->user.fused(&u, want, f_dev, gp_dev, nullptr, nullptr, x0, p, a, n);
        // c is needed for L and for the weights miu c - lambda of L'.p alike
        S->con.fused(&u, FLGPU_WANT_F | (want & FLGPU_WANT_GP), S->cpart, S->cdp, x0, p, a, S->m, n);
        if (S->comm && S->comm->nranks > 1) rank_sum(S->comm, s, S->cpart, S->m, S->cglob, S->gather, nullptr, 0);
        else FLGPU_CUDA_CHECK(cudaMemcpyAsync(S->cglob, S->cpart, sizeof(double) * S->m, cudaMemcpyDeviceToDevice, s));
        al_probe_terms_kernel<<<1, 32, 0, s>>>(f_dev, gp_dev, S->cglob, S->cdp, S->lambda_dev, S->miu, S->m,
                                               (want & FLGPU_WANT_F) && u.rank == 0, (want & FLGPU_WANT_GP) != 0);
        S->fused_probes++;
    }
    if (write) {
        S->user.fused(&u, write, nullptr, nullptr, x_out, g_out, x0, p, a, n);
        if (write & FLGPU_WRITE_G) {
            const double *xp = x_out;
            if (!(write & FLGPU_WRITE_X)) {          // f' at a point that is not stored: form it in a scratch vector
                if (!S->xtmp) S->xtmp = (double *)ws_alloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
                flgpu_vec_trial(S->xtmp, x0, p, a, n, u.stream);
                xp = S->xtmp;
            }
            eval_c(&u, S, xp, n);
            add_grad(&u, S, g_out, xp, n);
        }
    }
}

thread_local flgpu_al_stats tls_al{};

// ---- built-in constraint: unit sphere (test.f90:692-705)
void sphere_c(const flgpu_eval_ctx *ctx, double *c_dev, const double *x, int m, int64_t n) {
    (void)m;
    flgpu_vec_dot_sharded(x, x, n, ctx->n_global, c_dev, ctx->stream);   // this rank's root of x.x
    if (ctx->rank == 0) k::add_scalar_kernel<<<1, 1, 0, (cudaStream_t)ctx->stream>>>(c_dev, -1.0);
}
__global__ void __launch_bounds__(k::kThreads) scale2_kernel(double *out, const double *x, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * k::kThreads;
    for (int64_t i = (int64_t)blockIdx.x * k::kThreads + threadIdx.x; i < n; i += stride) out[i] = __dmul_rn(2.0, x[i]);
}
void sphere_cd(const flgpu_eval_ctx *ctx, double *cd_dev, const double *x, int m, int64_t n, int64_t ld) {
    (void)m; (void)ld;
    int64_t need = (n + k::kThreads - 1) / k::kThreads;
    const int grid = (int)(need < 1 ? 1 : (need > 148 * 8 ? 148 * 8 : need));
    scale2_kernel<<<grid, k::kThreads, 0, (cudaStream_t)ctx->stream>>>(cd_dev, x, n);
}
// fused form (flgpu_c_fused_fn): x.x - 1 and 2x.p at x = x0 + a*p; x.x in dot_kernel's order, hence the bits of sphere_c
// at the same point
__global__ void __launch_bounds__(k::kThreads, 4) sphere_fused_kernel(const double *__restrict__ x0, const double *__restrict__ p,
                                                                      double a, int64_t n, int64_t ch, k::Work w,
                                                                      double *out_c, double *out_cdp) {
    const k::Chunks C(n, ch);
    int parity = 0;
    for (int64_t c = blockIdx.x; c < C.nchunks; c += gridDim.x) {
        const int64_t hi = C.hi(c);
        double acc[2] = {0.0, 0.0};
        for (int64_t u = C.lo(c) + threadIdx.x; u < hi; u += k::kThreads) {
            const double2 xv = k::ld2(x0, u), pv = k::ld2(p, u);
            const double x = __dadd_rn(xv.x, __dmul_rn(a, pv.x)), y = __dadd_rn(xv.y, __dmul_rn(a, pv.y));
            acc[0] = fma(y, y, fma(x, x, acc[0]));
            acc[1] = fma(__dmul_rn(2.0, y), pv.y, fma(__dmul_rn(2.0, x), pv.x, acc[1]));
        }
        if (C.tail_here(c) && threadIdx.x == 0) {
            const double pv = p[n - 1], x = __dadd_rn(x0[n - 1], __dmul_rn(a, pv));
            acc[0] = fma(x, x, acc[0]);
            acc[1] = fma(__dmul_rn(2.0, x), pv, acc[1]);
        }
        red::chunk_flush<2>(acc, parity, w.partials, w.stride, c);
    }
    if (out_c) {
        double *const o[2] = {out_c, out_cdp};
        red::finish_in_kernel<2>(w.partials, w.stride, C.nchunks, w.tickets, o);
    }
}
void sphere_fused(const flgpu_eval_ctx *ctx, int flags, double *c_dev, double *cdp_dev, const double *x0, const double *p,
                  double a, int m, int64_t n) {
    (void)m; (void)flags;                          // both sums come out of the one pass
    cudaStream_t s = (cudaStream_t)ctx->stream;
    const int64_t n_global = ctx->n_global < n ? n : ctx->n_global;
    const int64_t ch = red::chunk_elems(n_global), nchunks = red::num_chunks(n, ch);
    const k::Work w = scratch_work(s, nchunks);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t grid = (int64_t)sms * 4;
    if (nchunks < grid) grid = nchunks < 1 ? 1 : nchunks;
    const bool in_kernel = nchunks <= red::kBlockChunks;
    sphere_fused_kernel<<<(int)grid, k::kThreads, 0, s>>>(x0, p, a, n, ch, w, in_kernel ? c_dev : nullptr, cdp_dev);
    if (!in_kernel) { double *out[2] = {c_dev, cdp_dev}; flgpu_reduce_tree(ctx->stream, nchunks, 2, out); }
    if (ctx->rank == 0) k::add_scalar_kernel<<<1, 1, 0, s>>>(c_dev, -1.0);
}
// reference-ABI form of the same constraint: device x / cdx, host cx
void ref_sphere_c(double *cx, const double *x, const int *M, const int *N) {
    (void)M;
    cudaStream_t s = (cudaStream_t)flgpu_current_stream();
    double *tmp = scratch_scalar(s) + 2;          // [0] is the f scalar of the built-in objectives
    flgpu_vec_dot(x, x, *N, tmp, s);
    double v = 0.0;
    FLGPU_CUDA_CHECK(cudaMemcpyAsync(&v, tmp, sizeof(double), cudaMemcpyDeviceToHost, s));
    FLGPU_CUDA_CHECK(cudaStreamSynchronize(s));
    cx[0] = v - 1.0;
}
void ref_sphere_cd(double *cdx, const double *x, const int *M, const int *N) {
    (void)M;
    int64_t need = (*N + k::kThreads - 1) / k::kThreads;
    const int grid = (int)(need < 1 ? 1 : (need > 148 * 8 ? 148 * 8 : need));
    scale2_kernel<<<grid, k::kThreads, 0, (cudaStream_t)flgpu_current_stream()>>>(cdx, x, *N);
}

// ---- adapters for reference-ABI constraint callbacks
struct RefConAdapter {
    flgpu_ref_c_fn c;
    flgpu_ref_cd_fn cd;
    int cb_space;
    int N, M;
    double *xh = nullptr, *cdh = nullptr, *ch = nullptr;   // pinned staging
};
struct RefALUser {              // what the composed callbacks see as the "user" of the wrapped problem
    RefAdapter obj;             // MUST stay first: ad_f / ad_fd / ad_ffd cast ctx->user to RefAdapter*
    RefConAdapter con;
};
void ad_c(const flgpu_eval_ctx *ctx, double *c_dev, const double *x, int m, int64_t n) {
    RefConAdapter *A = &((RefALUser *)ctx->user)->con;
    cudaStream_t s = (cudaStream_t)ctx->stream;
    int M = m, N = (int)n;
    if (A->cb_space == FLGPU_SPACE_HOST) {
        FLGPU_CUDA_CHECK(cudaMemcpyAsync(A->xh, x, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, s));
        FLGPU_CUDA_CHECK(cudaStreamSynchronize(s));
        A->c(A->ch, A->xh, &M, &N);
    } else {
        A->c(A->ch, x, &M, &N);
    }
    // the staging values are consumed by this copy before the next callback overwrites them (stream order +
    // the synchronisation every callback of this adapter performs)
    FLGPU_CUDA_CHECK(cudaMemcpyAsync(c_dev, A->ch, sizeof(double) * (size_t)m, cudaMemcpyHostToDevice, s));
    FLGPU_CUDA_CHECK(cudaStreamSynchronize(s));
}
void ad_cd(const flgpu_eval_ctx *ctx, double *cd_dev, const double *x, int m, int64_t n, int64_t ld) {
    RefConAdapter *A = &((RefALUser *)ctx->user)->con;
    cudaStream_t s = (cudaStream_t)ctx->stream;
    int M = m, N = (int)n;
    if (A->cb_space == FLGPU_SPACE_HOST) {
        FLGPU_CUDA_CHECK(cudaMemcpyAsync(A->xh, x, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, s));
        FLGPU_CUDA_CHECK(cudaStreamSynchronize(s));
        A->cd(A->cdh, A->xh, &M, &N);                                   // cdx(N,M), leading dimension N
        FLGPU_CUDA_CHECK(cudaMemcpy2DAsync(cd_dev, sizeof(double) * (size_t)ld, A->cdh, sizeof(double) * (size_t)n,
                                           sizeof(double) * (size_t)n, (size_t)m, cudaMemcpyHostToDevice, s));
        FLGPU_CUDA_CHECK(cudaStreamSynchronize(s));
    } else if (ld == n || m == 1) {
        A->cd(cd_dev, x, &M, &N);
    } else {
        fatal("AugmentedLagrangian: device constraint Jacobians need n to be a multiple of 32 when M > 1");
    }
}

bool str_is(const char *s, int len, const char *lit) {   // Fortran == on blank-padded strings
    const int n = (int)std::strlen(lit);
    if (!s) return false;
    for (int i = 0; i < n; i++) if (i >= len || s[i] != lit[i]) return false;
    for (int i = n; i < len; i++) if (s[i] != ' ') return false;
    return true;
}

}  // namespace

extern "C" {

void flgpu_al_options_default(flgpu_al_options *o, int solver) {
    std::memset(o, 0, sizeof *o);
    o->solver = solver;
    o->miu0 = 1.0;
    flgpu_options_default(&o->inner, solver == FLGPU_AL_CG);
}

void flgpu_last_al_stats(flgpu_al_stats *out) { *out = tls_al; }

int flgpu_builtin_constraints(int kind, flgpu_constraints *out) {
    if (kind != FLGPU_CON_SPHERE) return 1;
    out->c = sphere_c; out->cd = sphere_cd; out->m = 1; out->fused = sphere_fused;
    return 0;
}
int flgpu_builtin_ref_constraints(int kind, flgpu_ref_c_fn *c, flgpu_ref_cd_fn *cd) {
    if (kind != FLGPU_CON_SPHERE) return 1;
    *c = ref_sphere_c; *cd = ref_sphere_cd;
    return 0;
}

int flgpu_augmented_lagrangian(const flgpu_problem *prob, const flgpu_constraints *con, const flgpu_al_options *opt,
                               double *x, int64_t n, int x_space, flgpu_al_stats *stats) {
    require_device();
    if (!prob || !prob->f || !prob->fd) fatal("flgpu: f and fd callbacks are required");
    if (!con || !con->c || !con->cd || con->m < 1) fatal("flgpu: constraint callbacks c, cd and M >= 1 are required");
    if (con->m > kMaxConstraints) fatal("flgpu: AugmentedLagrangian supports at most 64 constraints");
    if (opt->solver != FLGPU_AL_LBFGS && opt->solver != FLGPU_AL_CG) fatal("flgpu: unknown AugmentedLagrangian solver");
    const int m = con->m;
    flgpu_options in = opt->inner;
    // f90:2036-2075: the common tunables are clamped once here and handed to every inner solve as they are
    const int maxit = in.max_iteration;
    const double tol = in.precision, incrmt = in.increment;
    in.wolfe_c1 = std::fmax(1e-15, in.wolfe_c1);
    in.wolfe_c2 = std::fmin(1.0 - 1e-15, std::fmax(in.wolfe_c1 + 1e-15, in.wolfe_c2));
    in.memory = in.memory > 1 ? in.memory : 1;
    in.no_clamp = 0;
    const double tolsq = tol * tol;
    double miu = std::fmax(1.0, opt->miu0);
    std::vector<double> lambda(m, 0.0), cx(m, 0.0);
    if (opt->lambda0) for (int j = 0; j < m; j++) lambda[j] = opt->lambda0[j];

    cudaStream_t s = (cudaStream_t)in.stream;
    bool own_stream = false;
    if (!s) { FLGPU_CUDA_CHECK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking)); own_stream = true; in.stream = (void *)s; }
    // every outer iteration runs a full inner solve: keep the work space between them (f90:2150-2185 re-enters
    // LBFGS / ConjugateGradient, which re-allocate; here the buffers are parked and reused)
    ws_arena_begin();                             // a per-call arena: the process-wide cache switch is not touched
    ALState S;
    S.user = *prob; S.con = *con; S.comm = in.comm; S.m = m;
    S.ld = (n + 31) / 32 * 32; if (S.ld == 0) S.ld = 32;
    const int G = S.comm ? S.comm->nranks : 1;
    int dev = 0, sms = 148;
    FLGPU_CUDA_CHECK(cudaGetDevice(&dev));
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t need = (n + k::kThreads - 1) / k::kThreads;
    S.grid = (int)(need < 1 ? 1 : (need > (int64_t)sms * 8 ? (int64_t)sms * 8 : need));
    S.lambda_dev = (double *)ws_alloc(sizeof(double) * kMaxConstraints);
    S.cpart = (double *)ws_alloc(sizeof(double) * kMaxConstraints);
    S.cglob = (double *)ws_alloc(sizeof(double) * kMaxConstraints);
    S.gather = (double *)ws_alloc(sizeof(double) * kMaxConstraints * (size_t)(G > 1 ? G : 1));
    S.cd_dev = (double *)ws_alloc(sizeof(double) * (size_t)S.ld * (size_t)m);
    S.cdp = (double *)ws_alloc(sizeof(double) * kMaxConstraints);
    S.n = n;
    double *xdev = x;
    if (x_space == FLGPU_SPACE_HOST) {          // x stays in HBM across the outer iterations
        xdev = (double *)ws_alloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
        FLGPU_CUDA_CHECK(cudaMemcpyAsync(xdev, x, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, s));
    }
    double *ch = nullptr;
    FLGPU_CUDA_CHECK(cudaMallocHost((void **)&ch, sizeof(double) * kMaxConstraints));

    flgpu_problem L;
    L.f = al_f; L.fd = al_fd; L.f_fd = al_ffd;       // the reference always passes f_fd = L_Ld / L_Ld_fdwithf
    // line-search trials without materialised points when both the objective and the constraints can be probed
    // (flgpu_options.no_fused switches it off like every other fused evaluation)
    const bool al_fuse = prob->fused && con->fused;
    L.user = &S; L.fused = al_fuse ? al_fused : nullptr; L.search = nullptr; L.search_caps = 0; L.update = nullptr; L.direction = nullptr; L.fused_multi = nullptr;
    flgpu_eval_ctx ctx;
    ctx.user = prob->user; ctx.stream = (void *)s; ctx.offset = in.offset; ctx.n_global = in.n_global ? in.n_global : n;
    ctx.rank = S.comm ? S.comm->rank : 0; ctx.nranks = G; ctx.device = dev;

    flgpu_al_stats A;
    std::memset(&A, 0, sizeof A);
    A.status = FLGPU_MAX_ITERATION;
    double cc = 0.0;
    int iIteration;
    for (iIteration = 1; iIteration <= maxit; iIteration++) {                          // f90:2150-2185
        FLGPU_CUDA_CHECK(cudaMemcpyAsync(S.lambda_dev, lambda.data(), sizeof(double) * m, cudaMemcpyHostToDevice, s));
        FLGPU_CUDA_CHECK(cudaStreamSynchronize(s));
        S.miu = miu;
        flgpu_stats st;
        run(opt->solver == FLGPU_AL_CG ? ALGO_CG : ALGO_LBFGS, &L, &in, xdev, n, FLGPU_SPACE_DEVICE, &st);
        A.inner_iterations += st.iterations;
        A.trials += st.n_trials;
        A.gpu_launches += st.gpu_launches;
        A.outer_iterations = iIteration;
        A.f = st.f;
        eval_c(&ctx, &S, xdev, n);                                                     // call c(cx,x,M,N)
        FLGPU_CUDA_CHECK(cudaMemcpyAsync(ch, S.cglob, sizeof(double) * m, cudaMemcpyDeviceToHost, s));
        FLGPU_CUDA_CHECK(cudaStreamSynchronize(s));
        cc = 0.0;
        for (int j = 0; j < m; j++) { cx[j] = ch[j]; cc += cx[j] * cx[j]; }
        A.cnorm2 = cc;
        if (cc < tolsq) { A.status = 0; break; }
        for (int j = 0; j < m; j++) lambda[j] = lambda[j] - miu * cx[j];
        miu = miu * incrmt;
    }
    A.miu = miu;
    if (iIteration > maxit && in.warning) {                                            // f90:2187-2190
        std::printf(" Failed augmented Lagrangian: max iteration exceeded!\n");
        std::printf(" Euclidean norm of constraint violation = %.17g\n", std::sqrt(cc));
    }
    if (x_space == FLGPU_SPACE_HOST) {
        FLGPU_CUDA_CHECK(cudaMemcpyAsync(x, xdev, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, s));
        FLGPU_CUDA_CHECK(cudaStreamSynchronize(s));
        ws_free(xdev, sizeof(double) * (size_t)(n > 0 ? n : 1));
    }
    FLGPU_CUDA_CHECK(cudaStreamSynchronize(s));
    cudaFreeHost(ch);
    ws_free(S.lambda_dev, sizeof(double) * kMaxConstraints);
    ws_free(S.cpart, sizeof(double) * kMaxConstraints);
    ws_free(S.cglob, sizeof(double) * kMaxConstraints);
    ws_free(S.gather, sizeof(double) * kMaxConstraints * (size_t)(G > 1 ? G : 1));
    ws_free(S.cd_dev, sizeof(double) * (size_t)S.ld * (size_t)m);
    ws_free(S.cdp, sizeof(double) * kMaxConstraints);
    if (S.xtmp) ws_free(S.xtmp, sizeof(double) * (size_t)(n > 0 ? n : 1));
    if (own_stream) { scratch_release(s); cudaStreamDestroy(s); }
    ws_arena_end();                                 // returns the parked work space to the driver
    tls_al = A;
    if (stats) *stats = A;
    return 0;
}

void __nonlinearoptimization_MOD_augmentedlagrangian(
    flgpu_ref_f_fn f, flgpu_ref_fd_fn fd, flgpu_ref_c_fn c, flgpu_ref_cd_fn cd, double *x, const int *N, const int *M,
    const char *UnconstrainedSolver, const double *lambda0, const double *miu0, void *fdd, void *cdd,
    const int *ExactStep, const int *Memory, const char *Method, flgpu_ref_f_fd_fn f_fd, const int32_t *Strong,
    const int32_t *Warning, const int *MaxIteration, const double *Precision, const double *MinStepLength,
    const double *WolfeConst1, const double *WolfeConst2, const double *Increment, int len_UnconstrainedSolver,
    int len_Method) {
    const bool is_cg = str_is(UnconstrainedSolver, len_UnconstrainedSolver, "ConjugateGradient");
    if (!is_cg && !str_is(UnconstrainedSolver, len_UnconstrainedSolver, "LBFGS")) {
        // 'BFGS' (the reference default) and 'NewtonRaphson' are dense-Hessian solvers outside the GPU hot path: hand
        // the call to the next definition of this symbol (libFL linked or loaded after libflgpu), if there is one
        typedef void (*next_fn)(flgpu_ref_f_fn, flgpu_ref_fd_fn, flgpu_ref_c_fn, flgpu_ref_cd_fn, double *, const int *,
                                const int *, const char *, const double *, const double *, void *, void *, const int *,
                                const int *, const char *, flgpu_ref_f_fd_fn, const int32_t *, const int32_t *,
                                const int *, const double *, const double *, const double *, const double *,
                                const double *, int, int);
        next_fn next = (next_fn)dlsym(RTLD_NEXT, "__nonlinearoptimization_MOD_augmentedlagrangian");
        if (next) {
            next(f, fd, c, cd, x, N, M, UnconstrainedSolver, lambda0, miu0, fdd, cdd, ExactStep, Memory, Method, f_fd,
                 Strong, Warning, MaxIteration, Precision, MinStepLength, WolfeConst1, WolfeConst2, Increment,
                 len_UnconstrainedSolver, len_Method);
            return;
        }
        std::string name(UnconstrainedSolver ? UnconstrainedSolver : "BFGS",
                         UnconstrainedSolver ? (size_t)(len_UnconstrainedSolver > 0 ? len_UnconstrainedSolver : 0) : 4);
        std::printf(" Program abort: unconstrained solver %s is a dense-Hessian method outside the GPU hot path; "
                    "libflgpu serves UnconstrainedSolver = LBFGS or ConjugateGradient (link libFL after libflgpu for "
                    "the others)\n", name.c_str());
        std::fflush(stdout);
        std::exit(1);
    }
    require_device();
    flgpu_al_options o;
    flgpu_al_options_default(&o, is_cg ? FLGPU_AL_CG : FLGPU_AL_LBFGS);
    o.lambda0 = lambda0;
    if (miu0) o.miu0 = *miu0;
    if (Memory) o.inner.memory = *Memory;
    if (Method) {                                    // character*32 type = Method; the inner solver reads 2 characters
        char t0 = len_Method > 0 ? Method[0] : ' ', t1 = len_Method > 1 ? Method[1] : ' ';
        if (t0 == 'P' && t1 == 'R') o.inner.method = FLGPU_CG_PR;
        else if (t0 == 'D' && t1 == 'Y') o.inner.method = FLGPU_CG_DY;
        else if (is_cg) {
            std::printf(" Program abort: unsupported conjugate gradient method %.*s\n", len_Method, Method);
            std::fflush(stdout);
            std::exit(1);
        }
    }
    fill_optional(o.inner, Strong, Warning, MaxIteration, Precision, MinStepLength, WolfeConst1, WolfeConst2, Increment);
    apply_thread_settings(o.inner);

    RefALUser U;
    flgpu_problem prob;
    ref_adapter_init(U.obj, f, fd, f_fd, *N, &prob);
    prob.user = &U;                                   // RefAdapter is the first member: ad_f / ad_fd / ad_ffd still work
    prob.fused = nullptr;
    prob.update = nullptr;
    prob.direction = nullptr;
    prob.fused_multi = nullptr;
    U.con.c = c; U.con.cd = cd; U.con.cb_space = U.obj.cb_space; U.con.N = *N; U.con.M = *M;
    FLGPU_CUDA_CHECK(cudaMallocHost((void **)&U.con.ch, sizeof(double) * (size_t)(*M > 0 ? *M : 1)));
    if (U.con.cb_space == FLGPU_SPACE_HOST) {
        FLGPU_CUDA_CHECK(cudaMallocHost((void **)&U.con.xh, sizeof(double) * (size_t)(*N > 0 ? *N : 1)));
        FLGPU_CUDA_CHECK(cudaMallocHost((void **)&U.con.cdh, sizeof(double) * (size_t)(*N > 0 ? *N : 1) * (size_t)(*M > 0 ? *M : 1)));
    }
    flgpu_constraints con;
    con.c = ad_c; con.cd = ad_cd; con.m = *M; con.fused = nullptr;
    flgpu_augmented_lagrangian(&prob, &con, &o, x, *N, x_space_now(), nullptr);
    ref_adapter_free(U.obj);
    if (U.con.ch) cudaFreeHost(U.con.ch);
    if (U.con.xh) cudaFreeHost(U.con.xh);
    if (U.con.cdh) cudaFreeHost(U.con.cdh);
}

void nonlinearoptimization_mp_augmentedlagrangian_(
    flgpu_ref_f_fn f, flgpu_ref_fd_fn fd, flgpu_ref_c_fn c, flgpu_ref_cd_fn cd, double *x, const int *N, const int *M,
    const char *UnconstrainedSolver, const double *lambda0, const double *miu0, void *fdd, void *cdd,
    const int *ExactStep, const int *Memory, const char *Method, flgpu_ref_f_fd_fn f_fd, const int32_t *Strong,
    const int32_t *Warning, const int *MaxIteration, const double *Precision, const double *MinStepLength,
    const double *WolfeConst1, const double *WolfeConst2, const double *Increment, int len_UnconstrainedSolver,
    int len_Method) {
    __nonlinearoptimization_MOD_augmentedlagrangian(f, fd, c, cd, x, N, M, UnconstrainedSolver, lambda0, miu0, fdd, cdd,
                                                    ExactStep, Memory, Method, f_fd, Strong, Warning, MaxIteration,
                                                    Precision, MinStepLength, WolfeConst1, WolfeConst2, Increment,
                                                    len_UnconstrainedSolver, len_Method);
}

}  // extern "C"
