// flgpu_objective.cuh -- write an objective once as a device functor, get all four libflgpu callbacks.
//
// The reference asks the user for f, fd and optionally f_fd (NonlinearOptimization.f90:33-43).  On the GPU those are
// kernels, and the line search is fastest when the objective kernel also forms the trial point x0 + a*p itself
// (flgpu_fused_fn, flgpu.h).  For objectives that are sums of terms over single elements or over pairs
// (x_{2j}, x_{2j+1}) -- the quartic of the reference's own tests, extended Rosenbrock, diagonal quadratics, any
// separable loss -- this header generates f, fd, f_fd AND the fused evaluation from one functor, with the library's
// deterministic reduction (fixed per-thread order, shuffle butterfly, fixed-order sum over blocks by the last block).
//
//   struct Quartic {                                   // f = sum x^4 (test/test.f90:630-663)
//       static constexpr int WIDTH = 1;                // 1: element-local, 2: pairs (x_{2j}, x_{2j+1})
//       __device__ void eval(int64_t i, double x, double &f, double &g) const { f = x*x*x*x; g = 4*x*x*x; }
//   };
//   flgpu_problem prob = flgpu_obj::make_problem<Quartic>(&my_functor_on_the_host);   // keeps the pointer, not a copy
//   flgpu_lbfgs(&prob, &opt, x, n, FLGPU_SPACE_DEVICE, &stats);
//
// WIDTH = 2 functors implement
//       __device__ void eval2(int64_t i, double xa, double xb, double &f, double &ga, double &gb) const;   // i even
//       __device__ void eval_tail(int64_t i, double x, double &f, double &g) const;                        // odd n
// `i` is the GLOBAL index (row-sharded runs: ctx->offset + local index; shards start on even indices).  The functor
// is passed to the kernels by value: keep it small and trivially copyable (pointers to device tables are fine).
// Compile the including file with nvcc -std=c++17 for sm_100a and link libflgpu.so.  Vectors are 16-byte aligned (they are the
// library's own work space).
#pragma once
#include <cooperative_groups.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#include "flgpu.h"
#include "flgpu_search_core.hpp"

namespace flgpu_obj {

constexpr int kThreads = 256;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block tree + last-block finish; out[i] receives accumulator i (fixed order: bitwise repeatable)
template <int NACC>
__device__ __forceinline__ void reduce_to(double (&acc)[NACC], double *(&out)[NACC], double *partials,
                                          unsigned int *ticket) {
    __shared__ double sh[NACC][kThreads / 32];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NACC; i++) {
        const double v = warp_sum(acc[i]);
        if (lane == 0) sh[i][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x < NACC) {
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < kThreads / 32; q++) s += sh[threadIdx.x][q];
        partials[(size_t)blockIdx.x * NACC + threadIdx.x] = s;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (warp < NACC) {
        double s = 0.0;
        for (unsigned b = lane; b < gridDim.x; b += 32) s += __ldcg(&partials[(size_t)b * NACC + warp]);
        s = warp_sum(s);
        if (lane == 0) *out[warp] = s;
    }
    if (threadIdx.x == 0) *ticket = 0u;
}

struct Args {
    const double *x;      // the point, or x0 when FUSED
    const double *p;      // FUSED
    double a;             // FUSED
    double *x_out, *g_out, *f_out, *gp_out;
    int64_t n, offset;
    double *partials;
    unsigned int *ticket;
};

// this thread's share of the units, in grid-stride order (shared by objective_kernel and search_kernel: same bits)
template <class Obj, bool FUSED, bool WANT_F, bool WANT_GP, bool WRITE_X, bool WRITE_G>
__device__ __forceinline__ void accumulate(const Obj &obj, const Args &a, double &fsum, double &gpsum) {
    const int64_t nu = a.n >> 1;
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    for (int64_t u = (int64_t)blockIdx.x * kThreads + threadIdx.x; u < nu; u += stride) {
        double2 x = __ldg(reinterpret_cast<const double2 *>(a.x) + u), pv = make_double2(0.0, 0.0);
        if (FUSED) {
            pv = __ldg(reinterpret_cast<const double2 *>(a.p) + u);
            x.x = __dadd_rn(x.x, __dmul_rn(a.a, pv.x));      // multiply, then add: the reference's x0+a*p (f90:1482)
            x.y = __dadd_rn(x.y, __dmul_rn(a.a, pv.y));
            if (WRITE_X) reinterpret_cast<double2 *>(a.x_out)[u] = x;
        }
        const int64_t i = a.offset + 2 * u;
        double2 g;
        if constexpr (Obj::WIDTH == 2) {
            double f = 0.0;
            obj.eval2(i, x.x, x.y, f, g.x, g.y);
            if (WANT_F) fsum += f;
        } else {
            double f0 = 0.0, f1 = 0.0;
            obj.eval(i, x.x, f0, g.x);
            obj.eval(i + 1, x.y, f1, g.y);
            if (WANT_F) { fsum += f0; fsum += f1; }
        }
        if (WRITE_G) reinterpret_cast<double2 *>(a.g_out)[u] = g;
        if (WANT_GP) gpsum = fma(g.y, pv.y, fma(g.x, pv.x, gpsum));
    }
    if ((a.n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {   // last element of an odd-length shard
        const int64_t k = a.n - 1;
        double x = a.x[k], pv = 0.0, f = 0.0, g = 0.0;
        if (FUSED) {
            pv = a.p[k];
            x = __dadd_rn(x, __dmul_rn(a.a, pv));
            if (WRITE_X) a.x_out[k] = x;
        }
        if constexpr (Obj::WIDTH == 2) obj.eval_tail(a.offset + k, x, f, g);
        else obj.eval(a.offset + k, x, f, g);
        if (WANT_F) fsum += f;
        if (WRITE_G) a.g_out[k] = g;
        if (WANT_GP) gpsum = fma(g, pv, gpsum);
    }
}

template <class Obj, bool FUSED, bool WANT_F, bool WANT_GP, bool WRITE_X, bool WRITE_G>
__global__ void __launch_bounds__(kThreads, 4) objective_kernel(Obj obj, Args a) {
    double fsum = 0.0, gpsum = 0.0;
    accumulate<Obj, FUSED, WANT_F, WANT_GP, WRITE_X, WRITE_G>(obj, a, fsum, gpsum);
    if (WANT_F && WANT_GP) {
        double acc[2] = {fsum, gpsum};
        double *out[2] = {a.f_out, a.gp_out};
        reduce_to<2>(acc, out, a.partials, a.ticket);
    } else if (WANT_F) {
        double acc[1] = {fsum};
        double *out[1] = {a.f_out};
        reduce_to<1>(acc, out, a.partials, a.ticket);
    } else if (WANT_GP) {
        double acc[1] = {gpsum};
        double *out[1] = {a.gp_out};
        reduce_to<1>(acc, out, a.partials, a.ticket);
    }
}

// ---- device-resident line search (flgpu_search_fn) for functor objectives: the whole Wolfe / Strong-Wolfe search in
// one cooperative kernel.  Every thread runs the same SearchCore state machine (flgpu_search_core.hpp, the source the
// host driver compiles) on the same values; an evaluation = accumulate() + block tree + one grid barrier + the
// fixed-order sum over blocks, repeated by every block.  Same grid as objective_kernel => same bits as the host-driven
// fused search => same decisions.  Single GPU (row-sharded runs fall back to the host-driven search).
struct SearchArgs {
    Args o;                                  // x = x0; x_out / g_out = accepted point / gradient
    double c1, c2abs, fx0, phid0, incr, a0;
    int strong, fdwithf;
    double *partials;                        // [2][gridDim.x][2]
    double *result;
};

constexpr double kEvalBudget = 100000.0;

template <class Obj>
struct DevSearch : flgpu::SearchCore<DevSearch<Obj>> {
    const Obj &obj;
    const SearchArgs &K;
    double (*sh)[kThreads / 32];
    double *bc;
    double f_cur = 0.0, gp_cur = 0.0, a_x = 0.0, a_g = 0.0;
    bool have_x = false, have_g = false;
    int parity = 0;
    double trials = 0.0, n_f = 0.0, n_fd = 0.0, n_ffd = 0.0, n_fonly = 0.0;
    __device__ DevSearch(const Obj &o, const SearchArgs &k, double (*s)[kThreads / 32], double *b)
        : obj(o), K(k), sh(s), bc(b) {}
    template <bool F, bool GP>
    __device__ void eval() {
        constexpr int NACC = (F && GP) ? 2 : 1;
        Args o = K.o;
        o.a = a_x;
        double fsum = 0.0, gpsum = 0.0;
        accumulate<Obj, true, F, GP, false, false>(obj, o, fsum, gpsum);
        double acc[2] = {F ? fsum : gpsum, gpsum};
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int i = 0; i < NACC; i++) {
            const double v = warp_sum(acc[i]);
            if (lane == 0) sh[i][warp] = v;
        }
        __syncthreads();
        double *part = K.partials + (size_t)parity * gridDim.x * 2;
        if (threadIdx.x < NACC) {
            double s = 0.0;
#pragma unroll
            for (int q = 0; q < kThreads / 32; q++) s += sh[threadIdx.x][q];
            part[(size_t)blockIdx.x * 2 + threadIdx.x] = s;
        }
        __threadfence();
        cooperative_groups::this_grid().sync();
        if (warp < NACC) {
            double s = 0.0;
            for (unsigned b = lane; b < gridDim.x; b += 32) s += __ldcg(&part[(size_t)b * 2 + warp]);
            s = warp_sum(s);
            if (lane == 0) bc[warp] = s;
        }
        __syncthreads();
        if (F) f_cur = bc[0];
        if (GP) gp_cur = bc[NACC - 1];
        __syncthreads();
        parity ^= 1;
    }
    __device__ void form(double step) { a_x = step; have_x = true; trials += 1.0; }
    __device__ void call_f() { eval<true, false>(); n_f += 1.0; }
    __device__ void call_fd() { eval<false, true>(); a_g = a_x; have_g = true; n_fd += 1.0; }
    __device__ void call_ffd() { eval<true, true>(); a_g = a_x; have_g = true; n_ffd += 1.0; }
    __device__ double slope() { return gp_cur; }
    __device__ double fx() { return f_cur; }
    __device__ void set_fx(double v) { f_cur = v; }
    __device__ void adopt_pre() {}
    __device__ void count_f_only() { n_fonly += 1.0; }
    // a kernel must terminate whatever the objective returns: after kEvalBudget evaluations the search gives up
    // (result[7] < 0 tells the host); a finite objective needs a few hundred at most
    __device__ bool aborted() const { return n_f + n_fd + n_ffd > kEvalBudget; }
};

// FAST = the FLGPU_LS_FAST searcher (SearchCore::fast), a template parameter: the reference-exact kernel stays as it is
template <class Obj, bool FAST>
__global__ void __launch_bounds__(kThreads, 4) search_kernel(Obj obj, SearchArgs K) {
    __shared__ double sh[2][kThreads / 32];
    __shared__ double bc[2];
    DevSearch<Obj> S(obj, K, sh, bc);
    S.c1 = K.c1; S.c2abs = K.c2abs; S.fx0 = K.fx0; S.phid0 = K.phid0; S.incr = K.incr;
    S.fdwithf = K.fdwithf != 0; S.a = K.a0; S.f_cur = K.fx0; S.pre = 0;
    if (FAST) S.fast(K.strong != 0);
    else if (K.strong) S.strongwolfe(); else S.wolfe();
    Args o = K.o;
    double f0 = 0.0, g0 = 0.0;
    if (S.have_x && S.have_g && S.a_x == S.a_g) {
        o.a = S.a_x;
        accumulate<Obj, true, false, false, true, true>(obj, o, f0, g0);
    } else {                                  // never taken by the reference's searchers; kept for fidelity
        if (S.have_x) { o.a = S.a_x; accumulate<Obj, true, false, false, true, false>(obj, o, f0, g0); }
        if (S.have_g) { o.a = S.a_g; accumulate<Obj, true, false, false, false, true>(obj, o, f0, g0); }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        K.result[0] = S.a; K.result[1] = S.f_cur; K.result[2] = S.trials; K.result[3] = S.n_f;
        K.result[4] = S.n_fd; K.result[5] = S.n_ffd; K.result[6] = S.n_fonly;
        K.result[7] = S.aborted() ? -1.0 : 0.0;
    }
}

template <class Obj>
struct Callbacks {
    template <bool FUSED, bool F, bool GP, bool WX, bool WG>
    static void launch(const flgpu_eval_ctx *ctx, double *f_dev, double *gp_dev, double *x_out, double *g_out,
                       const double *x, const double *p, double a, int64_t n) {
        cudaStream_t s = (cudaStream_t)ctx->stream;
        Args A;
        int max_blocks = 0;
        flgpu_reduction_workspace(ctx->stream, &A.partials, &A.ticket, &max_blocks);   // per-stream, library-owned
        A.x = x; A.p = p; A.a = a; A.x_out = x_out; A.g_out = g_out; A.f_out = f_dev; A.gp_out = gp_dev;
        A.n = n; A.offset = ctx->offset;
        objective_kernel<Obj, FUSED, F, GP, WX, WG><<<grid_for(n, max_blocks), kThreads, 0, s>>>(*(const Obj *)ctx->user, A);
    }
    static int grid_for(int64_t n, int max_blocks) {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        int64_t need = (n / 2 + kThreads) / kThreads, grid = (int64_t)sms * 4;       // one full wave of resident CTAs
        if (grid > max_blocks) grid = max_blocks;
        if (need < grid) grid = need < 1 ? 1 : need;
        return (int)grid;
    }
    static void search(const flgpu_eval_ctx *ctx, const flgpu_search_args *A, int64_t n) {
        if (A->policy == FLGPU_LS_FAST) search_policy<true>(ctx, A, n); else search_policy<false>(ctx, A, n);
    }
    template <bool FAST>
    static void search_policy(const flgpu_eval_ctx *ctx, const flgpu_search_args *A, int64_t n) {
        if (A->comm) { std::fprintf(stderr, "flgpu_obj: the header's device-resident search is single-GPU\n"); std::abort(); }
        SearchArgs K;
        int max_blocks = 0;
        flgpu_reduction_workspace(ctx->stream, &K.partials, &K.o.ticket, &max_blocks);
        K.o.partials = K.partials;
        K.o.x = A->x0_dev; K.o.p = A->p_dev; K.o.a = 0.0; K.o.x_out = A->x_out; K.o.g_out = A->g_out;
        K.o.f_out = nullptr; K.o.gp_out = nullptr; K.o.n = n; K.o.offset = ctx->offset;
        K.c1 = A->c1; K.c2abs = A->c2abs; K.fx0 = A->fx0; K.phid0 = A->phid0; K.incr = A->incr; K.a0 = A->a;
        K.strong = A->strong; K.fdwithf = A->fdwithf; K.result = A->result_dev;
        int resident = 0, sms = 148, dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, search_kernel<Obj, FAST>, kThreads, 0);
        int grid = grid_for(n, max_blocks);
        if (grid > resident * sms) grid = resident * sms;
        Obj obj = *(const Obj *)ctx->user;
        void *params[] = {&obj, &K};
        cudaError_t e = cudaLaunchCooperativeKernel((void *)search_kernel<Obj, FAST>, dim3(grid), dim3(kThreads), params, 0,
                                                    (cudaStream_t)ctx->stream);
        if (e != cudaSuccess) { std::fprintf(stderr, "flgpu_obj: cooperative launch failed: %s\n", cudaGetErrorString(e)); std::abort(); }
    }
    static void f(const flgpu_eval_ctx *ctx, double *f_dev, const double *x, int64_t n) {
        launch<false, true, false, false, false>(ctx, f_dev, nullptr, nullptr, nullptr, x, nullptr, 0.0, n);
    }
    static void fd(const flgpu_eval_ctx *ctx, double *g, const double *x, int64_t n) {
        launch<false, false, false, false, true>(ctx, nullptr, nullptr, nullptr, g, x, nullptr, 0.0, n);
    }
    static void f_fd(const flgpu_eval_ctx *ctx, double *f_dev, double *g, const double *x, int64_t n) {
        launch<false, true, false, false, true>(ctx, f_dev, nullptr, nullptr, g, x, nullptr, 0.0, n);
    }
    static void fused(const flgpu_eval_ctx *ctx, int flags, double *f_dev, double *gp_dev, double *x_out, double *g_out,
                      const double *x0, const double *p, double a, int64_t n) {
        switch (flags) {
        case FLGPU_WANT_F | FLGPU_WANT_GP:
            launch<true, true, true, false, false>(ctx, f_dev, gp_dev, nullptr, nullptr, x0, p, a, n); break;
        case FLGPU_WANT_F:
            launch<true, true, false, false, false>(ctx, f_dev, nullptr, nullptr, nullptr, x0, p, a, n); break;
        case FLGPU_WANT_GP:
            launch<true, false, true, false, false>(ctx, nullptr, gp_dev, nullptr, nullptr, x0, p, a, n); break;
        case FLGPU_WRITE_X | FLGPU_WRITE_G:
            launch<true, false, false, true, true>(ctx, nullptr, nullptr, x_out, g_out, x0, p, a, n); break;
        case FLGPU_WRITE_G:
            launch<true, false, false, false, true>(ctx, nullptr, nullptr, nullptr, g_out, x0, p, a, n); break;
        case FLGPU_WRITE_X:
            launch<true, false, false, true, false>(ctx, nullptr, nullptr, x_out, nullptr, x0, p, a, n); break;
        case FLGPU_WANT_F | FLGPU_WANT_GP | FLGPU_WRITE_X | FLGPU_WRITE_G:
            launch<true, true, true, true, true>(ctx, f_dev, gp_dev, x_out, g_out, x0, p, a, n); break;
        default:
            std::fprintf(stderr, "flgpu_obj: unsupported fused evaluation request %d\n", flags);
            std::abort();
        }
    }
};

// `obj` must stay alive while the problem is in use (the callbacks read it through ctx->user).
template <class Obj>
inline flgpu_problem make_problem(const Obj *obj, bool with_f_fd = true, bool with_fused = true, bool with_search = true) {
    flgpu_problem p;
    p.f = Callbacks<Obj>::f;
    p.fd = Callbacks<Obj>::fd;
    p.f_fd = with_f_fd ? Callbacks<Obj>::f_fd : nullptr;
    p.user = (void *)obj;
    p.fused = with_fused ? Callbacks<Obj>::fused : nullptr;
    p.search = (with_fused && with_search) ? Callbacks<Obj>::search : nullptr;
    p.search_caps = 0;                        // single GPU: row-sharded runs use the host-driven search
    return p;
}

}  // namespace flgpu_obj
