// flgpu_objective.cuh -- write an objective once as a device functor, get all four libflgpu callbacks.
//
// The reference asks the user for f, fd and optionally f_fd (NonlinearOptimization.f90:33-43).  On the GPU those are
// kernels, and the line search is fastest when the objective kernel also forms the trial point x0 + a*p itself
// (flgpu_fused_fn, flgpu.h).  For objectives that are sums of terms over single elements or over pairs
// (x_{2j}, x_{2j+1}) -- the quartic of the reference's own tests, extended Rosenbrock, diagonal quadratics, any
// separable loss -- this header generates f, fd, f_fd AND the fused evaluation from one functor, with the library's
// deterministic, partition-independent reduction (include/flgpu_reduce.cuh: fixed chunks, aligned binary tree).
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
#include "flgpu_exchange.cuh"
#include "flgpu_k1.cuh"
#include "flgpu_k3.cuh"
#include "flgpu_reduce.cuh"
#include "flgpu_search_core.hpp"

namespace flgpu_obj {

namespace red = flgpu::red;
constexpr int kThreads = red::kThreads;

struct Args {
    const double *x;      // the point, or x0 when FUSED
    const double *p;      // FUSED
    double a;             // FUSED
    double *x_out, *g_out;
    int64_t n, offset, ch;   // local rows, global index of row 0, chunk elements (flgpu_chunk_elems(n_global))
    double *partials;        // chunk sums: f -> row 0 (or the only row), f'.p -> the next row
    int64_t stride;
};

// One CHUNK of an evaluation (flgpu_reduce.cuh): this thread's 16-byte units of chunk c in order (shared by
// objective_kernel and search_kernel: same chunk sums, same bits)
template <class Obj, bool FUSED, bool WANT_F, bool WANT_GP, bool WRITE_X, bool WRITE_G>
__device__ __forceinline__ void chunk(const Obj &obj, const Args &a, int64_t c, int64_t nchunks, double &fsum, double &gpsum) {
    const int64_t nu = a.n >> 1, cu = a.ch >> 1;
    const int64_t lo = c * cu, hi = (lo + cu < nu) ? lo + cu : nu;
    for (int64_t u = lo + threadIdx.x; u < hi; u += kThreads) {
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
    if ((a.n & 1) && c == nchunks - 1 && threadIdx.x == 0) {   // last element of an odd-length shard
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
    const int64_t nchunks = red::num_chunks(a.n, a.ch);
    int parity = 0;
    for (int64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
        double fsum = 0.0, gpsum = 0.0;
        chunk<Obj, FUSED, WANT_F, WANT_GP, WRITE_X, WRITE_G>(obj, a, c, nchunks, fsum, gpsum);
        if (WANT_F && WANT_GP) {
            const double acc[2] = {fsum, gpsum};
            red::chunk_flush<2>(acc, parity, a.partials, a.stride, c);
        } else if (WANT_F) {
            const double acc[1] = {fsum};
            red::chunk_flush<1>(acc, parity, a.partials, a.stride, c);
        } else if (WANT_GP) {
            const double acc[1] = {gpsum};
            red::chunk_flush<1>(acc, parity, a.partials, a.stride, c);
        }
    }
}

// ---- batched fused evaluation (flgpu_fused_multi_fn): f and f'.p at x0 + steps[j]*p for FLGPU_MULTI_MAX steps in one pass
// over x0 and p.  Per step: chunk<Obj, true, true, true, false, false>'s arithmetic and accumulation order, hence its bits.
struct MultiArgs {
    const double *x, *p;
    double steps[FLGPU_MULTI_MAX];
    int64_t n, offset, ch;
    double *partials;        // chunk sums: f_j -> row 2j, (f'.p)_j -> row 2j+1
    int64_t stride;
};
template <class Obj>
__global__ void __launch_bounds__(kThreads, 3) objective_multi_kernel(Obj obj, MultiArgs a) {
    constexpr int J = FLGPU_MULTI_MAX;
    const int64_t nchunks = red::num_chunks(a.n, a.ch);
    const int64_t nu = a.n >> 1, cu = a.ch >> 1;
    int parity = 0;
    for (int64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
        double acc[2 * J];
#pragma unroll
        for (int i = 0; i < 2 * J; i++) acc[i] = 0.0;
        const int64_t lo = c * cu, hi = (lo + cu < nu) ? lo + cu : nu;
        for (int64_t u = lo + threadIdx.x; u < hi; u += kThreads) {
            const double2 x0 = __ldg(reinterpret_cast<const double2 *>(a.x) + u);
            const double2 pv = __ldg(reinterpret_cast<const double2 *>(a.p) + u);
            const int64_t i = a.offset + 2 * u;
#pragma unroll
            for (int j = 0; j < J; j++) {
                const double xa = __dadd_rn(x0.x, __dmul_rn(a.steps[j], pv.x)), xb = __dadd_rn(x0.y, __dmul_rn(a.steps[j], pv.y));
                double2 g;
                if constexpr (Obj::WIDTH == 2) {
                    double f = 0.0;
                    obj.eval2(i, xa, xb, f, g.x, g.y);
                    acc[2 * j] += f;
                } else {
                    double f0 = 0.0, f1 = 0.0;
                    obj.eval(i, xa, f0, g.x);
                    obj.eval(i + 1, xb, f1, g.y);
                    acc[2 * j] += f0; acc[2 * j] += f1;
                }
                acc[2 * j + 1] = fma(g.y, pv.y, fma(g.x, pv.x, acc[2 * j + 1]));
            }
        }
        if ((a.n & 1) && c == nchunks - 1 && threadIdx.x == 0) {
            const int64_t k = a.n - 1;
            const double x0 = a.x[k], pv = a.p[k];
#pragma unroll
            for (int j = 0; j < J; j++) {
                double f = 0.0, g = 0.0;
                const double x = __dadd_rn(x0, __dmul_rn(a.steps[j], pv));
                if constexpr (Obj::WIDTH == 2) obj.eval_tail(a.offset + k, x, f, g);
                else obj.eval(a.offset + k, x, f, g);
                acc[2 * j] += f;
                acc[2 * j + 1] = fma(g, pv, acc[2 * j + 1]);
            }
        }
        red::chunk_flush<2 * J>(acc, parity, a.partials, a.stride, c);
    }
}

// ---- device-resident line search (flgpu_search_fn) for functor objectives: the whole Wolfe / Strong-Wolfe search in
// one cooperative kernel.  Every thread runs the same SearchCore state machine (flgpu_search_core.hpp, the source the
// host driver compiles) on the same values; an evaluation = this block's chunks + one grid barrier + the tree over the
// chunk sums, repeated by every block.  Same chunk sums and tree as objective_kernel + flgpu_reduce_tree => same bits
// as the host-driven fused search => same decisions.  On row shards block 0 trades the rank's roots with the other ranks
// inside the kernel (flgpu_exchange.cuh: stores into the peers' search mailboxes, flags, rank tree) and a second grid
// barrier hands the sums to the other blocks -- as libflgpu's own search kernel does.
struct SearchArgs {
    Args o;                                  // x = x0; x_out / g_out = accepted point / gradient; partials rows [2*parity + i]
    double c1, c2abs, fx0, phid0, incr, a0;
    int strong, fdwithf, store;
    double *result;
    flgpu::k::SearchExchange ex;             // ex.G <= 1: single GPU
};

constexpr double kEvalBudget = 100000.0;

template <class Obj>
struct DevSearch : flgpu::SearchCore<DevSearch<Obj>> {
    const Obj &obj;
    const SearchArgs &K;
    double *rsh;                             // red::kWarps + red::kTopMax doubles of block scratch
    double f_cur = 0.0, gp_cur = 0.0, a_x = 0.0, a_g = 0.0;
    bool have_x = false, have_g = false;
    int parity = 0, fpar = 0;
    double trials = 0.0, n_f = 0.0, n_fd = 0.0, n_ffd = 0.0, n_fonly = 0.0;
    unsigned long long seq_base = 0, nexch = 0;   // exchanges made so far (uniform over the grid)
    __device__ DevSearch(const Obj &o, const SearchArgs &k, double *s) : obj(o), K(k), rsh(s) {
        if (K.ex.G > 1) seq_base = *K.ex.dseq;
    }
    template <bool F, bool GP>
    __device__ void eval() {
        constexpr int NACC = (F && GP) ? 2 : 1;
        Args o = K.o;
        o.a = a_x;
        const int64_t nchunks = red::num_chunks(o.n, o.ch);
        double *rows = o.partials + (int64_t)(2 * parity) * o.stride;
        for (int64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
            double fsum = 0.0, gpsum = 0.0;
            chunk<Obj, true, F, GP, false, false>(obj, o, c, nchunks, fsum, gpsum);
            if (NACC == 2) {
                const double acc[2] = {fsum, gpsum};
                red::chunk_flush<2>(acc, fpar, rows, o.stride, c);
            } else {
                const double acc[1] = {F ? fsum : gpsum};
                red::chunk_flush<1>(acc, fpar, rows, o.stride, c);
            }
        }
        __threadfence();
        cooperative_groups::this_grid().sync();
        double v[2] = {0.0, 0.0};
#pragma unroll
        for (int i = 0; i < NACC; i++) {
            v[i] = red::cta_root(rows + (int64_t)i * o.stride, nchunks, rsh);
            __syncthreads();
        }
        if (K.ex.G > 1) {
            nexch++;
            double *gl = K.ex.glob + parity * 2;
            if (blockIdx.x == 0) {
                __shared__ double mine[2], summed[2];
                if (threadIdx.x < NACC) mine[threadIdx.x] = v[threadIdx.x];
                __syncthreads();
                flgpu::k::mailbox_exchange_block(K.ex.peers, K.ex.me, K.ex.G, seq_base + nexch, mine, NACC, summed,
                                                 K.ex.timeout_ns);
                if (threadIdx.x < NACC) gl[threadIdx.x] = summed[threadIdx.x];
                __threadfence();
            }
            cooperative_groups::this_grid().sync();
            if (F) f_cur = __ldcg(&gl[0]);
            if (GP) gp_cur = __ldcg(&gl[NACC - 1]);
        } else {
            if (F) f_cur = v[0];
            if (GP) gp_cur = v[NACC - 1];
        }
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
__global__ void __launch_bounds__(kThreads, 3) search_kernel(Obj obj, SearchArgs K) {
    __shared__ double rsh[red::kWarps + red::kTopMax];
    DevSearch<Obj> S(obj, K, rsh);
    S.c1 = K.c1; S.c2abs = K.c2abs; S.fx0 = K.fx0; S.phid0 = K.phid0; S.incr = K.incr;
    S.fdwithf = K.fdwithf != 0; S.a = K.a0; S.f_cur = K.fx0; S.pre = 0;
    if (FAST) S.fast(K.strong != 0);
    else if (K.strong) S.strongwolfe(); else S.wolfe();
    Args o = K.o;
    const int64_t nchunks = red::num_chunks(o.n, o.ch);
    double f0 = 0.0, g0 = 0.0;
    if (K.store) {
        for (int64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
            if (S.have_x && S.have_g && S.a_x == S.a_g) {
                o.a = S.a_x;
                chunk<Obj, true, false, false, true, true>(obj, o, c, nchunks, f0, g0);
            } else {                                  // never taken by the reference's searchers; kept for fidelity
                if (S.have_x) { o.a = S.a_x; chunk<Obj, true, false, false, true, false>(obj, o, c, nchunks, f0, g0); }
                if (S.have_g) { o.a = S.a_g; chunk<Obj, true, false, false, false, true>(obj, o, c, nchunks, f0, g0); }
            }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        K.result[0] = S.a; K.result[1] = S.f_cur; K.result[2] = S.trials; K.result[3] = S.n_f;
        K.result[4] = S.n_fd; K.result[5] = S.n_ffd; K.result[6] = S.n_fonly;
        K.result[7] = S.aborted() ? -1.0 : (double)S.nexch;
        if (K.ex.G > 1) *K.ex.dseq = S.seq_base + S.nexch;
    }
}

// K1 source for functor objectives (flgpu_problem.update, flgpu_k1.cuh): x1 = x0 + a*p and f'(x1) formed in registers
template <class Obj>
struct FunctorSrc {
    Obj obj;
    __device__ void init(const flgpu::k::K1Args &) {}
    // f'(x0) is read back from memory here: a user objective may be expensive to evaluate twice (libflgpu's built-in
    // objectives re-evaluate it from x0 instead, which saves the load)
    __device__ __forceinline__ void unit(const flgpu::k::K1Args &a, int64_t u, bool own_new, double2 x0, double2 &x1,
                                         double2 &g1, double2 &g0) const {
        const double2 pv = __ldg(reinterpret_cast<const double2 *>(a.p) + u);
        g0 = __ldg(reinterpret_cast<const double2 *>(a.g0) + u);
        x1.x = __dadd_rn(x0.x, __dmul_rn(a.step, pv.x));
        x1.y = __dadd_rn(x0.y, __dmul_rn(a.step, pv.y));
        const int64_t i = a.offset + 2 * u;
        double f = 0.0;
        if constexpr (Obj::WIDTH == 2) obj.eval2(i, x1.x, x1.y, f, g1.x, g1.y);
        else { obj.eval(i, x1.x, f, g1.x); obj.eval(i + 1, x1.y, f, g1.y); }
        if (own_new) {
            reinterpret_cast<double2 *>(a.x1_out)[u] = x1;
            reinterpret_cast<double2 *>(a.g1_out)[u] = g1;
        }
    }
    __device__ __forceinline__ void tail(const flgpu::k::K1Args &a, int64_t i, bool own_new, double x0, double &x1,
                                         double &g1, double &g0) const {
        g0 = a.g0[i];
        x1 = __dadd_rn(x0, __dmul_rn(a.step, a.p[i]));
        double f = 0.0;
        if constexpr (Obj::WIDTH == 2) obj.eval_tail(a.offset + i, x1, f, g1);
        else obj.eval(a.offset + i, x1, f, g1);
        if (own_new) { a.x1_out[i] = x1; a.g1_out[i] = g1; }
    }
};

// K3 probe for functor objectives (flgpu_problem.direction, flgpu_k3.cuh): f and f'.p at the a = 1 trial point that K3
// forms in registers, accumulated exactly as chunk<Obj, true, true, WANT_GP, ...> does for the same units
template <class Obj>
struct FunctorProbe {
    static constexpr bool kOn = true;
    static constexpr int kSteps = flgpu::k::kProbeSteps;
    Obj obj;
    int64_t offset;
    __device__ void init(const flgpu::k::K3Args &, int) {}
    __device__ __forceinline__ void unit(const flgpu::k::K3Args &a, int64_t u, const double2 x1, const double2 pv,
                                         double *acc) const {
        const int64_t i = offset + 2 * u;
#pragma unroll
        for (int j = 0; j < kSteps; j++) {
            const double xa = __dadd_rn(x1.x, __dmul_rn(a.steps[j], pv.x)), xb = __dadd_rn(x1.y, __dmul_rn(a.steps[j], pv.y));
            double2 g;
            if constexpr (Obj::WIDTH == 2) {
                double f = 0.0;
                obj.eval2(i, xa, xb, f, g.x, g.y);
                acc[2 * j] += f;
            } else {
                double f0 = 0.0, f1 = 0.0;
                obj.eval(i, xa, f0, g.x);
                obj.eval(i + 1, xb, f1, g.y);
                acc[2 * j] += f0; acc[2 * j] += f1;
            }
            acc[2 * j + 1] = fma(g.y, pv.y, fma(g.x, pv.x, acc[2 * j + 1]));
        }
    }
    __device__ __forceinline__ void tail(const flgpu::k::K3Args &a, int64_t k, const double x1, const double pv, double *acc) const {
#pragma unroll
        for (int j = 0; j < kSteps; j++) {
            double f = 0.0, g = 0.0;
            const double x = __dadd_rn(x1, __dmul_rn(a.steps[j], pv));
            if constexpr (Obj::WIDTH == 2) obj.eval_tail(offset + k, x, f, g);
            else obj.eval(offset + k, x, f, g);
            acc[2 * j] += f;
            acc[2 * j + 1] = fma(g, pv, acc[2 * j + 1]);
        }
    }
};

template <class Obj>
struct Callbacks {
    static void geometry(const flgpu_eval_ctx *ctx, int64_t n, Args &A, int64_t &nchunks) {
        const int64_t n_global = ctx->n_global < n ? n : ctx->n_global;
        A.ch = flgpu_chunk_elems(n_global);
        nchunks = n > 0 ? (n + A.ch - 1) / A.ch : 1;
        flgpu_reduction_workspace(ctx->stream, nchunks, &A.partials, &A.stride);   // per-stream, library-owned
        A.n = n; A.offset = ctx->offset;
    }
    template <bool FUSED, bool F, bool GP, bool WX, bool WG>
    static void launch(const flgpu_eval_ctx *ctx, double *f_dev, double *gp_dev, double *x_out, double *g_out,
                       const double *x, const double *p, double a, int64_t n) {
        cudaStream_t s = (cudaStream_t)ctx->stream;
        Args A;
        int64_t nchunks = 1;
        geometry(ctx, n, A, nchunks);
        A.x = x; A.p = p; A.a = a; A.x_out = x_out; A.g_out = g_out;
        objective_kernel<Obj, FUSED, F, GP, WX, WG><<<grid_for(nchunks), kThreads, 0, s>>>(*(const Obj *)ctx->user, A);
        if (F || GP) {                     // chunk sums -> this rank's roots (the library's tree kernel)
            double *out[2] = {F ? f_dev : gp_dev, gp_dev};
            flgpu_reduce_tree(ctx->stream, nchunks, (F && GP) ? 2 : 1, out);
        }
    }
    static int grid_for(int64_t nchunks, int per_sm = 4) {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        int64_t grid = (int64_t)sms * per_sm;                                         // one full wave of resident CTAs
        if (nchunks < grid) grid = nchunks < 1 ? 1 : nchunks;
        return (int)grid;
    }
    static void search(const flgpu_eval_ctx *ctx, const flgpu_search_args *A, int64_t n) {
        if (A->policy == FLGPU_LS_FAST) search_policy<true>(ctx, A, n); else search_policy<false>(ctx, A, n);
    }
    template <bool FAST>
    static void search_policy(const flgpu_eval_ctx *ctx, const flgpu_search_args *A, int64_t n) {
        SearchArgs K;
        K.ex.G = 1; K.ex.me = 0; K.ex.dseq = nullptr; K.ex.glob = nullptr; K.ex.timeout_ns = 0;
        if (A->comm && flgpu_comm_search_exchange(A->comm, ctx->stream, &K.ex, sizeof K.ex) != 0) {
            std::fprintf(stderr, "flgpu_obj: the device-resident search on row shards needs the peer-memory exchange\n");
            std::abort();
        }
        int64_t nchunks = 1;
        geometry(ctx, n, K.o, nchunks);
        K.o.x = A->x0_dev; K.o.p = A->p_dev; K.o.a = 0.0; K.o.x_out = A->x_out; K.o.g_out = A->g_out;
        K.c1 = A->c1; K.c2abs = A->c2abs; K.fx0 = A->fx0; K.phid0 = A->phid0; K.incr = A->incr; K.a0 = A->a;
        K.strong = A->strong; K.fdwithf = A->fdwithf; K.store = A->no_store ? 0 : 1; K.result = A->result_dev;
        int resident = 0, sms = 148, dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, search_kernel<Obj, FAST>, kThreads, 0);
        int grid = grid_for(nchunks);
        if (grid > resident * sms) grid = resident * sms;
        Obj obj = *(const Obj *)ctx->user;
        void *params[] = {&obj, &K};
        cudaError_t e = cudaLaunchCooperativeKernel((void *)search_kernel<Obj, FAST>, dim3(grid), dim3(kThreads), params, 0,
                                                    (cudaStream_t)ctx->stream);
        if (e != cudaSuccess) { std::fprintf(stderr, "flgpu_obj: cooperative launch failed: %s\n", cudaGetErrorString(e)); std::abort(); }
    }
    static void update(const flgpu_eval_ctx *ctx, const flgpu_update_args *A, int64_t) {
        if (A->k1_bytes != sizeof(flgpu::k::K1Launch)) {
            std::fprintf(stderr, "flgpu_obj: K1Launch layout mismatch (header and libflgpu.so versions differ)\n");
            std::abort();
        }
        FunctorSrc<Obj> src{*(const Obj *)ctx->user};
        flgpu::k::launch_k1_pass(*(const flgpu::k::K1Launch *)A->k1, src);
    }
    static void fused_multi(const flgpu_eval_ctx *ctx, int count, const double *steps, double *out_dev, const double *x0,
                            const double *p, int64_t n) {
        if (count < 1 || count > FLGPU_MULTI_MAX) { std::fprintf(stderr, "flgpu_obj: batched evaluation of 1 to 4 steps\n"); std::abort(); }
        Args G;
        int64_t nchunks = 1;
        geometry(ctx, n, G, nchunks);
        MultiArgs A;
        A.x = x0; A.p = p; A.n = n; A.offset = G.offset; A.ch = G.ch; A.partials = G.partials; A.stride = G.stride;
        for (int j = 0; j < FLGPU_MULTI_MAX; j++) A.steps[j] = steps[j < count ? j : count - 1];
        objective_multi_kernel<Obj><<<grid_for(nchunks, 3), kThreads, 0, (cudaStream_t)ctx->stream>>>(*(const Obj *)ctx->user, A);
        double *out[2 * FLGPU_MULTI_MAX];
        for (int i = 0; i < 2 * FLGPU_MULTI_MAX; i++) out[i] = out_dev + i;
        flgpu_reduce_tree(ctx->stream, nchunks, 2 * count, out);
    }
    static void direction(const flgpu_eval_ctx *ctx, const flgpu_direction_args *A, int64_t) {
        if (A->k3_bytes != sizeof(flgpu::k::K3Launch)) {
            std::fprintf(stderr, "flgpu_obj: K3Launch layout mismatch (header and libflgpu.so versions differ)\n");
            std::abort();
        }
        FunctorProbe<Obj> probe{*(const Obj *)ctx->user, ctx->offset};
        flgpu::k::launch_k3_probe(*(const flgpu::k::K3Launch *)A->k3, probe);
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
    p.search_caps = FLGPU_SEARCH_ROW_SHARDS;  // the search kernel trades its sums with the other ranks itself
    p.update = with_fused ? Callbacks<Obj>::update : nullptr;
    p.direction = with_fused ? Callbacks<Obj>::direction : nullptr;
    p.fused_multi = with_fused ? Callbacks<Obj>::fused_multi : nullptr;
    return p;
}

}  // namespace flgpu_obj
