// kernels.cuh -- hand-written fp64 CUDA kernels (sm_100a) of the L-BFGS / CG hot path.
//
// Every kernel is a streaming pass bounded by HBM bandwidth (<= 0.25 flop/B): no tensor cores,
// no shared-memory tiling of the data (each element is used once), 128-bit coalesced accesses,
// grid = a multiple of the SM count, grid-stride loops with several independent 16-byte loads
// in flight per thread.  Reductions are deterministic: per-thread accumulation in a fixed
// element order, xor-butterfly warp shuffles, a fixed-order sum over warps in shared memory,
// one partial per block in global memory, and a fixed-order final sum performed by the last
// block to finish (ticket counter) -- the value never depends on which block that is.
//
// Element-wise results that the reference defines with separate multiply and add roundings
// (x0+a*p f90:1482, p-alpha*y f90:592, -g+beta*p f90:366) use __dmul_rn/__dadd_rn so they do
// not depend on FMA contraction; accumulations into reduction registers use FMA.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "backend.hpp"
#include "lbfgs_gram.hpp"

namespace flgpu {
namespace k {

constexpr int kThreads = 256;
constexpr int kMaxGrid = 148 * 8;     // partial buffers are sized for this many blocks
constexpr int kMaxMem = 64;           // largest LBFGS Memory supported by K1/K2/K3
constexpr int kResSlots = NSLOTS;     // R[0..16) = slots, R[16..) = K1 dots

// Reduction workspace shared by all library kernels of one backend (stream-ordered use).
struct Work {
    double *partials;        // [kMaxGrid][stride]
    unsigned int *ticket;    // zero between kernels
};

__device__ __forceinline__ double2 ld2(const double *p, int64_t u) {
    return __ldg(reinterpret_cast<const double2 *>(p) + u);
}
__device__ __forceinline__ void st2(double *p, int64_t u, double2 v) {
    reinterpret_cast<double2 *>(p)[u] = v;   // cache-streaming stores (__stcs) measured: no difference (profiles/r01_store_policy.md)
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// butterfly over aligned segments of SEG lanes (SEG = 32: the whole warp)
template <int SEG>
__device__ __forceinline__ double seg_sum(double v) {
#pragma unroll
    for (int o = SEG / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-level reduction of NACC per-thread accumulators followed by the grid-level finish.
// dest[i] = index into R receiving accumulator i.  All kThreads threads must call this.
template <int NACC>
__device__ __forceinline__ void reduce_finish(double (&acc)[NACC], const int (&dest)[NACC], Work w,
                                              double *R) {
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
        w.partials[(size_t)blockIdx.x * NACC + threadIdx.x] = s;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(w.ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // final: warp q handles accumulators q, q+8, ...; lanes stride over blocks, then butterfly
    for (int i = warp; i < NACC; i += kThreads / 32) {
        double s = 0.0;
        for (unsigned b = lane; b < gridDim.x; b += 32) s += __ldcg(&w.partials[(size_t)b * NACC + i]);
        s = warp_sum(s);
        if (lane == 0) R[dest[i]] = s;
    }
    if (threadIdx.x == 0) *w.ticket = 0u;
}

// Same reduction with one output pointer per accumulator (objective kernels: f and f'.p go to
// caller-supplied device scalars).
template <int NACC>
__device__ __forceinline__ void reduce_finish_to(double (&acc)[NACC], double *(&out)[NACC], Work w) {
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
        w.partials[(size_t)blockIdx.x * NACC + threadIdx.x] = s;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(w.ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (warp < NACC) {
        double s = 0.0;
        for (unsigned b = lane; b < gridDim.x; b += 32) s += __ldcg(&w.partials[(size_t)b * NACC + warp]);
        s = warp_sum(s);
        if (lane == 0) *out[warp] = s;
    }
    if (threadIdx.x == 0) *w.ticket = 0u;
}

// ------------------------------------------------------------------ K4a: x = x0 + a*p (f90:1482)
static __global__ void __launch_bounds__(kThreads) trial_kernel(double *__restrict__ x, const double *__restrict__ x0,
                                                         const double *__restrict__ p, double a, int64_t n) {
    const int64_t nu = n >> 1;
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    int64_t u = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    for (; u + 3 * stride < nu; u += 4 * stride) {
        double2 a0 = ld2(x0, u), a1 = ld2(x0, u + stride), a2 = ld2(x0, u + 2 * stride), a3 = ld2(x0, u + 3 * stride);
        double2 b0 = ld2(p, u), b1 = ld2(p, u + stride), b2 = ld2(p, u + 2 * stride), b3 = ld2(p, u + 3 * stride);
        st2(x, u, make_double2(__dadd_rn(a0.x, __dmul_rn(a, b0.x)), __dadd_rn(a0.y, __dmul_rn(a, b0.y))));
        st2(x, u + stride, make_double2(__dadd_rn(a1.x, __dmul_rn(a, b1.x)), __dadd_rn(a1.y, __dmul_rn(a, b1.y))));
        st2(x, u + 2 * stride, make_double2(__dadd_rn(a2.x, __dmul_rn(a, b2.x)), __dadd_rn(a2.y, __dmul_rn(a, b2.y))));
        st2(x, u + 3 * stride, make_double2(__dadd_rn(a3.x, __dmul_rn(a, b3.x)), __dadd_rn(a3.y, __dmul_rn(a, b3.y))));
    }
    for (; u < nu; u += stride) {
        double2 a0 = ld2(x0, u), b0 = ld2(p, u);
        st2(x, u, make_double2(__dadd_rn(a0.x, __dmul_rn(a, b0.x)), __dadd_rn(a0.y, __dmul_rn(a, b0.y))));
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) x[n - 1] = __dadd_rn(x0[n - 1], __dmul_rn(a, p[n - 1]));
}

// p = -g (f90:442, 369)
static __global__ void __launch_bounds__(kThreads) neg_kernel(double *__restrict__ p, const double *__restrict__ g, int64_t n) {
    const int64_t nu = n >> 1;
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    for (int64_t u = (int64_t)blockIdx.x * kThreads + threadIdx.x; u < nu; u += stride) {
        double2 v = ld2(g, u);
        st2(p, u, make_double2(-v.x, -v.y));
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) p[n - 1] = -g[n - 1];
}

// ------------------------------------------------------------------ K4b: dot_product(a,b) (f90:1485)
static __global__ void __launch_bounds__(kThreads) dot_kernel(const double *__restrict__ a, const double *__restrict__ b,
                                                       int64_t n, Work w, double *R, int dest) {
    const int64_t nu = n >> 1;
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    int64_t u = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    for (; u + 3 * stride < nu; u += 4 * stride) {
        double2 a0 = ld2(a, u), a1 = ld2(a, u + stride), a2 = ld2(a, u + 2 * stride), a3 = ld2(a, u + 3 * stride);
        double2 b0 = ld2(b, u), b1 = ld2(b, u + stride), b2 = ld2(b, u + 2 * stride), b3 = ld2(b, u + 3 * stride);
        s0 = fma(a0.y, b0.y, fma(a0.x, b0.x, s0));
        s1 = fma(a1.y, b1.y, fma(a1.x, b1.x, s1));
        s2 = fma(a2.y, b2.y, fma(a2.x, b2.x, s2));
        s3 = fma(a3.y, b3.y, fma(a3.x, b3.x, s3));
    }
    for (; u < nu; u += stride) {
        double2 a0 = ld2(a, u), b0 = ld2(b, u);
        s0 = fma(a0.y, b0.y, fma(a0.x, b0.x, s0));
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) s0 = fma(a[n - 1], b[n - 1], s0);
    double acc[1] = {(s0 + s1) + (s2 + s3)};
    const int d[1] = {dest};
    reduce_finish<1>(acc, d, w, R);
}

// ------------------------------------------------------------------ K5a: CG dots (f90:354-366, 375-387)
// one pass over f'new, f'old, p:  g.g, p.p, (g-gold).p, g.(g-gold), gold.gold
static __global__ void __launch_bounds__(kThreads) cg_dots_kernel(const double *__restrict__ g1, const double *__restrict__ g0,
                                                           const double *__restrict__ p, int64_t n, Work w, double *R) {
    const int64_t nu = n >> 1;
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    auto term = [&](double a, double b, double c) {
        const double d = a - b;                 // fdnew-fdold, rounded as in the reference
        acc[0] = fma(a, a, acc[0]);
        acc[1] = fma(c, c, acc[1]);
        acc[2] = fma(d, c, acc[2]);
        acc[3] = fma(a, d, acc[3]);
        acc[4] = fma(b, b, acc[4]);
    };
    int64_t u = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    for (; u + stride < nu; u += 2 * stride) {
        double2 a0 = ld2(g1, u), a1 = ld2(g1, u + stride);
        double2 b0 = ld2(g0, u), b1 = ld2(g0, u + stride);
        double2 c0 = ld2(p, u), c1 = ld2(p, u + stride);
        term(a0.x, b0.x, c0.x); term(a0.y, b0.y, c0.y);
        term(a1.x, b1.x, c1.x); term(a1.y, b1.y, c1.y);
    }
    for (; u < nu; u += stride) {
        double2 a0 = ld2(g1, u), b0 = ld2(g0, u), c0 = ld2(p, u);
        term(a0.x, b0.x, c0.x); term(a0.y, b0.y, c0.y);
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) term(g1[n - 1], g0[n - 1], p[n - 1]);
    const int d[5] = {SL_GG, SL_PP, SL_DGP, SL_GDG, SL_G0G0};
    reduce_finish<5>(acc, d, w, R);
}

// ------------------------------------------------------------------ K5b: p = -g + beta*p; g.p (f90:366-367)
static __global__ void __launch_bounds__(kThreads) cg_update_kernel(double *__restrict__ p, const double *__restrict__ g1,
                                                             double beta, int64_t n, Work w, double *R) {
    const int64_t nu = n >> 1;
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    double acc[1] = {0.0};
    int64_t u = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    for (; u + stride < nu; u += 2 * stride) {
        double2 a0 = ld2(g1, u), a1 = ld2(g1, u + stride);
        double2 c0 = reinterpret_cast<const double2 *>(p)[u], c1 = reinterpret_cast<const double2 *>(p)[u + stride];
        c0.x = __dadd_rn(-a0.x, __dmul_rn(beta, c0.x)); c0.y = __dadd_rn(-a0.y, __dmul_rn(beta, c0.y));
        c1.x = __dadd_rn(-a1.x, __dmul_rn(beta, c1.x)); c1.y = __dadd_rn(-a1.y, __dmul_rn(beta, c1.y));
        st2(p, u, c0); st2(p, u + stride, c1);
        acc[0] = fma(a0.y, c0.y, fma(a0.x, c0.x, acc[0]));
        acc[0] = fma(a1.y, c1.y, fma(a1.x, c1.x, acc[0]));
    }
    for (; u < nu; u += stride) {
        double2 a0 = ld2(g1, u);
        double2 c0 = reinterpret_cast<const double2 *>(p)[u];
        c0.x = __dadd_rn(-a0.x, __dmul_rn(beta, c0.x)); c0.y = __dadd_rn(-a0.y, __dmul_rn(beta, c0.y));
        st2(p, u, c0);
        acc[0] = fma(a0.y, c0.y, fma(a0.x, c0.x, acc[0]));
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const double v = __dadd_rn(-g1[n - 1], __dmul_rn(beta, p[n - 1]));
        p[n - 1] = v;
        acc[0] = fma(g1[n - 1], v, acc[0]);
    }
    const int d[1] = {SL_GP0};
    reduce_finish<1>(acc, d, w, R);
}

// ------------------------------------------------------------------ K1: ring update + all dots
// Replaces f90:609-624 (After: g.g, s=x-xold, y=g-gold, rho) and the 2k dot products of the next
// Before (f90:590-606).  Block = TX x NG threads: thread row tx walks the vector, thread group ty
// owns MT of the k_after-1 older columns; group 0 also owns the new column, which it builds in
// registers from x1,x0,g1,g0 and stores to ring slot new_slot.
struct K1Args {
    const double *x1, *x0, *g1, *g0;
    double *S, *Y;
    int64_t ld, n;
    int m, new_slot, k_after;
    int age_base;       // first age handled by this pass (1 for the first pass)
    int write_new;      // 1 on the first pass: store the new column and accumulate its dots
    Work w;
    double *R;          // dots land at R[kResSlots + d_*]
};

// Thread layout: every warp is cut into NG segments of SEG = 32/NG lanes; segment ty is column group
// ty, and the lanes of all segments of a warp address the SAME SEG consecutive double2 elements.  The loads
// of x1, x0, g1, g0 that every group needs are therefore issued with identical addresses inside one warp
// instruction and coalesce into a single request (no re-read of those four vectors per group), while
// each group's column loads stay contiguous runs of SEG*16 bytes.
template <int MT, int NG>
static __global__ void __launch_bounds__(kThreads) k1_update_dots_kernel(K1Args a) {
    constexpr int SEG = 32 / NG;                     // lanes per column group inside a warp
    constexpr int TX = kThreads / NG;                // double2 elements per block and loop trip
    constexpr int NW = kThreads / 32;                // every warp contributes to every group
    const int ty = (threadIdx.x & 31) / SEG;
    const int tx = (threadIdx.x >> 5) * SEG + (threadIdx.x & 31) % SEG;
    const int m = a.m;
    const double *cs[MT], *cy[MT];
    bool valid[MT];
    int slot[MT];
#pragma unroll
    for (int c = 0; c < MT; c++) {
        const int age = a.age_base + ty * MT + c;
        valid[c] = age < a.k_after;
        slot[c] = slot_of_age(a.new_slot, valid[c] ? age : 0, m);
        cs[c] = a.S + (size_t)slot[c] * a.ld;
        cy[c] = a.Y + (size_t)slot[c] * a.ld;
    }
    double acc[MT][4];
#pragma unroll
    for (int c = 0; c < MT; c++) acc[c][0] = acc[c][1] = acc[c][2] = acc[c][3] = 0.0;
    double ex[5] = {0.0, 0.0, 0.0, 0.0, 0.0};        // g.g, sn.g, yn.g, sn.yn, yn.yn
    const bool own_new = (ty == 0) && a.write_new;
    double *sn_col = a.S + (size_t)a.new_slot * a.ld, *yn_col = a.Y + (size_t)a.new_slot * a.ld;

    const int64_t nu = a.n >> 1;
    const int64_t stride = (int64_t)gridDim.x * TX;
    for (int64_t u = (int64_t)blockIdx.x * TX + tx; u < nu; u += stride) {
        const double2 g1 = ld2(a.g1, u), g0 = ld2(a.g0, u);
        double2 x1 = make_double2(0.0, 0.0), x0 = x1;
        if (a.write_new) { x1 = ld2(a.x1, u); x0 = ld2(a.x0, u); }   // later passes need only y_new = g1 - g0
        double2 s[MT], y[MT];
#pragma unroll
        for (int c = 0; c < MT; c++)
            if (valid[c]) { s[c] = ld2(cs[c], u); y[c] = ld2(cy[c], u); }
        const double2 sn = make_double2(x1.x - x0.x, x1.y - x0.y);   // s=x-xold  f90:623
        const double2 yn = make_double2(g1.x - g0.x, g1.y - g0.y);   // y=fdnew-fdold
        if (own_new) {
            st2(sn_col, u, sn); st2(yn_col, u, yn);
            ex[0] = fma(g1.y, g1.y, fma(g1.x, g1.x, ex[0]));
            ex[1] = fma(sn.y, g1.y, fma(sn.x, g1.x, ex[1]));
            ex[2] = fma(yn.y, g1.y, fma(yn.x, g1.x, ex[2]));
            ex[3] = fma(sn.y, yn.y, fma(sn.x, yn.x, ex[3]));
            ex[4] = fma(yn.y, yn.y, fma(yn.x, yn.x, ex[4]));
        }
#pragma unroll
        for (int c = 0; c < MT; c++)
            if (valid[c]) {
                acc[c][0] = fma(s[c].y, g1.y, fma(s[c].x, g1.x, acc[c][0]));
                acc[c][1] = fma(y[c].y, g1.y, fma(y[c].x, g1.x, acc[c][1]));
                acc[c][2] = fma(s[c].y, yn.y, fma(s[c].x, yn.x, acc[c][2]));
                acc[c][3] = fma(y[c].y, yn.y, fma(y[c].x, yn.x, acc[c][3]));
            }
    }
    if ((a.n & 1) && blockIdx.x == 0 && tx == 0) {   // odd tail element (one thread per column group)
        const int64_t i = a.n - 1;
        const double x1 = a.x1[i], x0 = a.x0[i], g1 = a.g1[i], g0 = a.g0[i];
        const double sn = x1 - x0, yn = g1 - g0;
        if (own_new) {
            sn_col[i] = sn; yn_col[i] = yn;
            ex[0] = fma(g1, g1, ex[0]); ex[1] = fma(sn, g1, ex[1]); ex[2] = fma(yn, g1, ex[2]);
            ex[3] = fma(sn, yn, ex[3]); ex[4] = fma(yn, yn, ex[4]);
        }
#pragma unroll
        for (int c = 0; c < MT; c++)
            if (valid[c]) {
                const double s = cs[c][i], y = cy[c][i];
                acc[c][0] = fma(s, g1, acc[c][0]); acc[c][1] = fma(y, g1, acc[c][1]);
                acc[c][2] = fma(s, yn, acc[c][2]); acc[c][3] = fma(y, yn, acc[c][3]);
            }
    }

    // ---- block reduction: within each column group, then one partial per (block, dot)
    const int nd = nd_of(m);
    __shared__ double sh[NG][4 * MT + 5][NW];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, wg = threadIdx.x >> 5;
    const bool seg_head = (lane % SEG) == 0;
#pragma unroll
    for (int c = 0; c < MT; c++)
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const double v = seg_sum<SEG>(acc[c][q]);
            if (seg_head) sh[ty][4 * c + q][wg] = v;
        }
#pragma unroll
    for (int q = 0; q < 5; q++) {
        const double v = seg_sum<SEG>(ex[q]);
        if (seg_head) sh[ty][4 * MT + q][wg] = v;
    }
    __syncthreads();
    double *part = a.w.partials + (size_t)blockIdx.x * nd;
    for (int idx = threadIdx.x; idx < NG * (4 * MT + 5); idx += kThreads) {
        const int g = idx / (4 * MT + 5), e = idx % (4 * MT + 5);
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < NW; q++) s += sh[g][e][q];
        if (e < 4 * MT) {
            const int c = e >> 2, q = e & 3;
            const int age = a.age_base + g * MT + c;
            if (age < a.k_after) {
                const int j = slot_of_age(a.new_slot, age, m);
                const int d = q == 0 ? d_A(m, j) : q == 1 ? d_B(m, j) : q == 2 ? d_SYN(m, j) : d_YYN(m, j);
                part[d] = s;
            }
        } else if (g == 0 && a.write_new) {
            const int q = e - 4 * MT, j = a.new_slot;
            const int d = q == 0 ? d_GG(m) : q == 1 ? d_A(m, j) : q == 2 ? d_B(m, j) : q == 3 ? d_SYN(m, j) : d_YYN(m, j);
            part[d] = s;
        }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(a.w.ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // final fixed-order sum over blocks for the dots this pass produced
    const int warp = threadIdx.x >> 5;
    for (int d = warp; d < nd; d += kThreads / 32) {
        // which dots belong to this pass: column j with age in [age_base, age_base+NG*MT) or the new column
        int j, is_col = 1;
        if (d == d_GG(m)) { is_col = 0; j = a.new_slot; }
        else j = d < 2 * m ? d % m : (d - 2 * m - 1) % m;
        const int age = (a.new_slot - j + m) % m;
        bool mine;
        if (age == 0) mine = a.write_new != 0;
        else mine = age >= a.age_base && age < a.age_base + NG * MT && age < a.k_after;
        (void)is_col;
        if (!mine) continue;
        double s = 0.0;
        for (unsigned b = lane; b < gridDim.x; b += 32) s += __ldcg(&a.w.partials[(size_t)b * nd + d]);
        s = warp_sum(s);
        if (lane == 0) {
            a.R[kResSlots + d] = s;
            if (d == d_GG(m)) a.R[SL_GG] = s;
        }
    }
    if (threadIdx.x == 0) *a.w.ticket = 0u;
}

// ------------------------------------------------------------------ K2: Gram-space two-loop, one warp
// Lane-parallel form of lbfgs_gram_solve(): lane L owns ring slots L and L+32.  Dall = dots of all
// ranks ([G][nd], rank-major) summed here in rank order, or the local dots when G == 1.
static __global__ void __launch_bounds__(32) k2_solve_kernel(int m, int k, int recent, const double *Dall, int G,
                                                      double *SY, double *YY, double *C) {
    extern __shared__ double smem[];
    const int nd = nd_of(m);
    double *sD = smem;                 // nd
    double *sSY = sD + nd;             // m*m
    double *sYY = sSY + m * m;         // m*m
    const int lane = threadIdx.x;
    for (int i = lane; i < nd; i += 32) {
        double s = Dall[i];
        for (int r = 1; r < G; r++) s += Dall[(size_t)r * nd + i];
        sD[i] = s;
    }
    __syncwarp();
    const int r = recent;
    // ages and validity of the (up to two) slots this lane owns
    int jj[2], age[2];
    bool ok[2];
#pragma unroll
    for (int c = 0; c < 2; c++) {
        jj[c] = lane + 32 * c;
        age[c] = jj[c] < m ? (recent - jj[c] + m) % m : 1 << 30;
        ok[c] = jj[c] < m && age[c] < k;
    }
    // load the persistent Gram blocks, then insert the newest pair's row/column (smem and global)
    for (int i = lane; i < m * m; i += 32) { sSY[i] = SY[i]; sYY[i] = YY[i]; }
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 2; c++)
        if (ok[c]) {
            const int j = jj[c];
            const double vs = sD[d_SYN(m, j)], vy = sD[d_YYN(m, j)];
            sSY[j * m + r] = vs; SY[j * m + r] = vs;
            sYY[j * m + r] = vy; YY[j * m + r] = vy;
            sYY[r * m + j] = vy; YY[r * m + j] = vy;
        }
    __syncwarp();
    double sq[2], yq[2], al[2] = {0.0, 0.0}, ee[2] = {0.0, 0.0};
#pragma unroll
    for (int c = 0; c < 2; c++) {
        sq[c] = ok[c] ? sD[d_A(m, jj[c])] : 0.0;
        yq[c] = ok[c] ? sD[d_B(m, jj[c])] : 0.0;
    }
    // first loop, newest -> oldest
    for (int t = 0; t < k; t++) {
        const int i = slot_of_age(recent, t, m);
        const double sqi = __shfl_sync(0xffffffffu, (i >> 5) ? sq[1] : sq[0], i & 31);
        const double rho = 1.0 / sSY[i * m + i];
        const double alpha = __dmul_rn(rho, sqi);
#pragma unroll
        for (int c = 0; c < 2; c++)
            if (ok[c]) {
                const int j = jj[c];
                if (age[c] > t) sq[c] = __dadd_rn(sq[c], -__dmul_rn(alpha, sSY[j * m + i]));
                yq[c] = __dadd_rn(yq[c], -__dmul_rn(alpha, sYY[j * m + i]));
                if (j == i) al[c] = alpha;
            }
    }
    const double rho_r = 1.0 / sSY[r * m + r];
    const double gamma = 1.0 / rho_r / sYY[r * m + r];
#pragma unroll
    for (int c = 0; c < 2; c++) yq[c] = __dmul_rn(gamma, yq[c]);
    // second loop, oldest -> newest
    for (int t = k - 1; t >= 0; t--) {
        const int i = slot_of_age(recent, t, m);
        const double yri = __shfl_sync(0xffffffffu, (i >> 5) ? yq[1] : yq[0], i & 31);
        const double ali = __shfl_sync(0xffffffffu, (i >> 5) ? al[1] : al[0], i & 31);
        const double rho = 1.0 / sSY[i * m + i];
        const double beta = __dmul_rn(rho, yri);
        const double e = __dadd_rn(ali, -beta);
#pragma unroll
        for (int c = 0; c < 2; c++)
            if (ok[c]) {
                const int j = jj[c];
                if (age[c] < t) yq[c] = __dadd_rn(yq[c], __dmul_rn(e, sSY[i * m + j]));
                if (j == i) ee[c] = e;
            }
    }
#pragma unroll
    for (int c = 0; c < 2; c++)
        if (jj[c] < m) {
            C[1 + jj[c]] = ok[c] ? al[c] : 0.0;
            C[1 + m + jj[c]] = ok[c] ? ee[c] : 0.0;
        }
    if (lane == 0) C[0] = gamma;
}

// ------------------------------------------------------------------ K3: direction + first trial point
// p = -( gamma (g - sum_newest..oldest alpha_i y_i) + sum_oldest..newest e_i s_i )   (f90:589-607)
// xt = x1 + p (the a=1 trial of the next line search, f90:607+1482), g.p and p.p reduced.
struct K3Args {
    double *p, *xt;
    const double *g1, *x1, *S, *Y, *C;
    int64_t ld, n;
    int m, k, recent;
    Work w;
    double *R;
};

// The 2k column operations are one list in the reference's order -- y_newest..y_oldest (q -= alpha y),
// the gamma scaling, s_oldest..s_newest (r += e s) -- walked in chunks of CH columns with two register
// buffers: the loads of chunk c+1 are in flight while chunk c is applied, for any k with a fixed
// register budget.  -(alpha*y) == (-alpha)*y exactly, so both phases are v = v + coef*col.
template <int CH>
static __global__ void __launch_bounds__(kThreads, 2) k3_direction_kernel(K3Args a) {
    __shared__ double coef[2 * kMaxMem];
    __shared__ const double *col[2 * kMaxMem];
    __shared__ double s_gamma;
    const int m = a.m, k = a.k;
    for (int t = threadIdx.x; t < k; t += kThreads) {
        const int j = slot_of_age(a.recent, t, m);
        coef[t] = -a.C[1 + j];                       // op t        : y of age t
        col[t] = a.Y + (size_t)j * a.ld;
        coef[2 * k - 1 - t] = a.C[1 + m + j];        // op 2k-1-t   : s of age t
        col[2 * k - 1 - t] = a.S + (size_t)j * a.ld;
    }
    if (threadIdx.x == 0) s_gamma = a.C[0];
    __syncthreads();
    const double gamma = s_gamma;
    const int nops = 2 * k;
    const int nchunks = (nops + CH - 1) / CH;
    double acc[2] = {0.0, 0.0};
    const int64_t nu = a.n >> 1;
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    for (int64_t u = (int64_t)blockIdx.x * kThreads + threadIdx.x; u < nu; u += stride) {
        double2 bufA[CH], bufB[CH];
        auto load = [&](double2 (&b)[CH], int c) {
#pragma unroll
            for (int i = 0; i < CH; i++) {
                const int o = c * CH + i;
                if (o < nops) b[i] = ld2(col[o], u);
            }
        };
        double2 v = ld2(a.g1, u);
        const double2 g = v;
        load(bufA, 0);
        double2 x = make_double2(0.0, 0.0);
        if (a.xt) x = ld2(a.x1, u);
        auto apply = [&](const double2 (&b)[CH], int c) {
#pragma unroll
            for (int i = 0; i < CH; i++) {
                const int o = c * CH + i;
                if (o < nops) {
                    if (o == k) { v.x = __dmul_rn(gamma, v.x); v.y = __dmul_rn(gamma, v.y); }
                    const double cf = coef[o];
                    v.x = __dadd_rn(v.x, __dmul_rn(cf, b[i].x));
                    v.y = __dadd_rn(v.y, __dmul_rn(cf, b[i].y));
                }
            }
        };
        for (int c = 0; c < nchunks; c += 2) {
            if (c + 1 < nchunks) load(bufB, c + 1);
            apply(bufA, c);
            if (c + 2 < nchunks) load(bufA, c + 2);
            if (c + 1 < nchunks) apply(bufB, c + 1);
        }
        const double2 pv = make_double2(-v.x, -v.y);
        st2(a.p, u, pv);
        if (a.xt) st2(a.xt, u, make_double2(x.x + pv.x, x.y + pv.y));
        acc[0] = fma(g.y, pv.y, fma(g.x, pv.x, acc[0]));
        acc[1] = fma(pv.y, pv.y, fma(pv.x, pv.x, acc[1]));
    }
    if ((a.n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const int64_t i = a.n - 1;
        const double g = a.g1[i];
        double v = g;
        for (int o = 0; o < nops; o++) {
            if (o == k) v = __dmul_rn(gamma, v);
            v = __dadd_rn(v, __dmul_rn(coef[o], col[o][i]));
        }
        const double pv = -v;
        a.p[i] = pv;
        if (a.xt) a.xt[i] = a.x1[i] + pv;
        acc[0] = fma(g, pv, acc[0]);
        acc[1] = fma(pv, pv, acc[1]);
    }
    const int d[2] = {SL_GP0, SL_PP};
    reduce_finish<2>(acc, d, a.w, a.R);
}

// ------------------------------------------------------------------ C1: rank exchange over peer memory
// Row-sharded runs combine each reduction's per-rank partial sums.  Instead of a library all-gather
// followed by a combine kernel, ONE single-block kernel per exchange does both over NVLink/NVSwitch
// peer memory: every rank stores its `count` partials straight into slot [me] of every peer's mailbox
// (plain st.global on IPC-mapped peer pointers), publishes a sequence number, waits until the G slots of
// its OWN mailbox carry this sequence number, and sums them in rank order -- identical bits on all ranks.
// Mailboxes are double-buffered on the sequence parity: a peer can be at most one exchange ahead (it
// needs this rank's flag of exchange seq+1 before it can start seq+2), so two buffers suffice.
constexpr int kMailWidth = 320;       // doubles per rank slot: >= NSLOTS + nd_of(kMaxMem)
constexpr int kMaxRanks = 16;

struct Mailbox {
    double data[2][kMaxRanks][kMailWidth];
    unsigned long long flag[2][kMaxRanks];
    unsigned long long error;         // set to the offending sequence number on a wait timeout
};

struct PeerTable { Mailbox *box[kMaxRanks]; };

// One IPC-shared allocation per rank: the mailbox of the host-driven exchanges (exchange_kernel), a second one for
// the exchanges a device-resident line search performs on its own, and that search's sequence counter (the host
// cannot know how many evaluations a search will make, so the counter lives on the device; every rank makes the
// same evaluations, so the counters agree without communication).
struct MailboxPair {
    Mailbox host_driven;
    Mailbox device_search;
    unsigned long long dseq;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Executed by ONE block: store vals[0..count) (count <= blockDim.x) into slot [me] of every rank's mailbox, publish
// `seq`, wait for every rank's slot of this rank's mailbox, return the rank-ordered sums in out[0..count).
__device__ __forceinline__ void mailbox_exchange_block(const PeerTable &peers, int me, int G, unsigned long long seq,
                                                       const double *vals, int count, double *out) {
    const int par = (int)(seq & 1ull), t = threadIdx.x;
    if (t < count) {
        const double v = vals[t];
        for (int r = 0; r < G; r++) peers.box[r]->data[par][me][t] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (t < G) st_release_sys(&peers.box[t]->flag[par][me], seq);
    Mailbox *mine = peers.box[me];
    if (t < G) {
        const unsigned long long t0 = global_timer_ns();
        while (ld_acquire_sys(&mine->flag[par][t]) < seq) {
            if (global_timer_ns() - t0 > 20000000000ull) {
                mine->error = seq;
                __threadfence_system();
                __trap();
            }
        }
    }
    __syncthreads();
    if (t < count) {
        double s = __ldcv(&mine->data[par][0][t]);
        for (int r = 1; r < G; r++) s += __ldcv(&mine->data[par][r][t]);
        out[t] = s;
    }
}


// src: this rank's `count` partial sums; out: their rank-ordered sum (count <= kMailWidth); host_out (optional):
// the same sums stored straight into pinned host memory followed by the flag word host_out[NSLOTS] = seq_host.
static __global__ void __launch_bounds__(kMailWidth) exchange_kernel(PeerTable peers, int me, int G,
                                                                       unsigned long long seq, const double *src,
                                                                       int count, double *out, double *host_out,
                                                                       unsigned long long seq_host, const double *extra) {
    __shared__ double vals[kMailWidth], sums[kMailWidth];
    const int t = threadIdx.x;
    if (t < count) vals[t] = src[t];
    __syncthreads();
    mailbox_exchange_block(peers, me, G, seq, vals, count, sums);
    if (t < count) {
        out[t] = sums[t];
        if (host_out) host_out[t] = sums[t];
    }
    if (host_out) {                    // publish: data, system-wide fence, then the flag the host polls
        if (extra && t < 8) host_out[NSLOTS + 8 + t] = extra[t];      // device-search scalars (identical on all ranks)
        __threadfence_system();
        __syncthreads();
        if (t == 0) *reinterpret_cast<volatile unsigned long long *>(host_out + NSLOTS) = seq_host;
    }
}

// Single GPU: the 16 result slots go to pinned host memory the same way (data, fence, flag), so the host polls a
// cache line instead of paying a DMA copy plus a driver synchronisation per line-search trial.
static __global__ void __launch_bounds__(32) publish_kernel(const double *src, double *host_out, unsigned long long seq_host,
                                                            const double *extra) {
    if (threadIdx.x < NSLOTS) host_out[threadIdx.x] = src[threadIdx.x];
    if (extra && threadIdx.x < 8) host_out[NSLOTS + 8 + threadIdx.x] = extra[threadIdx.x];   // device-search scalars
    __threadfence_system();
    __syncwarp();
    if (threadIdx.x == 0) *reinterpret_cast<volatile unsigned long long *>(host_out + NSLOTS) = seq_host;
}

// ------------------------------------------------------------------ multi-rank combine of the slots
static __global__ void combine_kernel(const double *all, int G, int count, double *out) {
    const int i = threadIdx.x;
    if (i < count) {
        double s = all[i];
        for (int r = 1; r < G; r++) s += all[(size_t)r * count + i];
        out[i] = s;
    }
}

static __global__ void set_scalar_kernel(double *dst, double v) { *dst = v; }
static __global__ void add_scalar_kernel(double *dst, double v) { *dst = __dadd_rn(*dst, v); }

}  // namespace k
}  // namespace flgpu
