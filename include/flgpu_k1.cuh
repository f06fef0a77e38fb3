// flgpu_k1.cuh -- K1 of the L-BFGS iteration (ring-buffer update + every dot product of the two-loop recursion in
// one pass), as a template over WHERE THE ACCEPTED POINT COMES FROM.
//
// libflgpu instantiates it with PlainSrc (x1 and f'(x1) are read from memory).  An objective that can evaluate its
// gradient inside a kernel instantiates it with a source that forms x1 = x0 + a*p and f'(x1) in registers and stores
// them, and RE-EVALUATES f'(x0) from the x0 it reads anyway instead of loading it (flgpu_problem.update): the line search
// then never stores its accepted point in a pass of its own, 10n -> 6n doubles of traffic per iteration.  libflgpu's built-in objectives do this in csrc/objectives.cu,
// include/flgpu_objective.cuh does it for user functors.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#include "flgpu_lbfgs_gram.hpp"
#include "flgpu_reduce.cuh"

namespace flgpu {
namespace k {

constexpr int kThreads = red::kThreads;
constexpr int kMaxGrid = 148 * 8;     // largest grid any streaming kernel is launched with
constexpr int kMaxMem = 64;           // largest LBFGS Memory supported by K1/K2/K3

// Reduction workspace shared by all library kernels of one backend / stream (stream-ordered use).
struct Work {
    double *partials;        // [rows][stride]: chunk sums, one row per accumulator
    int64_t stride;          // chunk capacity of a row
    double *blockvals;       // [rows][red::kTopMax]: roots of 4096-chunk blocks (rows with more than 4096 chunks)
    unsigned int *tickets;   // [rows], zero between kernels
};

__device__ __forceinline__ double2 ld2(const double *p, int64_t u) {
    return __ldg(reinterpret_cast<const double2 *>(p) + u);
}
__device__ __forceinline__ void st2(double *p, int64_t u, double2 v) {
    reinterpret_cast<double2 *>(p)[u] = v;   // cache-streaming stores (__stcs) measured: no difference (profiles/r01_store_policy.md)
}

// Chunk geometry of a streaming kernel: chunk c covers the 16-byte units [c*cu, min(nu, (c+1)*cu)); the odd last
// element of an odd-length shard belongs to the last chunk and is added by its thread 0 after that thread's units.
struct Chunks {
    int64_t nu, cu, nchunks;
    bool odd;
    __device__ Chunks(int64_t n, int64_t ch) : nu(n >> 1), cu(ch >> 1), nchunks(red::num_chunks(n, ch)), odd(n & 1) {}
    __device__ int64_t lo(int64_t c) const { return c * cu; }
    __device__ int64_t hi(int64_t c) const { const int64_t h = (c + 1) * cu; return h < nu ? h : nu; }
    __device__ bool tail_here(int64_t c) const { return odd && c == nchunks - 1; }
};

// ------------------------------------------------------------------ K1: ring update + all dots
// Replaces f90:609-624 (After: g.g, s=x-xold, y=g-gold, rho) and the 2k dot products of the next
// Before (f90:590-606).  Thread group ty owns MT of the k_after-1 older columns; group 0 also owns the new
// column, which it builds in registers and stores to ring slot new_slot.  The chunk sums of dot d go to row d of the
// partials (d = the index of flgpu_lbfgs_gram.hpp's layout); tree_kernel delivers row d to R[kResSlots + d].
//
// Where the accepted point comes from is a policy (Src): PlainSrc loads x1 and f'(x1) from memory; a fused source
// (objectives.cu, flgpu_objective.cuh) forms x1 = x0 + a*p and f'(x1) in registers and STORES them -- the separate
// "store the accepted point" pass of the line search disappears (10n -> 7n doubles per iteration).
struct K1Args {
    const double *x1, *x0, *g1, *g0;   // PlainSrc: x1, g1 read.  Fused: x1 = x1_out, g1 = g1_out are written by pass 1
    const double *p;                   // fused source only
    double step;                       // fused source only
    double *x1_out, *g1_out;           // fused source only
    double *S, *Y;
    int64_t ld, n, ch;
    int64_t offset, n_global;          // fused source only (index-dependent objectives)
    const double *tables;              // fused source only
    int m, new_slot, k_after;
    int age_base;       // first age handled by this pass (1 for the first pass)
    int write_new;      // 1 on the first pass: store the new column and accumulate its dots
    Work w;
};

struct PlainSrc {
    static constexpr bool kFused = false;
    __device__ void init(const K1Args &) {}
    // unit u: the accepted point, its gradient and the gradient at the previous point
    __device__ __forceinline__ void unit(const K1Args &a, int64_t u, bool, double2 x0, double2 &x1, double2 &g1,
                                         double2 &g0) const {
        (void)x0;
        g0 = ld2(a.g0, u);
        g1 = ld2(a.g1, u);
        if (a.write_new) x1 = ld2(a.x1, u);
    }
    __device__ __forceinline__ void tail(const K1Args &a, int64_t i, bool, double x0, double &x1, double &g1, double &g0) const {
        (void)x0;
        x1 = a.x1[i]; g1 = a.g1[i]; g0 = a.g0[i];
    }
};

// Thread layout: every warp is cut into NG segments of SEG = 32/NG lanes; segment ty is column group
// ty, and the lanes of all segments of a warp address the SAME SEG consecutive double2 elements.  The loads
// of x1, x0, g1, g0 that every group needs are therefore issued with identical addresses inside one warp
// instruction and coalesce into a single request (no re-read of those four vectors per group), while
// each group's column loads stay contiguous runs of SEG*16 bytes.
template <int MT, int NG, class Src>
static __global__ void __launch_bounds__(kThreads, 2) k1_update_dots_kernel(K1Args a, Src src) {
    constexpr int SEG = 32 / NG;                     // lanes per column group inside a warp
    constexpr int TX = kThreads / NG;                // double2 elements per block and loop trip
    constexpr int NW = kThreads / 32;                // every warp contributes to every group
    constexpr int NE = 4 * MT + 5;                   // sums per group (the 5 extra ones: group 0 only)
    __shared__ double sh[2][NG][NE][NW];
    src.init(a);
    const int ty = (threadIdx.x & 31) / SEG;
    const int tx = (threadIdx.x >> 5) * SEG + (threadIdx.x & 31) % SEG;
    const int m = a.m;
    const double *cs[MT], *cy[MT];
    bool valid[MT];
#pragma unroll
    for (int c = 0; c < MT; c++) {
        const int age = a.age_base + ty * MT + c;
        valid[c] = age < a.k_after;
        const int slot = slot_of_age(a.new_slot, valid[c] ? age : 0, m);
        cs[c] = a.S + (size_t)slot * a.ld;
        cy[c] = a.Y + (size_t)slot * a.ld;
    }
    const bool own_new = (ty == 0) && a.write_new;
    double *sn_col = a.S + (size_t)a.new_slot * a.ld, *yn_col = a.Y + (size_t)a.new_slot * a.ld;
    const int lane = threadIdx.x & 31, wg = threadIdx.x >> 5;
    const bool seg_head = (lane % SEG) == 0;
    const Chunks C(a.n, a.ch);
    int parity = 0;

    for (int64_t chunk = blockIdx.x; chunk < C.nchunks; chunk += gridDim.x) {
        const int64_t hi = C.hi(chunk);
        double acc[MT][4];
#pragma unroll
        for (int c = 0; c < MT; c++) acc[c][0] = acc[c][1] = acc[c][2] = acc[c][3] = 0.0;
        double ex[5] = {0.0, 0.0, 0.0, 0.0, 0.0};        // g.g, sn.g, yn.g, sn.yn, yn.yn
        for (int64_t u = C.lo(chunk) + tx; u < hi; u += TX) {
            double2 x0 = make_double2(0.0, 0.0);
            if (a.write_new) x0 = ld2(a.x0, u);          // later passes need only y_new = g1 - g0
            double2 s[MT], y[MT];
#pragma unroll
            for (int c = 0; c < MT; c++)
                if (valid[c]) { s[c] = ld2(cs[c], u); y[c] = ld2(cy[c], u); }
            double2 x1 = make_double2(0.0, 0.0), g1, g0;
            src.unit(a, u, own_new, x0, x1, g1, g0);
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
        if (C.tail_here(chunk) && tx == 0) {             // odd tail element (one thread per column group)
            const int64_t i = a.n - 1;
            const double x0 = a.write_new ? a.x0[i] : 0.0;
            double x1 = 0.0, g1, g0;
            src.tail(a, i, own_new, x0, x1, g1, g0);
            if (!a.write_new) x1 = 0.0;
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
        // ---- this chunk's sums: butterfly inside each column group's lanes, the 8 warps left to right
#pragma unroll
        for (int c = 0; c < MT; c++)
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const double v = red::seg_butterfly<SEG>(acc[c][q]);
                if (seg_head) sh[parity][ty][4 * c + q][wg] = v;
            }
#pragma unroll
        for (int q = 0; q < 5; q++) {
            const double v = red::seg_butterfly<SEG>(ex[q]);
            if (seg_head) sh[parity][ty][4 * MT + q][wg] = v;
        }
        __syncthreads();
        for (int idx = threadIdx.x; idx < NG * NE; idx += kThreads) {
            const int g = idx / NE, e = idx % NE;
            double s = 0.0;
#pragma unroll
            for (int q = 0; q < NW; q++) s += sh[parity][g][e][q];
            int d = -1;
            if (e < 4 * MT) {
                const int c = e >> 2, q = e & 3;
                const int age = a.age_base + g * MT + c;
                if (age < a.k_after) {
                    const int j = slot_of_age(a.new_slot, age, m);
                    d = q == 0 ? d_A(m, j) : q == 1 ? d_B(m, j) : q == 2 ? d_SYN(m, j) : d_YYN(m, j);
                }
            } else if (g == 0 && a.write_new) {
                const int q = e - 4 * MT, j = a.new_slot;
                d = q == 0 ? d_GG(m) : q == 1 ? d_A(m, j) : q == 2 ? d_B(m, j) : q == 3 ? d_SYN(m, j) : d_YYN(m, j);
            }
            if (d >= 0) a.w.partials[(int64_t)d * a.w.stride + chunk] = s;
        }
        parity ^= 1;
    }
}

// ---- launching one K1 pass.  The library chooses the thread shape (columns per group MT, groups per warp NG) from
// the number of columns the pass covers; the caller supplies the source.
struct K1Launch {
    K1Args a;
    int mt, ng;          // one of (2,1) (4,1) (5,1) (4,2) (5,2)
    int num_sms;
    int64_t nchunks;
    void *stream;
};

template <int MT, int NG, class Src>
inline void launch_k1_shape(const K1Launch &L, const Src &src) {
    // grid = SMs x CTAs actually resident for this instantiation (one full wave), capped by the number of chunks
    static int resident_of[64] = {0};            // per device: a process may drive several GPUs
    int dev = 0;
    cudaGetDevice(&dev);
    int &resident = resident_of[dev & 63];
    if (!resident) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, k1_update_dots_kernel<MT, NG, Src>, kThreads, 0);
        if (resident < 1) resident = 1;
    }
    int64_t grid = (int64_t)L.num_sms * resident;
    if (grid > kMaxGrid) grid = kMaxGrid;
    if (L.nchunks < grid) grid = L.nchunks < 1 ? 1 : L.nchunks;
    k1_update_dots_kernel<MT, NG, Src><<<(int)grid, kThreads, 0, (cudaStream_t)L.stream>>>(L.a, src);
}

template <class Src>
inline void launch_k1_pass(const K1Launch &L, const Src &src) {
    if (L.mt == 2 && L.ng == 1) launch_k1_shape<2, 1, Src>(L, src);
    else if (L.mt == 4 && L.ng == 1) launch_k1_shape<4, 1, Src>(L, src);
    else if (L.mt == 5 && L.ng == 1) launch_k1_shape<5, 1, Src>(L, src);
    else if (L.mt == 4 && L.ng == 2) launch_k1_shape<4, 2, Src>(L, src);
    else if (L.mt == 5 && L.ng == 2) launch_k1_shape<5, 2, Src>(L, src);
    else { std::fprintf(stderr, "flgpu: K1: unsupported (columns per group, groups) shape\n"); std::abort(); }
}

}  // namespace k
}  // namespace flgpu
