// kernels.cuh -- hand-written fp64 CUDA kernels (sm_100a) of the L-BFGS / CG hot path.
//
// Every kernel is a streaming pass bounded by HBM bandwidth (<= 0.25 flop/B): no tensor cores, 128-bit coalesced
// accesses, grid = a multiple of the SM count, several independent 16-byte loads in flight per thread (K3 stages its
// columns through a shared-memory ring filled by bulk async copies instead, see there).
//
// Reductions are deterministic AND partition-independent (include/flgpu_reduce.cuh): the vector is cut into chunks of
// CH = chunk_elems(n_global) elements; a thread block sums one chunk at a time in a fixed order and stores the chunk's
// sums; tree_kernel combines the chunk sums by the aligned binary tree over the chunk index and delivers the rank's
// root; ranks are combined by the same tree over the rank index.  Which block sums which chunk never matters.
//
// Element-wise results that the reference defines with separate multiply and add roundings
// (x0+a*p f90:1482, p-alpha*y f90:592, -g+beta*p f90:366) use __dmul_rn/__dadd_rn so they do
// not depend on FMA contraction; accumulations into reduction registers use FMA.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/flgpu_k1.cuh"
#include "../../include/flgpu_k3.cuh"
#include "../../include/flgpu_exchange.cuh"
#include "../../include/flgpu_reduce.cuh"
#include "backend.hpp"

namespace flgpu {
namespace k {

constexpr int kResSlots = NSLOTS;     // R[0..16) = slots, R[16..) = K1 dots

// ------------------------------------------------------------------ tree: chunk sums -> this rank's root
// grid (blocks of 4096 chunks, rows).  Row r of the partials is reduced to one value which goes to out[r] (rows <= 12),
// or to lin_out[r] (K1's dots); one row may be delivered to a second place (K1: g.g is also a result slot).
struct TreeArgs {
    Work w;
    int64_t nchunks;
    double *out[12];
    double *lin_out;
    int dup_row;
    double *dup_out;
};
static __global__ void __launch_bounds__(kThreads) tree_kernel(TreeArgs a) {
    __shared__ double sh[red::kWarps];
    __shared__ bool last;
    const int row = blockIdx.y, b = blockIdx.x, nblk = gridDim.x;
    const int64_t lo = (int64_t)b * red::kBlockChunks;
    const int64_t rem = a.nchunks - lo;
    double r = red::cta_tree(a.w.partials + (int64_t)row * a.w.stride + lo,
                             (int)(rem < red::kBlockChunks ? rem : red::kBlockChunks), sh);
    if (nblk > 1) {
        if (threadIdx.x == 0) {
            a.w.blockvals[row * red::kTopMax + b] = r;
            __threadfence();
            last = atomicAdd(&a.w.tickets[row], 1u) == (unsigned)nblk - 1;
        }
        __syncthreads();
        if (!last) return;
        __threadfence();
        r = red::top_tree<true>(a.w.blockvals + row * red::kTopMax, nblk);
        if (threadIdx.x == 0) a.w.tickets[row] = 0u;
    }
    if (threadIdx.x == 0) {
        double *dst = a.lin_out ? a.lin_out + row : a.out[row];
        if (dst) *dst = r;
        if (row == a.dup_row && a.dup_out) *a.dup_out = r;
    }
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
// chunk sums of a.b go to row `row` of the partials
// out != null (and at most 4096 chunks): the last block also forms the tree and stores the root (no tree_kernel launch)
static __global__ void __launch_bounds__(kThreads) dot_kernel(const double *__restrict__ a, const double *__restrict__ b,
                                                       int64_t n, int64_t ch, Work w, int row, double *out) {
    const Chunks C(n, ch);
    double *part = w.partials + (int64_t)row * w.stride;
    int parity = 0;
    for (int64_t c = blockIdx.x; c < C.nchunks; c += gridDim.x) {
        const int64_t hi = C.hi(c);
        double s = 0.0;
        int64_t u = C.lo(c) + threadIdx.x;
        for (; u + 3 * kThreads < hi; u += 4 * kThreads) {      // four units per trip, all loads issued first
            double2 a0 = ld2(a, u), a1 = ld2(a, u + kThreads), a2 = ld2(a, u + 2 * kThreads), a3 = ld2(a, u + 3 * kThreads);
            double2 b0 = ld2(b, u), b1 = ld2(b, u + kThreads), b2 = ld2(b, u + 2 * kThreads), b3 = ld2(b, u + 3 * kThreads);
            s = fma(a0.y, b0.y, fma(a0.x, b0.x, s));
            s = fma(a1.y, b1.y, fma(a1.x, b1.x, s));
            s = fma(a2.y, b2.y, fma(a2.x, b2.x, s));
            s = fma(a3.y, b3.y, fma(a3.x, b3.x, s));
        }
        for (; u < hi; u += kThreads) {
            double2 a0 = ld2(a, u), b0 = ld2(b, u);
            s = fma(a0.y, b0.y, fma(a0.x, b0.x, s));
        }
        if (C.tail_here(c) && threadIdx.x == 0) s = fma(a[n - 1], b[n - 1], s);
        const double acc[1] = {s};
        red::chunk_flush<1>(acc, parity, part, w.stride, c);
    }
    if (out) {
        double *const o[1] = {out};
        red::finish_in_kernel<1>(part, w.stride, C.nchunks, w.tickets, o);
    }
}

// ------------------------------------------------------------------ K5a: CG dots (f90:354-366, 375-387)
// one pass over f'new, f'old, p:  g.g, p.p, (g-gold).p, g.(g-gold), gold.gold  -> rows 0..4
struct Outs5 { double *p[5]; };   // destinations of an in-kernel finish (null p[0]: tree_kernel follows instead)
static __global__ void __launch_bounds__(kThreads) cg_dots_kernel(const double *__restrict__ g1, const double *__restrict__ g0,
                                                           const double *__restrict__ p, int64_t n, int64_t ch, Work w,
                                                           Outs5 outs) {
    const Chunks C(n, ch);
    int parity = 0;
    for (int64_t c = blockIdx.x; c < C.nchunks; c += gridDim.x) {
        const int64_t hi = C.hi(c);
        double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
        auto term = [&](double a, double b, double q) {
            const double d = a - b;                 // fdnew-fdold, rounded as in the reference
            acc[0] = fma(a, a, acc[0]);
            acc[1] = fma(q, q, acc[1]);
            acc[2] = fma(d, q, acc[2]);
            acc[3] = fma(a, d, acc[3]);
            acc[4] = fma(b, b, acc[4]);
        };
        int64_t u = C.lo(c) + threadIdx.x;
        for (; u + kThreads < hi; u += 2 * kThreads) {
            double2 a0 = ld2(g1, u), a1 = ld2(g1, u + kThreads);
            double2 b0 = ld2(g0, u), b1 = ld2(g0, u + kThreads);
            double2 c0 = ld2(p, u), c1 = ld2(p, u + kThreads);
            term(a0.x, b0.x, c0.x); term(a0.y, b0.y, c0.y);
            term(a1.x, b1.x, c1.x); term(a1.y, b1.y, c1.y);
        }
        for (; u < hi; u += kThreads) {
            double2 a0 = ld2(g1, u), b0 = ld2(g0, u), c0 = ld2(p, u);
            term(a0.x, b0.x, c0.x); term(a0.y, b0.y, c0.y);
        }
        if (C.tail_here(c) && threadIdx.x == 0) term(g1[n - 1], g0[n - 1], p[n - 1]);
        red::chunk_flush<5>(acc, parity, w.partials, w.stride, c);
    }
    if (outs.p[0]) {
        double *const o[5] = {outs.p[0], outs.p[1], outs.p[2], outs.p[3], outs.p[4]};
        red::finish_in_kernel<5>(w.partials, w.stride, C.nchunks, w.tickets, o);
    }
}

// ------------------------------------------------------------------ K5b: p = -g + beta*p; g.p (f90:366-367) -> row 0
static __global__ void __launch_bounds__(kThreads) cg_update_kernel(double *__restrict__ p, const double *__restrict__ g1,
                                                             double beta, int64_t n, int64_t ch, Work w, double *out) {
    const Chunks C(n, ch);
    int parity = 0;
    for (int64_t c = blockIdx.x; c < C.nchunks; c += gridDim.x) {
        const int64_t hi = C.hi(c);
        double acc[1] = {0.0};
        int64_t u = C.lo(c) + threadIdx.x;
        for (; u + kThreads < hi; u += 2 * kThreads) {
            double2 a0 = ld2(g1, u), a1 = ld2(g1, u + kThreads);
            double2 c0 = reinterpret_cast<const double2 *>(p)[u], c1 = reinterpret_cast<const double2 *>(p)[u + kThreads];
            c0.x = __dadd_rn(-a0.x, __dmul_rn(beta, c0.x)); c0.y = __dadd_rn(-a0.y, __dmul_rn(beta, c0.y));
            c1.x = __dadd_rn(-a1.x, __dmul_rn(beta, c1.x)); c1.y = __dadd_rn(-a1.y, __dmul_rn(beta, c1.y));
            st2(p, u, c0); st2(p, u + kThreads, c1);
            acc[0] = fma(a0.y, c0.y, fma(a0.x, c0.x, acc[0]));
            acc[0] = fma(a1.y, c1.y, fma(a1.x, c1.x, acc[0]));
        }
        for (; u < hi; u += kThreads) {
            double2 a0 = ld2(g1, u);
            double2 c0 = reinterpret_cast<const double2 *>(p)[u];
            c0.x = __dadd_rn(-a0.x, __dmul_rn(beta, c0.x)); c0.y = __dadd_rn(-a0.y, __dmul_rn(beta, c0.y));
            st2(p, u, c0);
            acc[0] = fma(a0.y, c0.y, fma(a0.x, c0.x, acc[0]));
        }
        if (C.tail_here(c) && threadIdx.x == 0) {
            const double v = __dadd_rn(-g1[n - 1], __dmul_rn(beta, p[n - 1]));
            p[n - 1] = v;
            acc[0] = fma(g1[n - 1], v, acc[0]);
        }
        red::chunk_flush<1>(acc, parity, w.partials, w.stride, c);
    }
    if (out) {
        double *const o[1] = {out};
        red::finish_in_kernel<1>(w.partials, w.stride, C.nchunks, w.tickets, o);
    }
}

// ------------------------------------------------------------------ K2: Gram-space two-loop, one warp
// Lane-parallel form of lbfgs_gram_solve(): lane L owns ring slots L and L+32.  Dall = dots of all
// ranks ([G][nd], rank-major) combined here by the rank tree, or the local dots when G == 1.
static __global__ void __launch_bounds__(32) k2_solve_kernel(int m, int k, int recent, const double *Dall, int G,
                                                      double *SY, double *YY, double *C) {
    extern __shared__ double smem[];
    const int nd = nd_of(m);
    double *sD = smem;                 // nd
    double *sSY = sD + nd;             // m*m
    double *sYY = sSY + m * m;         // m*m
    const int lane = threadIdx.x;
    for (int i = lane; i < nd; i += 32) sD[i] = G > 1 ? red::rank_tree(Dall + i, G, nd) : Dall[i];
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

// K3 (direction, optionally the first trial evaluation of the next line search): include/flgpu_k3.cuh

// ------------------------------------------------------------------ C1: rank exchange over peer memory
// Row-sharded runs combine each reduction's per-rank roots.  Instead of a library all-gather
// followed by a combine kernel, ONE single-block kernel per exchange does both over NVLink/NVSwitch
// peer memory: every rank stores its `count` values straight into slot [me] of every peer's mailbox
// (plain st.global on IPC-mapped peer pointers), publishes a sequence number, waits until the G slots of
// its OWN mailbox carry this sequence number, and combines them by the rank tree (red::rank_tree) -- identical bits
// on all ranks, and the bits of the single-GPU tree when the shards are aligned subtrees (flgpu_reduce.cuh).
// Mailboxes are double-buffered on the sequence parity: a peer can be at most one exchange ahead (it
// needs this rank's flag of exchange seq+1 before it can start seq+2), so two buffers suffice.
// (Mailbox, PeerTable, mailbox_exchange_block: include/flgpu_exchange.cuh -- shared with include/flgpu_objective.cuh)

// src: this rank's `count` values; out: their rank-tree sum (count <= kMailWidth); host_out (optional):
// the same sums stored straight into pinned host memory followed by the flag word host_out[NSLOTS] = seq_host.
static __global__ void __launch_bounds__(kMailWidth) exchange_kernel(PeerTable peers, int me, int G,
                                                                       unsigned long long seq, const double *src,
                                                                       int count, double *out, double *host_out,
                                                                       unsigned long long seq_host, const double *extra,
                                                                       unsigned long long timeout_ns) {
    __shared__ double vals[kMailWidth], sums[kMailWidth];
    const int t = threadIdx.x;
    if (t < count) vals[t] = src[t];
    __syncthreads();
    mailbox_exchange_block(peers, me, G, seq, vals, count, sums, timeout_ns);
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

// ------------------------------------------------------------------ multi-rank combine of the slots (NCCL fallback)
static __global__ void combine_kernel(const double *all, int G, int count, double *out) {
    const int i = threadIdx.x;
    if (i < count) out[i] = red::rank_tree(all + i, G, count);
}

static __global__ void set_scalar_kernel(double *dst, double v) { *dst = v; }
static __global__ void add_scalar_kernel(double *dst, double v) { *dst = __dadd_rn(*dst, v); }

}  // namespace k
}  // namespace flgpu
