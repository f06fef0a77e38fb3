// flgpu_k3.cuh -- K3 of the L-BFGS iteration (the new direction from the coefficients of K2), as a template over an
// optional PROBE: the first trial of the next line search (a = 1, f90:607 + 1482-1485) evaluated inside the same pass.
//
// libflgpu instantiates it with NoProbe (p only; with plain callbacks also the trial point x1 + p).  An objective that
// can evaluate f and f' inside a kernel instantiates it with a probe (flgpu_problem.direction): K3 then also loads x1,
// forms x1 + a*p in registers for the first steps of the search's bracketing walk -- a = 1 (1*p is exact: this IS the
// fused evaluation's x0 + a*p), Increment, Increment^2, Increment^3 (f90:1499-1501) -- and reduces f and f'.p at each in
// the chunk order every other kernel uses: the bits separate probe launches would deliver, for 1n instead of 2n doubles
// per evaluated trial and no launch at all.  libflgpu's built-in objectives do this
// in csrc/objectives.cu, include/flgpu_objective.cuh does it for user functors.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#include "flgpu_k1.cuh"
#include "flgpu_lbfgs_gram.hpp"
#include "flgpu_reduce.cuh"

namespace flgpu {
namespace k {

// ------------------------------------------------------------------ K3: direction + first trial point
// p = -( gamma (g - sum_newest..oldest alpha_i y_i) + sum_oldest..newest e_i s_i )   (f90:589-607)
// xt = x1 + p (the a=1 trial of the next line search, f90:607+1482); chunk sums of g.p -> row 0, p.p -> row 1.
// With a probe of S steps: x1 is read, rows 2+2j and 3+2j of the chunk sums carry f and f'.p at x1 + steps[j]*p.
constexpr int kProbeSteps = 4;
struct K3Args {
    double *p, *xt;
    const double *g1, *x1, *S, *Y, *C;
    int64_t ld, n, ch;
    int m, k, recent;
    int64_t offset, n_global;   // probe only (index-dependent objectives)
    double steps[kProbeSteps];  // probe only: 1, Increment, Increment^2, ... as the host will form them
    Work w;
};

// No first-trial evaluation: what libflgpu launches on its own.
// A probe accumulates, for j < kSteps, f(x + steps[j]*p) into acc[2j] and f'(x + steps[j]*p).p into acc[2j+1] (x + a*p:
// multiply, then add) -- per unit, in the order the fused evaluation uses.
struct NoProbe {
    static constexpr bool kOn = false;
    static constexpr int kSteps = 0;
    __device__ void init(const K3Args &, int) {}
    __device__ __forceinline__ void unit(const K3Args &, int64_t, double2, double2, double *) const {}
    __device__ __forceinline__ void tail(const K3Args &, int64_t, double, double, double *) const {}
};

// The 2k column operations are one list in the reference's order -- y_newest..y_oldest (q -= alpha y),
// the gamma scaling, s_oldest..s_newest (r += e s); -(alpha*y) == (-alpha)*y exactly, so both phases are
// v = v + coef*col.  Coefficient / column tables in shared memory, built by every block.
struct K3Tables {
    double coef[2 * kMaxMem];
    const double *col[2 * kMaxMem];
    double gamma;
};
__device__ __forceinline__ void k3_build_tables(const K3Args &a, K3Tables &T, int nthreads) {
    const int m = a.m, k = a.k;
    for (int t = threadIdx.x; t < k; t += nthreads) {
        const int j = slot_of_age(a.recent, t, m);
        T.coef[t] = -a.C[1 + j];                       // op t        : y of age t
        T.col[t] = a.Y + (size_t)j * a.ld;
        T.coef[2 * k - 1 - t] = a.C[1 + m + j];        // op 2k-1-t   : s of age t
        T.col[2 * k - 1 - t] = a.S + (size_t)j * a.ld;
    }
    if (threadIdx.x == 0) T.gamma = a.C[0];
}
template <int NACC, class Probe>
__device__ __forceinline__ void k3_tail(const K3Args &a, const K3Tables &T, const Probe &probe, double (&acc)[NACC]) {
    const int64_t i = a.n - 1;
    const int nops = 2 * a.k;
    const double g = a.g1[i];
    double v = g;
    for (int o = 0; o < nops; o++) {
        if (o == a.k) v = __dmul_rn(T.gamma, v);
        v = __dadd_rn(v, __dmul_rn(T.coef[o], T.col[o][i]));
    }
    const double pv = -v;
    a.p[i] = pv;
    if (a.xt) a.xt[i] = a.x1[i] + pv;
    acc[0] = fma(g, pv, acc[0]);
    acc[1] = fma(pv, pv, acc[1]);
    if (Probe::kOn) probe.tail(a, i, a.x1[i], pv, &acc[2]);
}

// ---- K3, register version (FLGPU_K3=regs; the r01 kernel on the chunked reduction): the column list is walked in
// chunks of CH columns with two register buffers, the loads of group c+1 in flight while group c is applied.
template <int CH>
static __global__ void __launch_bounds__(kThreads, 2) k3_direction_kernel(K3Args a) {
    __shared__ K3Tables T;
    k3_build_tables(a, T, kThreads);
    __syncthreads();
    const int k = a.k;
    const double gamma = T.gamma;
    const int nops = 2 * k;
    const int ngroups = (nops + CH - 1) / CH;
    const Chunks C(a.n, a.ch);
    int parity = 0;
    for (int64_t chunk = blockIdx.x; chunk < C.nchunks; chunk += gridDim.x) {
        const int64_t hi = C.hi(chunk);
        double acc[2] = {0.0, 0.0};
        for (int64_t u = C.lo(chunk) + threadIdx.x; u < hi; u += kThreads) {
            double2 bufA[CH], bufB[CH];
            auto load = [&](double2 (&b)[CH], int c) {
#pragma unroll
                for (int i = 0; i < CH; i++) {
                    const int o = c * CH + i;
                    if (o < nops) b[i] = ld2(T.col[o], u);
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
                        const double cf = T.coef[o];
                        v.x = __dadd_rn(v.x, __dmul_rn(cf, b[i].x));
                        v.y = __dadd_rn(v.y, __dmul_rn(cf, b[i].y));
                    }
                }
            };
            for (int c = 0; c < ngroups; c += 2) {
                if (c + 1 < ngroups) load(bufB, c + 1);
                apply(bufA, c);
                if (c + 2 < ngroups) load(bufA, c + 2);
                if (c + 1 < ngroups) apply(bufB, c + 1);
            }
            const double2 pv = make_double2(-v.x, -v.y);
            st2(a.p, u, pv);
            if (a.xt) st2(a.xt, u, make_double2(x.x + pv.x, x.y + pv.y));
            acc[0] = fma(g.y, pv.y, fma(g.x, pv.x, acc[0]));
            acc[1] = fma(pv.y, pv.y, fma(pv.x, pv.x, acc[1]));
        }
        if (C.tail_here(chunk) && threadIdx.x == 0) k3_tail(a, T, NoProbe(), acc);
        red::chunk_flush<2>(acc, parity, a.w.partials, a.w.stride, chunk);
    }
}

// ---- K3, bulk-async version (default).  The register version keeps at most two buffers of 8 columns x 16 B per
// thread in flight (its register budget), which left it 10 % below K1's per-byte rate (profiles/r01c_kernels.md:
// 79.6 % DRAM utilisation, 10 long-scoreboard stalls per issue).  Here one elected thread of a producer warp streams
// the 2k+1 (+1) input vectors of a tile of 256 units as 4 KB pieces into a shared-memory ring with 1-D bulk async
// copies (cp.async.bulk, SASS UBLKCP) that signal an mbarrier per stage; the 8 consumer warps apply the pieces in the
// reference's order out of shared memory.  Bytes in flight per SM are then the ring size (2 x NST x P x 4 KB), not a
// register budget, and the consumers need ~40 registers.  Same arithmetic in the same order as the register version:
// identical bits.
namespace tma {
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completes on `bar`
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
}  // namespace tma

template <int P, int NST, int MINB, class Probe>
static __global__ void __launch_bounds__(kThreads + 32, MINB) k3_direction_tma_kernel(K3Args a, Probe probe) {
    extern __shared__ __align__(128) unsigned char dyn[];
    double2 *ring = reinterpret_cast<double2 *>(dyn);            // [NST][P][kThreads]
    __shared__ K3Tables T;
    __shared__ __align__(8) uint64_t full[NST], empty[NST];
    constexpr int NACC = 2 + 2 * Probe::kSteps;                  // g.p, p.p (, f and f'.p at each probed step)
    probe.init(a, kThreads + 32);
    k3_build_tables(a, T, kThreads + 32);
    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; s++) { tma::mbar_init(&full[s], 1); tma::mbar_init(&empty[s], red::kWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int k = a.k;
    const int npieces = 2 * k + 1 + ((a.xt || Probe::kOn) ? 1 : 0);   // g, the 2k columns, (x)
    const int ngroups = (npieces + P - 1) / P;
    const Chunks C(a.n, a.ch);
    const bool producer = threadIdx.x >= kThreads;

    if (producer) {
        if (threadIdx.x != kThreads) return;                     // one elected thread issues every copy
        uint32_t it = 0;
        for (int64_t chunk = blockIdx.x; chunk < C.nchunks; chunk += gridDim.x) {
            const int64_t hi = C.hi(chunk);
            for (int64_t t0 = C.lo(chunk); t0 < hi; t0 += kThreads) {
                const uint32_t units = (uint32_t)((hi - t0) < kThreads ? (hi - t0) : kThreads);
                const uint32_t bytes = units * 16u;
                for (int grp = 0; grp < ngroups; grp++, it++) {
                    const int s = it % NST;
                    tma::mbar_wait(&empty[s], ((it / NST) & 1u) ^ 1u);
                    const int cnt = (npieces - grp * P) < P ? (npieces - grp * P) : P;
                    tma::mbar_expect_tx(&full[s], bytes * (uint32_t)cnt);
                    for (int i = 0; i < cnt; i++) {
                        const int j = grp * P + i;
                        const double *src = j == 0 ? a.g1 : (j <= 2 * k ? T.col[j - 1] : a.x1);
                        tma::bulk_g2s(ring + ((size_t)s * P + i) * kThreads, reinterpret_cast<const double2 *>(src) + t0,
                                      bytes, &full[s]);
                    }
                }
            }
        }
        return;
    }

    // ---- consumers (threads 0..255): barrier 1 is theirs alone
    const double gamma = T.gamma;
    const int lane = threadIdx.x & 31;
    uint32_t it = 0;
    int parity = 0;
    for (int64_t chunk = blockIdx.x; chunk < C.nchunks; chunk += gridDim.x) {
        const int64_t hi = C.hi(chunk);
        double acc[NACC];
#pragma unroll
        for (int i = 0; i < NACC; i++) acc[i] = 0.0;
        for (int64_t t0 = C.lo(chunk); t0 < hi; t0 += kThreads) {
            const int64_t u = t0 + threadIdx.x;
            const bool active = u < hi;
            double2 v = make_double2(0.0, 0.0), g = v, x = v;
            for (int grp = 0; grp < ngroups; grp++, it++) {
                const int s = it % NST;
                tma::mbar_wait(&full[s], (it / NST) & 1u);
                if (active) {
                    const double2 *st = ring + (size_t)s * P * kThreads + threadIdx.x;
#pragma unroll
                    for (int i = 0; i < P; i++) {
                        const int j = grp * P + i;
                        if (j < npieces) {
                            const double2 b = st[(size_t)i * kThreads];
                            if (j == 0) { v = b; g = b; }
                            else if (j <= 2 * k) {
                                const int o = j - 1;
                                if (o == k) { v.x = __dmul_rn(gamma, v.x); v.y = __dmul_rn(gamma, v.y); }
                                const double cf = T.coef[o];
                                v.x = __dadd_rn(v.x, __dmul_rn(cf, b.x));
                                v.y = __dadd_rn(v.y, __dmul_rn(cf, b.y));
                            } else {
                                x = b;
                            }
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) tma::mbar_arrive(&empty[s]);      // this warp is done with the stage
            }
            if (active) {
                const double2 pv = make_double2(-v.x, -v.y);
                st2(a.p, u, pv);
                if (a.xt) st2(a.xt, u, make_double2(x.x + pv.x, x.y + pv.y));
                acc[0] = fma(g.y, pv.y, fma(g.x, pv.x, acc[0]));
                acc[1] = fma(pv.y, pv.y, fma(pv.x, pv.x, acc[1]));
                // the first trial points of the next search, x1 + a*p with the roundings of the fused evaluation
                if (Probe::kOn) probe.unit(a, u, x, pv, &acc[2]);
            }
        }
        if (C.tail_here(chunk) && threadIdx.x == 0) k3_tail(a, T, probe, acc);
        red::chunk_flush<NACC, 1>(acc, parity, a.w.partials, a.w.stride, chunk);
    }
}


// ---- launching K3 with a probe (flgpu_problem.direction).  The library prepares everything but the probe; the ring
// shape is the default one (11 pieces x 2 stages, 2 CTAs/SM: with x1 as the extra piece 2k+2 = 22 pieces at k = 10).
struct K3Launch {
    K3Args a;
    int grid;            // SMs x 2 resident CTAs, capped by the number of chunks
    void *stream;
};

template <class Probe>
inline void launch_k3_probe(const K3Launch &L, const Probe &probe) {
    constexpr int P = 11, NST = 2, MINB = 2;
    constexpr size_t smem = (size_t)NST * P * kThreads * sizeof(double2);
    static bool attr_set[64] = {false};          // per device: a process may drive several GPUs
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev & 63]) {
        if (cudaFuncSetAttribute(k3_direction_tma_kernel<P, NST, MINB, Probe>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem) != cudaSuccess) {
            std::fprintf(stderr, "flgpu: K3: cannot reserve %zu bytes of shared memory\n", smem);
            std::abort();
        }
        attr_set[dev & 63] = true;
    }
    k3_direction_tma_kernel<P, NST, MINB, Probe><<<L.grid, kThreads + 32, smem, (cudaStream_t)L.stream>>>(L.a, probe);
}

}  // namespace k
}  // namespace flgpu
