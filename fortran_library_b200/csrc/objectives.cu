// objectives.cu -- K6: the benchmark objectives of BASELINE.json as CUDA kernels (f and f' in one
// pass over x, f reduced deterministically), plus index-generated start vectors.
//
//   quartic     f = sum x^4, f' = 4 x^3            (the reference's test objective, test/test.f90:630-663;
//                                                    x**4 = (x*x)*(x*x), x**3 = (x*x)*x as gfortran expands them)
//   Rosenbrock  f = sum_j 100 (x_{2j+1} - x_{2j}^2)^2 + (1 - x_{2j})^2   (extended, pairwise)
//   diag quad   f = 1/2 sum d_i (x_i - 1)^2, d_i log-uniform in [1, 1e6]
//   quartic1    f = sum (x-1)^4 + (x-1)^2       (the quartic with a non-zero, well-conditioned minimiser x* = 1)
//
// Element-wise arithmetic uses separate multiply/add roundings in the same order as the CPU
// oracle's objectives so that f' agrees bit for bit for the same x.
#include <cooperative_groups.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <utility>

#include "backend_cuda.cuh"
#include "../../include/flgpu_search_core.hpp"

namespace flgpu {

// ---- per-stream scratch for library kernels launched outside a CudaBackend (callbacks, primitives)
struct Scratch {
    k::Work work;         // 8 rows of chunk sums (grown on demand), block values, tickets
    double *scalar;       // device double[4]
    double *host_scalar;  // pinned double[4]
    double *tables;       // device double[768], diag-quad factors
};
constexpr int kScratchRows = 8;
static std::mutex g_scratch_mu;
static std::map<std::pair<int, void *>, Scratch> g_scratch;

// nchunks: chunk sums per row the caller is about to produce (the partial buffer grows to hold them)
Scratch &scratch_for(cudaStream_t s, int64_t nchunks) {
    int dev = 0;
    FLGPU_CUDA_CHECK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_scratch_mu);
    auto key = std::make_pair(dev, (void *)s);
    auto it = g_scratch.find(key);
    if (it == g_scratch.end()) {
        Scratch sc{};
        FLGPU_CUDA_CHECK(cudaMalloc((void **)&sc.work.blockvals, (size_t)kScratchRows * red::kTopMax * sizeof(double)));
        FLGPU_CUDA_CHECK(cudaMalloc((void **)&sc.work.tickets, kScratchRows * sizeof(unsigned int)));
        FLGPU_CUDA_CHECK(cudaMemset(sc.work.tickets, 0, kScratchRows * sizeof(unsigned int)));
        FLGPU_CUDA_CHECK(cudaMalloc((void **)&sc.scalar, 4 * sizeof(double)));
        FLGPU_CUDA_CHECK(cudaMallocHost((void **)&sc.host_scalar, 4 * sizeof(double)));
        FLGPU_CUDA_CHECK(cudaMalloc((void **)&sc.tables, 768 * sizeof(double)));
        double h[768];
        for (int q = 0; q < 256; q++) {
            h[q] = std::pow(10.0, 6.0 * (double)q / 16777216.0);
            h[256 + q] = std::pow(10.0, 6.0 * (double)q / 65536.0);
            h[512 + q] = std::pow(10.0, 6.0 * (double)q / 256.0);
        }
        FLGPU_CUDA_CHECK(cudaMemcpy(sc.tables, h, sizeof h, cudaMemcpyHostToDevice));
        it = g_scratch.emplace(key, sc).first;
    }
    Scratch &sc = it->second;
    if (nchunks > sc.work.stride) {        // grow (cudaFree waits for work still using the old buffer)
        int64_t cap = sc.work.stride > 0 ? sc.work.stride : 1024;
        while (cap < nchunks) cap *= 2;
        if (sc.work.partials) cudaFree(sc.work.partials);
        FLGPU_CUDA_CHECK(cudaMalloc((void **)&sc.work.partials, (size_t)kScratchRows * (size_t)cap * sizeof(double)));
        sc.work.stride = cap;
    }
    return sc;
}

// Called when a library-owned stream is destroyed: its scratch would otherwise stay in the map forever.
void scratch_release(cudaStream_t s) {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(g_scratch_mu);
    auto it = g_scratch.find(std::make_pair(dev, (void *)s));
    if (it == g_scratch.end()) return;
    cudaFree(it->second.work.partials);
    cudaFree(it->second.work.blockvals);
    cudaFree(it->second.work.tickets);
    cudaFree(it->second.scalar);
    cudaFreeHost(it->second.host_scalar);
    cudaFree(it->second.tables);
    g_scratch.erase(it);
}
double *scratch_scalar(cudaStream_t s) { return scratch_for(s, 1).scalar; }
k::Work scratch_work(cudaStream_t s, int64_t nchunks) { return scratch_for(s, nchunks < 1 ? 1 : nchunks).work; }

// chunk sums of rows [0, nrows) of `w` -> out[row] (device pointers; null = skip)
void launch_tree(const k::Work &w, int64_t nchunks, int nrows, double *const *out, cudaStream_t s) {
    k::TreeArgs a;
    a.w = w; a.nchunks = nchunks; a.lin_out = nullptr; a.dup_row = -1; a.dup_out = nullptr;
    for (int i = 0; i < 12; i++) a.out[i] = i < nrows ? out[i] : nullptr;
    const int nblk = (int)((nchunks + red::kBlockChunks - 1) / red::kBlockChunks);
    k::tree_kernel<<<dim3(nblk, nrows), k::kThreads, 0, s>>>(a);
}

namespace k {

__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dadd_rn(a, -b); }

struct ObjArgs {
    const double *x;    // the point (unfused) or x0 (fused: the point is x0 + a*p)
    const double *p;    // fused only
    double a;           // fused only
    double *x_out;      // fused + FLGPU_WRITE_X
    double *g;          // may be null (f only)
    int64_t n, offset, n_global, ch;
    double scale;       // diag quad: 2^24/(n_global-1)
    const double *tables;
    Work w;             // chunk sums: f -> row 0 (or the only row), f'.p -> the next row
    double *out[2];     // in-kernel finish (at most 4096 chunks): where the roots of those rows go; null: tree_kernel follows
};

// What an objective needs to know about where a unit sits in the global vector.
struct ObjIndex {
    double offset_d, scale;
    int64_t n_global;
    const double *tab;   // the diag-quad factor tables in shared memory (unused otherwise)
    // d_i = 10^(6 q / 2^24), q = trunc(i * scale) <= 2^24 (diagonal quadratic); the index arrives
    // as a double (exact below 2^53) and the truncation is a 32-bit conversion: same bits, no 64-bit I2F/F2I
    __device__ __forceinline__ double coeff(double i) const {
        if (n_global <= 1) return 1.0;
        const unsigned int q = __double2uint_rz(mul(i, scale));
        if (q >> 24) return 1.0e6;
        return mul(mul(tab[512 + ((q >> 16) & 255)], tab[256 + ((q >> 8) & 255)]), tab[q & 255]);
    }
};

// One 16-byte unit of objective KIND at local unit index u: f terms are ADDED to fsum in element order, f' -> g.
template <int KIND, bool WANT_F, bool NEED_G>
__device__ __forceinline__ void objective_unit(const ObjIndex &ix, int64_t u, const double2 x, double &fsum, double2 &g) {
    if (KIND == FLGPU_OBJ_QUARTIC) {
        const double x2 = mul(x.x, x.x), y2 = mul(x.y, x.y);
        if (WANT_F) { fsum += mul(x2, x2); fsum += mul(y2, y2); }
        if (NEED_G) { g.x = mul(4.0, mul(x2, x.x)); g.y = mul(4.0, mul(y2, x.y)); }
    } else if (KIND == FLGPU_OBJ_ROSENBROCK) {
        const double t1 = sub(x.y, mul(x.x, x.x)), t2 = sub(1.0, x.x);
        if (WANT_F) fsum += add(mul(mul(100.0, t1), t1), mul(t2, t2));
        if (NEED_G) {
            g.x = sub(mul(mul(-400.0, x.x), t1), mul(2.0, t2));
            g.y = mul(200.0, t1);
        }
    } else if (KIND == FLGPU_OBJ_QUARTIC_SHIFTED) {
        const double t0 = sub(x.x, 1.0), t1 = sub(x.y, 1.0);
        const double a2 = mul(t0, t0), b2 = mul(t1, t1);
        if (WANT_F) { fsum += add(mul(a2, a2), a2); fsum += add(mul(b2, b2), b2); }
        if (NEED_G) { g.x = add(mul(4.0, mul(a2, t0)), mul(2.0, t0)); g.y = add(mul(4.0, mul(b2, t1)), mul(2.0, t1)); }
    } else {
        const double i = ix.offset_d + (double)(2 * u);   // exact: both terms and the sum are integers < 2^53
        const double d0 = ix.coeff(i), d1 = ix.coeff(i + 1.0);
        const double t0 = sub(x.x, 1.0), t1 = sub(x.y, 1.0);
        if (WANT_F) { fsum += mul(mul(mul(0.5, d0), t0), t0); fsum += mul(mul(mul(0.5, d1), t1), t1); }
        if (NEED_G) { g.x = mul(d0, t0); g.y = mul(d1, t1); }
    }
}
// Several evaluations of the same unit (the batched probe, K3's probe): what does not depend on the point -- the diagonal
// quadratic's factors d_i and the products 0.5 d_i -- is formed once per unit; every point then sees the operations of
// objective_unit in the same order on the same values.
template <int KIND>
struct UnitInvariants {
    double d0 = 0.0, d1 = 0.0, h0 = 0.0, h1 = 0.0;
    __device__ __forceinline__ void prepare(const ObjIndex &ix, int64_t u) {
        if (KIND == FLGPU_OBJ_DIAGQUAD) {
            const double i = ix.offset_d + (double)(2 * u);
            d0 = ix.coeff(i); d1 = ix.coeff(i + 1.0);
            h0 = mul(0.5, d0); h1 = mul(0.5, d1);
        }
    }
    __device__ __forceinline__ void eval(const ObjIndex &ix, int64_t u, const double2 x, double &fsum, double2 &g) const {
        if (KIND == FLGPU_OBJ_DIAGQUAD) {
            const double t0 = sub(x.x, 1.0), t1 = sub(x.y, 1.0);
            fsum += mul(mul(h0, t0), t0); fsum += mul(mul(h1, t1), t1);
            g.x = mul(d0, t0); g.y = mul(d1, t1);
        } else {
            objective_unit<KIND, true, true>(ix, u, x, fsum, g);
        }
    }
};
// the unpaired last element of an odd-length shard (local element index i)
template <int KIND, bool WANT_F>
__device__ __forceinline__ void objective_tail(const ObjIndex &ix, int64_t i, const double x, double &fsum, double &g) {
    if (KIND == FLGPU_OBJ_QUARTIC) {
        const double x2 = mul(x, x);
        if (WANT_F) fsum += mul(x2, x2);
        g = mul(4.0, mul(x2, x));
    } else if (KIND == FLGPU_OBJ_ROSENBROCK) {
        const double t2 = sub(1.0, x);
        if (WANT_F) fsum += mul(t2, t2);
        g = mul(-2.0, t2);
    } else if (KIND == FLGPU_OBJ_QUARTIC_SHIFTED) {
        const double t = sub(x, 1.0), t2 = mul(t, t);
        if (WANT_F) fsum += add(mul(t2, t2), t2);
        g = add(mul(4.0, mul(t2, t)), mul(2.0, t));
    } else {
        const double d = ix.coeff(ix.offset_d + (double)i), t = sub(x, 1.0);
        if (WANT_F) fsum += mul(mul(mul(0.5, d), t), t);
        g = mul(d, t);
    }
}

// FUSED: the point is formed as x0 + a*p (multiply, then add: f90:1482) instead of being loaded; WANT_GP:
// f'(x).p is reduced alongside f; WRITE_X / WRITE_G: store the point / gradient.
// One CHUNK of an objective evaluation (flgpu_reduce.cuh): this thread's units of chunk c in order, f and f'.p
// accumulated into fsum / gpsum.  Shared by objective_kernel (one evaluation per launch) and search_kernel (a whole
// line search per launch), so both produce the same chunk sums; the accumulation order of f does not depend on the
// flags either, so f has the same bits on the fused and the unfused path.
template <int KIND, bool FUSED, bool WANT_F, bool WANT_GP, bool WRITE_X, bool WRITE_G>
__device__ __forceinline__ void objective_chunk(const ObjArgs &a, const ObjIndex &ix, const Chunks &C, int64_t c,
                                                double &fsum, double &gpsum) {
    constexpr bool NEED_G = WANT_GP || WRITE_G;
    const double step = a.a;
    auto unit = [&](int64_t u, double2 x, const double2 pv) {
        if (FUSED) {
            x.x = add(x.x, mul(step, pv.x));
            x.y = add(x.y, mul(step, pv.y));
            if (WRITE_X) st2(a.x_out, u, x);
        }
        double2 g = make_double2(0.0, 0.0);
        objective_unit<KIND, WANT_F, NEED_G>(ix, u, x, fsum, g);
        if (WRITE_G) st2(a.g, u, g);
        if (WANT_GP) gpsum = fma(g.y, pv.y, fma(g.x, pv.x, gpsum));
    };
    // four units per trip with all loads issued first (8 x 16 B in flight per thread on the fused path)
    const double2 zero2 = make_double2(0.0, 0.0);
    const int64_t hi = C.hi(c);
    int64_t u = C.lo(c) + threadIdx.x;
    for (; u + 3 * kThreads < hi; u += 4 * kThreads) {
        const double2 x0v = ld2(a.x, u), x1v = ld2(a.x, u + kThreads), x2v = ld2(a.x, u + 2 * kThreads),
                      x3v = ld2(a.x, u + 3 * kThreads);
        double2 p0v = zero2, p1v = zero2, p2v = zero2, p3v = zero2;
        if (FUSED) {
            p0v = ld2(a.p, u); p1v = ld2(a.p, u + kThreads); p2v = ld2(a.p, u + 2 * kThreads); p3v = ld2(a.p, u + 3 * kThreads);
        }
        unit(u, x0v, p0v); unit(u + kThreads, x1v, p1v); unit(u + 2 * kThreads, x2v, p2v); unit(u + 3 * kThreads, x3v, p3v);
    }
    for (; u < hi; u += kThreads) {
        const double2 xv = ld2(a.x, u);
        double2 pv = zero2;
        if (FUSED) pv = ld2(a.p, u);
        unit(u, xv, pv);
    }
    if (C.tail_here(c) && threadIdx.x == 0) {
        const int64_t i = a.n - 1;
        double x = a.x[i], pv = 0.0;
        if (FUSED) {
            pv = a.p[i];
            x = add(x, mul(step, pv));
            if (WRITE_X) a.x_out[i] = x;
        }
        double g = 0.0;
        objective_tail<KIND, WANT_F>(ix, i, x, fsum, g);
        if (WRITE_G) a.g[i] = g;
        if (WANT_GP) gpsum = fma(g, pv, gpsum);
    }
}

template <int KIND>
__device__ __forceinline__ ObjIndex load_tables(int64_t offset, int64_t n_global, double scale, const double *tables, double *tab,
                                                int nthreads) {
    if (KIND == FLGPU_OBJ_DIAGQUAD) {
        for (int i = threadIdx.x; i < 768; i += nthreads) tab[i] = tables[i];
        __syncthreads();
    }
    ObjIndex ix;
    ix.offset_d = (double)offset; ix.scale = scale; ix.n_global = n_global; ix.tab = tab;
    return ix;
}

// K1 source for the built-in objectives (flgpu_problem.update, include/flgpu_k1.cuh): the accepted point
// x1 = x0 + a*p (multiply, then add: f90:1482) and f'(x1) are formed in registers -- the same roundings as the fused
// evaluation that would otherwise store them -- and stored by the column group that owns the new column.
template <int KIND>
struct BuiltinSrc {
    const double *tables;
    double scale;
    ObjIndex ix;
    __device__ void init(const K1Args &a) {
        __shared__ double tab[KIND == FLGPU_OBJ_DIAGQUAD ? 768 : 1];
        ix = load_tables<KIND>(a.offset, a.n_global, scale, tables, tab, kThreads);
    }
    // f'(x0) is re-evaluated from the x0 that is loaded anyway (same operations on the same bits as when it was first
    // formed: identical value) instead of being read back: one n-vector less traffic per iteration
    __device__ __forceinline__ void unit(const K1Args &a, int64_t u, bool own_new, double2 x0, double2 &x1, double2 &g1,
                                         double2 &g0) const {
        const double2 pv = ld2(a.p, u);
        x1.x = add(x0.x, mul(a.step, pv.x));
        x1.y = add(x0.y, mul(a.step, pv.y));
        double f = 0.0;
        objective_unit<KIND, false, true>(ix, u, x1, f, g1);
        objective_unit<KIND, false, true>(ix, u, x0, f, g0);
        if (own_new) { st2(a.x1_out, u, x1); st2(a.g1_out, u, g1); }
    }
    __device__ __forceinline__ void tail(const K1Args &a, int64_t i, bool own_new, double x0, double &x1, double &g1,
                                         double &g0) const {
        x1 = add(x0, mul(a.step, a.p[i]));
        double f = 0.0;
        objective_tail<KIND, false>(ix, i, x1, f, g1);
        objective_tail<KIND, false>(ix, i, x0, f, g0);
        if (own_new) { a.x1_out[i] = x1; a.g1_out[i] = g1; }
    }
};

// K3 probe for the built-in objectives (flgpu_problem.direction, include/flgpu_k3.cuh): f and f'.p at the first trial
// points of the next search, formed in registers from the x1 and p K3 holds -- objective_unit on the same points in the
// same per-thread unit order as objective_chunk, hence the chunk sums (and everything above them) of separate fused
// evaluations.
template <int KIND>
struct BuiltinProbe {
    static constexpr bool kOn = true;
    static constexpr int kSteps = kProbeSteps;
    const double *tables;
    double scale;
    ObjIndex ix;
    __device__ void init(const K3Args &a, int nthreads) {
        __shared__ double tab[KIND == FLGPU_OBJ_DIAGQUAD ? 768 : 1];
        ix = load_tables<KIND>(a.offset, a.n_global, scale, tables, tab, nthreads);
    }
    __device__ __forceinline__ void unit(const K3Args &a, int64_t u, const double2 x1, const double2 pv, double *acc) const {
        UnitInvariants<KIND> inv;
        inv.prepare(ix, u);
#pragma unroll
        for (int j = 0; j < kSteps; j++) {
            double2 x, g = make_double2(0.0, 0.0);
            x.x = add(x1.x, mul(a.steps[j], pv.x));
            x.y = add(x1.y, mul(a.steps[j], pv.y));
            inv.eval(ix, u, x, acc[2 * j], g);
            acc[2 * j + 1] = fma(g.y, pv.y, fma(g.x, pv.x, acc[2 * j + 1]));
        }
    }
    __device__ __forceinline__ void tail(const K3Args &a, int64_t i, const double x1, const double pv, double *acc) const {
#pragma unroll
        for (int j = 0; j < kSteps; j++) {
            double g = 0.0;
            objective_tail<KIND, true>(ix, i, add(x1, mul(a.steps[j], pv)), acc[2 * j], g);
            acc[2 * j + 1] = fma(g, pv, acc[2 * j + 1]);
        }
    }
};

// One kernel serves the plain callbacks (f, fd, f_fd) and the fused line-search evaluation (flgpu_fused_fn).
template <int KIND, bool FUSED, bool WANT_F, bool WANT_GP, bool WRITE_X, bool WRITE_G>
__global__ void __launch_bounds__(kThreads, 4) objective_kernel(ObjArgs a) {
    __shared__ double tab[KIND == FLGPU_OBJ_DIAGQUAD ? 768 : 1];
    const ObjIndex ix = load_tables<KIND>(a.offset, a.n_global, a.scale, a.tables, tab, kThreads);
    const Chunks C(a.n, a.ch);
    int parity = 0;
    for (int64_t c = blockIdx.x; c < C.nchunks; c += gridDim.x) {
        double fsum = 0.0, gpsum = 0.0;
        objective_chunk<KIND, FUSED, WANT_F, WANT_GP, WRITE_X, WRITE_G>(a, ix, C, c, fsum, gpsum);
        if (WANT_F && WANT_GP) {
            const double acc[2] = {fsum, gpsum};
            red::chunk_flush<2>(acc, parity, a.w.partials, a.w.stride, c);
        } else if (WANT_F) {
            const double acc[1] = {fsum};
            red::chunk_flush<1>(acc, parity, a.w.partials, a.w.stride, c);
        } else if (WANT_GP) {
            const double acc[1] = {gpsum};
            red::chunk_flush<1>(acc, parity, a.w.partials, a.w.stride, c);
        }
    }
    if (a.out[0]) {                      // small reduction: the last block forms the tree(s) itself
        if (WANT_F && WANT_GP) {
            double *const o[2] = {a.out[0], a.out[1]};
            red::finish_in_kernel<2>(a.w.partials, a.w.stride, C.nchunks, a.w.tickets, o);
        } else if (WANT_F || WANT_GP) {
            double *const o[1] = {a.out[0]};
            red::finish_in_kernel<1>(a.w.partials, a.w.stride, C.nchunks, a.w.tickets, o);
        }
    }
}

// ------------------------------------------------------------------ batched fused evaluation (flgpu_fused_multi_fn)
// f and f'.p at x0 + steps[j]*p for J steps in ONE pass over x0 and p: the traffic of one probe, J times its (few)
// flops.  Per step the arithmetic and the accumulation order are objective_chunk's (thread t takes the chunk's units
// t, t+256, ... in order; f terms added in element order, f'.p by FMA), so every pair of sums carries the bits of a
// separate objective_kernel<KIND, FUSED, F, GP> launch with a = steps[j].  Chunk sums: f_j -> row 2j, (f'.p)_j -> row 2j+1.
struct MultiArgs {
    const double *x, *p;
    double steps[FLGPU_MULTI_MAX];
    int64_t n, offset, n_global, ch;
    double scale;
    const double *tables;
    Work w;
    double *out;        // in-kernel finish (at most 4096 chunks): out[0 .. 2J); null: tree_kernel follows
};

template <int KIND, int J>
__global__ void __launch_bounds__(kThreads, 3) objective_multi_kernel(MultiArgs a) {
    __shared__ double tab[KIND == FLGPU_OBJ_DIAGQUAD ? 768 : 1];
    const ObjIndex ix = load_tables<KIND>(a.offset, a.n_global, a.scale, a.tables, tab, kThreads);
    const Chunks C(a.n, a.ch);
    int parity = 0;
    for (int64_t c = blockIdx.x; c < C.nchunks; c += gridDim.x) {
        double acc[2 * J];
#pragma unroll
        for (int i = 0; i < 2 * J; i++) acc[i] = 0.0;
        auto unit = [&](int64_t u, const double2 x0v, const double2 pv) {
            UnitInvariants<KIND> inv;
            inv.prepare(ix, u);
#pragma unroll
            for (int j = 0; j < J; j++) {
                double2 x, g = make_double2(0.0, 0.0);
                x.x = add(x0v.x, mul(a.steps[j], pv.x));
                x.y = add(x0v.y, mul(a.steps[j], pv.y));
                inv.eval(ix, u, x, acc[2 * j], g);
                acc[2 * j + 1] = fma(g.y, pv.y, fma(g.x, pv.x, acc[2 * j + 1]));
            }
        };
        const int64_t hi = C.hi(c);
        int64_t u = C.lo(c) + threadIdx.x;
        for (; u + 3 * kThreads < hi; u += 4 * kThreads) {       // four units per trip, all eight loads issued first
            const double2 x0v = ld2(a.x, u), x1v = ld2(a.x, u + kThreads), x2v = ld2(a.x, u + 2 * kThreads),
                          x3v = ld2(a.x, u + 3 * kThreads);
            const double2 p0v = ld2(a.p, u), p1v = ld2(a.p, u + kThreads), p2v = ld2(a.p, u + 2 * kThreads),
                          p3v = ld2(a.p, u + 3 * kThreads);
            unit(u, x0v, p0v); unit(u + kThreads, x1v, p1v); unit(u + 2 * kThreads, x2v, p2v); unit(u + 3 * kThreads, x3v, p3v);
        }
        for (; u < hi; u += kThreads) unit(u, ld2(a.x, u), ld2(a.p, u));
        if (C.tail_here(c) && threadIdx.x == 0) {
            const int64_t i = a.n - 1;
            const double x0 = a.x[i], pv = a.p[i];
#pragma unroll
            for (int j = 0; j < J; j++) {
                double g = 0.0;
                objective_tail<KIND, true>(ix, i, add(x0, mul(a.steps[j], pv)), acc[2 * j], g);
                acc[2 * j + 1] = fma(g, pv, acc[2 * j + 1]);
            }
        }
        red::chunk_flush<2 * J>(acc, parity, a.w.partials, a.w.stride, c);
    }
    if (a.out) {
        double *o[2 * J];
#pragma unroll
        for (int i = 0; i < 2 * J; i++) o[i] = a.out + i;
        double *const(&oc)[2 * J] = o;
        red::finish_in_kernel<2 * J>(a.w.partials, a.w.stride, C.nchunks, a.w.tickets, oc);
    }
}

// ------------------------------------------------------------------ device-resident line search (flgpu_search_fn)
// The whole Wolfe / Strong-Wolfe search in ONE cooperative kernel.  Every thread of every block runs the same state
// machine (SearchCore, the source the host driver compiles too) on the same values, so control flow is uniform across
// the grid; an evaluation is this block's chunks (objective_chunk, as in objective_kernel), one grid-wide barrier, and
// the tree over the chunk sums (flgpu_reduce.cuh) -- repeated by every block, which saves the second barrier a
// broadcast would need.  Chunk sums are double-buffered on the evaluation parity: a block can be at most one
// evaluation ahead of the slowest reader.  Chunk sums and tree are those of objective_kernel + tree_kernel, so f and
// f'.p carry the same bits as on the host-driven fused path and both paths take the same decisions.
struct SearchKArgs {
    ObjArgs o;              // x = x0, p, x_out / g = accepted point / gradient; a is set per evaluation;
                            // o.w.partials: rows [2 * parity + i]
    double c1, c2abs, fx0, phid0, incr, a0;
    int strong, fdwithf, store;   // store = 0: the caller's K1 forms and stores the accepted point itself
    double *result;         // FLGPU_SEARCH_RESULT_DOUBLES
    // row-sharded runs: the rank exchange happens inside the kernel (block 0) over the search mailboxes
    PeerTable peers;
    int me, G;
    unsigned long long *dseq;   // this rank's sequence counter for those mailboxes (device memory)
    double *glob;               // [2][2] rank-combined values for the other blocks
    unsigned long long timeout_ns;
};

constexpr double kEvalBudget = 100000.0;

template <int KIND>
struct DevSearch : SearchCore<DevSearch<KIND>> {
    const SearchKArgs &K;
    const ObjIndex &ix;
    double *rsh;                   // red::kWarps + red::kTopMax doubles of block scratch
    double f_cur = 0.0, gp_cur = 0.0, a_x = 0.0, a_g = 0.0;
    bool have_x = false, have_g = false;
    int parity = 0, fpar = 0;
    double trials = 0.0, n_f = 0.0, n_fd = 0.0, n_ffd = 0.0, n_fonly = 0.0;
    unsigned long long seq_base = 0, nexch = 0;   // exchanges made so far (uniform over the grid)

    __device__ DevSearch(const SearchKArgs &k, const ObjIndex &i, double *s) : K(k), ix(i), rsh(s) {
        if (K.G > 1) seq_base = *K.dseq;
    }

    template <bool F, bool GP>
    __device__ void eval() {
        constexpr int NACC = (F && GP) ? 2 : 1;
        ObjArgs o = K.o;
        o.a = a_x;
        const Chunks C(o.n, o.ch);
        double *rows = o.w.partials + (int64_t)(2 * parity) * o.w.stride;
        for (int64_t c = blockIdx.x; c < C.nchunks; c += gridDim.x) {
            double fsum = 0.0, gpsum = 0.0;
            objective_chunk<KIND, true, F, GP, false, false>(o, ix, C, c, fsum, gpsum);
            if (NACC == 2) {
                const double acc[2] = {fsum, gpsum};
                red::chunk_flush<2>(acc, fpar, rows, o.w.stride, c);
            } else {
                const double acc[1] = {F ? fsum : gpsum};
                red::chunk_flush<1>(acc, fpar, rows, o.w.stride, c);
            }
        }
        __threadfence();
        cooperative_groups::this_grid().sync();
        double v[2] = {0.0, 0.0};
#pragma unroll
        for (int i = 0; i < NACC; i++) {
            v[i] = red::cta_root(rows + (int64_t)i * o.w.stride, C.nchunks, rsh);
            __syncthreads();
        }
        if (K.G > 1) {
            // every block holds this rank's roots; block 0 trades them with the other ranks (stores into their
            // mailboxes, flags, rank tree) and a second barrier hands the result to the rest of the grid
            nexch++;
            double *gl = K.glob + parity * 2;
            if (blockIdx.x == 0) {
                __shared__ double mine[2], summed[2];
                if (threadIdx.x < NACC) mine[threadIdx.x] = v[threadIdx.x];
                __syncthreads();
                mailbox_exchange_block(K.peers, K.me, K.G, seq_base + nexch, mine, NACC, summed, K.timeout_ns);
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

// FAST = the FLGPU_LS_FAST searcher (SearchCore::fast); a template parameter so that the reference-exact kernel's code
// and register allocation do not depend on it
template <int KIND, bool FAST>
__global__ void __launch_bounds__(kThreads, 3) search_kernel(SearchKArgs K) {
    __shared__ double tab[KIND == FLGPU_OBJ_DIAGQUAD ? 768 : 1];
    __shared__ double rsh[red::kWarps + red::kTopMax];
    const ObjIndex ix = load_tables<KIND>(K.o.offset, K.o.n_global, K.o.scale, K.o.tables, tab, kThreads);
    DevSearch<KIND> S(K, ix, rsh);
    S.c1 = K.c1; S.c2abs = K.c2abs; S.fx0 = K.fx0; S.phid0 = K.phid0; S.incr = K.incr;
    S.fdwithf = K.fdwithf != 0; S.a = K.a0; S.f_cur = K.fx0; S.pre = 0;
    if (FAST) S.fast(K.strong != 0);
    else if (K.strong) S.strongwolfe(); else S.wolfe();
    // the point and gradient the reference leaves in x / fdx
    ObjArgs o = K.o;
    const Chunks C(o.n, o.ch);
    double f0 = 0.0, g0 = 0.0;
    if (K.store) {
        for (int64_t c = blockIdx.x; c < C.nchunks; c += gridDim.x) {
            if (S.have_x && S.have_g && S.a_x == S.a_g) {
                o.a = S.a_x;
                objective_chunk<KIND, true, false, false, true, true>(o, ix, C, c, f0, g0);
            } else {                                   // never taken by the reference's searchers; kept for fidelity
                if (S.have_x) { o.a = S.a_x; objective_chunk<KIND, true, false, false, true, false>(o, ix, C, c, f0, g0); }
                if (S.have_g) { o.a = S.a_g; objective_chunk<KIND, true, false, false, false, true>(o, ix, C, c, f0, g0); }
            }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        K.result[0] = S.a; K.result[1] = S.f_cur; K.result[2] = S.trials; K.result[3] = S.n_f;
        K.result[4] = S.n_fd; K.result[5] = S.n_ffd; K.result[6] = S.n_fonly;
        K.result[7] = S.aborted() ? -1.0 : (double)S.nexch;
        if (K.G > 1) *K.dseq = S.seq_base + S.nexch;
    }
}

__device__ __forceinline__ double splitmix_u(unsigned long long z) {
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    return mul((double)(z >> 11), 1.0 / 9007199254740992.0);
}

__global__ void __launch_bounds__(kThreads) start_kernel(int kind, unsigned long long seed, double *x, int64_t offset,
                                                         int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    for (int64_t k = (int64_t)blockIdx.x * kThreads + threadIdx.x; k < n; k += stride) {
        const int64_t i = offset + k;
        const double u = splitmix_u((unsigned long long)i + seed);
        double v;
        switch (kind) {
        case FLGPU_START_QUARTIC_U: v = u; break;
        case FLGPU_START_ROSEN_STD: v = (i & 1) ? 1.0 : -1.2; break;
        case FLGPU_START_ROSEN_PERT: v = add((i & 1) ? 1.0 : -1.2, mul(0.1, sub(u, 0.5))); break;
        default: v = 0.0; break;
        }
        x[k] = v;
    }
}

}  // namespace k

// ---- launcher shared by both callback flavours
static int device_sms() {
    static std::mutex mu;
    static std::map<int, int> sms_of;            // per device: a process may drive several GPUs
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    int &sms = sms_of[dev];
    if (!sms) {
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}
// grid of a reducing kernel: one full wave of resident CTAs (`per_sm` per SM), never more blocks than chunks
static int chunk_grid(int64_t nchunks, int per_sm) {
    int64_t g = (int64_t)device_sms() * per_sm;
    if (g > k::kMaxGrid) g = k::kMaxGrid;
    return (int)(nchunks < g ? (nchunks < 1 ? 1 : nchunks) : g);
}
static int unit_grid(int64_t n) {
    int64_t need = (n / 2 + k::kThreads) / k::kThreads;
    int64_t g = (int64_t)device_sms() * 8;
    if (g > k::kMaxGrid) g = k::kMaxGrid;
    return (int)(need < g ? (need < 1 ? 1 : need) : g);
}

// flags = FLGPU_WANT_* | FLGPU_WRITE_*; fused: x_dev is x0 and the point is x0 + a*p
void launch_objective(int kind, bool fused, int flags, double *f_dev, double *gp_dev, double *x_out, double *g_dev,
                      const double *x_dev, const double *p_dev, double step, int64_t n, int64_t offset,
                      int64_t n_global, cudaStream_t s) {
    require_aligned16(x_dev, "objective: x"); require_aligned16(p_dev, "objective: p");
    require_aligned16(x_out, "objective: x_out"); require_aligned16(g_dev, "objective: f'");
    if (n_global < n) n_global = n;
    const int64_t ch = red::chunk_elems(n_global), nchunks = red::num_chunks(n, ch);
    Scratch &sc = scratch_for(s, nchunks);
    k::ObjArgs a;
    a.x = x_dev; a.p = p_dev; a.a = step; a.x_out = x_out; a.g = g_dev;
    a.n = n; a.offset = offset; a.n_global = n_global; a.ch = ch;
    a.scale = n_global > 1 ? 16777216.0 / (double)(n_global - 1) : 0.0;
    a.tables = sc.tables; a.w = sc.work;
    // chunk sums -> this rank's roots: f in row 0 (or f'.p when only that was asked for), f'.p in the next row; up to 4096
    // chunks the producing kernel's last block forms the trees itself, above that tree_kernel follows
    const bool wf = (flags & FLGPU_WANT_F) != 0, wgp = (flags & FLGPU_WANT_GP) != 0;
    double *out[2] = {wf ? f_dev : gp_dev, gp_dev};
    const bool in_kernel = (wf || wgp) && nchunks <= red::kBlockChunks;
    a.out[0] = in_kernel ? out[0] : nullptr; a.out[1] = in_kernel ? out[1] : nullptr;
    const int grid = chunk_grid(nchunks, 4);     // = resident CTAs per SM (__launch_bounds__(256, 4)): one full wave
#define FLGPU_OBJ_CASE(KIND, FU, F, GP, WX, WG)                                                                 \
    k::objective_kernel<KIND, FU, F, GP, WX, WG><<<grid, k::kThreads, 0, s>>>(a)
#define FLGPU_OBJ_LAUNCH(KIND)                                                                                  \
    do {                                                                                                        \
        if (!fused) {                                                                                           \
            if (flags == (FLGPU_WANT_F | FLGPU_WRITE_G)) FLGPU_OBJ_CASE(KIND, false, true, false, false, true); \
            else if (flags == FLGPU_WANT_F) FLGPU_OBJ_CASE(KIND, false, true, false, false, false);             \
            else if (flags == FLGPU_WRITE_G) FLGPU_OBJ_CASE(KIND, false, false, false, false, true);            \
            else fatal("built-in objective: unsupported evaluation request");                                   \
        } else {                                                                                                \
            if (flags == (FLGPU_WANT_F | FLGPU_WANT_GP)) FLGPU_OBJ_CASE(KIND, true, true, true, false, false);  \
            else if (flags == FLGPU_WANT_F) FLGPU_OBJ_CASE(KIND, true, true, false, false, false);              \
            else if (flags == FLGPU_WANT_GP) FLGPU_OBJ_CASE(KIND, true, false, true, false, false);             \
            else if (flags == (FLGPU_WRITE_X | FLGPU_WRITE_G)) FLGPU_OBJ_CASE(KIND, true, false, false, true, true); \
            else if (flags == FLGPU_WRITE_G) FLGPU_OBJ_CASE(KIND, true, false, false, false, true);             \
            else if (flags == FLGPU_WRITE_X) FLGPU_OBJ_CASE(KIND, true, false, false, true, false);             \
            else if (flags == (FLGPU_WANT_F | FLGPU_WANT_GP | FLGPU_WRITE_X | FLGPU_WRITE_G))                   \
                FLGPU_OBJ_CASE(KIND, true, true, true, true, true);                                             \
            else fatal("built-in objective: unsupported fused evaluation request");                             \
        }                                                                                                       \
    } while (0)
    switch (kind) {
    case FLGPU_OBJ_QUARTIC: FLGPU_OBJ_LAUNCH(FLGPU_OBJ_QUARTIC); break;
    case FLGPU_OBJ_ROSENBROCK: FLGPU_OBJ_LAUNCH(FLGPU_OBJ_ROSENBROCK); break;
    case FLGPU_OBJ_DIAGQUAD: FLGPU_OBJ_LAUNCH(FLGPU_OBJ_DIAGQUAD); break;
    case FLGPU_OBJ_QUARTIC_SHIFTED: FLGPU_OBJ_LAUNCH(FLGPU_OBJ_QUARTIC_SHIFTED); break;
    default: fatal("unknown built-in objective");
    }
#undef FLGPU_OBJ_LAUNCH
#undef FLGPU_OBJ_CASE
    if ((wf || wgp) && !in_kernel) launch_tree(sc.work, nchunks, wf && wgp ? 2 : 1, out, s);
}

// batched fused evaluation: out_dev[2j], out_dev[2j+1] = this rank's roots of f and f'.p at x0 + steps[j]*p
void launch_objective_multi(int kind, int count, const double *steps, double *out_dev, const double *x_dev,
                            const double *p_dev, int64_t n, int64_t offset, int64_t n_global, cudaStream_t s) {
    constexpr int J = FLGPU_MULTI_MAX;
    if (count < 1 || count > J) fatal("built-in objective: batched evaluation of 1 to 4 steps");
    require_aligned16(x_dev, "objective: x"); require_aligned16(p_dev, "objective: p");
    if (n_global < n) n_global = n;
    const int64_t ch = red::chunk_elems(n_global), nchunks = red::num_chunks(n, ch);
    Scratch &sc = scratch_for(s, nchunks);
    k::MultiArgs a;
    a.x = x_dev; a.p = p_dev;
    for (int j = 0; j < J; j++) a.steps[j] = steps[j < count ? j : count - 1];   // unused lanes repeat the last step
    a.n = n; a.offset = offset; a.n_global = n_global; a.ch = ch;
    a.scale = n_global > 1 ? 16777216.0 / (double)(n_global - 1) : 0.0;
    a.tables = sc.tables; a.w = sc.work;
    const bool in_kernel = nchunks <= red::kBlockChunks;
    a.out = in_kernel ? out_dev : nullptr;
    // (staging x0 and p through a shared-memory ring as K3 does was built and measured: 1.07 ms against 1.04 ms at
    // n = 2^28 for Rosenbrock, 0.80 against 0.80 for the quartic -- the pass is bound by fp64 issue, not by exposed
    // memory latency -- so the register version is the only one kept)
    const int grid = chunk_grid(nchunks, 3);     // = resident CTAs per SM (__launch_bounds__(256, 3)): one full wave
    switch (kind) {
    case FLGPU_OBJ_QUARTIC: k::objective_multi_kernel<FLGPU_OBJ_QUARTIC, J><<<grid, k::kThreads, 0, s>>>(a); break;
    case FLGPU_OBJ_ROSENBROCK: k::objective_multi_kernel<FLGPU_OBJ_ROSENBROCK, J><<<grid, k::kThreads, 0, s>>>(a); break;
    case FLGPU_OBJ_DIAGQUAD: k::objective_multi_kernel<FLGPU_OBJ_DIAGQUAD, J><<<grid, k::kThreads, 0, s>>>(a); break;
    case FLGPU_OBJ_QUARTIC_SHIFTED: k::objective_multi_kernel<FLGPU_OBJ_QUARTIC_SHIFTED, J><<<grid, k::kThreads, 0, s>>>(a); break;
    default: fatal("unknown built-in objective");
    }

    if (!in_kernel) {
        double *out[2 * J];
        for (int i = 0; i < 2 * J; i++) out[i] = i < 2 * count ? out_dev + i : nullptr;
        launch_tree(sc.work, nchunks, 2 * count, out, s);
    }
}

// 64-bit device-callback flavour
template <int KIND>
static void dev_f(const flgpu_eval_ctx *c, double *f, const double *x, int64_t n) {
    launch_objective(KIND, false, FLGPU_WANT_F, f, nullptr, nullptr, nullptr, x, nullptr, 0.0, n, c->offset,
                     c->n_global, (cudaStream_t)c->stream);
}
template <int KIND>
static void dev_fd(const flgpu_eval_ctx *c, double *g, const double *x, int64_t n) {
    launch_objective(KIND, false, FLGPU_WRITE_G, nullptr, nullptr, nullptr, g, x, nullptr, 0.0, n, c->offset,
                     c->n_global, (cudaStream_t)c->stream);
}
template <int KIND>
static void dev_ffd(const flgpu_eval_ctx *c, double *f, double *g, const double *x, int64_t n) {
    launch_objective(KIND, false, FLGPU_WANT_F | FLGPU_WRITE_G, f, nullptr, nullptr, g, x, nullptr, 0.0, n,
                     c->offset, c->n_global, (cudaStream_t)c->stream);
}
template <int KIND>
static void dev_fused(const flgpu_eval_ctx *c, int flags, double *f, double *gp, double *x_out, double *g_out,
                      const double *x0, const double *p, double a, int64_t n) {
    launch_objective(KIND, true, flags, f, gp, x_out, g_out, x0, p, a, n, c->offset, c->n_global,
                     (cudaStream_t)c->stream);
}

template <int KIND>
static void dev_fused_multi(const flgpu_eval_ctx *c, int count, const double *steps, double *out_dev, const double *x0,
                            const double *p, int64_t n) {
    launch_objective_multi(KIND, count, steps, out_dev, x0, p, n, c->offset, c->n_global, (cudaStream_t)c->stream);
}

// flgpu_problem.update: the first K1 pass with the accepted point formed in the kernel
template <int KIND>
static void dev_update(const flgpu_eval_ctx *c, const flgpu_update_args *A, int64_t n) {
    (void)n;
    if (A->k1_bytes != sizeof(k::K1Launch)) fatal("flgpu_problem.update: K1Launch layout mismatch (header / library versions differ)");
    const k::K1Launch &L = *(const k::K1Launch *)A->k1;
    Scratch &sc = scratch_for((cudaStream_t)c->stream, 1);
    k::BuiltinSrc<KIND> src;
    src.tables = sc.tables;
    src.scale = L.a.n_global > 1 ? 16777216.0 / (double)(L.a.n_global - 1) : 0.0;
    k::launch_k1_pass(L, src);
}

// flgpu_problem.direction: K3 with the first trial of the next search evaluated inside the kernel
template <int KIND>
static void dev_direction(const flgpu_eval_ctx *c, const flgpu_direction_args *A, int64_t n) {
    (void)n;
    if (A->k3_bytes != sizeof(k::K3Launch)) fatal("flgpu_problem.direction: K3Launch layout mismatch (header / library versions differ)");
    const k::K3Launch &L = *(const k::K3Launch *)A->k3;
    Scratch &sc = scratch_for((cudaStream_t)c->stream, 1);
    k::BuiltinProbe<KIND> probe;
    probe.tables = sc.tables;
    probe.scale = L.a.n_global > 1 ? 16777216.0 / (double)(L.a.n_global - 1) : 0.0;
    k::launch_k3_probe(L, probe);
}

// device-resident search: same chunk sums and tree as the probes (bit-identical f, f'.p); the grid is capped by what
// can be co-resident (any grid gives the same bits: the chunk sums do not depend on which block forms them)
template <int KIND, bool FAST>
static void dev_search_policy(const flgpu_eval_ctx *c, const flgpu_search_args *A, int64_t n) {
    cudaStream_t s = (cudaStream_t)c->stream;
    const int64_t n_global = c->n_global < n ? n : c->n_global;
    const int64_t ch = red::chunk_elems(n_global), nchunks = red::num_chunks(n, ch);
    Scratch &sc = scratch_for(s, nchunks);
    int dev = 0, resident = 0, coop = 0;
    FLGPU_CUDA_CHECK(cudaGetDevice(&dev));
    FLGPU_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, k::search_kernel<KIND, FAST>, k::kThreads, 0));
    FLGPU_CUDA_CHECK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    if (!coop || resident < 1) fatal("device-resident line search needs cooperative kernel launch");
    int grid = chunk_grid(nchunks, 4);
    if (grid > resident * device_sms()) grid = resident * device_sms();
    k::SearchKArgs K;
    K.o.x = A->x0_dev; K.o.p = A->p_dev; K.o.a = 0.0; K.o.x_out = A->x_out; K.o.g = A->g_out;
    K.o.n = n; K.o.offset = c->offset; K.o.n_global = n_global; K.o.ch = ch;
    K.o.scale = n_global > 1 ? 16777216.0 / (double)(n_global - 1) : 0.0;
    K.o.tables = sc.tables; K.o.w = sc.work;       // rows 0..3: [evaluation parity][f, f'.p]
    K.o.out[0] = K.o.out[1] = nullptr;
    K.c1 = A->c1; K.c2abs = A->c2abs; K.fx0 = A->fx0; K.phid0 = A->phid0; K.incr = A->incr; K.a0 = A->a;
    K.strong = A->strong; K.fdwithf = A->fdwithf; K.store = A->no_store ? 0 : 1;
    K.glob = sc.work.blockvals + (size_t)(kScratchRows - 1) * red::kTopMax;   // 4 doubles nobody else uses during a search
    K.result = A->result_dev;
    const flgpu_comm *comm = (const flgpu_comm *)A->comm;
    K.G = 1; K.me = 0; K.dseq = nullptr; K.timeout_ns = 0;
    if (comm && comm->nranks > 1) {
        if (!comm->p2p) fatal("device-resident line search on row shards needs the peer-memory exchange");
        K.peers = comm->peers_search; K.me = comm->rank; K.G = comm->nranks; K.dseq = &comm->local->dseq;
        K.timeout_ns = comm->timeout_ns;
    }
    void *params[] = {&K};
    FLGPU_CUDA_CHECK(cudaLaunchCooperativeKernel((void *)k::search_kernel<KIND, FAST>, dim3(grid), dim3(k::kThreads), params, 0, s));
}
template <int KIND>
static void dev_search(const flgpu_eval_ctx *c, const flgpu_search_args *A, int64_t n) {
    if (A->policy == FLGPU_LS_FAST) dev_search_policy<KIND, true>(c, A, n);
    else dev_search_policy<KIND, false>(c, A, n);
}

// reference-ABI flavour: device x / f' pointers, host f, runs on the current call's stream
static double ref_eval(int kind, bool want_f, double *g, const double *x, int dim) {
    cudaStream_t s = (cudaStream_t)flgpu_current_stream();
    Scratch &sc = scratch_for(s, 1);
    launch_objective(kind, false, (want_f ? FLGPU_WANT_F : 0) | (g ? FLGPU_WRITE_G : 0), want_f ? sc.scalar : nullptr,
                     nullptr, nullptr, g, x, nullptr, 0.0, dim, 0, dim, s);
    if (!want_f) return 0.0;
    FLGPU_CUDA_CHECK(cudaMemcpyAsync(sc.host_scalar, sc.scalar, sizeof(double), cudaMemcpyDeviceToHost, s));
    FLGPU_CUDA_CHECK(cudaStreamSynchronize(s));
    return sc.host_scalar[0];
}
template <int KIND>
static void ref_f(double *fx, const double *x, const int *dim) { *fx = ref_eval(KIND, true, nullptr, x, *dim); }
template <int KIND>
static void ref_fd(double *fdx, const double *x, const int *dim) { ref_eval(KIND, false, fdx, x, *dim); }
template <int KIND>
static int ref_ffd(double *fx, double *fdx, const double *x, const int *dim) {
    *fx = ref_eval(KIND, true, fdx, x, *dim);
    return 0;
}

}  // namespace flgpu

using namespace flgpu;

extern "C" int flgpu_comm_search_exchange(const flgpu_comm *c, void *stream, void *out, size_t out_bytes) {
    if (out_bytes != sizeof(k::SearchExchange)) return 2;
    if (!c || !c->p2p) return 1;
    Scratch &sc = scratch_for((cudaStream_t)stream, 1);
    k::SearchExchange E;
    E.peers = c->peers_search; E.me = c->rank; E.G = c->nranks; E.dseq = &c->local->dseq; E.timeout_ns = c->timeout_ns;
    E.glob = sc.work.blockvals + (size_t)(kScratchRows - 1) * red::kTopMax;   // the four doubles the built-in search uses
    std::memcpy(out, &E, sizeof E);
    return 0;
}

extern "C" int flgpu_builtin_problem(int kind, flgpu_problem *out) {
    out->user = nullptr;
    switch (kind) {
    case FLGPU_OBJ_QUARTIC: out->f = dev_f<0>; out->fd = dev_fd<0>; out->f_fd = dev_ffd<0>; out->fused = dev_fused<0>; out->search = dev_search<0>; out->search_caps = FLGPU_SEARCH_ROW_SHARDS; out->update = dev_update<0>; out->direction = dev_direction<0>; out->fused_multi = dev_fused_multi<0>; return 0;
    case FLGPU_OBJ_ROSENBROCK: out->f = dev_f<1>; out->fd = dev_fd<1>; out->f_fd = dev_ffd<1>; out->fused = dev_fused<1>; out->search = dev_search<1>; out->search_caps = FLGPU_SEARCH_ROW_SHARDS; out->update = dev_update<1>; out->direction = dev_direction<1>; out->fused_multi = dev_fused_multi<1>; return 0;
    case FLGPU_OBJ_DIAGQUAD: out->f = dev_f<2>; out->fd = dev_fd<2>; out->f_fd = dev_ffd<2>; out->fused = dev_fused<2>; out->search = dev_search<2>; out->search_caps = FLGPU_SEARCH_ROW_SHARDS; out->update = dev_update<2>; out->direction = dev_direction<2>; out->fused_multi = dev_fused_multi<2>; return 0;
    case FLGPU_OBJ_QUARTIC_SHIFTED: out->f = dev_f<3>; out->fd = dev_fd<3>; out->f_fd = dev_ffd<3>; out->fused = dev_fused<3>; out->search = dev_search<3>; out->search_caps = FLGPU_SEARCH_ROW_SHARDS; out->update = dev_update<3>; out->direction = dev_direction<3>; out->fused_multi = dev_fused_multi<3>; return 0;
    }
    return 1;
}

// fused evaluation registered for the built-in reference-ABI callbacks (flgpu_register_fused)
namespace flgpu {
flgpu_fused_fn builtin_fused_for(flgpu_ref_f_fn f) {
    if (f == ref_f<0>) return dev_fused<0>;
    if (f == ref_f<1>) return dev_fused<1>;
    if (f == ref_f<2>) return dev_fused<2>;
    if (f == ref_f<3>) return dev_fused<3>;
    return nullptr;
}
flgpu_fused_multi_fn builtin_fused_multi_for(flgpu_ref_f_fn f) {
    if (f == ref_f<0>) return dev_fused_multi<0>;
    if (f == ref_f<1>) return dev_fused_multi<1>;
    if (f == ref_f<2>) return dev_fused_multi<2>;
    if (f == ref_f<3>) return dev_fused_multi<3>;
    return nullptr;
}
flgpu_direction_fn builtin_direction_for(flgpu_ref_f_fn f) {
    if (f == ref_f<0>) return dev_direction<0>;
    if (f == ref_f<1>) return dev_direction<1>;
    if (f == ref_f<2>) return dev_direction<2>;
    if (f == ref_f<3>) return dev_direction<3>;
    return nullptr;
}
flgpu_update_fn builtin_update_for(flgpu_ref_f_fn f) {
    if (f == ref_f<0>) return dev_update<0>;
    if (f == ref_f<1>) return dev_update<1>;
    if (f == ref_f<2>) return dev_update<2>;
    if (f == ref_f<3>) return dev_update<3>;
    return nullptr;
}
}  // namespace flgpu

extern "C" int flgpu_builtin_ref_callbacks(int kind, flgpu_ref_f_fn *f, flgpu_ref_fd_fn *fd, flgpu_ref_f_fd_fn *f_fd) {
    switch (kind) {
    case FLGPU_OBJ_QUARTIC: *f = ref_f<0>; *fd = ref_fd<0>; *f_fd = ref_ffd<0>; return 0;
    case FLGPU_OBJ_ROSENBROCK: *f = ref_f<1>; *fd = ref_fd<1>; *f_fd = ref_ffd<1>; return 0;
    case FLGPU_OBJ_DIAGQUAD: *f = ref_f<2>; *fd = ref_fd<2>; *f_fd = ref_ffd<2>; return 0;
    case FLGPU_OBJ_QUARTIC_SHIFTED: *f = ref_f<3>; *fd = ref_fd<3>; *f_fd = ref_ffd<3>; return 0;
    }
    return 1;
}

extern "C" int flgpu_fill_start(int start_kind, uint64_t seed, double *x_dev, int64_t offset, int64_t n_local,
                                int64_t n_global, void *stream) {
    (void)n_global;
    require_device();
    int64_t need = (n_local + k::kThreads - 1) / k::kThreads;
    if (need < 1) need = 1;
    const int grid = (int)(need < k::kMaxGrid ? need : k::kMaxGrid);
    k::start_kernel<<<grid, k::kThreads, 0, (cudaStream_t)stream>>>(start_kind, seed, x_dev, offset, n_local);
    FLGPU_CUDA_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int flgpu_reduction_workspace(void *stream, int64_t nchunks, double **partials, int64_t *stride) {
    require_device();
    Scratch &sc = scratch_for((cudaStream_t)stream, nchunks < 1 ? 1 : nchunks);
    *partials = sc.work.partials;
    *stride = sc.work.stride;
    return 0;
}

extern "C" int flgpu_reduce_tree(void *stream, int64_t nchunks, int nrows, double *const *out_dev) {
    require_device();
    if (nrows < 1 || nrows > kScratchRows) fatal("flgpu_reduce_tree: between 1 and 8 rows");
    Scratch &sc = scratch_for((cudaStream_t)stream, nchunks < 1 ? 1 : nchunks);
    launch_tree(sc.work, nchunks < 1 ? 1 : nchunks, nrows, out_dev, (cudaStream_t)stream);
    FLGPU_CUDA_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int64_t flgpu_chunk_elems(int64_t n_global) { return red::chunk_elems(n_global); }

// a.b over a shard of a vector of n_global elements: this rank's root of the partition-independent tree
extern "C" int flgpu_vec_dot_sharded(const double *a_dev, const double *b_dev, int64_t n, int64_t n_global, double *out_dev,
                                     void *stream) {
    require_device();
    require_aligned16(a_dev, "flgpu_vec_dot: a"); require_aligned16(b_dev, "flgpu_vec_dot: b");
    cudaStream_t s = (cudaStream_t)stream;
    if (n_global < n) n_global = n;
    const int64_t ch = red::chunk_elems(n_global), nchunks = red::num_chunks(n, ch);
    Scratch &sc = scratch_for(s, nchunks);
    const bool in_kernel = nchunks <= red::kBlockChunks;
    k::dot_kernel<<<chunk_grid(nchunks, 8), k::kThreads, 0, s>>>(a_dev, b_dev, n, ch, sc.work, 0, in_kernel ? out_dev : nullptr);
    if (!in_kernel) { double *out[1] = {out_dev}; launch_tree(sc.work, nchunks, 1, out, s); }
    FLGPU_CUDA_CHECK(cudaGetLastError());
    return 0;
}
extern "C" int flgpu_vec_dot(const double *a_dev, const double *b_dev, int64_t n, double *out_dev, void *stream) {
    return flgpu_vec_dot_sharded(a_dev, b_dev, n, n, out_dev, stream);
}

extern "C" int flgpu_vec_trial(double *x_dev, const double *x0_dev, const double *p_dev, double a, int64_t n,
                               void *stream) {
    require_device();
    require_aligned16(x_dev, "flgpu_vec_trial: x"); require_aligned16(x0_dev, "flgpu_vec_trial: x0");
    require_aligned16(p_dev, "flgpu_vec_trial: p");
    k::trial_kernel<<<unit_grid(n), k::kThreads, 0, (cudaStream_t)stream>>>(x_dev, x0_dev, p_dev, a, n);
    FLGPU_CUDA_CHECK(cudaGetLastError());
    return 0;
}
