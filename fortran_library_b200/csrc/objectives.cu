// objectives.cu -- K6: the benchmark objectives of BASELINE.json as CUDA kernels (f and f' in one
// pass over x, f reduced deterministically), plus index-generated start vectors.
//
//   quartic     f = sum x^4, f' = 4 x^3            (the reference's test objective, test/test.f90:630-663;
//                                                    x**4 = (x*x)*(x*x), x**3 = (x*x)*x as gfortran expands them)
//   Rosenbrock  f = sum_j 100 (x_{2j+1} - x_{2j}^2)^2 + (1 - x_{2j})^2   (extended, pairwise)
//   diag quad   f = 1/2 sum d_i (x_i - 1)^2, d_i log-uniform in [1, 1e6]
//
// Element-wise arithmetic uses separate multiply/add roundings in the same order as the CPU
// oracle's objectives so that f' agrees bit for bit for the same x.
#include <cmath>
#include <map>
#include <mutex>

#include "backend_cuda.cuh"

namespace flgpu {

// ---- per-stream scratch for library kernels launched outside a CudaBackend (callbacks, primitives)
struct Scratch {
    k::Work work;
    double *scalar;       // device double[4]
    double *host_scalar;  // pinned double[4]
    double *tables;       // device double[768], diag-quad factors
};
static std::mutex g_scratch_mu;
static std::map<std::pair<int, void *>, Scratch> g_scratch;

Scratch &scratch_for(cudaStream_t s) {
    int dev = 0;
    FLGPU_CUDA_CHECK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_scratch_mu);
    auto key = std::make_pair(dev, (void *)s);
    auto it = g_scratch.find(key);
    if (it != g_scratch.end()) return it->second;
    Scratch sc;
    FLGPU_CUDA_CHECK(cudaMalloc((void **)&sc.work.partials, (size_t)k::kMaxGrid * 8 * sizeof(double)));
    FLGPU_CUDA_CHECK(cudaMalloc((void **)&sc.work.ticket, 64));
    FLGPU_CUDA_CHECK(cudaMemset(sc.work.ticket, 0, 64));
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
    return g_scratch.emplace(key, sc).first->second;
}

namespace k {

__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dadd_rn(a, -b); }

struct ObjArgs {
    const double *x;
    double *g;          // may be null (f only)
    double *f_out;      // device scalar, may be null (f' only)
    int64_t n, offset, n_global;
    double scale;       // diag quad: 2^24/(n_global-1)
    const double *tables;
    Work w;
};

template <int KIND, bool WANT_F, bool WANT_G>
__global__ void __launch_bounds__(kThreads) objective_kernel(ObjArgs a) {
    __shared__ double tab[KIND == FLGPU_OBJ_DIAGQUAD ? 768 : 1];
    if (KIND == FLGPU_OBJ_DIAGQUAD) {
        for (int i = threadIdx.x; i < 768; i += kThreads) tab[i] = a.tables[i];
        __syncthreads();
    }
    auto coeff = [&](int64_t i) -> double {
        if (a.n_global <= 1) return 1.0;
        const unsigned long long q = (unsigned long long)mul((double)i, a.scale);
        if (q >> 24) return 1.0e6;
        return mul(mul(tab[512 + ((q >> 16) & 255)], tab[256 + ((q >> 8) & 255)]), tab[q & 255]);
    };
    double fsum = 0.0;
    const int64_t nu = a.n >> 1;
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    for (int64_t u = (int64_t)blockIdx.x * kThreads + threadIdx.x; u < nu; u += stride) {
        const double2 x = ld2(a.x, u);
        double2 g;
        if (KIND == FLGPU_OBJ_QUARTIC) {
            const double x2 = mul(x.x, x.x), y2 = mul(x.y, x.y);
            if (WANT_F) { fsum += mul(x2, x2); fsum += mul(y2, y2); }
            g.x = mul(4.0, mul(x2, x.x)); g.y = mul(4.0, mul(y2, x.y));
        } else if (KIND == FLGPU_OBJ_ROSENBROCK) {
            const double t1 = sub(x.y, mul(x.x, x.x)), t2 = sub(1.0, x.x);
            if (WANT_F) fsum += add(mul(mul(100.0, t1), t1), mul(t2, t2));
            g.x = sub(mul(mul(-400.0, x.x), t1), mul(2.0, t2));
            g.y = mul(200.0, t1);
        } else {
            const int64_t i = a.offset + 2 * u;
            const double d0 = coeff(i), d1 = coeff(i + 1);
            const double t0 = sub(x.x, 1.0), t1 = sub(x.y, 1.0);
            if (WANT_F) { fsum += mul(mul(mul(0.5, d0), t0), t0); fsum += mul(mul(mul(0.5, d1), t1), t1); }
            g.x = mul(d0, t0); g.y = mul(d1, t1);
        }
        if (WANT_G) st2(a.g, u, g);
    }
    if ((a.n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const int64_t i = a.n - 1;
        const double x = a.x[i];
        double g;
        if (KIND == FLGPU_OBJ_QUARTIC) {
            const double x2 = mul(x, x);
            if (WANT_F) fsum += mul(x2, x2);
            g = mul(4.0, mul(x2, x));
        } else if (KIND == FLGPU_OBJ_ROSENBROCK) {   // unpaired last element
            const double t2 = sub(1.0, x);
            if (WANT_F) fsum += mul(t2, t2);
            g = mul(-2.0, t2);
        } else {
            const double d = coeff(a.offset + i), t = sub(x, 1.0);
            if (WANT_F) fsum += mul(mul(mul(0.5, d), t), t);
            g = mul(d, t);
        }
        if (WANT_G) a.g[i] = g;
    }
    if (WANT_F) {
        double acc[1] = {fsum};
        const int d[1] = {0};
        reduce_finish<1>(acc, d, a.w, a.f_out);
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
static int obj_grid(int64_t n) {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    int64_t need = (n / 2 + k::kThreads) / k::kThreads;
    int64_t g = (int64_t)sms * 8;
    if (g > k::kMaxGrid) g = k::kMaxGrid;
    return (int)(need < g ? (need < 1 ? 1 : need) : g);
}

void launch_objective(int kind, double *f_dev, double *g_dev, const double *x_dev, int64_t n, int64_t offset,
                      int64_t n_global, cudaStream_t s) {
    Scratch &sc = scratch_for(s);
    k::ObjArgs a;
    a.x = x_dev; a.g = g_dev; a.f_out = f_dev; a.n = n; a.offset = offset; a.n_global = n_global;
    a.scale = n_global > 1 ? 16777216.0 / (double)(n_global - 1) : 0.0;
    a.tables = sc.tables; a.w = sc.work;
    const int grid = obj_grid(n);
#define FLGPU_OBJ_LAUNCH(KIND)                                                                              \
    do {                                                                                                    \
        if (f_dev && g_dev) k::objective_kernel<KIND, true, true><<<grid, k::kThreads, 0, s>>>(a);          \
        else if (f_dev) k::objective_kernel<KIND, true, false><<<grid, k::kThreads, 0, s>>>(a);             \
        else k::objective_kernel<KIND, false, true><<<grid, k::kThreads, 0, s>>>(a);                        \
    } while (0)
    switch (kind) {
    case FLGPU_OBJ_QUARTIC: FLGPU_OBJ_LAUNCH(FLGPU_OBJ_QUARTIC); break;
    case FLGPU_OBJ_ROSENBROCK: FLGPU_OBJ_LAUNCH(FLGPU_OBJ_ROSENBROCK); break;
    case FLGPU_OBJ_DIAGQUAD: FLGPU_OBJ_LAUNCH(FLGPU_OBJ_DIAGQUAD); break;
    default: fatal("unknown built-in objective");
    }
#undef FLGPU_OBJ_LAUNCH
}

// 64-bit device-callback flavour
template <int KIND>
static void dev_f(const flgpu_eval_ctx *c, double *f, const double *x, int64_t n) {
    launch_objective(KIND, f, nullptr, x, n, c->offset, c->n_global, (cudaStream_t)c->stream);
}
template <int KIND>
static void dev_fd(const flgpu_eval_ctx *c, double *g, const double *x, int64_t n) {
    launch_objective(KIND, nullptr, g, x, n, c->offset, c->n_global, (cudaStream_t)c->stream);
}
template <int KIND>
static void dev_ffd(const flgpu_eval_ctx *c, double *f, double *g, const double *x, int64_t n) {
    launch_objective(KIND, f, g, x, n, c->offset, c->n_global, (cudaStream_t)c->stream);
}

// reference-ABI flavour: device x / f' pointers, host f, runs on the current call's stream
static double ref_eval(int kind, bool want_f, double *g, const double *x, int dim) {
    cudaStream_t s = (cudaStream_t)flgpu_current_stream();
    Scratch &sc = scratch_for(s);
    launch_objective(kind, want_f ? sc.scalar : nullptr, g, x, dim, 0, dim, s);
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

extern "C" int flgpu_builtin_problem(int kind, flgpu_problem *out) {
    out->user = nullptr;
    switch (kind) {
    case FLGPU_OBJ_QUARTIC: out->f = dev_f<0>; out->fd = dev_fd<0>; out->f_fd = dev_ffd<0>; return 0;
    case FLGPU_OBJ_ROSENBROCK: out->f = dev_f<1>; out->fd = dev_fd<1>; out->f_fd = dev_ffd<1>; return 0;
    case FLGPU_OBJ_DIAGQUAD: out->f = dev_f<2>; out->fd = dev_fd<2>; out->f_fd = dev_ffd<2>; return 0;
    }
    return 1;
}

extern "C" int flgpu_builtin_ref_callbacks(int kind, flgpu_ref_f_fn *f, flgpu_ref_fd_fn *fd, flgpu_ref_f_fd_fn *f_fd) {
    switch (kind) {
    case FLGPU_OBJ_QUARTIC: *f = ref_f<0>; *fd = ref_fd<0>; *f_fd = ref_ffd<0>; return 0;
    case FLGPU_OBJ_ROSENBROCK: *f = ref_f<1>; *fd = ref_fd<1>; *f_fd = ref_ffd<1>; return 0;
    case FLGPU_OBJ_DIAGQUAD: *f = ref_f<2>; *fd = ref_fd<2>; *f_fd = ref_ffd<2>; return 0;
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

extern "C" int flgpu_vec_dot(const double *a_dev, const double *b_dev, int64_t n, double *out_dev, void *stream) {
    require_device();
    cudaStream_t s = (cudaStream_t)stream;
    Scratch &sc = scratch_for(s);
    k::dot_kernel<<<obj_grid(n), k::kThreads, 0, s>>>(a_dev, b_dev, n, sc.work, out_dev, 0);
    FLGPU_CUDA_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int flgpu_vec_trial(double *x_dev, const double *x0_dev, const double *p_dev, double a, int64_t n,
                               void *stream) {
    require_device();
    k::trial_kernel<<<obj_grid(n), k::kThreads, 0, (cudaStream_t)stream>>>(x_dev, x0_dev, p_dev, a, n);
    FLGPU_CUDA_CHECK(cudaGetLastError());
    return 0;
}
