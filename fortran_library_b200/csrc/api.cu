// api.cu -- the C-ABI of libflgpu.so (include/flgpu.h): flgpu_* entry points and the reference's
// own compiled symbol names (__nonlinearoptimization_MOD_* / nonlinearoptimization_mp_*_).
#include <atomic>
#include <map>
#include <mutex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <vector>

#include "api_internal.hpp"
#include "backend_cuda.cuh"
#include "driver.hpp"

using namespace flgpu;

namespace flgpu {
flgpu_fused_fn builtin_fused_for(flgpu_ref_f_fn f);   // objectives.cu
flgpu_update_fn builtin_update_for(flgpu_ref_f_fn f);
flgpu_direction_fn builtin_direction_for(flgpu_ref_f_fn f);
flgpu_fused_multi_fn builtin_fused_multi_for(flgpu_ref_f_fn f);
extern int g_k1_shape[2];                             // backend_cuda.cu
}

namespace flgpu_api {

// flgpu_register_fused: reference-ABI objective -> fused line-search evaluation
struct FusedEntry { flgpu_fused_fn fn; void *user; };
std::mutex g_fused_mu;
std::map<flgpu_ref_f_fn, FusedEntry> g_fused;

struct ThreadState {
    void *stream = nullptr;
    int device = -1;
    flgpu_stats last{};
    std::vector<KernelTime> times;
    flgpu_observer_fn observer = nullptr;
    void *observer_user = nullptr;
    CudaBackend *backend = nullptr;   // the call currently executing on this thread
    int line_search = -1;             // flgpu_set_line_search; -1 = FLGPU_LINE_SEARCH decides
};
thread_local ThreadState tls;

std::atomic<int> g_x_space{-1}, g_cb_space{-1};

int space_from_env(const char *name, int dflt) {
    const char *v = std::getenv(name);
    if (!v) return dflt;
    if (!std::strcmp(v, "host") || !std::strcmp(v, "HOST")) return FLGPU_SPACE_HOST;
    if (!std::strcmp(v, "device") || !std::strcmp(v, "DEVICE")) return FLGPU_SPACE_DEVICE;
    return dflt;
}
int x_space_now() {
    int v = g_x_space.load();
    return v >= 0 ? v : space_from_env("FLGPU_X_SPACE", FLGPU_SPACE_HOST);
}
// -1 = not chosen (flgpu_set_callback_space / FLGPU_CALLBACK_SPACE): ref_adapter_init then picks HOST -- what code
// compiled against the reference expects: x and fdx it can dereference -- unless the objective is one the library knows
// to be a device callback (its built-ins, anything registered with flgpu_register_fused)
int cb_space_now() {
    int v = g_cb_space.load();
    return v >= 0 ? v : space_from_env("FLGPU_CALLBACK_SPACE", -1);
}

double wall_ms() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

int run(int algo, const flgpu_problem *prob, const flgpu_options *opt, double *x, int64_t n, int x_space,
        flgpu_stats *stats) {
    require_device();
    if (!prob || !prob->f || !prob->fd) fatal("flgpu: f and fd callbacks are required (f90:40)");
    if (n < 0) fatal("flgpu: negative dimension");
    if (algo == ALGO_LBFGS && opt->memory > FLGPU_MAX_MEMORY) {
        // the reference accepts any Memory >= 1 (f90:419, 435); K2 / K3 hold their tables for at most 64 pairs
        std::fprintf(stderr, "flgpu: LBFGS Memory = %d exceeds FLGPU_MAX_MEMORY = %d; nothing was done (x is unchanged)\n",
                     opt->memory, FLGPU_MAX_MEMORY);
        if (stats) { std::memset(stats, 0, sizeof *stats); stats->status = FLGPU_INVALID_ARGUMENT; }
        return FLGPU_ERR_MEMORY_LIMIT;
    }
    const char *tr = std::getenv("FLGPU_TRACE_PHASES");   // wall-clock phases of one call, to stderr
    const bool trace = tr && tr[0] && tr[0] != '0';
    const double t0 = wall_ms();
    double t1, t2;
    flgpu_stats st;
    ThreadState saved = tls;
    {
        CudaBackend B(*prob, n, *opt);
        Params P = params_from_options(*opt, algo == ALGO_CG, prob->f_fd != nullptr);
        tls.stream = B.stream_handle();
        tls.backend = &B;
        FLGPU_CUDA_CHECK(cudaGetDevice(&tls.device));
        t1 = wall_ms();
        if (algo == ALGO_CG) run_cg(B, P, x, x_space, &st);
        else if (algo == ALGO_SD) run_sd(B, P, x, x_space, &st);
        else run_lbfgs(B, P, x, x_space, &st);
        B.resolve_times();
        tls.times = B.times;
        t2 = wall_ms();
        if (trace)
            std::fprintf(stderr, "flgpu phases: cudaMalloc %.1f ms, x upload %.1f ms, x download %.1f ms\n", B.alloc_ms,
                         B.upload_ms, B.download_ms);
    }   // work space released here (~CudaBackend)
    const double t3 = wall_ms();
    tls.stream = saved.stream;
    tls.device = saved.device;
    tls.backend = saved.backend;
    tls.last = st;
    if (stats) *stats = st;
    if (trace)
        std::fprintf(stderr, "flgpu phases: setup %.1f ms, optimise (incl. work-space allocation, x transfers) %.1f ms, "
                             "release %.1f ms; %lld iterations, %lld trials\n",
                     t1 - t0, t2 - t1, t3 - t2, (long long)st.iterations, (long long)st.n_trials);
    return 0;
}

// ---- adapter: reference-ABI callbacks (f90:33-38) behind the 64-bit device-callback interface
// (struct RefAdapter: api_internal.hpp)
void ad_fused(const flgpu_eval_ctx *c, int flags, double *f_dev, double *gp_dev, double *x_out, double *g_out,
              const double *x0, const double *p, double a, int64_t n) {
    const RefAdapter *A = (const RefAdapter *)c->user;
    flgpu_eval_ctx inner = *c;
    inner.user = A->fused_user;
    A->fused(&inner, flags, f_dev, gp_dev, x_out, g_out, x0, p, a, n);
}
void ad_update(const flgpu_eval_ctx *c, const flgpu_update_args *args, int64_t n) {
    const RefAdapter *A = (const RefAdapter *)c->user;
    flgpu_eval_ctx inner = *c;
    inner.user = A->fused_user;
    A->update(&inner, args, n);
}
void ad_direction(const flgpu_eval_ctx *c, const flgpu_direction_args *args, int64_t n) {
    const RefAdapter *A = (const RefAdapter *)c->user;
    flgpu_eval_ctx inner = *c;
    inner.user = A->fused_user;
    A->direction(&inner, args, n);
}
void ad_fused_multi(const flgpu_eval_ctx *c, int count, const double *steps, double *out_dev, const double *x0,
                    const double *p, int64_t n) {
    const RefAdapter *A = (const RefAdapter *)c->user;
    flgpu_eval_ctx inner = *c;
    inner.user = A->fused_user;
    A->fused_multi(&inner, count, steps, out_dev, x0, p, n);
}
void to_host(const RefAdapter *A, const double *x_dev, int64_t n, cudaStream_t s) {
    FLGPU_CUDA_CHECK(cudaMemcpyAsync(A->xh, x_dev, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s));
    FLGPU_CUDA_CHECK(cudaStreamSynchronize(s));
}
void ad_f(const flgpu_eval_ctx *c, double *f_dev, const double *x, int64_t n) {
    const RefAdapter *A = (const RefAdapter *)c->user;
    cudaStream_t s = (cudaStream_t)c->stream;
    int dim = (int)n;
    double fx = 0.0;
    if (A->cb_space == FLGPU_SPACE_HOST) { to_host(A, x, n, s); A->f(&fx, A->xh, &dim); }
    else A->f(&fx, x, &dim);
    k::set_scalar_kernel<<<1, 1, 0, s>>>(f_dev, fx);
}
void ad_fd(const flgpu_eval_ctx *c, double *g, const double *x, int64_t n) {
    const RefAdapter *A = (const RefAdapter *)c->user;
    cudaStream_t s = (cudaStream_t)c->stream;
    int dim = (int)n;
    if (A->cb_space == FLGPU_SPACE_HOST) {
        to_host(A, x, n, s);
        A->fd(A->gh, A->xh, &dim);
        FLGPU_CUDA_CHECK(cudaMemcpyAsync(g, A->gh, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, s));
    } else {
        A->fd(g, x, &dim);
    }
}
void ad_ffd(const flgpu_eval_ctx *c, double *f_dev, double *g, const double *x, int64_t n) {
    const RefAdapter *A = (const RefAdapter *)c->user;
    cudaStream_t s = (cudaStream_t)c->stream;
    int dim = (int)n;
    double fx = 0.0;
    if (A->cb_space == FLGPU_SPACE_HOST) {
        to_host(A, x, n, s);
        (void)A->f_fd(&fx, A->gh, A->xh, &dim);
        FLGPU_CUDA_CHECK(cudaMemcpyAsync(g, A->gh, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, s));
    } else {
        (void)A->f_fd(&fx, g, x, &dim);
    }
    k::set_scalar_kernel<<<1, 1, 0, s>>>(f_dev, fx);
}

void ref_adapter_init(RefAdapter &A, flgpu_ref_f_fn f, flgpu_ref_fd_fn fd, flgpu_ref_f_fd_fn f_fd, int dim,
                      flgpu_problem *prob) {
    A.f = f; A.fd = fd; A.f_fd = f_fd; A.cb_space = cb_space_now();
    if (A.cb_space < 0) {
        bool known_device = builtin_fused_for(f) != nullptr;
        if (!known_device) {
            std::lock_guard<std::mutex> lock(g_fused_mu);
            known_device = g_fused.find(f) != g_fused.end();
        }
        A.cb_space = known_device ? FLGPU_SPACE_DEVICE : FLGPU_SPACE_HOST;
    }
    if (A.cb_space == FLGPU_SPACE_HOST) {
        FLGPU_CUDA_CHECK(cudaMallocHost((void **)&A.xh, sizeof(double) * (size_t)(dim > 0 ? dim : 1)));
        FLGPU_CUDA_CHECK(cudaMallocHost((void **)&A.gh, sizeof(double) * (size_t)(dim > 0 ? dim : 1)));
    }
    prob->f = ad_f; prob->fd = ad_fd; prob->f_fd = f_fd ? ad_ffd : nullptr; prob->user = &A;
    prob->fused = nullptr;
    prob->search = nullptr;
    prob->search_caps = 0;
    prob->update = nullptr;
    prob->direction = nullptr;
    prob->fused_multi = nullptr;
    if (A.cb_space == FLGPU_SPACE_DEVICE) {
        {
            std::lock_guard<std::mutex> lock(g_fused_mu);
            auto it = g_fused.find(f);
            if (it != g_fused.end()) { A.fused = it->second.fn; A.fused_user = it->second.user; }
        }
        if (!A.fused) { A.fused = builtin_fused_for(f); A.update = builtin_update_for(f); A.direction = builtin_direction_for(f);
                        A.fused_multi = builtin_fused_multi_for(f); }
        if (A.fused) prob->fused = ad_fused;
        if (A.fused && A.update) prob->update = ad_update;
        if (A.fused && A.direction) prob->direction = ad_direction;
        if (A.fused && A.fused_multi) prob->fused_multi = ad_fused_multi;
    }
}
void ref_adapter_free(RefAdapter &A) {
    if (A.xh) cudaFreeHost(A.xh);
    if (A.gh) cudaFreeHost(A.gh);
    A.xh = A.gh = nullptr;
}
void apply_thread_settings(flgpu_options &o) {
    o.observer = tls.observer;
    o.observer_user = tls.observer_user;
    const char *nf = std::getenv("FLGPU_NO_FUSED");
    if (nf && nf[0] && nf[0] != '0') o.no_fused = 1;
    const char *ds = std::getenv("FLGPU_DEVICE_SEARCH");   // 0 / 1 / 2 as flgpu_options.device_search
    if (ds && ds[0] >= '0' && ds[0] <= '2') o.device_search = ds[0] - '0';
    if (tls.line_search >= 0) {
        o.line_search = tls.line_search;
    } else {
        const char *ls = std::getenv("FLGPU_LINE_SEARCH");   // reference | fast
        if (ls && (!std::strcmp(ls, "fast") || !std::strcmp(ls, "FAST") || !std::strcmp(ls, "1")))
            o.line_search = FLGPU_LS_FAST;
    }
}

void run_ref(int algo, flgpu_ref_f_fn f, flgpu_ref_fd_fn fd, flgpu_ref_f_fd_fn f_fd, double *x, int dim,
             flgpu_options &o) {
    require_device();
    RefAdapter A;
    flgpu_problem prob;
    ref_adapter_init(A, f, fd, f_fd, dim, &prob);
    apply_thread_settings(o);
    run(algo, &prob, &o, x, dim, x_space_now(), nullptr);
    ref_adapter_free(A);
}

// Fortran OPTIONAL -> options (NULL = absent keeps the default, f90:417-434 / 212-229)
void fill_optional(flgpu_options &o, const int32_t *Strong, const int32_t *Warning, const int *MaxIteration,
                   const double *Precision, const double *MinStepLength, const double *WolfeConst1,
                   const double *WolfeConst2, const double *Increment) {
    if (Strong) o.strong = *Strong != 0;
    if (Warning) o.warning = *Warning != 0;
    if (MaxIteration) o.max_iteration = *MaxIteration;
    if (Precision) o.precision = *Precision;
    if (MinStepLength) o.min_step_length = *MinStepLength;
    if (WolfeConst1) o.wolfe_c1 = *WolfeConst1;
    if (WolfeConst2) o.wolfe_c2 = *WolfeConst2;
    if (Increment) o.increment = *Increment;
}

// character(*) Method compared as 'DY' / 'PR' (f90:207,214,240,311,345; f90:2273-2297)
int parse_method(const char *Method, int len, bool whole_string) {
    char t[2] = {' ', ' '};
    if (Method) { if (len > 0) t[0] = Method[0]; if (len > 1) t[1] = Method[1]; }
    bool tail_blank = true;
    if (whole_string && Method) for (int j = 2; j < len; j++) if (Method[j] != ' ') tail_blank = false;
    if (t[0] == 'D' && t[1] == 'Y' && tail_blank) return FLGPU_CG_DY;
    if (t[0] == 'P' && t[1] == 'R' && tail_blank) return FLGPU_CG_PR;
    std::string name(Method ? Method : "", Method ? (size_t)(len > 0 ? len : 0) : 0);
    std::printf(" Program abort: unsupported conjugate gradient method %s\n", name.c_str());
    std::fflush(stdout);
    std::exit(1);  // the reference executes `stop` (f90:345)
}

void set_last_stats(const flgpu_stats &st) { tls.last = st; }

}  // namespace flgpu_api
using namespace flgpu_api;

extern "C" {

const char *flgpu_version(void) { return "flgpu 0.1 (sm_100a, fp64)"; }

void flgpu_options_default(flgpu_options *o, int for_cg) {
    std::memset(o, 0, sizeof *o);
    o->memory = 10;
    o->method = FLGPU_CG_DY;
    o->strong = 1;
    o->warning = 1;
    o->max_iteration = 1000;
    o->precision = 1e-15;
    o->min_step_length = 1e-15;
    o->wolfe_c1 = 1e-4;
    o->wolfe_c2 = for_cg ? 0.45 : 0.9;
    o->increment = 1.05;
    o->device_search = 2;
}

int flgpu_lbfgs(const flgpu_problem *prob, const flgpu_options *opt, double *x, int64_t n_local, int x_space,
                flgpu_stats *stats) {
    return run(ALGO_LBFGS, prob, opt, x, n_local, x_space, stats);
}
int flgpu_conjugate_gradient(const flgpu_problem *prob, const flgpu_options *opt, double *x, int64_t n_local,
                             int x_space, flgpu_stats *stats) {
    return run(ALGO_CG, prob, opt, x, n_local, x_space, stats);
}
int flgpu_steepest_descent(const flgpu_problem *prob, const flgpu_options *opt, double *x, int64_t n_local,
                           int x_space, flgpu_stats *stats) {
    return run(ALGO_SD, prob, opt, x, n_local, x_space, stats);
}

void flgpu_set_x_space(int space) { g_x_space.store(space); }
void flgpu_set_callback_space(int space) { g_cb_space.store(space); }
void *flgpu_current_stream(void) { return tls.stream; }
int flgpu_current_device(void) { return tls.device; }
void flgpu_last_stats(flgpu_stats *out) { *out = tls.last; }
void flgpu_set_observer(flgpu_observer_fn fn, void *user) { tls.observer = fn; tls.observer_user = user; }
void flgpu_set_line_search(int policy) { tls.line_search = policy; }
void flgpu_register_fused(flgpu_ref_f_fn f, flgpu_fused_fn fused, void *user) {
    std::lock_guard<std::mutex> lock(g_fused_mu);
    if (fused) g_fused[f] = FusedEntry{fused, user};
    else g_fused.erase(f);
}

void flgpu_set_workspace_cache(int on) { flgpu::ws_set_enabled(on != 0); }
void flgpu_release_workspace(void) { flgpu::ws_release(); }

void flgpu_debug_set_k1_shape(int columns_per_group, int groups) {
    flgpu::g_k1_shape[0] = columns_per_group;
    flgpu::g_k1_shape[1] = groups;
}

void flgpu_reset_kernel_times(void) {
    if (!tls.backend) return;
    tls.backend->resolve_times();
    for (auto &kt : tls.backend->times) { kt.ms = 0.0; kt.launches = 0; kt.bytes = 0.0; }
}

int flgpu_kernel_times(const char **names, double *ms, int64_t *launches, double *bytes, int cap) {
    int n = 0;
    for (auto &kt : tls.times) {
        if (n >= cap) break;
        names[n] = kt.name.c_str();
        ms[n] = kt.ms;
        launches[n] = kt.launches;
        bytes[n] = kt.bytes;
        n++;
    }
    return n;
}

void *flgpu_malloc(size_t bytes) {
    require_device();
    void *p = nullptr;
    FLGPU_CUDA_CHECK(cudaMalloc(&p, bytes ? bytes : 1));
    return p;
}
void flgpu_free(void *dev_ptr) { if (dev_ptr) cudaFree(dev_ptr); }
int flgpu_memcpy(void *dst, const void *src, size_t bytes, int dst_space, int src_space, void *stream) {
    require_device();
    cudaMemcpyKind kind = dst_space == FLGPU_SPACE_DEVICE
                              ? (src_space == FLGPU_SPACE_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice)
                              : (src_space == FLGPU_SPACE_DEVICE ? cudaMemcpyDeviceToHost : cudaMemcpyHostToHost);
    FLGPU_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, kind, (cudaStream_t)stream));
    FLGPU_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
}
int flgpu_device_count(void) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) { cudaGetLastError(); return 0; }
    return count;
}

// ---------------------------------------------------------------- Fortran ABI
void __nonlinearoptimization_MOD_lbfgs(flgpu_ref_f_fn f, flgpu_ref_fd_fn fd, double *x, const int *dim,
                                       const int *Memory, flgpu_ref_f_fd_fn f_fd, const int32_t *Strong,
                                       const int32_t *Warning, const int *MaxIteration, const double *Precision,
                                       const double *MinStepLength, const double *WolfeConst1,
                                       const double *WolfeConst2, const double *Increment) {
    flgpu_options o;
    flgpu_options_default(&o, 0);
    if (Memory) o.memory = *Memory;
    fill_optional(o, Strong, Warning, MaxIteration, Precision, MinStepLength, WolfeConst1, WolfeConst2, Increment);
    if (o.memory > FLGPU_MAX_MEMORY) {
        // this signature has no error channel and the reference would simply run (f90:419): run with the largest memory
        // the kernels hold and say so, rather than end the host program
        if (o.warning)
            std::printf(" L-BFGS warning: Memory = %d is above the %d pairs the GPU kernels hold; running with Memory = %d\n",
                        o.memory, FLGPU_MAX_MEMORY, FLGPU_MAX_MEMORY);
        o.memory = FLGPU_MAX_MEMORY;
    }
    run_ref(ALGO_LBFGS, f, fd, f_fd, x, *dim, o);
}

void __nonlinearoptimization_MOD_steepestdescent(flgpu_ref_f_fn f, flgpu_ref_fd_fn fd, double *x, const int *dim,
                                                 flgpu_ref_f_fd_fn f_fd, const int32_t *Strong,
                                                 const int32_t *Warning, const int *MaxIteration,
                                                 const double *Precision, const double *MinStepLength,
                                                 const double *WolfeConst1, const double *WolfeConst2,
                                                 const double *Increment) {
    flgpu_options o;
    flgpu_options_default(&o, 0);   // c2 default 0.9 (f90:84)
    fill_optional(o, Strong, Warning, MaxIteration, Precision, MinStepLength, WolfeConst1, WolfeConst2, Increment);
    run_ref(ALGO_SD, f, fd, f_fd, x, *dim, o);
}
void nonlinearoptimization_mp_steepestdescent_(flgpu_ref_f_fn f, flgpu_ref_fd_fn fd, double *x, const int *dim,
                                               flgpu_ref_f_fd_fn f_fd, const int32_t *Strong, const int32_t *Warning,
                                               const int *MaxIteration, const double *Precision,
                                               const double *MinStepLength, const double *WolfeConst1,
                                               const double *WolfeConst2, const double *Increment) {
    __nonlinearoptimization_MOD_steepestdescent(f, fd, x, dim, f_fd, Strong, Warning, MaxIteration, Precision,
                                                MinStepLength, WolfeConst1, WolfeConst2, Increment);
}

void __nonlinearoptimization_MOD_conjugategradient(flgpu_ref_f_fn f, flgpu_ref_fd_fn fd, double *x, const int *dim,
                                                   const char *Method, flgpu_ref_f_fd_fn f_fd, const int32_t *Strong,
                                                   const int32_t *Warning, const int *MaxIteration,
                                                   const double *Precision, const double *MinStepLength,
                                                   const double *WolfeConst1, const double *WolfeConst2,
                                                   const double *Increment, int len_Method) {
    flgpu_options o;
    flgpu_options_default(&o, 1);
    if (Method) o.method = parse_method(Method, len_Method, false);
    fill_optional(o, Strong, Warning, MaxIteration, Precision, MinStepLength, WolfeConst1, WolfeConst2, Increment);
    run_ref(ALGO_CG, f, fd, f_fd, x, *dim, o);
}

void __nonlinearoptimization_MOD_conjugategradient_basic(flgpu_ref_f_fn f, flgpu_ref_fd_fn fd, double *x,
                                                         const int *dim, const char *Method, const int32_t *Strong,
                                                         const int32_t *Warning, const int *MaxIteration,
                                                         const double *Precision, const double *MinStepLength,
                                                         const double *WolfeConst1, const double *WolfeConst2,
                                                         const double *Increment, int len_Method) {
    flgpu_options o;
    flgpu_options_default(&o, 1);
    o.method = parse_method(Method, len_Method, true);
    o.no_clamp = 1;  // f90:2265-2278: tunables used as given
    fill_optional(o, Strong, Warning, MaxIteration, Precision, MinStepLength, WolfeConst1, WolfeConst2, Increment);
    run_ref(ALGO_CG, f, fd, nullptr, x, *dim, o);
}

void nonlinearoptimization_mp_lbfgs_(flgpu_ref_f_fn f, flgpu_ref_fd_fn fd, double *x, const int *dim,
                                     const int *Memory, flgpu_ref_f_fd_fn f_fd, const int32_t *Strong,
                                     const int32_t *Warning, const int *MaxIteration, const double *Precision,
                                     const double *MinStepLength, const double *WolfeConst1,
                                     const double *WolfeConst2, const double *Increment) {
    __nonlinearoptimization_MOD_lbfgs(f, fd, x, dim, Memory, f_fd, Strong, Warning, MaxIteration, Precision,
                                      MinStepLength, WolfeConst1, WolfeConst2, Increment);
}
void nonlinearoptimization_mp_conjugategradient_(flgpu_ref_f_fn f, flgpu_ref_fd_fn fd, double *x, const int *dim,
                                                 const char *Method, flgpu_ref_f_fd_fn f_fd, const int32_t *Strong,
                                                 const int32_t *Warning, const int *MaxIteration,
                                                 const double *Precision, const double *MinStepLength,
                                                 const double *WolfeConst1, const double *WolfeConst2,
                                                 const double *Increment, int len_Method) {
    __nonlinearoptimization_MOD_conjugategradient(f, fd, x, dim, Method, f_fd, Strong, Warning, MaxIteration,
                                                  Precision, MinStepLength, WolfeConst1, WolfeConst2, Increment,
                                                  len_Method);
}
void nonlinearoptimization_mp_conjugategradient_basic_(flgpu_ref_f_fn f, flgpu_ref_fd_fn fd, double *x,
                                                       const int *dim, const char *Method, const int32_t *Strong,
                                                       const int32_t *Warning, const int *MaxIteration,
                                                       const double *Precision, const double *MinStepLength,
                                                       const double *WolfeConst1, const double *WolfeConst2,
                                                       const double *Increment, int len_Method) {
    __nonlinearoptimization_MOD_conjugategradient_basic(f, fd, x, dim, Method, Strong, Warning, MaxIteration,
                                                        Precision, MinStepLength, WolfeConst1, WolfeConst2,
                                                        Increment, len_Method);
}

}  // extern "C"

// ---------------------------------------------------------------- two-loop recursion as an operator
struct flgpu_history {
    CudaBackend *B;
    History *H;
};

extern "C" flgpu_history *flgpu_history_create(int64_t n_local, int memory, void *stream, flgpu_comm *comm) {
    require_device();
    flgpu_problem none{};
    flgpu_options o;
    flgpu_options_default(&o, 0);
    o.stream = stream;
    o.comm = comm;
    flgpu_history *h = new flgpu_history;
    h->B = new CudaBackend(none, n_local, o);
    h->H = new History(*h->B, memory);
    return h;
}
extern "C" int flgpu_history_push(flgpu_history *h, const double *x1_dev, const double *x0_dev, const double *g1_dev,
                                  const double *g0_dev) {
    require_aligned16(x1_dev, "flgpu_history_push: x1"); require_aligned16(x0_dev, "flgpu_history_push: x0");
    require_aligned16(g1_dev, "flgpu_history_push: g1"); require_aligned16(g0_dev, "flgpu_history_push: g0");
    h->H->push(x1_dev, x0_dev, g1_dev, g0_dev);
    return 0;
}
extern "C" int flgpu_history_direction(flgpu_history *h, const double *g1_dev, const double *x1_dev, double *p_dev,
                                       double *xt_dev, double *gp, double *pp) {
    require_aligned16(g1_dev, "flgpu_history_direction: g1"); require_aligned16(x1_dev, "flgpu_history_direction: x1");
    require_aligned16(p_dev, "flgpu_history_direction: p"); require_aligned16(xt_dev, "flgpu_history_direction: xt");
    h->H->direction(g1_dev, x1_dev, p_dev, xt_dev, gp, pp);
    return 0;
}
extern "C" void flgpu_history_destroy(flgpu_history *h) {
    if (!h) return;
    delete h->H;
    delete h->B;
    delete h;
}
