// backend_cuda.cu -- CudaBackend: memory, launch geometry, stream ordering, NCCL exchange.
#include "backend_cuda.cuh"

#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <cstdint>
#include <cstring>
#include <ctime>
#include <map>
#include <mutex>

namespace flgpu {

static double now_ms() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

// ---- work-space cache (opt-in): device buffers of a finished call are kept for the next call on the same
// device instead of going back to the driver (cudaFree of a 50 GiB work space costs 0.25-0.55 s on B200:
// profiles/r01_e2e_phases.md).  Off by default: like the reference, nothing then persists between calls.
// Two scopes: the process-wide cache (flgpu_set_workspace_cache / FLGPU_WORKSPACE_CACHE) and a per-thread ARENA that
// a composite call (AugmentedLagrangian: one inner solve per outer iteration) opens for its own duration -- it never
// touches the process-wide switch, so concurrent users of the library are unaffected.
namespace {
typedef std::multimap<std::pair<int, size_t>, void *> FreeList;
struct WsCache {
    std::mutex mu;
    FreeList free_;
    std::map<void *, int> device_of;     // owning device, recorded when the buffer was allocated
    bool enabled = false, env_read = false;
} g_ws;
struct WsArena { int depth = 0; FreeList free_; };
thread_local WsArena t_arena;
bool ws_enabled() {
    if (!g_ws.env_read) {
        const char *v = std::getenv("FLGPU_WORKSPACE_CACHE");
        if (v && v[0] && v[0] != '0') g_ws.enabled = true;
        g_ws.env_read = true;
    }
    return g_ws.enabled;
}
void release_list(FreeList &fl) {
    for (auto &kv : fl) {
        cudaFree(kv.second);
        std::lock_guard<std::mutex> lock(g_ws.mu);
        g_ws.device_of.erase(kv.second);
    }
    fl.clear();
}
}  // namespace
void *ws_alloc(size_t bytes) {
    int dev = 0;
    cudaGetDevice(&dev);
    const auto key = std::make_pair(dev, bytes);
    if (t_arena.depth > 0) {
        auto it = t_arena.free_.find(key);
        if (it != t_arena.free_.end()) { void *p = it->second; t_arena.free_.erase(it); return p; }
    }
    {
        std::lock_guard<std::mutex> lock(g_ws.mu);
        auto it = g_ws.free_.find(key);
        if (it != g_ws.free_.end()) { void *p = it->second; g_ws.free_.erase(it); return p; }
    }
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {            // out of memory with buffers parked in a cache: release them and retry
        cudaGetLastError();
        release_list(t_arena.free_);
        ws_release();
        e = cudaMalloc(&p, bytes);
    }
    if (e != cudaSuccess) cuda_fail("cudaMalloc (work space)", e, __FILE__, __LINE__);
    std::lock_guard<std::mutex> lock(g_ws.mu);
    g_ws.device_of[p] = dev;
    return p;
}
void ws_free(void *p, size_t bytes) {
    if (!p) return;
    int dev = 0;
    {
        std::lock_guard<std::mutex> lock(g_ws.mu);
        auto it = g_ws.device_of.find(p);
        if (it != g_ws.device_of.end()) dev = it->second; else cudaGetDevice(&dev);
    }
    if (t_arena.depth > 0) { t_arena.free_.emplace(std::make_pair(dev, bytes), p); return; }
    std::lock_guard<std::mutex> lock(g_ws.mu);
    if (ws_enabled()) {
        g_ws.free_.emplace(std::make_pair(dev, bytes), p);
    } else {
        g_ws.device_of.erase(p);
        cudaFree(p);
    }
}
void ws_release() {
    FreeList taken;
    { std::lock_guard<std::mutex> lock(g_ws.mu); taken.swap(g_ws.free_); }
    release_list(taken);
}
bool ws_is_enabled() {
    std::lock_guard<std::mutex> lock(g_ws.mu);
    return ws_enabled();
}
void ws_set_enabled(bool on) {
    { std::lock_guard<std::mutex> lock(g_ws.mu); g_ws.enabled = on; g_ws.env_read = true; }
    if (!on) ws_release();
}
void ws_arena_begin() { t_arena.depth++; }
void ws_arena_end() {
    if (t_arena.depth > 0 && --t_arena.depth == 0) release_list(t_arena.free_);
}

void cuda_fail(const char *what, cudaError_t e, const char *file, int line) {
    std::fprintf(stderr, "flgpu: CUDA failure %s (%s) at %s:%d -- no CPU fallback exists, aborting\n",
                 cudaGetErrorString(e), what, file, line);
    std::abort();
}
void require_aligned16(const void *p, const char *what) {
    if (reinterpret_cast<uintptr_t>(p) & 15u) {
        std::fprintf(stderr, "flgpu: %s must be 16-byte aligned (the kernels use 128-bit accesses)\n", what);
        std::abort();
    }
}
void fatal(const char *msg) {
    std::fprintf(stderr, "flgpu: %s\n", msg);
    std::abort();
}
int require_device() {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) {
        std::fprintf(stderr,
                     "flgpu: no usable CUDA device (%s). This library is CUDA-only (sm_100a); there is no "
                     "CPU fallback.\n",
                     e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        std::abort();
    }
    int dev = 0;
    FLGPU_CUDA_CHECK(cudaGetDevice(&dev));
    return dev;
}

// --------------------------------------------------------------------------- NCCL via dlopen
namespace {
struct UniqueId { char internal[128]; };
struct NcclApi {
    void *handle = nullptr;
    int (*GetUniqueId)(UniqueId *) = nullptr;
    int (*CommInitRank)(void **, int, UniqueId, int) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;  // datatype: 0 int8, 8 f64
    int (*CommDestroy)(void *) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
NcclApi &nccl() {
    static NcclApi api;
    if (!api.handle) {
        api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!api.handle) api.handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!api.handle) fatal("multi-GPU run requested but libnccl.so.2 cannot be loaded");
        api.GetUniqueId = (int (*)(UniqueId *))dlsym(api.handle, "ncclGetUniqueId");
        api.CommInitRank = (int (*)(void **, int, UniqueId, int))dlsym(api.handle, "ncclCommInitRank");
        api.AllGather = (int (*)(const void *, void *, size_t, int, void *, cudaStream_t))dlsym(api.handle, "ncclAllGather");
        api.CommDestroy = (int (*)(void *))dlsym(api.handle, "ncclCommDestroy");
        api.GetErrorString = (const char *(*)(int))dlsym(api.handle, "ncclGetErrorString");
        if (!api.GetUniqueId || !api.CommInitRank || !api.AllGather || !api.CommDestroy)
            fatal("libnccl.so.2 lacks a required symbol");
    }
    return api;
}
void nccl_check(int rc, const char *what) {
    if (rc != 0) {
        std::fprintf(stderr, "flgpu: NCCL failure in %s: %s -- aborting\n", what,
                     nccl().GetErrorString ? nccl().GetErrorString(rc) : "?");
        std::abort();
    }
}
}  // namespace

void nccl_allgather_f64(flgpu_comm *c, const double *send, double *recv, size_t count, cudaStream_t s) {
    nccl_check(nccl().AllGather(send, recv, count, /*ncclFloat64*/ 8, c->nccl_comm, s), "ncclAllGather");
}

// ---- peer-memory mailboxes: allocate, exchange CUDA IPC handles through NCCL, map every peer.
// Any failure (IPC unavailable in this container, no peer access) leaves p2p = false on ALL ranks and the
// exchange falls back to ncclAllGather + combine_kernel; FLGPU_EXCHANGE=nccl forces that.
namespace {
struct HandleMsg { cudaIpcMemHandle_t h; int ok; int pad; };
void setup_p2p(flgpu_comm *c) {
    const int G = c->nranks;
    const char *mode = std::getenv("FLGPU_EXCHANGE");
    int want = !(mode && !std::strcmp(mode, "nccl")) && G <= k::kMaxRanks;
    HandleMsg mine;
    std::memset(&mine, 0, sizeof mine);
    if (want) {
        if (cudaMalloc((void **)&c->local, sizeof(k::MailboxPair)) == cudaSuccess &&
            cudaMemset(c->local, 0, sizeof(k::MailboxPair)) == cudaSuccess &&
            cudaIpcGetMemHandle(&mine.h, c->local) == cudaSuccess)
            mine.ok = 1;
        else
            cudaGetLastError();
    }
    cudaStream_t s;
    FLGPU_CUDA_CHECK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    char *dsend = nullptr, *drecv = nullptr;
    FLGPU_CUDA_CHECK(cudaMalloc((void **)&dsend, sizeof(HandleMsg)));
    FLGPU_CUDA_CHECK(cudaMalloc((void **)&drecv, sizeof(HandleMsg) * G));
    std::vector<HandleMsg> all(G);
    auto gather = [&]() {
        FLGPU_CUDA_CHECK(cudaMemcpyAsync(dsend, &mine, sizeof mine, cudaMemcpyHostToDevice, s));
        nccl_check(nccl().AllGather(dsend, drecv, sizeof(HandleMsg), /*ncclInt8*/ 0, c->nccl_comm, s), "ncclAllGather");
        FLGPU_CUDA_CHECK(cudaMemcpyAsync(all.data(), drecv, sizeof(HandleMsg) * G, cudaMemcpyDeviceToHost, s));
        FLGPU_CUDA_CHECK(cudaStreamSynchronize(s));
    };
    gather();                                       // round 1: handles
    int ok = 1;
    for (int r = 0; r < G; r++) ok = ok && all[r].ok;
    if (ok) {
        for (int r = 0; r < G; r++) {
            if (r == c->rank) {
                c->peers.box[r] = &c->local->host_driven;
                c->peers_search.box[r] = &c->local->device_search;
                continue;
            }
            void *ptr = nullptr;
            if (cudaIpcOpenMemHandle(&ptr, all[r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                ok = 0;
                break;
            }
            c->peers.box[r] = &((k::MailboxPair *)ptr)->host_driven;
            c->peers_search.box[r] = &((k::MailboxPair *)ptr)->device_search;
        }
    }
    mine.ok = ok;
    gather();                                       // round 2: did every rank map every peer?
    int all_ok = 1;
    for (int r = 0; r < G; r++) all_ok = all_ok && all[r].ok;
    c->p2p = all_ok != 0;
    if (!c->p2p && want && c->rank == 0)
        std::fprintf(stderr, "flgpu: peer-memory exchange unavailable (CUDA IPC / peer access); using ncclAllGather\n");
    if (!c->p2p) {
        for (int r = 0; r < G; r++)
            if (r != c->rank && c->peers.box[r]) { cudaIpcCloseMemHandle(c->peers.box[r]); c->peers.box[r] = nullptr; }
        // (host_driven is the first member: its address is the mapping's base address)
    }
    cudaFree(dsend);
    cudaFree(drecv);
    cudaStreamDestroy(s);
}
}  // namespace

}  // namespace flgpu

extern "C" int flgpu_comm_unique_id(void *id128) {
    flgpu::require_device();
    flgpu::UniqueId id;
    flgpu::nccl_check(flgpu::nccl().GetUniqueId(&id), "ncclGetUniqueId");
    std::memcpy(id128, &id, 128);
    return 0;
}
extern "C" flgpu_comm *flgpu_comm_create(const void *id128, int rank, int nranks) {
    flgpu::require_device();
    flgpu_comm *c = new flgpu_comm;
    c->rank = rank;
    c->nranks = nranks;
    flgpu::UniqueId id;
    std::memcpy(&id, id128, 128);
    flgpu::nccl_check(flgpu::nccl().CommInitRank(&c->nccl_comm, nranks, id, rank), "ncclCommInitRank");
    if (const char *ts = std::getenv("FLGPU_EXCHANGE_TIMEOUT_S")) {      // e.g. under a debugger or a slow host callback
        const double sec = std::atof(ts);
        if (sec > 0.0) c->timeout_ns = (unsigned long long)(sec * 1e9);
    }
    flgpu::setup_p2p(c);
    return c;
}
extern "C" int flgpu_comm_uses_peer_memory(const flgpu_comm *c) { return c && c->p2p ? 1 : 0; }
extern "C" void flgpu_comm_destroy(flgpu_comm *c) {
    if (!c) return;
    cudaDeviceSynchronize();
    if (c->p2p)
        for (int r = 0; r < c->nranks; r++)
            if (r != c->rank && c->peers.box[r]) cudaIpcCloseMemHandle(c->peers.box[r]);
    if (c->local) cudaFree(c->local);
    if (c->nccl_comm) flgpu::nccl().CommDestroy(c->nccl_comm);
    delete c;
}

namespace flgpu {

// --------------------------------------------------------------------------- CudaBackend
CudaBackend::CudaBackend(const flgpu_problem &prob_, int64_t n_local, const flgpu_options &opt)
    : prob(prob_) {
    device = require_device();
    n = n_local;
    comm = opt.comm;
    timing = opt.time_kernels != 0;
    cudaDeviceProp props;
    FLGPU_CUDA_CHECK(cudaGetDeviceProperties(&props, device));
    num_sms = props.multiProcessorCount;
    if (props.major < 10)
        std::fprintf(stderr, "flgpu: warning: device %s is sm_%d%d; kernels are built for sm_100a only\n",
                     props.name, props.major, props.minor);
    if (opt.stream) {
        stream = (cudaStream_t)opt.stream;
    } else {
        FLGPU_CUDA_CHECK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        own_stream = true;
    }
    ld = (n + 31) / 32 * 32;  // 256-byte aligned columns
    if (ld == 0) ld = 32;
    ctx.user = prob.user;
    ctx.stream = (void *)stream;
    ctx.offset = opt.offset;
    ctx.n_global = opt.n_global ? opt.n_global : n_local;
    ctx.rank = comm ? comm->rank : 0;
    ctx.nranks = comm ? comm->nranks : 1;
    ctx.device = device;
    const int G = ctx.nranks;
    const size_t nres = NSLOTS + nd_of(k::kMaxMem);
    auto dalloc = [&](size_t bytes) {
        void *p = ws_alloc(bytes);
        FLGPU_CUDA_CHECK(cudaMemsetAsync(p, 0, bytes, stream));
        owned.emplace_back(p, bytes);
        return p;
    };
    R = (double *)dalloc(nres * sizeof(double));
    Rall = (double *)dalloc(nres * sizeof(double) * G);
    Rglob = (double *)dalloc(NSLOTS * sizeof(double));
    Dsum = (double *)dalloc(nd_of(k::kMaxMem) * sizeof(double));
    Rsearch = (double *)dalloc(FLGPU_SEARCH_RESULT_DOUBLES * sizeof(double));
    // reduction geometry: chunk size from the GLOBAL dimension (flgpu_reduce.cuh), one partial per chunk and accumulator
    ch = red::chunk_elems(ctx.n_global);
    nchunks = red::num_chunks(n, ch);
    alloc_work(12);
    FLGPU_CUDA_CHECK(cudaMallocHost((void **)&host_pinned, (NSLOTS + 16) * sizeof(double)));
    std::memset(host_pinned, 0, (NSLOTS + 16) * sizeof(double));
    const char *sync_mode = std::getenv("FLGPU_SYNC");
    poll_sync = !(sync_mode && !std::strcmp(sync_mode, "stream"));
}

CudaBackend::~CudaBackend() {
    cudaStreamSynchronize(stream);
    for (auto &pb : owned) ws_free(pb.first, pb.second);
    if (host_pinned) cudaFreeHost(host_pinned);
    for (auto &pe : pending) { cudaEventDestroy(pe.a); cudaEventDestroy(pe.b); }
    for (auto e : event_pool) cudaEventDestroy(e);
    if (own_stream) { scratch_release(stream); cudaStreamDestroy(stream); }
}

// rows of chunk sums (one per accumulator of the widest kernel), block values and tickets of the tree kernel
void CudaBackend::alloc_work(int rows) {
    if (rows <= work_rows) return;
    auto dalloc = [&](size_t bytes) {
        void *p = ws_alloc(bytes);
        FLGPU_CUDA_CHECK(cudaMemsetAsync(p, 0, bytes, stream));
        owned.emplace_back(p, bytes);
        return p;
    };
    work.stride = nchunks;
    work.partials = (double *)dalloc((size_t)rows * (size_t)nchunks * sizeof(double));
    work.blockvals = (double *)dalloc((size_t)rows * red::kTopMax * sizeof(double));
    work.tickets = (unsigned int *)dalloc((size_t)rows * sizeof(unsigned int));
    work_rows = rows;
}

// chunk sums of rows [0, nrows) -> this rank's roots at out[row] (nrows <= 12)
void CudaBackend::tree(int nrows, double *const *out) {
    k::TreeArgs a;
    a.w = work; a.nchunks = nchunks; a.lin_out = nullptr; a.dup_row = -1; a.dup_out = nullptr;
    for (int i = 0; i < 12; i++) a.out[i] = i < nrows ? out[i] : nullptr;
    const int nblk = (int)((nchunks + red::kBlockChunks - 1) / red::kBlockChunks);
    k::tree_kernel<<<dim3(nblk, nrows), k::kThreads, 0, stream>>>(a);
    launches++;
}

double *CudaBackend::vec_alloc() {
    void *p = nullptr;
    const double t0 = now_ms();
    p = ws_alloc((size_t)ld * sizeof(double));
    alloc_ms += now_ms() - t0;
    FLGPU_CUDA_CHECK(cudaMemsetAsync(p, 0, (size_t)ld * sizeof(double), stream));
    owned.emplace_back(p, (size_t)ld * sizeof(double));
    return (double *)p;
}

void CudaBackend::lbfgs_alloc(int m) {
    if (m > k::kMaxMem) fatal("LBFGS Memory above FLGPU_MAX_MEMORY reached the backend (internal error: the API checks it)");
    mem = m;
    alloc_work(nd_of(m));
    auto dalloc = [&](size_t bytes) {
        void *p = nullptr;
        const double t0 = now_ms();
        p = ws_alloc(bytes);
        alloc_ms += now_ms() - t0;
        owned.emplace_back(p, bytes);
        return p;
    };
    // ring buffers: column-major (ld, m) like the reference's s(dim,0:mem), y(dim,0:mem) (f90:435)
    S = (double *)dalloc((size_t)ld * m * sizeof(double));
    Y = (double *)dalloc((size_t)ld * m * sizeof(double));
    SY = (double *)dalloc((size_t)m * m * sizeof(double));
    YY = (double *)dalloc((size_t)m * m * sizeof(double));
    C = (double *)dalloc((size_t)nc_of(m) * sizeof(double));
    FLGPU_CUDA_CHECK(cudaMemsetAsync(SY, 0, (size_t)m * m * sizeof(double), stream));
    FLGPU_CUDA_CHECK(cudaMemsetAsync(YY, 0, (size_t)m * m * sizeof(double), stream));
    FLGPU_CUDA_CHECK(cudaMemsetAsync(C, 0, (size_t)nc_of(m) * sizeof(double), stream));
}

void CudaBackend::upload(double *dst, const double *user_x, int x_space) {
    const double t0 = now_ms();
    FLGPU_CUDA_CHECK(cudaMemcpyAsync(dst, user_x, (size_t)n * sizeof(double),
                                     x_space == FLGPU_SPACE_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                                     stream));
    FLGPU_CUDA_CHECK(cudaStreamSynchronize(stream));
    upload_ms += now_ms() - t0;
}
void CudaBackend::download(double *user_x, const double *src, int x_space) {
    FLGPU_CUDA_CHECK(cudaStreamSynchronize(stream));
    const double t0 = now_ms();
    FLGPU_CUDA_CHECK(cudaMemcpyAsync(user_x, src, (size_t)n * sizeof(double),
                                     x_space == FLGPU_SPACE_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                                     stream));
    FLGPU_CUDA_CHECK(cudaStreamSynchronize(stream));
    download_ms += now_ms() - t0;
    resolve_times();
}

// ---- launch geometry: one full wave of resident CTAs (a multiple of the SM count), never more blocks than chunks
int CudaBackend::grid_for(int blocks_per_sm) const {
    int64_t g = (int64_t)num_sms * blocks_per_sm;
    if (g > k::kMaxGrid) g = k::kMaxGrid;
    if (nchunks < g) g = nchunks;
    return (int)(g < 1 ? 1 : g);
}
// element-wise kernels without a reduction: plain grid-stride over 16-byte units
int CudaBackend::grid_units(int64_t units, int blocks_per_sm) const {
    int64_t need = (units + k::kThreads - 1) / k::kThreads;
    int64_t g = (int64_t)num_sms * blocks_per_sm;
    if (g > k::kMaxGrid) g = k::kMaxGrid;
    if (need < g) g = need;
    return (int)(g < 1 ? 1 : g);
}

// ---- timing
cudaEvent_t CudaBackend::get_event() {
    if (!event_pool.empty()) { cudaEvent_t e = event_pool.back(); event_pool.pop_back(); return e; }
    cudaEvent_t e;
    FLGPU_CUDA_CHECK(cudaEventCreate(&e));
    return e;
}
int CudaBackend::time_index(const char *name) {
    for (size_t i = 0; i < times.size(); i++) if (times[i].name == name) return (int)i;
    KernelTime kt; kt.name = name; times.push_back(kt);
    return (int)times.size() - 1;
}
int CudaBackend::time_begin(const char *name, double bytes) {
    if (!timing) return -1;
    const int idx = time_index(name);
    times[idx].launches++;
    times[idx].bytes += bytes;
    Pending pe; pe.idx = idx; pe.a = get_event(); pe.b = get_event();
    FLGPU_CUDA_CHECK(cudaEventRecord(pe.a, stream));
    pending.push_back(pe);
    return (int)pending.size() - 1;
}
void CudaBackend::time_end(int token) {
    if (token < 0) return;
    FLGPU_CUDA_CHECK(cudaEventRecord(pending[token].b, stream));
}
void CudaBackend::resolve_times() {
    if (pending.empty()) return;
    for (auto &pe : pending) {
        float ms = 0.f;
        FLGPU_CUDA_CHECK(cudaEventSynchronize(pe.b));
        FLGPU_CUDA_CHECK(cudaEventElapsedTime(&ms, pe.a, pe.b));
        times[pe.idx].ms += ms;
        event_pool.push_back(pe.a); event_pool.push_back(pe.b);
    }
    pending.clear();
}

// ---- callbacks
void CudaBackend::eval_f(const double *x) {
    const int t = time_begin("callback:f", 8.0 * n);
    prob.f(&ctx, R + SL_F, x, n);
    time_end(t);
    callback_launches++;
}
void CudaBackend::eval_g(const double *x, double *g) {
    const int t = time_begin("callback:fd", 16.0 * n);
    prob.fd(&ctx, g, x, n);
    time_end(t);
    callback_launches++;
}
void CudaBackend::eval_fg(const double *x, double *g) {
    const int t = time_begin("callback:f_fd", 16.0 * n);
    prob.f_fd(&ctx, R + SL_F, g, x, n);
    time_end(t);
    callback_launches++;
}

void CudaBackend::fused_eval(int flags, double a, const double *x0, const double *p, double *x_out,
                             double *g_out) {
    const char *name = (flags & (FLGPU_WRITE_X | FLGPU_WRITE_G)) ? "callback:fused_store" : "callback:fused_probe";
    const double words = 2.0 + ((flags & FLGPU_WRITE_X) ? 1.0 : 0.0) + ((flags & FLGPU_WRITE_G) ? 1.0 : 0.0);
    const int t = time_begin(name, 8.0 * n * words);
    prob.fused(&ctx, flags, R + SL_F, R + SL_GP, x_out, g_out, x0, p, a, n);
    time_end(t);
    callback_launches++;
}

void CudaBackend::fused_eval_multi(int count, const double *steps, const double *x0, const double *p) {
    const int t = time_begin("callback:fused_probe_multi", 16.0 * n);   // one pass over x0 and p, whatever the count
    prob.fused_multi(&ctx, count, steps, R + SL_AUX, x0, p, n);
    time_end(t);
    callback_launches++;
}

void CudaBackend::device_search(int policy, bool strong, bool fdwithf, double c1, double c2abs, double fx0, double phid0,
                                double incr, double a, const double *x0, const double *p, double *xt, double *gt,
                                bool no_store) {
    flgpu_search_args A;
    A.x0_dev = x0; A.p_dev = p; A.x_out = xt; A.g_out = gt;
    A.c1 = c1; A.c2abs = c2abs; A.fx0 = fx0; A.phid0 = phid0; A.incr = incr; A.a = a;
    A.strong = strong ? 1 : 0; A.fdwithf = fdwithf ? 1 : 0;
    A.policy = policy;
    A.no_store = no_store ? 1 : 0;
    A.result_dev = Rsearch;
    A.comm = ctx.nranks > 1 ? comm : nullptr;
    const int t = time_begin("callback:device_search", 0.0);   // bytes depend on the trial count: see flgpu_stats
    prob.search(&ctx, &A, n);
    time_end(t);
    callback_launches++;
}
void CudaBackend::credit_search_bytes(double bytes) {
    if (timing) times[time_index("callback:device_search")].bytes += bytes;
}
void CudaBackend::search_result(double *out) {
    std::memcpy(out, host_pinned + NSLOTS + 8, FLGPU_SEARCH_RESULT_DOUBLES * sizeof(double));
}

// ---- primitives
void CudaBackend::trial_x(double *x, const double *x0, const double *p, double a) {
    const int t = time_begin("trial_x", 24.0 * n);
    k::trial_kernel<<<grid_units(n / 8 + 1, 8), k::kThreads, 0, stream>>>(x, x0, p, a, n);
    time_end(t);
    launches++;
}
void CudaBackend::dot(const double *a, const double *b, int slot) {
    const int t = time_begin("dot", (a == b ? 8.0 : 16.0) * n);
    const bool in_kernel = nchunks <= red::kBlockChunks;     // small reductions finish inside the producing kernel
    k::dot_kernel<<<grid_for(8), k::kThreads, 0, stream>>>(a, b, n, ch, work, 0, in_kernel ? R + slot : nullptr);
    launches++;
    if (!in_kernel) { double *out[1] = {R + slot}; tree(1, out); }
    time_end(t);
}
void CudaBackend::neg(double *p, const double *g) {
    const int t = time_begin("neg", 16.0 * n);
    k::neg_kernel<<<grid_units(n / 2 + 1, 8), k::kThreads, 0, stream>>>(p, g, n);
    time_end(t);
    launches++;
}

// ---- L-BFGS
int g_k1_shape[2] = {0, 0};
namespace {
// per-device launch facts (a process may drive several GPUs: never cache these per process)
struct DeviceFacts { bool k2_attr = false; int k3_tma_attr[16] = {0}; int k3_mode = -1; };
DeviceFacts &facts(int device) {
    static std::mutex mu;
    static std::map<int, DeviceFacts> m;
    std::lock_guard<std::mutex> lock(mu);
    return m[device];
}
}  // namespace

// K1 shape for `rem` older columns still to be covered (k1_pass_shape, flgpu_lbfgs_gram.hpp), or the tuning override
void k1_shape_for(int rem, int &mt, int &ng) {
    if (g_k1_shape[0] > 0) { mt = g_k1_shape[0]; ng = g_k1_shape[1]; }   // flgpu_debug_set_k1_shape
    else k1_pass_shape(rem, mt, ng);
}

// The K1 passes over the k_after-1 older columns (NG*MT per pass; the first pass also builds the new column).  With a
// fused source the first pass is launched by the objective (flgpu_problem.update); later passes read the x1, g1 it stored.
void CudaBackend::k1_passes(k::K1Args &a, int nother, bool fused, int t) {
    int age = 1;
    bool first = true;
    do {
        k::K1Launch L;
        a.age_base = age;
        a.write_new = first ? 1 : 0;
        k1_shape_for(nother - (age - 1), L.mt, L.ng);
        L.a = a; L.num_sms = num_sms; L.nchunks = nchunks; L.stream = (void *)stream;
        if (first && fused) {
            flgpu_update_args ua;
            ua.k1 = &L; ua.k1_bytes = sizeof L;
            prob.update(&ctx, &ua, n);
            callback_launches++;
        } else {
            k::launch_k1_pass(L, k::PlainSrc());
            launches++;
        }
        age += L.mt * L.ng;
        first = false;
    } while (age - 1 < nother);
    lbfgs_dots_tree();
    time_end(t);
}

void CudaBackend::lbfgs_update_dots(const double *x1, const double *x0, const double *g1,
                                    const double *g0, int new_slot, int k_after) {
    const int nother = k_after - 1;
    // algorithmic bytes: one ideal pass (the g1, g0 re-reads of extra passes at m > 11 are overhead, not credit)
    const int t = time_begin("k1_update_dots", 8.0 * n * (2.0 * nother + 6.0));
    k::K1Args a{};
    a.x1 = x1; a.x0 = x0; a.g1 = g1; a.g0 = g0; a.S = S; a.Y = Y; a.ld = ld; a.n = n; a.ch = ch;
    a.m = mem; a.new_slot = new_slot; a.k_after = k_after; a.w = work;
    k1_passes(a, nother, false, t);
}

void CudaBackend::lbfgs_update_dots_fused(double step, const double *x0, const double *p, const double *g0, double *x1,
                                          double *g1, int new_slot, int k_after) {
    const int nother = k_after - 1;
    // reads x0, p (and g0, unless the source re-evaluates it: the built-in objectives do) and the older columns; writes
    // x1, g1 and the new column pair.  Credited as the leaner of the two: 2 + 2(k-1) in, 4 out.
    const int t = time_begin("k1_update_dots_fused", 8.0 * n * (2.0 * nother + 6.0));
    k::K1Args a{};
    a.x1 = x1; a.x0 = x0; a.g1 = g1; a.g0 = g0; a.p = p; a.step = step; a.x1_out = x1; a.g1_out = g1;
    a.S = S; a.Y = Y; a.ld = ld; a.n = n; a.ch = ch; a.offset = ctx.offset; a.n_global = ctx.n_global;
    a.m = mem; a.new_slot = new_slot; a.k_after = k_after; a.w = work;
    k1_passes(a, nother, true, t);
}

// chunk sums of all nd dots -> R[kResSlots + d]; g.g also to its result slot
void CudaBackend::lbfgs_dots_tree() {
    k::TreeArgs a;
    a.w = work; a.nchunks = nchunks; a.lin_out = R + k::kResSlots; a.dup_row = d_GG(mem); a.dup_out = R + SL_GG;
    for (int i = 0; i < 12; i++) a.out[i] = nullptr;
    const int nblk = (int)((nchunks + red::kBlockChunks - 1) / red::kBlockChunks);
    k::tree_kernel<<<dim3(nblk, nd_of(mem)), k::kThreads, 0, stream>>>(a);
    launches++;
}

void CudaBackend::lbfgs_solve(int kk, int recent) {
    const int nd = nd_of(mem);
    const int G = ctx.nranks;
    const double *Dall = R + NSLOTS;
    if (G > 1) {
        exchange(R + NSLOTS, nd, Dsum, nullptr);
        Dall = Dsum;
    }
    const int t = time_begin("k2_solve", 0.0);
    const size_t smem = (size_t)(nd + 2 * mem * mem) * sizeof(double);
    if (smem > 48 * 1024) {
        DeviceFacts &f = facts(device);
        if (!f.k2_attr) {
            FLGPU_CUDA_CHECK(cudaFuncSetAttribute(k::k2_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            f.k2_attr = true;
        }
    }
    k::k2_solve_kernel<<<1, 32, smem, stream>>>(mem, kk, recent, Dall, 1, SY, YY, C);
    time_end(t);
    launches++;
}

void CudaBackend::lbfgs_direction(double *p, double *xt, const double *g1, const double *x1, int kk,
                                  int recent) {
    const int t = time_begin("k3_direction", 8.0 * n * (2.0 * kk + (xt ? 4.0 : 2.0)));
    k::K3Args a;
    a.p = p; a.xt = xt; a.g1 = g1; a.x1 = x1; a.S = S; a.Y = Y; a.C = C; a.ld = ld; a.n = n; a.ch = ch;
    a.m = mem; a.k = kk; a.recent = recent; a.offset = ctx.offset; a.n_global = ctx.n_global; a.w = work;
    DeviceFacts &f = facts(device);
    if (f.k3_mode < 0) {
        // FLGPU_K3 = regs (register double buffers) | tma (default: bulk-async ring, 11 pieces x 2 stages, 2 CTAs/SM:
        // profiles/r02_k3_ring_shapes.md) |
        // "P,NST": another ring shape (tuning; the instantiated ones are listed below)
        const char *v = std::getenv("FLGPU_K3");
        f.k3_mode = 1;
        if (v && !std::strcmp(v, "regs")) f.k3_mode = 0;
        else if (v && std::strchr(v, ',')) f.k3_mode = 100 * std::atoi(v) + std::atoi(std::strchr(v, ',') + 1);
    }
    if (f.k3_mode == 0) {
        k::k3_direction_kernel<8><<<grid_for(2), k::kThreads, 0, stream>>>(a);
    } else {
        bool done = false;
#define FLGPU_K3_CASE(P, NST, MINB, ID)                                                                                  \
        if (!done && (f.k3_mode == 100 * P + NST || (ID == 0 && f.k3_mode == 1))) {                                        \
            constexpr size_t smem = (size_t)NST * P * k::kThreads * sizeof(double2);                                       \
            if (!f.k3_tma_attr[ID]) {                                                                                      \
                FLGPU_CUDA_CHECK(cudaFuncSetAttribute(k::k3_direction_tma_kernel<P, NST, MINB, k::NoProbe>,                           \
                                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));            \
                f.k3_tma_attr[ID] = 1;                                                                                     \
            }                                                                                                              \
            k::k3_direction_tma_kernel<P, NST, MINB, k::NoProbe><<<grid_for(MINB), k::kThreads + 32, smem, stream>>>(a, k::NoProbe()); \
            done = true;                                                                                                   \
        }
        // the shapes of profiles/r02_k3_ring_shapes.md worth keeping selectable (the one-CTA-per-SM and 3-CTA ones lost)
        FLGPU_K3_CASE(11, 2, 2, 0) FLGPU_K3_CASE(9, 3, 2, 1) FLGPU_K3_CASE(8, 3, 2, 2) FLGPU_K3_CASE(7, 3, 2, 3)
        FLGPU_K3_CASE(4, 6, 2, 4)
#undef FLGPU_K3_CASE
        if (!done) fatal("FLGPU_K3: this ring shape is not instantiated");
    }
    launches++;
    double *out[2] = {R + SL_GP0, R + SL_PP};
    tree(2, out);
    time_end(t);
}

static_assert(k::kProbeSteps == FLGPU_MULTI_MAX, "the K3 probe fills one batch of the line search");
// K3 launched by the objective with its probe (flgpu_problem.direction): reads g1, x1 and the 2k columns, writes p; the
// chunk sums of g1.p, p.p go to rows 0, 1 and those of f, f'.p at x1 + steps[j]*p to rows 2+2j, 3+2j.  Step 0 (= 1, the
// first trial) is delivered to SL_F / SL_GP like a probe; steps 1.. to the batch slots SL_AUX+2j, SL_AUX+2j+1.
void CudaBackend::lbfgs_direction_probe(double *p, const double *g1, const double *x1, int kk, int recent, int flags,
                                        const double *steps) {
    const int t = time_begin("k3_direction_probe", 8.0 * n * (2.0 * kk + 3.0));
    k::K3Launch L;
    k::K3Args &a = L.a;
    a.p = p; a.xt = nullptr; a.g1 = g1; a.x1 = x1; a.S = S; a.Y = Y; a.C = C; a.ld = ld; a.n = n; a.ch = ch;
    a.m = mem; a.k = kk; a.recent = recent; a.offset = ctx.offset; a.n_global = ctx.n_global; a.w = work;
    for (int j = 0; j < k::kProbeSteps; j++) a.steps[j] = steps[j];
    L.grid = grid_for(2); L.stream = (void *)stream;
    flgpu_direction_args da;
    da.k3 = &L; da.k3_bytes = sizeof L; da.flags = flags;
    prob.direction(&ctx, &da, n);
    callback_launches++;
    double *out[2 + 2 * k::kProbeSteps] = {R + SL_GP0, R + SL_PP, R + SL_F, R + SL_GP};
    for (int j = 1; j < k::kProbeSteps; j++) { out[2 + 2 * j] = R + SL_AUX + 2 * j; out[3 + 2 * j] = R + SL_AUX + 2 * j + 1; }
    tree(2 + 2 * k::kProbeSteps, out);
    time_end(t);
}

// ---- CG
void CudaBackend::cg_dots(const double *g1, const double *g0, const double *p) {
    const int t = time_begin("cg_dots", 24.0 * n);
    double *out[5] = {R + SL_GG, R + SL_PP, R + SL_DGP, R + SL_GDG, R + SL_G0G0};
    const bool in_kernel = nchunks <= red::kBlockChunks;
    k::Outs5 o{};
    if (in_kernel) for (int i = 0; i < 5; i++) o.p[i] = out[i];
    k::cg_dots_kernel<<<grid_for(6), k::kThreads, 0, stream>>>(g1, g0, p, n, ch, work, o);
    launches++;
    if (!in_kernel) tree(5, out);
    time_end(t);
}
void CudaBackend::cg_update(double *p, const double *g1, double beta) {
    const int t = time_begin("cg_update", 24.0 * n);
    const bool in_kernel = nchunks <= red::kBlockChunks;
    k::cg_update_kernel<<<grid_for(6), k::kThreads, 0, stream>>>(p, g1, beta, n, ch, work, in_kernel ? R + SL_GP0 : nullptr);
    launches++;
    if (!in_kernel) { double *out[1] = {R + SL_GP0}; tree(1, out); }
    time_end(t);
}

// ---- ranks: out[i] = rank tree (red::rank_tree) of src_r[i]: identical bits on every rank.  One kernel over peer memory, or the
// ncclAllGather + combine fallback (gather: [G][count] device scratch).  host_out (optional, peer-memory path
// only): pinned host array that receives the sums followed by the flag word host_seq_next; returns whether it did.
bool rank_sum(flgpu_comm *c, cudaStream_t s, const double *src, int count, double *out, double *gather,
              double *host_out, unsigned long long host_seq_next, const double *extra) {
    if (count > k::kMailWidth) fatal("rank_sum: more values than one mailbox slot holds");
    if (c->p2p) {
        k::exchange_kernel<<<1, k::kMailWidth, 0, s>>>(c->peers, c->rank, c->nranks, ++c->seq, src, count, out, host_out,
                                                      host_seq_next, extra, c->timeout_ns);
        return host_out != nullptr;
    }
    nccl_allgather_f64(c, src, gather, (size_t)count, s);
    k::combine_kernel<<<1, k::kMailWidth, 0, s>>>(gather, c->nranks, count, out);
    return false;
}

bool CudaBackend::exchange(const double *src, int count, double *out, double *host_out) {
    const int t = time_begin("c1_exchange", 0.0);
    const bool host_written = rank_sum(comm, stream, src, count, out, Rall, host_out, host_seq + 1,
                                       host_out ? Rsearch : nullptr);
    time_end(t);
    launches++;
    return host_written;
}

// ---- host <- device.  The last kernel of the chain stores the 16 slots and then a sequence number into pinned
// host memory; the host polls that word (no DMA copy, no driver synchronisation: ~20 us -> a few us per trial).
// FLGPU_SYNC=stream selects the plain cudaMemcpyAsync + cudaStreamSynchronize path.
void CudaBackend::fetch(double *host_slots) {
    const double *src = R;
    bool on_host = false;
    if (ctx.nranks > 1) {
        on_host = exchange(R, NSLOTS, Rglob, poll_sync ? host_pinned : nullptr);
        src = Rglob;
    }
    if (poll_sync) {
        if (!on_host) {
            k::publish_kernel<<<1, 32, 0, stream>>>(src, host_pinned, host_seq + 1, Rsearch);
            launches++;
        }
        host_seq++;
        volatile unsigned long long *flag = reinterpret_cast<volatile unsigned long long *>(host_pinned + NSLOTS);
        unsigned long long spins = 0;
        while (*flag != host_seq) {
            if ((++spins & 0xffffu) == 0) {          // every 65536 polls make sure the stream is still healthy
                cudaError_t q = cudaStreamQuery(stream);
                if (q != cudaSuccess && q != cudaErrorNotReady) cuda_fail("waiting for results", q, __FILE__, __LINE__);
                if (q == cudaSuccess && *flag != host_seq) fatal("result flag was not published (internal error)");
            }
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
        }
        std::atomic_thread_fence(std::memory_order_acquire);
    } else {
        FLGPU_CUDA_CHECK(cudaMemcpyAsync(host_pinned, src, NSLOTS * sizeof(double), cudaMemcpyDeviceToHost, stream));
        FLGPU_CUDA_CHECK(cudaMemcpyAsync(host_pinned + NSLOTS + 8, Rsearch, FLGPU_SEARCH_RESULT_DOUBLES * sizeof(double),
                                         cudaMemcpyDeviceToHost, stream));
        FLGPU_CUDA_CHECK(cudaStreamSynchronize(stream));
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) cuda_fail("kernel launch", e, __FILE__, __LINE__);
    std::memcpy(host_slots, host_pinned, NSLOTS * sizeof(double));
    syncs++;
    if (timing && pending.size() > 4096) resolve_times();
}

}  // namespace flgpu
