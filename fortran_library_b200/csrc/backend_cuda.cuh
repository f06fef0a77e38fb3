// backend_cuda.cuh -- the product's only Backend: hand-written sm_100a kernels on one CUDA stream.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <utility>
#include <vector>

#include "../../include/flgpu.h"
#include "backend.hpp"
#include "kernels.cuh"

#define FLGPU_CUDA_CHECK(expr)                                                                     \
    do {                                                                                           \
        cudaError_t e__ = (expr);                                                                  \
        if (e__ != cudaSuccess) ::flgpu::cuda_fail(#expr, e__, __FILE__, __LINE__);                \
    } while (0)

struct flgpu_comm {
    void *nccl_comm = nullptr;  // ncclComm_t (plumbing: rendezvous, IPC-handle exchange, fallback all-gather)
    int rank = 0, nranks = 1;
    // peer-memory exchange (kernels.cuh C1): this rank's mailbox and the IPC-mapped mailboxes of its peers
    bool p2p = false;
    flgpu::k::MailboxPair *local = nullptr;
    flgpu::k::PeerTable peers{};        // host-driven exchanges (exchange_kernel); seq = last sequence number used
    flgpu::k::PeerTable peers_search{}; // exchanges made inside a device-resident line search; counter local->dseq
    unsigned long long seq = 0;
    unsigned long long timeout_ns = 20000000000ull;   // a peer silent for this long is reported and the exchange traps
};

namespace flgpu {

[[noreturn]] void cuda_fail(const char *what, cudaError_t e, const char *file, int line);
[[noreturn]] void fatal(const char *msg);
void require_aligned16(const void *p, const char *what);   // aborts with a message otherwise
// Aborts with a message unless a CUDA device is usable (the library has no CPU path).
int require_device();

// ---- work-space allocation (optionally cached between calls: flgpu_set_workspace_cache)
void *ws_alloc(size_t bytes);
void ws_free(void *p, size_t bytes);
void ws_release();
void ws_set_enabled(bool on);
bool ws_is_enabled();
// per-thread arena: buffers freed between begin/end are parked for this thread's next allocation, released at end
void ws_arena_begin();
void ws_arena_end();

// ---- per-stream scratch of the built-in objectives / primitives (objectives.cu)
void scratch_release(cudaStream_t s);
double *scratch_scalar(cudaStream_t s);   // device double[4] private to the stream
k::Work scratch_work(cudaStream_t s, int64_t nchunks);   // the stream's reduction rows (8), tickets zero between kernels

// ---- NCCL (resolved with dlopen so single-GPU use has no NCCL dependency)
void nccl_allgather_f64(flgpu_comm *c, const double *send, double *recv, size_t count, cudaStream_t s);
// rank-ordered sum of `count` (<= k::kMailWidth) doubles over the communicator (backend_cuda.cu)
bool rank_sum(flgpu_comm *c, cudaStream_t s, const double *src, int count, double *out, double *gather,
              double *host_out, unsigned long long host_seq_next, const double *extra = nullptr);

// ---- per-kernel CUDA-event timing (flgpu_options.time_kernels)
struct KernelTime {
    std::string name;
    double ms = 0.0;
    int64_t launches = 0;
    double bytes = 0.0;  // algorithmic bytes credited
};

class CudaBackend : public Backend {
public:
    CudaBackend(const flgpu_problem &prob, int64_t n_local, const flgpu_options &opt);
    ~CudaBackend() override;

    double *vec_alloc() override;
    void lbfgs_alloc(int mem) override;
    void upload(double *dst, const double *user_x, int x_space) override;
    void download(double *user_x, const double *src, int x_space) override;
    void eval_f(const double *x) override;
    void eval_g(const double *x, double *g) override;
    void eval_fg(const double *x, double *g) override;
    bool fused_available() const override { return prob.fused != nullptr; }
    void fused_eval(int flags, double a, const double *x0, const double *p, double *x_out, double *g_out) override;
    bool device_search_available() const override {
        return prob.search != nullptr &&
               (ctx.nranks == 1 || (comm && comm->p2p && (prob.search_caps & FLGPU_SEARCH_ROW_SHARDS)));
    }
    void device_search(int policy, bool strong, bool fdwithf, double c1, double c2abs, double fx0, double phid0, double incr,
                       double a, const double *x0, const double *p, double *xt, double *gt, bool no_store = false) override;
    bool fused_update_available() const override { return prob.update != nullptr && prob.fused != nullptr; }
    void lbfgs_update_dots_fused(double a, const double *x0, const double *p, const double *g0, double *x1, double *g1,
                                 int new_slot, int k_after) override;
    void search_result(double *out) override;
    void credit_search_bytes(double bytes) override;
    void trial_x(double *x, const double *x0, const double *p, double a) override;
    void dot(const double *a, const double *b, int slot) override;
    void neg(double *p, const double *g) override;
    void lbfgs_update_dots(const double *x1, const double *x0, const double *g1, const double *g0,
                           int new_slot, int k_after) override;
    void lbfgs_solve(int k, int recent) override;
    bool fused_multi_available() const override { return prob.fused_multi != nullptr && prob.fused != nullptr; }
    void fused_eval_multi(int count, const double *steps, const double *x0, const double *p) override;
    bool fused_direction_available() const override { return prob.direction != nullptr && prob.fused != nullptr; }
    void lbfgs_direction_probe(double *p, const double *g1, const double *x1, int k, int recent, int flags,
                               const double *steps) override;
    void lbfgs_direction(double *p, double *xt, const double *g1, const double *x1, int k,
                         int recent) override;
    void cg_dots(const double *g1, const double *g0, const double *p) override;
    void cg_update(double *p, const double *g1, double beta) override;
    void fetch(double *host_slots) override;
    void *stream_handle() override { return (void *)stream; }

    // accessors for the history API / tests
    double *S = nullptr, *Y = nullptr, *SY = nullptr, *YY = nullptr, *C = nullptr;
    int mem = 0;
    int64_t ld = 0;
    int64_t ch = 1024, nchunks = 1;   // reduction geometry (flgpu_reduce.cuh): chunk elements, local chunks
    double *R = nullptr;          // [NSLOTS + nd] device results
    cudaStream_t stream = nullptr;
    std::vector<KernelTime> times;
    void resolve_times();
    double alloc_ms = 0.0, upload_ms = 0.0, download_ms = 0.0;   // wall clock, for FLGPU_TRACE_PHASES

private:
    flgpu_problem prob;
    flgpu_eval_ctx ctx;
    flgpu_comm *comm = nullptr;
    bool own_stream = false;
    int device = 0, num_sms = 148;
    std::vector<std::pair<void *, size_t>> owned;
    k::Work work{};
    double *Rall = nullptr;       // [G][NSLOTS + nd] all-gathered results (NCCL fallback only)
    double *Dsum = nullptr;       // [nd] rank-ordered sum of the K1 dots
    double *Rsearch = nullptr;    // [FLGPU_SEARCH_RESULT_DOUBLES] scalars of the last device-resident search
    // out = sum over ranks, in rank order; returns true when the sums were also stored into host_out
    bool exchange(const double *src, int count, double *out, double *host_out);
    double *Rglob = nullptr;      // [NSLOTS] combined slots
    // pinned, device-addressable under UVA: [0,NSLOTS) result slots, [NSLOTS] flag word, [NSLOTS+8, +16) search scalars
    double *host_pinned = nullptr;
    unsigned long long host_seq = 0; // value of the flag word after the last fetch()
    bool poll_sync = true;
    bool timing = false;
    struct Pending { int idx; cudaEvent_t a, b; };
    std::vector<Pending> pending;
    std::vector<cudaEvent_t> event_pool;
    int grid_for(int blocks_per_sm) const;                       // reducing kernels: <= one block per chunk
    int grid_units(int64_t units, int blocks_per_sm) const;      // element-wise kernels
    void alloc_work(int rows);
    void tree(int nrows, double *const *out);                    // chunk sums of rows [0, nrows) -> out[row]
    void lbfgs_dots_tree();
    void k1_passes(k::K1Args &a, int nother, bool fused, int t);
    int work_rows = 0;
    int time_begin(const char *name, double bytes);
    void time_end(int token);
    cudaEvent_t get_event();
    int time_index(const char *name);
};

}  // namespace flgpu
