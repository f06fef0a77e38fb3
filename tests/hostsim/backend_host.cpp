// tests/hostsim/backend_host.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A host-memory implementation of fortran_library_b200/csrc/backend.hpp, so that the host
// control flow in driver.cpp (the same translation unit the product links) can be checked against
// the oracle on a machine without a GPU, including the world_size>1 row-shard combination
// over a caller-supplied all-gather (gloo in the tests).  Built into
// tests/hostsim/libflgpu_hostsim.so; libflgpu.so never links or loads it and has no CPU path.
//
// Element-wise arithmetic matches the CUDA kernels (separate multiply and add roundings);
// reductions are blocked sums, so results agree with the GPU up to summation order only.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../fortran_library_b200/csrc/backend.hpp"
#include "../../fortran_library_b200/csrc/driver.hpp"
#include "../../include/flgpu_lbfgs_gram.hpp"
#include "../../include/flgpu_reduce_geom.h"
#include "../../include/flgpu_search_core.hpp"
#include "../../oracle/oracle.h"

namespace {

typedef void (*allgather_fn)(void *user, const double *local, double *all, int count);
allgather_fn g_allgather = nullptr;
void *g_allgather_user = nullptr;
int g_rank = 0, g_nranks = 1;

const long BLK = 1024;

template <class F>
double blocked_sum(long n, F term) {
    double total = 0.0;
    for (long b = 0; b < n; b += BLK) {
        double s = 0.0;
        const long e = b + BLK < n ? b + BLK : n;
        for (long i = b; i < e; i++) s += term(i);
        total += s;
    }
    return total;
}

class HostBackend : public flgpu::Backend {
public:
    flgpu_problem prob;
    flgpu_eval_ctx ctx;
    std::vector<double *> owned;
    double slots[flgpu::NSLOTS];
    int mem = 0;
    double *S = nullptr, *Y = nullptr;
    std::vector<double> SY, YY, D, C, w1, w2, w3;

    HostBackend(const flgpu_problem &p, int64_t n_, int64_t offset, int64_t n_global) : prob(p) {
        n = n_;
        std::memset(slots, 0, sizeof slots);
        ctx.user = p.user; ctx.stream = nullptr; ctx.offset = offset;
        ctx.n_global = n_global ? n_global : n_; ctx.rank = g_rank; ctx.nranks = g_nranks; ctx.device = -1;
    }
    ~HostBackend() override { for (double *v : owned) std::free(v); }

    double *vec_alloc() override {
        double *v = (double *)std::calloc((size_t)(n > 0 ? n : 1), sizeof(double));
        owned.push_back(v);
        return v;
    }
    void lbfgs_alloc(int m) override {
        mem = m;
        S = (double *)std::calloc((size_t)(n > 0 ? n : 1) * m, sizeof(double));
        Y = (double *)std::calloc((size_t)(n > 0 ? n : 1) * m, sizeof(double));
        owned.push_back(S); owned.push_back(Y);
        SY.assign((size_t)m * m, 0.0); YY.assign((size_t)m * m, 0.0);
        D.assign(flgpu::nd_of(m), 0.0); C.assign(flgpu::nc_of(m), 0.0);
        w1.assign(m, 0.0); w2.assign(m, 0.0); w3.assign(m, 0.0);
    }
    void upload(double *dst, const double *user_x, int) override { std::memcpy(dst, user_x, sizeof(double) * n); }
    void download(double *user_x, const double *src, int) override { std::memcpy(user_x, src, sizeof(double) * n); }

    void eval_f(const double *x) override { prob.f(&ctx, &slots[flgpu::SL_F], x, n); callback_launches++; }
    void eval_g(const double *x, double *g) override { prob.fd(&ctx, g, x, n); callback_launches++; }
    void eval_fg(const double *x, double *g) override { prob.f_fd(&ctx, &slots[flgpu::SL_F], g, x, n); callback_launches++; }

    bool fused_available() const override { return prob.fused != nullptr; }
    void fused_eval(int flags, double a, const double *x0, const double *p, double *x_out, double *g_out) override {
        prob.fused(&ctx, flags, &slots[flgpu::SL_F], &slots[flgpu::SL_GP], x_out, g_out, x0, p, a, n);
        callback_launches++;
    }
    // "device-resident" search on the host: the same SearchCore the CUDA search kernels instantiate, with EAGER
    // evaluations (every call computes f / f'.p at once, as a cooperative kernel does) instead of the driver's lazy
    // ones -- so the driver's device-search branch and the eager evaluator semantics are testable without a GPU.
    struct EagerSearch : flgpu::SearchCore<EagerSearch> {
        HostBackend &B;
        const double *x0, *p;
        double f_cur = 0.0, gp_cur = 0.0, a_x = 0.0, a_g = 0.0;
        bool have_x = false, have_g = false;
        double trials = 0, n_f = 0, n_fd = 0, n_ffd = 0, n_fonly = 0;
        EagerSearch(HostBackend &b, const double *x0_, const double *p_) : B(b), x0(x0_), p(p_) {}
        void eval(int flags) {
            double fv = 0.0, gv = 0.0;
            B.prob.fused(&B.ctx, flags, &fv, &gv, nullptr, nullptr, x0, p, a_x, B.n);
            if (flags & FLGPU_WANT_F) f_cur = fv;
            if (flags & FLGPU_WANT_GP) gp_cur = gv;
        }
        void form(double step) { a_x = step; have_x = true; trials += 1; }
        void call_f() { eval(FLGPU_WANT_F); n_f += 1; }
        void call_fd() { eval(FLGPU_WANT_GP); a_g = a_x; have_g = true; n_fd += 1; }
        void call_ffd() { eval(FLGPU_WANT_F | FLGPU_WANT_GP); a_g = a_x; have_g = true; n_ffd += 1; }
        double slope() { return gp_cur; }
        double fx() { return f_cur; }
        void set_fx(double v) { f_cur = v; }
        void adopt_pre() {}
        void count_f_only() { n_fonly += 1; }
        static bool aborted() { return false; }
    };
    double search_res[FLGPU_SEARCH_RESULT_DOUBLES] = {0};
    bool device_search_available() const override { return prob.fused != nullptr && g_nranks <= 1; }
    void device_search(int policy, bool strong, bool fdwithf, double c1, double c2abs, double fx0, double phid0, double incr,
                       double a, const double *x0, const double *p, double *xt, double *gt, bool = false) override {
        callback_launches++;
        EagerSearch S(*this, x0, p);
        S.c1 = c1; S.c2abs = c2abs; S.fx0 = fx0; S.phid0 = phid0; S.incr = incr; S.fdwithf = fdwithf;
        S.a = a; S.f_cur = fx0; S.pre = 0;
        if (policy == FLGPU_LS_FAST) S.fast(strong);
        else if (strong) S.strongwolfe(); else S.wolfe();
        double fv, gv;
        if (S.have_x && S.have_g && S.a_x == S.a_g) {
            prob.fused(&ctx, FLGPU_WRITE_X | FLGPU_WRITE_G, &fv, &gv, xt, gt, x0, p, S.a_x, n);
        } else {
            if (S.have_x) prob.fused(&ctx, FLGPU_WRITE_X, &fv, &gv, xt, nullptr, x0, p, S.a_x, n);
            if (S.have_g) prob.fused(&ctx, FLGPU_WRITE_G, &fv, &gv, nullptr, gt, x0, p, S.a_g, n);
        }
        const double r[FLGPU_SEARCH_RESULT_DOUBLES] = {S.a, S.f_cur, S.trials, S.n_f, S.n_fd, S.n_ffd, S.n_fonly, 0.0};
        std::memcpy(search_res, r, sizeof r);
    }
    void search_result(double *out) override { std::memcpy(out, search_res, sizeof search_res); }

    void trial_x(double *x, const double *x0, const double *p, double a) override {
        launches++;
        for (long i = 0; i < n; i++) x[i] = x0[i] + a * p[i];
    }
    void dot(const double *a, const double *b, int slot) override {
        launches++;
        slots[slot] = blocked_sum(n, [&](long i) { return a[i] * b[i]; });
    }
    void neg(double *p, const double *g) override {
        launches++;
        for (long i = 0; i < n; i++) p[i] = -g[i];
    }

    void lbfgs_update_dots(const double *x1, const double *x0, const double *g1, const double *g0,
                           int new_slot, int k_after) override {
        launches++;
        const int m = mem;
        double *sn = S + (long)new_slot * n, *yn = Y + (long)new_slot * n;
        for (long i = 0; i < n; i++) { sn[i] = x1[i] - x0[i]; yn[i] = g1[i] - g0[i]; }
        for (int t = 0; t < k_after; t++) {
            const int j = flgpu::slot_of_age(new_slot, t, m);
            const double *sj = S + (long)j * n, *yj = Y + (long)j * n;
            D[flgpu::d_A(m, j)] = blocked_sum(n, [&](long i) { return sj[i] * g1[i]; });
            D[flgpu::d_B(m, j)] = blocked_sum(n, [&](long i) { return yj[i] * g1[i]; });
            D[flgpu::d_SYN(m, j)] = blocked_sum(n, [&](long i) { return sj[i] * yn[i]; });
            D[flgpu::d_YYN(m, j)] = blocked_sum(n, [&](long i) { return yj[i] * yn[i]; });
        }
        D[flgpu::d_GG(m)] = blocked_sum(n, [&](long i) { return g1[i] * g1[i]; });
        slots[flgpu::SL_GG] = D[flgpu::d_GG(m)];
    }
    void lbfgs_solve(int k, int recent) override {
        launches++;
        std::vector<double> G(D);
        combine(G.data(), (int)G.size());
        flgpu::lbfgs_gram_solve(mem, k, recent, G.data(), SY.data(), YY.data(), C.data(), w1.data(),
                                w2.data(), w3.data());
    }
    void lbfgs_direction(double *p, double *xt, const double *g1, const double *x1, int k,
                         int recent) override {  // xt may be null (fused line search)
        launches++;
        const int m = mem;
        const double gamma = C[0];
        for (long i = 0; i < n; i++) {
            double q = g1[i];
            for (int t = 0; t < k; t++) {
                const int j = flgpu::slot_of_age(recent, t, m);
                q = q - C[1 + j] * Y[(long)j * n + i];
            }
            double r = gamma * q;
            for (int t = k - 1; t >= 0; t--) {
                const int j = flgpu::slot_of_age(recent, t, m);
                r = r + C[1 + m + j] * S[(long)j * n + i];
            }
            p[i] = -r;
            if (xt) xt[i] = x1[i] + p[i];
        }
        slots[flgpu::SL_GP0] = blocked_sum(n, [&](long i) { return g1[i] * p[i]; });
        slots[flgpu::SL_PP] = blocked_sum(n, [&](long i) { return p[i] * p[i]; });
    }

    void cg_dots(const double *g1, const double *g0, const double *p) override {
        launches++;
        slots[flgpu::SL_GG] = blocked_sum(n, [&](long i) { return g1[i] * g1[i]; });
        slots[flgpu::SL_PP] = blocked_sum(n, [&](long i) { return p[i] * p[i]; });
        slots[flgpu::SL_DGP] = blocked_sum(n, [&](long i) { return (g1[i] - g0[i]) * p[i]; });
        slots[flgpu::SL_GDG] = blocked_sum(n, [&](long i) { return g1[i] * (g1[i] - g0[i]); });
        slots[flgpu::SL_G0G0] = blocked_sum(n, [&](long i) { return g0[i] * g0[i]; });
    }
    void cg_update(double *p, const double *g1, double beta) override {
        launches++;
        for (long i = 0; i < n; i++) p[i] = -g1[i] + beta * p[i];
        slots[flgpu::SL_GP0] = blocked_sum(n, [&](long i) { return g1[i] * p[i]; });
    }

    // the product's rank tree (flgpu_reduce_geom.h) over the gathered per-rank values: bitwise identical on every rank
    void combine(double *v, int count) {
        if (g_nranks <= 1 || !g_allgather) return;
        std::vector<double> all((size_t)count * g_nranks);
        g_allgather(g_allgather_user, v, all.data(), count);
        for (int i = 0; i < count; i++) v[i] = flgpu::red::rank_tree(all.data() + i, g_nranks, count);
    }
    void fetch(double *host) override {
        syncs++;
        std::memcpy(host, slots, sizeof slots);
        combine(host, flgpu::NSLOTS);
    }
};

// ---- objective adapters over oracle/objectives.c (host pointers)
// The simulator sums f pairwise (a GPU callback cannot sum sequentially either), so CPU tests see the
// same kind of objective-value noise the CUDA objective kernels produce.
int g_obj_sum_mode = 2;
void obj_f(const flgpu_eval_ctx *c, double *f, const double *x, int64_t n) {
    int d = (int)n;
    orc_obj_select((int)(intptr_t)c->user, c->offset, c->n_global);
    orc_obj_set_sum_mode(g_obj_sum_mode);
    orc_obj_f(f, x, &d);
}
void obj_fd(const flgpu_eval_ctx *c, double *g, const double *x, int64_t n) {
    int d = (int)n;
    orc_obj_select((int)(intptr_t)c->user, c->offset, c->n_global);
    orc_obj_fd(g, x, &d);
}
void obj_ffd(const flgpu_eval_ctx *c, double *f, double *g, const double *x, int64_t n) {
    int d = (int)n;
    orc_obj_select((int)(intptr_t)c->user, c->offset, c->n_global);
    orc_obj_set_sum_mode(g_obj_sum_mode);
    orc_obj_f_fd(f, g, x, &d);
}

// fused evaluation (flgpu_fused_fn) on host memory: the point is formed element-wise, multiply then add
void obj_fused(const flgpu_eval_ctx *c, int flags, double *f, double *gp, double *x_out, double *g_out,
               const double *x0, const double *p, double a, int64_t n) {
    std::vector<double> x((size_t)(n > 0 ? n : 1)), g((size_t)(n > 0 ? n : 1));
    for (int64_t i = 0; i < n; i++) x[i] = x0[i] + a * p[i];
    int d = (int)n;
    double fv = 0.0;
    orc_obj_select((int)(intptr_t)c->user, c->offset, c->n_global);
    orc_obj_set_sum_mode(g_obj_sum_mode);
    orc_obj_f_fd(&fv, g.data(), x.data(), &d);
    if (flags & FLGPU_WANT_F) *f = fv;
    if (flags & FLGPU_WANT_GP) *gp = blocked_sum(n, [&](long i) { return g[i] * p[i]; });
    if (flags & FLGPU_WRITE_X) std::memcpy(x_out, x.data(), sizeof(double) * n);
    if (flags & FLGPU_WRITE_G) std::memcpy(g_out, g.data(), sizeof(double) * n);
}

}  // namespace

extern "C" {

void flgpu_hostsim_set_comm(allgather_fn fn, void *user, int rank, int nranks) {
    g_allgather = fn; g_allgather_user = user; g_rank = rank; g_nranks = nranks;
}

void flgpu_hostsim_set_obj_sum_mode(int mode) { g_obj_sum_mode = mode; }

void flgpu_hostsim_builtin_problem(int kind, flgpu_problem *out) {
    out->f = obj_f; out->fd = obj_fd; out->f_fd = obj_ffd; out->user = (void *)(intptr_t)kind;
    out->fused = obj_fused;
    out->search = nullptr;
    out->search_caps = 0;
    out->update = nullptr;
}

void flgpu_hostsim_options_default(flgpu_options *o, int for_cg) {
    std::memset(o, 0, sizeof *o);
    o->memory = 10; o->method = FLGPU_CG_DY; o->strong = 1; o->warning = 1; o->max_iteration = 1000;
    o->precision = 1e-15; o->min_step_length = 1e-15; o->wolfe_c1 = 1e-4;
    o->wolfe_c2 = for_cg ? 0.45 : 0.9; o->increment = 1.05; o->device_search = 2;
}

int flgpu_hostsim_lbfgs(const flgpu_problem *prob, const flgpu_options *opt, double *x, int64_t n,
                        flgpu_stats *stats) {
    HostBackend B(*prob, n, opt->offset, opt->n_global);
    flgpu::Params P = flgpu::params_from_options(*opt, false, prob->f_fd != nullptr);
    flgpu_stats st;
    flgpu::run_lbfgs(B, P, x, 0, &st);
    if (stats) *stats = st;
    return 0;
}

int flgpu_hostsim_sd(const flgpu_problem *prob, const flgpu_options *opt, double *x, int64_t n,
                     flgpu_stats *stats) {
    HostBackend B(*prob, n, opt->offset, opt->n_global);
    flgpu::Params P = flgpu::params_from_options(*opt, false, prob->f_fd != nullptr);
    flgpu_stats st;
    flgpu::run_sd(B, P, x, 0, &st);
    if (stats) *stats = st;
    return 0;
}

int flgpu_hostsim_cg(const flgpu_problem *prob, const flgpu_options *opt, double *x, int64_t n,
                     flgpu_stats *stats) {
    HostBackend B(*prob, n, opt->offset, opt->n_global);
    flgpu::Params P = flgpu::params_from_options(*opt, true, prob->f_fd != nullptr);
    flgpu_stats st;
    flgpu::run_cg(B, P, x, 0, &st);
    if (stats) *stats = st;
    return 0;
}

}  // extern "C"

// ---- two-loop recursion as an operator (mirror of flgpu_history_* in libflgpu.so)
struct hostsim_history {
    HostBackend *B;
    flgpu::History *H;
};
extern "C" {
hostsim_history *flgpu_hostsim_history_create(int64_t n, int memory) {
    flgpu_problem none{};
    hostsim_history *h = new hostsim_history;
    h->B = new HostBackend(none, n, 0, n);
    h->H = new flgpu::History(*h->B, memory);
    return h;
}
int flgpu_hostsim_history_push(hostsim_history *h, const double *x1, const double *x0, const double *g1,
                               const double *g0) {
    h->H->push(x1, x0, g1, g0);
    return 0;
}
int flgpu_hostsim_history_direction(hostsim_history *h, const double *g1, const double *x1, double *p, double *xt,
                                    double *gp, double *pp) {
    h->H->direction(g1, x1, p, xt, gp, pp);
    return 0;
}
void flgpu_hostsim_history_destroy(hostsim_history *h) {
    delete h->H;
    delete h->B;
    delete h;
}
}
