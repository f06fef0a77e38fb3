// tests/hostsim/backend_host.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A host-memory implementation of fortran_library_b200/csrc/backend.hpp, so that the host
// control flow in driver.cpp (the same translation unit the product links) can be checked against
// the oracle on a machine without a GPU, including the world_size>1 row-shard combination
// over a caller-supplied all-gather (gloo in the tests).  Built into
// tests/hostsim/libflgpu_hostsim.so; libflgpu.so never links or loads it and has no CPU path.
//
// Element-wise arithmetic matches the CUDA kernels (separate multiply and add roundings), and so do the
// REDUCTIONS: namespace model below restates include/flgpu_reduce.cuh and each kernel's per-chunk thread order in
// scalar C++ (which thread adds which 16-byte unit, FMA accumulation, lane butterfly, warps left to right, aligned
// binary tree over chunks, rank tree) -- an independent statement of the same arithmetic, so a run of the host
// simulator and a run of libflgpu.so on the GPU agree BIT FOR BIT (tests/test_gpu.py::test_gpu_equals_host_simulator).
#include <algorithm>
#include <cmath>
#include <cstdint>
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

// ---- scalar model of the CUDA reductions (include/flgpu_reduce.cuh + the kernels' per-chunk thread order)
namespace model {

// aligned binary tree over v[0 .. count) with +0.0 for the missing leaves (cta_tree + top_tree: 2^18 leaves)
double node(const std::vector<double> &v, int64_t lo, int64_t hi) {
    if (lo >= (int64_t)v.size()) return 0.0;
    if (hi - lo == 1) return v[(size_t)lo];
    const int64_t mid = lo + (hi - lo) / 2;
    return node(v, lo, mid) + node(v, mid, hi);
}
double tree(const std::vector<double> &chunk_sums) {
    return node(chunk_sums, 0, (int64_t)flgpu::red::kTopMax * flgpu::red::kBlockChunks);
}

// A reducing kernel: NACC accumulators per thread.  Each chunk of ch elements is summed by a block of 8 warps; a warp
// holds SEG lanes of the accumulator's column group (SEG = 32 except in K1 with two groups per warp), so TX = 8 * SEG
// threads share the chunk's 16-byte units: thread tx takes units lo+tx, lo+tx+TX, ... in order (unit(u, acc)); the odd
// last element of the shard is added by thread 0 of the last chunk after its units (tail(acc)); lanes are combined by an
// xor butterfly (far first), the 8 warps left to right; the chunk sums by the aligned binary tree.
template <int NACC, class Unit, class Tail>
void reduce(int64_t n, int64_t ch, int SEG, Unit unit, Tail tail, double (&out)[NACC]) {
    const int TX = 8 * SEG;
    const int64_t nu = n >> 1, cu = ch >> 1, nchunks = flgpu::red::num_chunks(n, ch);
    std::vector<double> sums[NACC];
    for (int i = 0; i < NACC; i++) sums[i].assign((size_t)nchunks, 0.0);
    std::vector<double> acc((size_t)TX * NACC);
    for (int64_t c = 0; c < nchunks; c++) {
        const int64_t lo = c * cu, hi = (lo + cu < nu) ? lo + cu : nu;
        std::fill(acc.begin(), acc.end(), 0.0);
        for (int tx = 0; tx < TX; tx++)
            for (int64_t u = lo + tx; u < hi; u += TX) unit(u, &acc[(size_t)tx * NACC]);
        if ((n & 1) && c == nchunks - 1) tail(&acc[0]);
        for (int i = 0; i < NACC; i++) {
            double s = 0.0;
            for (int q = 0; q < 8; q++) {
                double v[32], t[32];
                for (int l = 0; l < SEG; l++) v[l] = acc[(size_t)(q * SEG + l) * NACC + i];
                for (int o = SEG / 2; o > 0; o >>= 1) {
                    for (int l = 0; l < SEG; l++) t[l] = v[l] + v[l ^ o];
                    for (int l = 0; l < SEG; l++) v[l] = t[l];
                }
                s += v[0];
            }
            sums[i][(size_t)c] = s;
        }
    }
    for (int i = 0; i < NACC; i++) out[i] = tree(sums[i]);
}

// dot_kernel: a.b
double dot(const double *a, const double *b, int64_t n, int64_t ch) {
    double out[1];
    reduce<1>(n, ch, 32,
              [&](int64_t u, double *acc) { acc[0] = std::fma(a[2 * u + 1], b[2 * u + 1], std::fma(a[2 * u], b[2 * u], acc[0])); },
              [&](double *acc) { acc[0] = std::fma(a[n - 1], b[n - 1], acc[0]); }, out);
    return out[0];
}
// objective kernels: f = sum of per-element terms (a unit adds its first element's term, then its second's;
// Rosenbrock has one term per unit), f'.p as a dot
double fsum(const double *terms, int64_t n, int64_t ch, bool one_term_per_unit) {
    double out[1];
    reduce<1>(n, ch, 32,
              [&](int64_t u, double *acc) { acc[0] += terms[2 * u]; if (!one_term_per_unit) acc[0] += terms[2 * u + 1]; },
              [&](double *acc) { acc[0] += terms[n - 1]; }, out);
    return out[0];
}

}  // namespace model

class HostBackend : public flgpu::Backend {
public:
    flgpu_problem prob;
    flgpu_eval_ctx ctx;
    std::vector<double *> owned;
    double slots[flgpu::NSLOTS];
    int mem = 0;
    double *S = nullptr, *Y = nullptr;
    std::vector<double> SY, YY, D, C, w1, w2, w3;
    int64_t ch = 1024;     // chunk elements, from the GLOBAL dimension (flgpu_reduce_geom.h)

    HostBackend(const flgpu_problem &p, int64_t n_, int64_t offset, int64_t n_global) : prob(p) {
        n = n_;
        std::memset(slots, 0, sizeof slots);
        ctx.user = p.user; ctx.stream = nullptr; ctx.offset = offset;
        ctx.n_global = n_global ? n_global : n_; ctx.rank = g_rank; ctx.nranks = g_nranks; ctx.device = -1;
        ch = flgpu::red::chunk_elems(ctx.n_global);
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
    bool fused_multi_available() const override { return prob.fused_multi != nullptr; }
    void fused_eval_multi(int count, const double *steps, const double *x0, const double *p) override {
        prob.fused_multi(&ctx, count, steps, &slots[flgpu::SL_AUX], x0, p, n);
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
        slots[slot] = model::dot(a, b, n, ch);
    }
    void neg(double *p, const double *g) override {
        launches++;
        for (long i = 0; i < n; i++) p[i] = -g[i];
    }

    // K1 (include/flgpu_k1.cuh): the new column pair, then every dot in the thread order of the pass that owns it --
    // pass shapes from k1_pass_shape(); the dots of the new column and g.g belong to group 0 of the first pass
    void lbfgs_update_dots(const double *x1, const double *x0, const double *g1, const double *g0,
                           int new_slot, int k_after) override {
        launches++;
        const int m = mem;
        double *sn = S + (long)new_slot * n, *yn = Y + (long)new_slot * n;
        for (long i = 0; i < n; i++) { sn[i] = x1[i] - x0[i]; yn[i] = g1[i] - g0[i]; }
        const int nother = k_after - 1;
        int age = 1;
        bool first = true;
        do {
            int mt, ng;
            flgpu::k1_pass_shape(nother - (age - 1), mt, ng);
            const int SEG = 32 / ng;
            auto four = [&](int j) {          // s_j.g1, y_j.g1, s_j.y_new, y_j.y_new
                const double *sj = S + (long)j * n, *yj = Y + (long)j * n;
                double out[4];
                model::reduce<4>(n, ch, SEG,
                    [&](int64_t u, double *acc) {
                        const long a = 2 * u, b = 2 * u + 1;
                        acc[0] = std::fma(sj[b], g1[b], std::fma(sj[a], g1[a], acc[0]));
                        acc[1] = std::fma(yj[b], g1[b], std::fma(yj[a], g1[a], acc[1]));
                        acc[2] = std::fma(sj[b], yn[b], std::fma(sj[a], yn[a], acc[2]));
                        acc[3] = std::fma(yj[b], yn[b], std::fma(yj[a], yn[a], acc[3]));
                    },
                    [&](double *acc) {
                        const long i = n - 1;
                        acc[0] = std::fma(sj[i], g1[i], acc[0]); acc[1] = std::fma(yj[i], g1[i], acc[1]);
                        acc[2] = std::fma(sj[i], yn[i], acc[2]); acc[3] = std::fma(yj[i], yn[i], acc[3]);
                    }, out);
                D[flgpu::d_A(m, j)] = out[0]; D[flgpu::d_B(m, j)] = out[1];
                D[flgpu::d_SYN(m, j)] = out[2]; D[flgpu::d_YYN(m, j)] = out[3];
            };
            if (first) {                      // g.g, sn.g, yn.g, sn.yn, yn.yn
                double out[5];
                model::reduce<5>(n, ch, SEG,
                    [&](int64_t u, double *acc) {
                        const long a = 2 * u, b = 2 * u + 1;
                        acc[0] = std::fma(g1[b], g1[b], std::fma(g1[a], g1[a], acc[0]));
                        acc[1] = std::fma(sn[b], g1[b], std::fma(sn[a], g1[a], acc[1]));
                        acc[2] = std::fma(yn[b], g1[b], std::fma(yn[a], g1[a], acc[2]));
                        acc[3] = std::fma(sn[b], yn[b], std::fma(sn[a], yn[a], acc[3]));
                        acc[4] = std::fma(yn[b], yn[b], std::fma(yn[a], yn[a], acc[4]));
                    },
                    [&](double *acc) {
                        const long i = n - 1;
                        acc[0] = std::fma(g1[i], g1[i], acc[0]); acc[1] = std::fma(sn[i], g1[i], acc[1]);
                        acc[2] = std::fma(yn[i], g1[i], acc[2]); acc[3] = std::fma(sn[i], yn[i], acc[3]);
                        acc[4] = std::fma(yn[i], yn[i], acc[4]);
                    }, out);
                D[flgpu::d_GG(m)] = out[0];
                D[flgpu::d_A(m, new_slot)] = out[1]; D[flgpu::d_B(m, new_slot)] = out[2];
                D[flgpu::d_SYN(m, new_slot)] = out[3]; D[flgpu::d_YYN(m, new_slot)] = out[4];
            }
            for (int t = age; t < age + mt * ng && t < k_after; t++) four(flgpu::slot_of_age(new_slot, t, m));
            age += mt * ng;
            first = false;
        } while (age - 1 < nother);
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
        double out[2];                      // K3: g.p and p.p in one kernel (256 threads per chunk)
        model::reduce<2>(n, ch, 32,
            [&](int64_t u, double *acc) {
                const long a = 2 * u, b = 2 * u + 1;
                acc[0] = std::fma(g1[b], p[b], std::fma(g1[a], p[a], acc[0]));
                acc[1] = std::fma(p[b], p[b], std::fma(p[a], p[a], acc[1]));
            },
            [&](double *acc) { const long i = n - 1; acc[0] = std::fma(g1[i], p[i], acc[0]); acc[1] = std::fma(p[i], p[i], acc[1]); },
            out);
        slots[flgpu::SL_GP0] = out[0];
        slots[flgpu::SL_PP] = out[1];
    }
    // K3 with the first trials of the next search (flgpu_problem.direction on the GPU): by definition the direction
    // followed by separate fused evaluations at x1 + steps[j]*p -- what the CUDA kernel must reproduce bit for bit
    bool fused_direction_available() const override { return prob.fused != nullptr; }
    void lbfgs_direction_probe(double *p, const double *g1, const double *x1, int k, int recent, int /*flags*/,
                               const double *steps) override {
        lbfgs_direction(p, nullptr, g1, x1, k, recent);
        for (int j = 0; j < FLGPU_MULTI_MAX; j++) {
            double *f = j == 0 ? &slots[flgpu::SL_F] : &slots[flgpu::SL_AUX + 2 * j];
            double *gp = j == 0 ? &slots[flgpu::SL_GP] : &slots[flgpu::SL_AUX + 2 * j + 1];
            prob.fused(&ctx, FLGPU_WANT_F | FLGPU_WANT_GP, f, gp, nullptr, nullptr, x1, p, steps[j], n);
        }
        callback_launches++;
    }

    void cg_dots(const double *g1, const double *g0, const double *p) override {
        launches++;
        auto term = [&](long i, double *acc) {        // cg_dots_kernel's `term`
            const double a = g1[i], b = g0[i], q = p[i], d = a - b;
            acc[0] = std::fma(a, a, acc[0]); acc[1] = std::fma(q, q, acc[1]); acc[2] = std::fma(d, q, acc[2]);
            acc[3] = std::fma(a, d, acc[3]); acc[4] = std::fma(b, b, acc[4]);
        };
        double out[5];
        model::reduce<5>(n, ch, 32, [&](int64_t u, double *acc) { term(2 * u, acc); term(2 * u + 1, acc); },
                         [&](double *acc) { term(n - 1, acc); }, out);
        slots[flgpu::SL_GG] = out[0]; slots[flgpu::SL_PP] = out[1]; slots[flgpu::SL_DGP] = out[2];
        slots[flgpu::SL_GDG] = out[3]; slots[flgpu::SL_G0G0] = out[4];
    }
    void cg_update(double *p, const double *g1, double beta) override {
        launches++;
        for (long i = 0; i < n; i++) p[i] = -g1[i] + beta * p[i];
        slots[flgpu::SL_GP0] = model::dot(g1, p, n, ch);
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

// ---- objective adapters over oracle/objectives.c (host pointers): element-wise values from the oracle's formulas,
// f and f'.p reduced in the CUDA objective kernels' order (model::fsum / model::dot)
int g_obj_sum_mode = 2;    // (kept for ABI compatibility of the test helper; the model has one summation order)
double model_f(const flgpu_eval_ctx *c, const double *x, int64_t n) {
    int d = (int)n;
    const int kind = (int)(intptr_t)c->user;
    orc_obj_select(kind, c->offset, c->n_global);
    std::vector<double> terms((size_t)(n > 0 ? n : 1));
    orc_obj_terms(terms.data(), x, &d);
    return model::fsum(terms.data(), n, flgpu::red::chunk_elems(c->n_global ? c->n_global : n), kind == ORC_OBJ_ROSENBROCK);
}
void obj_f(const flgpu_eval_ctx *c, double *f, const double *x, int64_t n) { *f = model_f(c, x, n); }
void obj_fd(const flgpu_eval_ctx *c, double *g, const double *x, int64_t n) {
    int d = (int)n;
    orc_obj_select((int)(intptr_t)c->user, c->offset, c->n_global);
    orc_obj_fd(g, x, &d);
}
void obj_ffd(const flgpu_eval_ctx *c, double *f, double *g, const double *x, int64_t n) {
    obj_fd(c, g, x, n);
    *f = model_f(c, x, n);
}

// fused evaluation (flgpu_fused_fn) on host memory: the point is formed element-wise, multiply then add
void obj_fused(const flgpu_eval_ctx *c, int flags, double *f, double *gp, double *x_out, double *g_out,
               const double *x0, const double *p, double a, int64_t n) {
    std::vector<double> x((size_t)(n > 0 ? n : 1)), g((size_t)(n > 0 ? n : 1));
    for (int64_t i = 0; i < n; i++) x[i] = x0[i] + a * p[i];
    obj_fd(c, g.data(), x.data(), n);
    if (flags & FLGPU_WANT_F) *f = model_f(c, x.data(), n);
    if (flags & FLGPU_WANT_GP) *gp = model::dot(g.data(), p, n, flgpu::red::chunk_elems(c->n_global ? c->n_global : n));
    if (flags & FLGPU_WRITE_X) std::memcpy(x_out, x.data(), sizeof(double) * n);
    if (flags & FLGPU_WRITE_G) std::memcpy(g_out, g.data(), sizeof(double) * n);
}

// batched fused evaluation (flgpu_fused_multi_fn): by definition the values of `count` separate fused evaluations
void obj_fused_multi(const flgpu_eval_ctx *c, int count, const double *steps, double *out, const double *x0,
                     const double *p, int64_t n) {
    for (int j = 0; j < count; j++)
        obj_fused(c, FLGPU_WANT_F | FLGPU_WANT_GP, &out[2 * j], &out[2 * j + 1], nullptr, nullptr, x0, p, steps[j], n);
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
    out->direction = nullptr;
    out->fused_multi = obj_fused_multi;
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
