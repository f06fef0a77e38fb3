// driver.cpp -- host control flow of the hot path: LBFGS (f90:398-625), ConjugateGradient
// (f90:193-394, 2249-2346) and the Wolfe / Strong-Wolfe line searchers (f90:1286-1698), written
// against the asynchronous vector operations of backend.hpp.  "f90:" = the reference's
// source/NonlinearOptimization.f90.
//
// What differs from the reference, by design:
//  * vectors never leave the device: x0/x and f'old/f'new are pairs of buffers that swap roles
//    when a step is accepted, so `x0=x` (f90:1480), `xold=x; fdold=fdnew` (f90:587) cost nothing;
//  * every host decision is taken from scalars delivered by Backend::fetch(), one
//    synchronisation per line-search trial;
//  * LBFGS: Before()/After() (f90:586-624) become one stream-ordered chain K1 -> K2 -> K3 ->
//    first trial evaluation that is enqueued speculatively right after a step is accepted;
//    the convergence tests of After() are applied to the scalars that chain returns.
// Every branch, comparison, default and exit test follows the reference statement by statement.
#include "driver.hpp"
#include "../../include/flgpu_search_core.hpp"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <utility>

namespace flgpu {

Params params_from_options(const flgpu_options &o, bool for_cg, bool has_f_fd) {
    Params P;
    P.mem = o.memory > 1 ? o.memory : 1;                                    // f90:419
    P.method = o.method;
    P.strong = o.strong != 0;
    P.warn = o.warning != 0;
    P.maxit = o.max_iteration;
    P.tol = o.precision * o.precision;                                      // f90:427
    P.minstep = o.min_step_length * o.min_step_length;                      // f90:429
    if (o.no_clamp) {                                                       // f90:2265-2278
        P.c1 = o.wolfe_c1;
        P.c2 = o.wolfe_c2;
    } else {
        P.c1 = std::fmax(1e-15, o.wolfe_c1);                                // f90:431
        P.c2 = std::fmin(1.0 - 1e-15, std::fmax(P.c1 + 1e-15, o.wolfe_c2)); // f90:433
    }
    P.incr = std::fmax(1.0 + 1e-15, o.increment);                           // f90:1478
    P.has_f_fd = has_f_fd;
    P.fused = !o.no_fused;
    P.device_search = o.device_search;   // 0 off, 1 on, 2 auto (by size, see use_device_search)
    P.line_search = o.line_search == FLGPU_LS_FAST ? FLGPU_LS_FAST : FLGPU_LS_REFERENCE;
    P.observer = o.observer;
    P.observer_user = o.observer_user;
    (void)for_cg;
    return P;
}

namespace {

// One line search along p from x0 (device buffers); trial points go to xt, trial gradients to gt.  The reference's
// control flow is SearchCore (flgpu_search_core.hpp); this class supplies the evaluations: asynchronous kernels plus one
// host round trip whenever a value steers a branch.
struct Search : SearchCore<Search> {
    Backend &B;
    flgpu_stats &st;
    const double *x0;
    double *xt, *gt;
    const double *p;
    int64_t trials = 0;

    Search(Backend &b, flgpu_stats &s) : B(b), st(s) {}

    // ---- Fortran `fx`, filled lazily from the device
    double fx_ = 0.0;
    bool f_pending = false;
    double slots[NSLOTS];

    // ---- batched evaluation (flgpu_fused_multi_fn).  While the reference's searchers bracket, every next step is the
    // previous one times or divided by `incr` (f90:1499-1501, 1488-1490, 1518, 1308-1310, 1325): when a trial continues
    // such a walk -- or is the first of a search -- the next FLGPU_MULTI_MAX steps of the walk are evaluated in the same
    // pass over x0 and p.  The state machine then finds the following trials already evaluated (same bits as separate
    // probes) and takes its decisions without a launch or a host round trip; what it never asks for is dropped.
    bool multi = false;
    int nb = 0;                                   // steps of the current batch
    double b_a[FLGPU_MULTI_MAX], b_f[FLGPU_MULTI_MAX], b_gp[FLGPU_MULTI_MAX];
    bool b_inflight = false;                      // launched; values not on the host yet
    int f_idx = -1, gp_idx = -1;                  // source of the pending f / of the current f'.p: batch entry, or -1 = slots
    double prev_a = 0.0;                          // the step formed before the current one
    bool have_prev = false;

    void sync() {
        B.fetch(slots);
        st.host_syncs++;
        if (b_inflight) {
            for (int j = 0; j < nb; j++) { b_f[j] = slots[SL_AUX + 2 * j]; b_gp[j] = slots[SL_AUX + 2 * j + 1]; }
            b_inflight = false;
        }
        if (f_pending) { fx_ = f_idx >= 0 ? b_f[f_idx] : slots[SL_F]; f_pending = false; }
    }
    double fx() { if (f_pending) sync(); return fx_; }
    void set_fx(double v) { f_pending = false; fx_ = v; }
    void count_f_only() { st.n_f_only_trials++; }
    static bool aborted() { return false; }   // the reference searches until its own exit tests fire

    // Fused mode (flgpu_fused_fn): a trial point exists only as its step a_x until the search returns;
    // each callback of the reference becomes one probe kernel that forms x0+a_x*p on the fly, and
    // finish() stores the point and gradient the reference would have left in x / fdx.
    bool fused = false;
    double a_x = 0.0, a_g = 0.0;   // step of the last point formed / of the last gradient evaluated
    bool have_x = false, have_g = false;

    void form(double step) {                                                         // x=x0+a*p
        if (fused) {
            if (have_x) { prev_a = a_x; have_prev = true; }
            a_x = step; have_x = true;
        }
        else B.trial_x(xt, x0, p, step);
        trials++; st.n_trials++;
    }
    // The batch entry that holds f and f'.p at a_x (launching a batch if a_x starts or continues a walk); -1: none
    int batch_entry() {
        if (!multi) return -1;
        for (int j = 0; j < nb; j++) if (b_a[j] == a_x) return j;
        int dir = 0;
        if (!have_prev) dir = 1;                          // first trial: growth follows whenever it is acceptable
        else if (a_x == prev_a * incr) dir = 1;
        else if (a_x == prev_a / incr) dir = -1;
        if (dir == 0) return -1;                          // a zoom step: nothing to predict
        if (b_inflight || (f_pending && f_idx >= 0)) sync();   // the old batch is about to be overwritten
        nb = FLGPU_MULTI_MAX;
        b_a[0] = a_x;
        for (int j = 1; j < nb; j++) b_a[j] = dir > 0 ? b_a[j - 1] * incr : b_a[j - 1] / incr;
        B.fused_eval_multi(nb, b_a, x0, p);
        b_inflight = true;
        st.n_batched_passes++;
        return 0;
    }
    void take_f(int j) {
        f_idx = j;
        if (j >= 0 && !b_inflight) { fx_ = b_f[j]; f_pending = false; }
        else f_pending = true;
    }
    void call_f() {
        if (fused) {
            const int j = batch_entry();
            if (j < 0) B.fused_eval(FLGPU_WANT_F, a_x, x0, p, nullptr, nullptr);
            take_f(j);
        } else { B.eval_f(xt); f_pending = true; }
        st.n_f++;
    }
    void call_fd() {
        if (fused) {
            const int j = batch_entry();
            if (j < 0) B.fused_eval(FLGPU_WANT_GP, a_x, x0, p, nullptr, nullptr);
            gp_idx = j; a_g = a_x; have_g = true;
        }
        else B.eval_g(xt, gt);
        st.n_fd++;
    }
    void call_ffd() {
        if (fused) {
            const int j = batch_entry();
            if (j < 0) B.fused_eval(FLGPU_WANT_F | FLGPU_WANT_GP, a_x, x0, p, nullptr, nullptr);
            take_f(j);
            gp_idx = j; a_g = a_x; have_g = true;
        } else { B.eval_fg(xt, gt); f_pending = true; }
        st.n_f_fd++;
    }
    double slope() {                                                                 // dot_product(fdx,p)
        if (!fused) B.dot(gt, p, SL_GP);       // fused: f'.p was reduced by the probe that evaluated f'
        if (fused && gp_idx >= 0) {            // ... or by the batch that did
            if (b_inflight || f_pending) sync();
            return b_gp[gp_idx];
        }
        sync();
        return slots[SL_GP];
    }
    // the caller's chain already formed the first trial point (and evaluated it)
    void adopt_pre() {
        trials++; set_fx(pre_f);
        a_x = a; have_x = true;
        if (pre == 3) { a_g = a; have_g = true; gp_idx = -1; }
    }
    // the accepted point is x0 + a_x*p and its gradient was evaluated there: the reference's searchers always end so
    bool can_defer() const { return fused && have_x && have_g && a_x == a_g; }
    void finish() {
        if (!fused) return;
        if (have_x && have_g && a_x == a_g) {
            B.fused_eval(FLGPU_WRITE_X | FLGPU_WRITE_G, a_x, x0, p, xt, gt);
        } else {                               // never taken by the reference's searchers; kept for fidelity
            if (have_x) B.trial_x(xt, x0, p, a_x);
            if (have_g) B.fused_eval(FLGPU_WRITE_G, a_g, x0, p, nullptr, gt);
        }
    }
};

// Device-resident search (flgpu_search_fn) gives the same bits as the host-driven fused search; it pays when the
// host round trip per trial (~15 us) is large next to a probe kernel.  Measured with the chunked reductions of round 2
// (profiles/r02_device_search.md): 1.34x at 2^18 rows, 0.87x at 2^20, 0.73x at 2^22, 0.93x at 2^26 on one GPU, and
// 0.88-0.95x at 2^20..2^25 rows per GPU on two -- every block re-forms the tree over all chunk sums after the grid
// barrier, which the separate tree kernel of the host-driven path does once.  Auto mode therefore uses it up to 2^18
// rows per GPU.
bool use_device_search(const Params &P, Backend &B) {
    if (!P.fused || P.device_search == 0 || !B.device_search_available()) return false;
    if (P.device_search == 1) return true;
    // FLGPU_LS_FAST accepts the first trial most of the time, and LBFGS evaluates that trial inside the speculative
    // K1->K2->K3 chain: the host-driven search already costs one round trip per iteration, a search kernel would add one
    if (P.line_search == FLGPU_LS_FAST) return false;
    return B.n <= ((int64_t)1 << 18);
}

// FLGPU_LS_FAST asks for f and f' at every trial.  A fused probe delivers both in one pass whether or not the problem
// has an f_fd callback; on the unfused path f_fd is used when present, else f then fd.
bool fast_uses_ffd(const Params &P, Backend &B) { return P.has_f_fd || (P.fused && B.fused_available()); }

// Runs one line search; on return xt/gt hold the accepted point and gradient -- unless the caller asked to defer the
// store (L-BFGS with flgpu_problem.update: K1 forms x0 + a*p and f' itself) and the result says `deferred`.
struct SearchResult { double a, fx; int64_t trials; bool deferred; };

// pre_steps / pre_vals (optional): the caller's chain evaluated the first FLGPU_MULTI_MAX steps of the bracketing walk
// already (K3 with a probe): steps, and f, f'.p at each.
SearchResult line_search(Backend &B, flgpu_stats &st, const Params &P, bool strong, bool fdwithf,
                         const double *x0, double *xt, double *gt, const double *p, double a,
                         double fx0, double phid0, int pre, double pre_f, double pre_gp, bool defer_store = false,
                         const double *pre_steps = nullptr, const double *pre_vals = nullptr) {
    Search S(B, st);
    S.x0 = x0; S.xt = xt; S.gt = gt; S.p = p;
    S.c1 = P.c1; S.c2abs = P.c2 * std::fabs(phid0); S.fx0 = fx0; S.phid0 = phid0; S.incr = P.incr;
    S.a = a; S.fx_ = fx0;
    S.pre = pre; S.pre_f = pre_f; S.pre_gp = pre_gp;
    S.fused = P.fused && B.fused_available();
    const bool fast = P.line_search == FLGPU_LS_FAST;
    // the fast policy's steps come from interpolation, not from a walk: nothing to batch (FLGPU_FUSED_MULTI=0: one probe
    // per trial -- the same bits, more passes; kept for comparison)
    const char *fm = std::getenv("FLGPU_FUSED_MULTI");
    S.multi = S.fused && !fast && B.fused_multi_available() && !(fm && fm[0] == '0');
    if (S.multi && pre != 0 && pre_steps) {       // the walk's first steps are on the host already
        S.nb = FLGPU_MULTI_MAX;
        for (int j = 0; j < S.nb; j++) { S.b_a[j] = pre_steps[j]; S.b_f[j] = pre_vals[2 * j]; S.b_gp[j] = pre_vals[2 * j + 1]; }
    }
    // FLGPU_LS_FAST evaluates f and f' together at every trial: one fused probe, else f_fd when the problem has one
    if (fast) fdwithf = fast_uses_ffd(P, B);
    S.fdwithf = fdwithf;
    st.n_linesearch++;
    if (S.fused && pre == 0 && use_device_search(P, B)) {
        // the same SearchCore, run by every thread of one cooperative kernel; one host round trip per search
        double res[FLGPU_SEARCH_RESULT_DOUBLES], slots[NSLOTS];
        B.device_search(P.line_search, strong, fdwithf, S.c1, S.c2abs, fx0, phid0, S.incr, a, x0, p, xt, gt, defer_store);
        B.fetch(slots); st.host_syncs++;
        B.search_result(res);
        st.n_trials += (int64_t)res[2]; st.n_f += (int64_t)res[3]; st.n_fd += (int64_t)res[4];
        st.n_f_fd += (int64_t)res[5]; st.n_f_only_trials += (int64_t)res[6];
        B.credit_search_bytes(8.0 * (double)B.n * (2.0 * (res[3] + res[4] + res[5]) + 4.0));   // 2n per evaluation + the store
        if (res[7] < 0.0)
            std::printf(" Line search warning: the device-resident search gave up after %.0f evaluations "
                        "(does the objective return NaN?)\n", res[3] + res[4] + res[5]);
        SearchResult r;
        r.a = res[0]; r.fx = res[1]; r.trials = (int64_t)res[2]; r.deferred = defer_store;
        return r;
    }
    if (fast) S.fast(strong);
    else if (strong) S.strongwolfe(); else S.wolfe();
    SearchResult r;
    r.deferred = defer_store && S.can_defer() && S.a_x == S.a;
    if (!r.deferred) S.finish();
    r.fx = S.fx();
    r.a = S.a;
    r.trials = S.trials;
    return r;
}

bool observe(const Params &P, Backend &B, const flgpu_stats &st, int64_t it, double a, double f, double phid0,
             int64_t trials, const double *p, const double *x, const double *g) {
    if (!P.observer) return false;
    flgpu_iter_info info;
    info.iteration = it; info.n_local = B.n; info.step = a; info.f = f; info.phid0 = phid0;
    info.trials = trials; info.p_dev = p; info.x_dev = x; info.g_dev = g; info.stream = B.stream_handle();
    info.gpu_launches = B.launches; info.callbacks = B.callback_launches; info.total_trials = st.n_trials;
    return P.observer(P.observer_user, &info) != 0;
}

void step_warning(const char *who, double gg) {
    std::printf(" %s warning: step length has converged, but gradient norm has not met accuracy goal\n", who);
    std::printf(" Euclidean norm of gradient = %.17g\n", std::sqrt(gg));
}

}  // namespace

// --------------------------------------------------------------------------- LBFGS
void run_lbfgs(Backend &B, const Params &P, double *x_user, int x_space, flgpu_stats *stp) {
    flgpu_stats &st = *stp;
    std::memset(&st, 0, sizeof st);
    double slots[NSLOTS];
    double *xc = B.vec_alloc(), *xo = B.vec_alloc();   // current (accepted) x / the other buffer
    double *gc = B.vec_alloc(), *go = B.vec_alloc();   // f' at xc / the other buffer
    double *p = B.vec_alloc();
    const int mem = P.mem;
    const bool fused = P.fused && B.fused_available();
    const bool dsearch = fused && use_device_search(P, B);   // the search kernel does every trial
    // K1 forms and stores the accepted point itself (FLGPU_FUSED_UPDATE=0: the search stores it, K1 reads it back --
    // the same bits, 3n more doubles of traffic; kept for comparison)
    const char *fu = std::getenv("FLGPU_FUSED_UPDATE");
    const bool fuse_k1 = fused && B.fused_update_available() && !(fu && fu[0] == '0');
    // K3 evaluates the first trial (a = 1) of the next search while it writes p (FLGPU_FUSED_DIRECTION=0: a separate
    // probe launch follows K3 -- the same bits, n more doubles of traffic; kept for comparison)
    const char *fd_env = std::getenv("FLGPU_FUSED_DIRECTION");
    const bool fuse_k3 = fused && B.fused_direction_available() && !(fd_env && fd_env[0] == '0');
    B.lbfgs_alloc(mem);
    B.upload(xc, x_user, x_space);

    // f90:436-445
    if (P.has_f_fd) { B.eval_fg(xc, gc); st.n_f_fd++; }
    else { B.eval_f(xc); st.n_f++; B.eval_g(xc, gc); st.n_fd++; }
    B.dot(gc, gc, SL_GG);
    B.fetch(slots); st.host_syncs++;
    double fnew = slots[SL_F];
    double gg = slots[SL_GG];
    st.f = fnew; st.gnorm2 = gg;
    if (gg < P.tol) { st.status = FLGPU_INITIAL_CONVERGED; goto finish; }    // f90:443 (-phidnew<tol)
    {
        B.neg(p, gc);                                                        // p=-fdnew
        double phid0 = -gg;
        double a = (fnew == 0.0) ? 1.0 : std::fabs(fnew) / std::sqrt(gg);    // f90:444-445
        double pp = gg;                                                      // dot_product(p,p), p=-f'
        int recent = -1, k = 0;
        int64_t it = 0;                                                      // accepted steps so far
        const int64_t total = 1 + (int64_t)(mem - 1) + (int64_t)P.maxit;     // f90:448,472,511
        int pre = 0; double pre_f = 0, pre_gp = 0;
        // the bracketing walk every main-loop search starts with (a = 1, then a = a*Increment: f90:607, 1499-1501), formed
        // as the search will form it; K3 with a probe evaluates these steps while it writes p
        double walk[FLGPU_MULTI_MAX], walk_vals[2 * FLGPU_MULTI_MAX];
        walk[0] = 1.0;
        for (int j = 1; j < FLGPU_MULTI_MAX; j++) walk[j] = walk[j - 1] * P.incr;
        bool have_walk = false;
        for (;;) {
            // which searcher this outer iteration uses: never _fdwithf before the main loop (f90:448-498)
            const bool in_main = it >= mem;
            const bool fdwithf = in_main && P.has_f_fd;
            // the step after this search ends the run by count: no K1 follows, the search stores its point itself
            const bool defer = fuse_k1 && (it + 1 < total);
            SearchResult r = line_search(B, st, P, P.strong, fdwithf && P.strong, xc, xo, go, p, a, fnew,
                                         phid0, pre, pre_f, pre_gp, defer, have_walk ? walk : nullptr, walk_vals);
            std::swap(xc, xo); std::swap(gc, go);       // accepted point becomes current; xo/go = xold/fdold
            a = r.a; fnew = r.fx;
            st.iterations = ++it;
            st.f = fnew;
            int new_slot, k_after;
            if (it <= mem) { new_slot = recent + 1; k_after = k + 1; }       // f90:470,508 (append)
            else { new_slot = (recent + 1) % mem; k_after = mem; }           // f90:622 (overwrite oldest)
            // ---- After() f90:609-624 fused with the next Before() f90:586-608
            bool k1_done = false;
            if (r.deferred) {            // K1 with x1 = xo + a*p and f'(x1) formed in registers, stored to xc / gc
                B.lbfgs_update_dots_fused(a, xo, p, go, xc, gc, new_slot, k_after);
                k1_done = true;
            }
            const bool stop = observe(P, B, st, it - 1, a, fnew, phid0, r.trials, p, xc, gc);
            const bool last = (it >= total) || stop;
            const bool fast = P.line_search == FLGPU_LS_FAST;
            // fast: f and f' at the first trial of every search; reference: f_fd only in the main loop (f90:448-498)
            const bool next_both = fast;
            const bool next_fdwithf = fast ? fast_uses_ffd(P, B) : (it >= mem) && P.has_f_fd && P.strong;
            if (last) {
                if (!k1_done) B.dot(gc, gc, SL_GG);                          // (K1 delivers g.g as well)
            } else {
                if (!k1_done) B.lbfgs_update_dots(xc, xo, gc, go, new_slot, k_after);   // K1
                B.lbfgs_solve(k_after, new_slot);                            // K2
                if (dsearch) {                                               // K3: new p only
                    B.lbfgs_direction(p, nullptr, gc, xc, k_after, new_slot);
                } else if (fuse_k3) {                                        // K3: new p and the first trial (a=1) in one pass
                    B.lbfgs_direction_probe(p, gc, xc, k_after, new_slot,
                                            next_fdwithf ? (FLGPU_WANT_F | FLGPU_WANT_GP) : FLGPU_WANT_F, walk);
                    st.n_trials++;
                    if (next_fdwithf) st.n_f_fd++; else st.n_f++;
                } else if (fused) {                                          // K3: new p; first trial (a=1) probed
                    B.lbfgs_direction(p, nullptr, gc, xc, k_after, new_slot);
                    st.n_trials++;
                    if (next_fdwithf) { B.fused_eval(FLGPU_WANT_F | FLGPU_WANT_GP, 1.0, xc, p, nullptr, nullptr); st.n_f_fd++; }
                    else { B.fused_eval(FLGPU_WANT_F, 1.0, xc, p, nullptr, nullptr); st.n_f++; }
                } else {
                    B.lbfgs_direction(p, xo, gc, xc, k_after, new_slot);     // K3: new p, first trial point (a=1)
                    st.n_trials++;
                    if (next_fdwithf) { B.eval_fg(xo, go); st.n_f_fd++; B.dot(go, p, SL_GP); }
                    else if (next_both) { B.eval_f(xo); st.n_f++; B.eval_g(xo, go); st.n_fd++; B.dot(go, p, SL_GP); }
                    else { B.eval_f(xo); st.n_f++; }
                }
            }
            B.fetch(slots); st.host_syncs++;
            gg = slots[SL_GG];
            st.gnorm2 = gg;
            // the first trial of the next search was evaluated speculatively; if the run ends here the reference
            // never makes that evaluation, so it is taken back out of the statistics
            auto uncount_speculative = [&]() {
                if (last || dsearch) return;
                st.n_trials--;
                if (next_fdwithf) st.n_f_fd--;
                else { st.n_f--; if (next_both) st.n_fd--; }
            };
            if (gg < P.tol) { uncount_speculative(); st.status = FLGPU_CONVERGED; break; }   // f90:611-614
            if (pp * a * a < P.minstep) {                                    // f90:615-621
                if (P.warn) step_warning(it <= mem ? "BFGS" : "L-BFGS", gg);
                uncount_speculative();
                st.status = FLGPU_STEP_CONVERGED; break;
            }
            if (stop) { uncount_speculative(); st.status = FLGPU_STOPPED_BY_OBSERVER; break; }
            if (last) {                                                      // f90:580-583
                st.status = FLGPU_MAX_ITERATION;
                if (P.warn) {
                    std::printf(" Failed L-BFGS: max iteration exceeded!\n");
                    std::printf(" Euclidean norm of gradient = %.17g\n", std::sqrt(gg));
                }
                break;
            }
            recent = new_slot; k = k_after;
            phid0 = slots[SL_GP0];                                           // f90:607
            pp = slots[SL_PP];
            a = 1.0;
            pre = dsearch ? 0 : ((next_fdwithf || next_both) ? 3 : 2);
            pre_f = slots[SL_F];
            pre_gp = slots[SL_GP];
            have_walk = fuse_k3 && !dsearch;
            if (have_walk) {
                walk_vals[0] = slots[SL_F]; walk_vals[1] = slots[SL_GP];
                for (int j = 1; j < FLGPU_MULTI_MAX; j++) {
                    walk_vals[2 * j] = slots[SL_AUX + 2 * j]; walk_vals[2 * j + 1] = slots[SL_AUX + 2 * j + 1];
                }
            }
        }
    }
finish:
    B.download(x_user, xc, x_space);
    st.gpu_launches = B.launches;
}

// --------------------------------------------------------------------------- ConjugateGradient
void run_cg(Backend &B, const Params &P, double *x_user, int x_space, flgpu_stats *stp) {
    flgpu_stats &st = *stp;
    std::memset(&st, 0, sizeof st);
    double slots[NSLOTS];
    double *xc = B.vec_alloc(), *xo = B.vec_alloc();
    double *gc = B.vec_alloc(), *go = B.vec_alloc();
    double *p = B.vec_alloc();
    B.upload(xc, x_user, x_space);
    const bool is_pr = P.method == FLGPU_CG_PR;
    const bool strong = is_pr ? true : P.strong;                             // f90:311-344
    const char *who = is_pr ? "Polak-Ribiere+ conjugate gradient" : "Dai-Yuan conjugate gradient";

    if (P.has_f_fd) { B.eval_fg(xc, gc); st.n_f_fd++; }                      // f90:230-234
    else { B.eval_f(xc); st.n_f++; B.eval_g(xc, gc); st.n_fd++; }
    B.dot(gc, gc, SL_GG);
    B.fetch(slots); st.host_syncs++;
    double fnew = slots[SL_F];
    double gg = slots[SL_GG];
    st.f = fnew; st.gnorm2 = gg;
    if (gg < P.tol) { st.status = FLGPU_INITIAL_CONVERGED; goto finish; }    // f90:237
    {
        B.neg(p, gc);
        double phidnew = -gg;                                                // f90:236
        double a = (fnew == 0.0) ? 1.0 : std::fabs(fnew) / std::sqrt(gg);
        st.status = FLGPU_MAX_ITERATION;
        int64_t it = 0;
        for (int iIteration = 1; iIteration <= P.maxit; iIteration++) {
            const double phidold = phidnew;                                  // fdold=fdnew by buffer swap
            SearchResult r = line_search(B, st, P, strong, P.has_f_fd && strong, xc, xo, go, p, a, fnew,
                                         phidnew, 0, 0.0, 0.0);
            std::swap(xc, xo); std::swap(gc, go);
            a = r.a; fnew = r.fx;
            st.iterations = ++it;
            st.f = fnew;
            const bool stop = observe(P, B, st, it - 1, a, fnew, phidold, r.trials, p, xc, gc);
            // DY() f90:352-372 / PR() f90:373-393
            B.cg_dots(gc, go, p);
            B.fetch(slots); st.host_syncs++;
            gg = slots[SL_GG];
            st.gnorm2 = gg;
            if (gg < P.tol) { st.status = FLGPU_CONVERGED; break; }
            if (slots[SL_PP] * a * a < P.minstep) {
                if (P.warn) step_warning(who, gg);
                st.status = FLGPU_STEP_CONVERGED; break;
            }
            if (stop) { st.status = FLGPU_STOPPED_BY_OBSERVER; break; }
            const double beta = is_pr ? slots[SL_GDG] / slots[SL_G0G0]       // f90:387
                                      : gg / slots[SL_DGP];                  // f90:366
            B.cg_update(p, gc, beta);
            B.fetch(slots); st.host_syncs++;
            phidnew = slots[SL_GP0];
            if (phidnew > 0.0) {                                             // f90:368-370
                B.neg(p, gc);
                phidnew = -gg;
            }
            a = a * phidold / phidnew;                                       // f90:371
        }
        if (st.status == FLGPU_MAX_ITERATION && P.warn) {                    // f90:347-350
            std::printf(" Failed conjugate gradient: max iteration exceeded!\n");
            std::printf(" Euclidean norm of gradient = %.17g\n", std::sqrt(gg));
        }
    }
finish:
    B.download(x_user, xc, x_space);
    st.gpu_launches = B.launches;
}

// --------------------------------------------------------------------------- SteepestDescent
// f90:55-188.  After() (f90:172-187): g.g, p.p a^2, p = -g, a = a*phidold/phidnew.  p.p needs no pass:
// p = -f'old, so dot_product(p,p) is the previous g.g bit for bit.
void run_sd(Backend &B, const Params &P, double *x_user, int x_space, flgpu_stats *stp) {
    flgpu_stats &st = *stp;
    std::memset(&st, 0, sizeof st);
    double slots[NSLOTS];
    double *xc = B.vec_alloc(), *xo = B.vec_alloc();
    double *gc = B.vec_alloc(), *go = B.vec_alloc();
    double *p = B.vec_alloc();
    B.upload(xc, x_user, x_space);
    if (P.has_f_fd) { B.eval_fg(xc, gc); st.n_f_fd++; }                      // f90:86-90
    else { B.eval_f(xc); st.n_f++; B.eval_g(xc, gc); st.n_fd++; }
    B.dot(gc, gc, SL_GG);
    B.fetch(slots); st.host_syncs++;
    double fnew = slots[SL_F];
    double gg = slots[SL_GG];
    st.f = fnew; st.gnorm2 = gg;
    if (gg < P.tol) { st.status = FLGPU_INITIAL_CONVERGED; goto finish; }    // f90:93
    {
        B.neg(p, gc);                                                        // f90:92
        double phidnew = -gg;
        double a = (fnew == 0.0) ? 1.0 : std::fabs(fnew) / std::sqrt(gg);    // f90:94-95
        st.status = FLGPU_MAX_ITERATION;
        int64_t it = 0;
        for (int iIteration = 1; iIteration <= P.maxit; iIteration++) {
            const double phidold = phidnew;
            const double pp = gg;                                            // dot_product(p,p), p = -f'old
            // Strong -> StrongWolfe(_fdwithf); else Wolfe / Wolfe_fdwithf, which never calls f_fd (f90:1373)
            SearchResult r = line_search(B, st, P, P.strong, P.has_f_fd && P.strong, xc, xo, go, p, a, fnew,
                                         phidnew, 0, 0.0, 0.0);
            std::swap(xc, xo); std::swap(gc, go);
            a = r.a; fnew = r.fx;
            st.iterations = ++it;
            st.f = fnew;
            const bool stop = observe(P, B, st, it - 1, a, fnew, phidold, r.trials, p, xc, gc);
            B.dot(gc, gc, SL_GG);                                            // After() f90:172-187
            B.fetch(slots); st.host_syncs++;
            gg = slots[SL_GG];
            st.gnorm2 = gg;
            if (gg < P.tol) { st.status = FLGPU_CONVERGED; break; }
            if (pp * a * a < P.minstep) {
                if (P.warn) step_warning("Steepest descent", gg);
                st.status = FLGPU_STEP_CONVERGED; break;
            }
            if (stop) { st.status = FLGPU_STOPPED_BY_OBSERVER; break; }
            B.neg(p, gc);
            phidnew = -gg;
            a = a * phidold / phidnew;
        }
        if (st.status == FLGPU_MAX_ITERATION && P.warn) {                    // f90:168-171
            std::printf(" Failed steepest descent: max iteration exceeded!\n");
            std::printf(" Euclidean norm of gradient = %.17g\n", std::sqrt(gg));
        }
    }
finish:
    B.download(x_user, xc, x_space);
    st.gpu_launches = B.launches;
}

}  // namespace flgpu

namespace flgpu {

void History::push(const double *x1, const double *x0, const double *g1, const double *g0) {
    int new_slot, k_after;
    if (k_ < mem_) { new_slot = recent_ + 1; k_after = k_ + 1; }   // f90:470,508
    else { new_slot = (recent_ + 1) % mem_; k_after = mem_; }      // f90:622
    B.lbfgs_update_dots(x1, x0, g1, g0, new_slot, k_after);
    B.lbfgs_solve(k_after, new_slot);
    recent_ = new_slot;
    k_ = k_after;
}

void History::direction(const double *g1, const double *x1, double *p, double *xt, double *gp, double *pp) {
    double slots[NSLOTS];
    B.lbfgs_direction(p, xt, g1, x1, k_, recent_);
    B.fetch(slots);
    if (gp) *gp = slots[SL_GP0];
    if (pp) *pp = slots[SL_PP];
}

}  // namespace flgpu
