// driver.hpp -- host control flow of LBFGS / ConjugateGradient and their line searchers.
#pragma once
#include "../../include/flgpu.h"
#include "backend.hpp"

namespace flgpu {

struct Params {
    int mem = 10;        // f90:419-420
    int method = FLGPU_CG_DY;
    bool strong = true;  // f90:421-422
    bool warn = true;    // f90:423-424
    int maxit = 1000;    // f90:425-426
    double tol = 1e-30;  // Precision^2, f90:427-428
    double minstep = 1e-30;
    double c1 = 1e-4, c2 = 0.9, incr = 1.05;
    bool has_f_fd = false;
    bool fused = true;   // use flgpu_problem.fused when the problem supplies it
    int device_search = 2;   // flgpu_problem.search (whole line search in one cooperative kernel): 0 off, 1 on, 2 auto
    int line_search = FLGPU_LS_REFERENCE;   // FLGPU_LS_FAST: accept-at-first-Wolfe-point searcher (not a reference routine)
    flgpu_observer_fn observer = nullptr;
    void *observer_user = nullptr;
};

// Applies the reference defaults and fail-safe clamps (f90:417-434, 212-229, 1478-1479).
Params params_from_options(const flgpu_options &o, bool for_cg, bool has_f_fd);

void run_lbfgs(Backend &B, const Params &P, double *x_user, int x_space, flgpu_stats *st);
void run_cg(Backend &B, const Params &P, double *x_user, int x_space, flgpu_stats *st);
void run_sd(Backend &B, const Params &P, double *x_user, int x_space, flgpu_stats *st);

}  // namespace flgpu

namespace flgpu {

// The two-loop recursion as a standalone operator (reference: LBFGS::Before f90:586-608 and the
// ring-buffer update of After f90:622-623) over the same K1/K2/K3 the optimizer uses.
class History {
public:
    History(Backend &b, int mem) : B(b), mem_(mem > 1 ? mem : 1) { B.lbfgs_alloc(mem_); }
    // append (or overwrite the oldest) pair s = x1-x0, y = g1-g0; refresh the coefficients for g1
    void push(const double *x1, const double *x0, const double *g1, const double *g0);
    // p = -H g1, xt = x1 + p; returns g1.p and p.p
    void direction(const double *g1, const double *x1, double *p, double *xt, double *gp, double *pp);
    int count() const { return k_; }

private:
    Backend &B;
    int mem_, k_ = 0, recent_ = -1;
};

}  // namespace flgpu
