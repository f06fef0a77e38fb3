// backend.hpp -- the set of vector operations the optimizer drivers are written against.
//
// The product has exactly one implementation: CudaBackend (backend_cuda.cu), hand-written
// sm_100a kernels.  tests/hostsim/ holds a second, test-only implementation on host memory so
// that the HOST control flow in driver.cpp (line-search state machines, ring-buffer
// bookkeeping, multi-rank combination) can be checked against the oracle on a machine without
// a GPU.  libflgpu.so never contains or loads the host simulator.
//
// All operations are asynchronous "enqueue" calls in stream order.  Scalar results land in a
// small array of SLOTS holding this rank's partial sums; fetch() combines the slots over the
// ranks of the row-shard communicator in rank order, copies them to the host and synchronises.
#pragma once
#include <cstdint>

namespace flgpu {

enum Slot {
    SL_F = 0,     // objective partial written by the f / f_fd callback
    SL_GP = 1,    // f'(trial) . p                      (dot_product(fdx,p), f90:1485)
    SL_GG = 2,    // f'(x) . f'(x) at the accepted point (f90:611, 354)
    SL_PP = 3,    // p . p                              (f90:615, 358)
    SL_GP0 = 4,   // f'(x) . p_new = phi'(0)            (f90:607, 367)
    SL_DGP = 5,   // (f'new - f'old) . p                (f90:366)
    SL_GDG = 6,   // f'new . (f'new - f'old)            (f90:387)
    SL_G0G0 = 7,  // f'old . f'old                      (f90:387)
    SL_AUX = 8,
    NSLOTS = 16
};

class Backend {
public:
    virtual ~Backend() {}
    int64_t n = 0;          // local rows
    int64_t launches = 0;   // library kernels enqueued
    int64_t callback_launches = 0;   // objective callbacks invoked (f, fd, f_fd, fused, search): one kernel each for the built-ins
    int64_t syncs = 0;      // host synchronisations

    // ---- memory (library-owned work space, f90:413-415, 435, 1476)
    virtual double *vec_alloc() = 0;
    virtual void lbfgs_alloc(int mem) = 0;  // ring buffers S, Y (n x mem each), Gram blocks, coefficients
    virtual void upload(double *dst, const double *user_x, int x_space) = 0;
    virtual void download(double *user_x, const double *src, int x_space) = 0;

    // ---- user callbacks (f90:33-38): f -> SL_F
    virtual void eval_f(const double *x) = 0;
    virtual void eval_g(const double *x, double *g) = 0;
    virtual void eval_fg(const double *x, double *g) = 0;

    // ---- fused line-search evaluation (flgpu_fused_fn): x = x0 + a*p formed inside the objective kernel;
    // f -> SL_F, f'(x).p -> SL_GP, x / f' stored only when asked (flags = FLGPU_WANT_* | FLGPU_WRITE_*)
    virtual bool fused_available() const { return false; }
    virtual void fused_eval(int /*flags*/, double /*a*/, const double * /*x0*/, const double * /*p*/,
                            double * /*x_out*/, double * /*g_out*/) {}

    // ---- batched fused evaluation (flgpu_fused_multi_fn): f and f'.p at x0 + steps[j]*p for j < count <= FLGPU_MULTI_MAX
    // in ONE pass over x0 and p; f_j -> SL_AUX + 2j, (f'.p)_j -> SL_AUX + 2j + 1, each with the bits of fused_eval
    virtual bool fused_multi_available() const { return false; }
    virtual void fused_eval_multi(int /*count*/, const double * /*steps (host)*/, const double * /*x0*/,
                                  const double * /*p*/) {}

    // ---- device-resident line search (flgpu_search_fn): the whole search in one cooperative kernel; the accepted
    // point / gradient land in xt / gt, the scalars in search_result() after the next fetch()
    virtual bool device_search_available() const { return false; }
    virtual void device_search(int /*policy: FLGPU_LS_* */, bool /*strong*/, bool /*fdwithf*/, double /*c1*/,
                               double /*c2abs*/, double /*fx0*/,
                               double /*phid0*/, double /*incr*/, double /*a*/, const double * /*x0*/,
                               const double * /*p*/, double * /*xt*/, double * /*gt*/, bool /*no_store*/ = false) {}
    virtual void search_result(double * /*out: FLGPU_SEARCH_RESULT_DOUBLES*/) {}
    // algorithmic bytes of the last device_search(), known only once its evaluation count is (kernel timing)
    virtual void credit_search_bytes(double /*bytes*/) {}

    // ---- primitives
    virtual void trial_x(double *x, const double *x0, const double *p, double a) = 0;  // x = x0 + a*p
    virtual void dot(const double *a, const double *b, int slot) = 0;
    virtual void neg(double *p, const double *g) = 0;                                  // p = -g

    // ---- L-BFGS (f90:586-624 restructured, DESIGN.md "compact two-loop")
    // K1: s_new = x1-x0, y_new = g1-g0 written to ring slot new_slot; all dots of the k_after valid
    //     columns against g1 and y_new; g1.g1 -> SL_GG.
    virtual void lbfgs_update_dots(const double *x1, const double *x0, const double *g1,
                                   const double *g0, int new_slot, int k_after) = 0;
    // K1 with the accepted point formed inside the kernel (flgpu_problem.update): x1 = x0 + a*p and g1 = f'(x1) are
    // computed in registers and STORED to x1 / g1 together with the new column; same dots as lbfgs_update_dots.
    virtual bool fused_update_available() const { return false; }
    virtual void lbfgs_update_dots_fused(double /*a*/, const double * /*x0*/, const double * /*p*/, const double * /*g0*/,
                                         double * /*x1*/, double * /*g1*/, int /*new_slot*/, int /*k_after*/) {}
    // K2: two-loop recursion carried out on the (2k+1)-dimensional Gram representation.
    virtual void lbfgs_solve(int k, int recent) = 0;
    // K3: p = -H g1 from the coefficients of K2, xt = x1 + p (skipped when xt is null), g1.p -> SL_GP0,
    //     p.p -> SL_PP.
    virtual void lbfgs_direction(double *p, double *xt, const double *g1, const double *x1, int k,
                                 int recent) = 0;

    // K3 with the first trials of the next search evaluated inside the kernel (flgpu_problem.direction): p is written,
    // x1 + steps[j]*p (steps[0] = 1; FLGPU_MULTI_MAX steps) is formed in registers only; f, f'.p at step 0 -> SL_F, SL_GP
    // and at step j >= 1 -> SL_AUX + 2j, SL_AUX + 2j + 1, each with the bits fused_eval(.., steps[j], x1, p) would deliver.
    virtual bool fused_direction_available() const { return false; }
    virtual void lbfgs_direction_probe(double * /*p*/, const double * /*g1*/, const double * /*x1*/, int /*k*/,
                                       int /*recent*/, int /*flags*/, const double * /*steps*/) {}

    // ---- CG (f90:352-393)
    virtual void cg_dots(const double *g1, const double *g0, const double *p) = 0;
    virtual void cg_update(double *p, const double *g1, double beta) = 0;  // p = -g1 + beta*p; g1.p -> SL_GP0

    // ---- host <- device
    virtual void fetch(double *host_slots /*[NSLOTS]*/) = 0;
    virtual void *stream_handle() { return nullptr; }
};

}  // namespace flgpu
