// api_internal.hpp -- pieces of api.cu shared with al.cu (not part of the C-ABI).
#pragma once
#include <cstdint>

#include "../../include/flgpu.h"

namespace flgpu_api {

enum Algo { ALGO_LBFGS = 0, ALGO_CG = 1, ALGO_SD = 2 };

// One optimizer call on a fresh CudaBackend (work space allocated inside, released on return).
int run(int algo, const flgpu_problem *prob, const flgpu_options *opt, double *x, int64_t n, int x_space,
        flgpu_stats *stats);

// Reference-ABI callbacks (f90:33-38) behind the 64-bit device-callback interface.  With callback space HOST the
// library stages x / f' through the pinned buffers xh / gh so unmodified host callbacks keep working.
struct RefAdapter {
    flgpu_ref_f_fn f = nullptr;
    flgpu_ref_fd_fn fd = nullptr;
    flgpu_ref_f_fd_fn f_fd = nullptr;
    int cb_space = FLGPU_SPACE_DEVICE;
    double *xh = nullptr, *gh = nullptr;
    flgpu_fused_fn fused = nullptr;       // registered with flgpu_register_fused for this f
    void *fused_user = nullptr;
    flgpu_update_fn update = nullptr;     // fused accepted-point update of the built-in objectives
    flgpu_direction_fn direction = nullptr;   // K3 with the first trial evaluated inside (built-in objectives)
    flgpu_fused_multi_fn fused_multi = nullptr;   // batched fused evaluation (built-in objectives)
};
void ref_adapter_init(RefAdapter &A, flgpu_ref_f_fn f, flgpu_ref_fd_fn fd, flgpu_ref_f_fd_fn f_fd, int dim,
                      flgpu_problem *prob);
void ref_adapter_free(RefAdapter &A);
void apply_thread_settings(flgpu_options &o);   // observer / FLGPU_NO_FUSED of the calling thread
int x_space_now();
int cb_space_now();
void fill_optional(flgpu_options &o, const int32_t *Strong, const int32_t *Warning, const int *MaxIteration,
                   const double *Precision, const double *MinStepLength, const double *WolfeConst1,
                   const double *WolfeConst2, const double *Increment);
void set_last_stats(const flgpu_stats &st);

}  // namespace flgpu_api
