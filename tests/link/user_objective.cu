// A user's objective written with include/flgpu_objective.cuh (test fixture; compiled by the GPU test with nvcc):
// which = 0: element-local, index-dependent   f = sum_i w_i (x_i - 1)^4 + (x_i - 1)^2,  w_i = 1 + (i mod 7)
// which = 1: pairwise                          f = sum_j (x_{2j} - 2)^2 + 5 (x_{2j+1} - x_{2j}^2)^2   (+ (x-2)^2 tail)
// Arithmetic uses separate multiplies and adds (no FMA contraction: built with -fmad=false) so that the gradient is
// bit-identical to the NumPy statement of the same formulas in tests/test_gpu.py.
#include "../../include/flgpu_objective.cuh"

struct Weighted {
    static constexpr int WIDTH = 1;
    __device__ void eval(int64_t i, double x, double &f, double &g) const {
        const double w = 1.0 + (double)(i % 7), t = x - 1.0, t2 = t * t;
        f = w * (t2 * t2) + t2;
        g = (4.0 * w) * (t2 * t) + 2.0 * t;
    }
};
struct Pairs {
    static constexpr int WIDTH = 2;
    __device__ void eval2(int64_t, double a, double b, double &f, double &ga, double &gb) const {
        const double t1 = a - 2.0, t2 = b - a * a;
        f = t1 * t1 + (5.0 * t2) * t2;
        ga = 2.0 * t1 - (20.0 * a) * t2;
        gb = 10.0 * t2;
    }
    __device__ void eval_tail(int64_t, double x, double &f, double &g) const {
        const double t = x - 2.0;
        f = t * t;
        g = 2.0 * t;
    }
};

static Weighted g_weighted;
static Pairs g_pairs;

extern "C" int user_problem(int which, flgpu_problem *out) {
    if (which == 0) *out = flgpu_obj::make_problem<Weighted>(&g_weighted);
    else *out = flgpu_obj::make_problem<Pairs>(&g_pairs);
    return 0;
}
