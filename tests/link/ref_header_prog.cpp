// A user program written against the UNMODIFIED reference header (/root/reference/cpp/NonlinearOptimization.hpp is
// included by absolute path, never copied): the optimizer calls of the reference's test/test.cpp:84-124 that lie on
// the hot path, with host callbacks exactly as a libFL user writes them.  __graft_entry__.build() compiles it here,
// where the reference tree is mounted, against libflgpu.so; the binary travels to the GPU box and
// tests/test_gpu.py::test_reference_header_program_runs executes it as it is (host callbacks are the default of the reference-named symbols).
// "Correct routines should print close to 0" (test.cpp:74): exit status = number of results that are not.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <tuple>
#include <vector>

#include "/root/reference/cpp/NonlinearOptimization.hpp"

static void f(double & fx, const double * x, const int & dim) {
    fx = 0.0;
    for (int i = 0; i < dim; i++) fx += std::pow(x[i], 4);
}
static void fd(double * fdx, const double * x, const int & dim) {
    for (int i = 0; i < dim; i++) fdx[i] = 4.0 * std::pow(x[i], 3);
}
static int f_fd(double & fx, double * fdx, const double * x, const int & dim) {
    f(fx, x, dim); fd(fdx, x, dim);
    return 0;
}
static void c(double * cx, const double * x, const int & M, const int & N) {
    (void)M;
    cx[0] = -1.0;
    for (int i = 0; i < N; i++) cx[0] += x[i] * x[i];
}
static void cd(double * cdx, const double * x, const int & M, const int & N) {
    (void)M;
    for (int i = 0; i < N; i++) cdx[i] = 2.0 * x[i];
}
static double norm(const double * x, int dim) {
    double s = 0.0;
    for (int i = 0; i < dim; i++) s += x[i] * x[i];
    return std::sqrt(s);
}
static void start(double * x, int dim) {
    for (int i = 0; i < dim; i++) x[i] = (double)std::rand() / (double)RAND_MAX;
}

int main() {
    const int dim = 10;
    double x[dim];
    int bad = 0;
    const char * names[] = {"Steepest descent", "Dai-Yuan conjugate gradient: basic version", "Dai-Yuan conjugate gradient",
                            "Polak-Ribiere+ conjugate gradient", "augmented Lagrangian based on LBFGS",
                            "augmented Lagrangian based on conjugate gradient"};
    for (int which = 0; which < 6; which++) {
        start(x, dim);
        switch (which) {
        case 0: FL::NO::SteepestDescent(f, fd, f_fd, x, dim); break;
        case 1: FL::NO::ConjugateGradient(f, fd, x, dim); break;
        case 2: FL::NO::ConjugateGradient(f, fd, f_fd, x, dim); break;
        case 3: FL::NO::ConjugateGradient(f, fd, f_fd, x, dim, "PR"); break;
        case 4: FL::NO::AugmentedLagrangian(f, fd, f_fd, nullptr, c, cd, nullptr, x, dim, 1, "LBFGS", {}, 1.0, 20, 10, "DY",
                                            true, false, 100, 1e-8); break;
        case 5: FL::NO::AugmentedLagrangian(f, fd, f_fd, nullptr, c, cd, nullptr, x, dim, 1, "ConjugateGradient", {}, 1.0, 20,
                                            10, "DY", true, false, 100, 1e-8); break;
        }
        double r = norm(x, dim);
        if (which >= 4) r = std::fabs(r - 1.0);
        std::printf("%s\n%.6e\n\n", names[which], r);
        if (!(r < 1e-3)) bad++;
    }
    return bad;
}
