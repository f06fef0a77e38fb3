// Drop-in check in the manner of the reference's test/test.cpp:84-101: host callbacks written the way
// a user of libFL.so writes them (f = sum x^4, f' = 4 x^3, dim = 10, start in [0,1)^10), called
// through the FL::NO wrappers, linked against libflgpu.so instead of libFL.so.
// Host callbacks are the default for the reference-named symbols: no setting is needed.
// "Correct routines should print close to 0" (test.cpp:74): exit status 0 iff every norm < 1e-3.
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "../../include/NonlinearOptimization_flgpu.hpp"

static void f(double & fx, const double * x, const int & dim) {
    fx = 0.0;
    for (int i = 0; i < dim; i++) fx += std::pow(x[i], 4);
}
static void fd(double * fdx, const double * x, const int & dim) {
    for (int i = 0; i < dim; i++) fdx[i] = 4.0 * std::pow(x[i], 3);
}
static int f_fd(double & fx, double * fdx, const double * x, const int & dim) {
    fx = 0.0;
    for (int i = 0; i < dim; i++) { fx += std::pow(x[i], 4); fdx[i] = 4.0 * std::pow(x[i], 3); }
    return 0;
}
// the reference's test constraint (test.cpp / test.f90:692-705): unit sphere
static void con(double * cx, const double * x, const int & M, const int & N) {
    (void)M;
    cx[0] = -1.0;
    for (int i = 0; i < N; i++) cx[0] += x[i] * x[i];
}
static void cond(double * cdx, const double * x, const int & M, const int & N) {
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
    struct { const char * name; int which; } cases[] = {
        {"Dai-Yuan conjugate gradient: basic version", 0}, {"Dai-Yuan conjugate gradient", 1},
        {"Polak-Ribiere+ conjugate gradient: basic version", 2}, {"Polak-Ribiere+ conjugate gradient", 3},
        {"L-BFGS", 4}, {"L-BFGS: fd with f, Memory=5", 5}, {"Steepest descent", 6},
        {"augmented Lagrangian based on LBFGS", 7}, {"augmented Lagrangian based on conjugate gradient", 8}};
    for (auto & c : cases) {
        start(x, dim);
        switch (c.which) {
        case 0: FL::NO::ConjugateGradient(f, fd, x, dim); break;
        case 1: FL::NO::ConjugateGradient(f, fd, f_fd, x, dim); break;
        case 2: FL::NO::ConjugateGradient(f, fd, x, dim, "PR"); break;
        case 3: FL::NO::ConjugateGradient(f, fd, f_fd, x, dim, "PR"); break;
        case 4: FL::NO::LBFGS(f, fd, x, dim); break;
        case 5: FL::NO::LBFGS(f, fd, f_fd, x, dim, 5); break;
        case 6: FL::NO::SteepestDescent(f, fd, f_fd, x, dim); break;
        case 7: FL::NO::AugmentedLagrangian(f, fd, f_fd, nullptr, con, cond, nullptr, x, dim, 1, "LBFGS", {}, 1.0, 20, 10,
                                            "DY", true, false, 100, 1e-8); break;
        case 8: FL::NO::AugmentedLagrangian(f, fd, f_fd, nullptr, con, cond, nullptr, x, dim, 1, "ConjugateGradient", {},
                                            1.0, 20, 10, "DY", true, false, 100, 1e-8); break;
        }
        double r = norm(x, dim);
        if (c.which >= 7) r = std::fabs(r - 1.0);      // constrained: "norm2(x)-1 should print close to 0"
        std::printf("%s\n%.6e\n\n", c.name, r);
        if (!(r < 1e-3)) bad++;
    }
    return bad;
}
