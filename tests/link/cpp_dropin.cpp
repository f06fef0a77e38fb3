// Drop-in check in the manner of the reference's test/test.cpp:84-101: host callbacks written the way
// a user of libFL.so writes them (f = sum x^4, f' = 4 x^3, dim = 10, start in [0,1)^10), called
// through the FL::NO wrappers, linked against libflgpu.so instead of libFL.so.
// Run with FLGPU_CALLBACK_SPACE=host (the callbacks dereference host pointers).
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
        {"L-BFGS", 4}, {"L-BFGS: fd with f, Memory=5", 5}};
    for (auto & c : cases) {
        start(x, dim);
        switch (c.which) {
        case 0: FL::NO::ConjugateGradient(f, fd, x, dim); break;
        case 1: FL::NO::ConjugateGradient(f, fd, f_fd, x, dim); break;
        case 2: FL::NO::ConjugateGradient(f, fd, x, dim, "PR"); break;
        case 3: FL::NO::ConjugateGradient(f, fd, f_fd, x, dim, "PR"); break;
        case 4: FL::NO::LBFGS(f, fd, x, dim); break;
        case 5: FL::NO::LBFGS(f, fd, f_fd, x, dim, 5); break;
        }
        const double r = norm(x, dim);
        std::printf("%s\n%.6e\n\n", c.name, r);
        if (!(r < 1e-3)) bad++;
    }
    return bad;
}
