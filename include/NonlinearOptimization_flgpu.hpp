// NonlinearOptimization_flgpu.hpp -- C++ interface to libflgpu.so in the style of the reference's
// cpp/NonlinearOptimization.hpp (namespace FL::NO, default-argument wrappers over the compiled
// Fortran symbol names; reference hpp:292-324 declarations, hpp:414-454 wrappers).  A program written
// against the reference header for ConjugateGradient links unchanged against libflgpu.so; LBFGS,
// which the reference header does not declare, is added in the same style.
//
// Logical arguments travel as 4-byte integers (-1/0 as the reference wrappers send, hpp:423-425;
// any non-zero is true), CHARACTER(*) Method as char* plus a trailing hidden length.
#ifndef NonlinearOptimization_flgpu_hpp
#define NonlinearOptimization_flgpu_hpp

#include <algorithm>
#include <cstdint>
#include <string>
#include <vector>

namespace FL { namespace NO {

typedef void (*f_t)(double &, const double *, const int &);
typedef void (*fd_t)(double *, const double *, const int &);
typedef int (*f_fd_t)(double &, double *, const double *, const int &);
typedef void (*c_t)(double *, const double *, const int &, const int &);    // c(cx, x, M, N)
typedef void (*cd_t)(double *, const double *, const int &, const int &);   // cd(cdx(N,M), x, M, N)
typedef int (*fdd_t)(double *, const double *, const int &);
typedef int (*cdd_t)(double *, const double *, const int &, const int &);

#ifdef __INTEL_COMPILER
#define FLGPU_SYM(lower) nonlinearoptimization_mp_##lower##_
#else
#define FLGPU_SYM(lower) __nonlinearoptimization_MOD_##lower
#endif

extern "C" {
    void FLGPU_SYM(steepestdescent)(
        f_t f, fd_t fd, double * x, const int & dim, f_fd_t f_fd,
        const int32_t & Strong, const int32_t & Warning, const int & MaxIteration,
        const double & Precision, const double & MinStepLength,
        const double & WolfeConst1, const double & WolfeConst2, const double & Increment);
    void FLGPU_SYM(conjugategradient_basic)(
        f_t f, fd_t fd, double * x, const int & dim, const char * Method,
        const int32_t & Strong, const int32_t & Warning, const int & MaxIteration,
        const double & Precision, const double & MinStepLength,
        const double & WolfeConst1, const double & WolfeConst2, const double & Increment,
        int len_Method);
    void FLGPU_SYM(conjugategradient)(
        f_t f, fd_t fd, double * x, const int & dim, const char * Method, f_fd_t f_fd,
        const int32_t & Strong, const int32_t & Warning, const int & MaxIteration,
        const double & Precision, const double & MinStepLength,
        const double & WolfeConst1, const double & WolfeConst2, const double & Increment,
        int len_Method);
    void FLGPU_SYM(augmentedlagrangian)(
        f_t f, fd_t fd, c_t c, cd_t cd, double * x, const int & N, const int & M,
        const char * UnconstrainedSolver, const double * lambda0, const double & miu0,
        fdd_t fdd, cdd_t cdd, const int & ExactStep, const int & Memory, const char * Method, f_fd_t f_fd,
        const int32_t & Strong, const int32_t & Warning, const int & MaxIteration,
        const double & Precision, const double & MinStepLength,
        const double & WolfeConst1, const double & WolfeConst2, const double & Increment,
        int len_UnconstrainedSolver, int len_Method);
    void FLGPU_SYM(lbfgs)(
        f_t f, fd_t fd, double * x, const int & dim, const int & Memory, f_fd_t f_fd,
        const int32_t & Strong, const int32_t & Warning, const int & MaxIteration,
        const double & Precision, const double & MinStepLength,
        const double & WolfeConst1, const double & WolfeConst2, const double & Increment);
}

// SteepestDescent (f90:55), as reference hpp:395-411
inline void SteepestDescent(f_t f, fd_t fd, f_fd_t f_fd, double * x, const int & dim,
    const bool & Strong = true, const bool & Warning = true,
    const int & MaxIteration = 1000, const double & Precision = 1e-15, const double & MinStepLength = 1e-15,
    const double & WolfeConst1 = 1e-4, const double & WolfeConst2 = 0.9, const double & Increment = 1.05) {
    const int32_t s = Strong ? -1 : 0, w = Warning ? -1 : 0;
    FLGPU_SYM(steepestdescent)(f, fd, x, dim, f_fd, s, w, MaxIteration, Precision, MinStepLength,
        WolfeConst1, WolfeConst2, Increment);
}

// f and f' evaluated separately -> ConjugateGradient_basic (f90:2249), as reference hpp:414-432
inline void ConjugateGradient(f_t f, fd_t fd, double * x, const int & dim,
    const std::string & Method = "DY", const bool & Strong = true, const bool & Warning = true,
    const int & MaxIteration = 1000, const double & Precision = 1e-15, const double & MinStepLength = 1e-15,
    const double & WolfeConst1 = 1e-4, const double & WolfeConst2 = 0.45, const double & Increment = 1.05) {
    const int32_t s = Strong ? -1 : 0, w = Warning ? -1 : 0;
    FLGPU_SYM(conjugategradient_basic)(f, fd, x, dim, Method.c_str(), s, w, MaxIteration, Precision,
        MinStepLength, WolfeConst1, WolfeConst2, Increment, (int)Method.size());
}

// f_fd available -> ConjugateGradient (f90:193) with every optional present, as reference hpp:434-454
inline void ConjugateGradient(f_t f, fd_t fd, f_fd_t f_fd, double * x, const int & dim,
    const std::string & Method = "DY", const bool & Strong = true, const bool & Warning = true,
    const int & MaxIteration = 1000, const double & Precision = 1e-15, const double & MinStepLength = 1e-15,
    const double & WolfeConst1 = 1e-4, const double & WolfeConst2 = 0.45, const double & Increment = 1.05) {
    const int32_t s = Strong ? -1 : 0, w = Warning ? -1 : 0;
    FLGPU_SYM(conjugategradient)(f, fd, x, dim, Method.c_str(), f_fd, s, w, MaxIteration, Precision,
        MinStepLength, WolfeConst1, WolfeConst2, Increment, (int)Method.size());
}

// LBFGS (f90:398); f_fd may be nullptr = absent optional (f90:406, 512)
inline void LBFGS(f_t f, fd_t fd, f_fd_t f_fd, double * x, const int & dim, const int & Memory = 10,
    const bool & Strong = true, const bool & Warning = true,
    const int & MaxIteration = 1000, const double & Precision = 1e-15, const double & MinStepLength = 1e-15,
    const double & WolfeConst1 = 1e-4, const double & WolfeConst2 = 0.9, const double & Increment = 1.05) {
    const int32_t s = Strong ? -1 : 0, w = Warning ? -1 : 0;
    FLGPU_SYM(lbfgs)(f, fd, x, dim, Memory, f_fd, s, w, MaxIteration, Precision, MinStepLength,
        WolfeConst1, WolfeConst2, Increment);
}
inline void LBFGS(f_t f, fd_t fd, double * x, const int & dim, const int & Memory = 10,
    const bool & Strong = true, const bool & Warning = true,
    const int & MaxIteration = 1000, const double & Precision = 1e-15, const double & MinStepLength = 1e-15,
    const double & WolfeConst1 = 1e-4, const double & WolfeConst2 = 0.9, const double & Increment = 1.05) {
    LBFGS(f, fd, nullptr, x, dim, Memory, Strong, Warning, MaxIteration, Precision, MinStepLength,
          WolfeConst1, WolfeConst2, Increment);
}

// AugmentedLagrangian (f90:2005), as reference hpp:514-545.  libflgpu serves UnconstrainedSolver = "LBFGS" and
// "ConjugateGradient" (the default here is "LBFGS": the reference's default "BFGS" is a dense-Hessian solver outside
// the GPU path); fdd / cdd are accepted for signature compatibility and ignored.
inline void AugmentedLagrangian(f_t f, fd_t fd, f_fd_t f_fd, fdd_t fdd, c_t c, cd_t cd, cdd_t cdd,
    double * x, const int & N, const int & M,
    const std::string & UnconstrainedSolver = "LBFGS", std::vector<double> lambda0 = {}, const double & miu0 = 1.0,
    const int & ExactStep = 20, const int & Memory = 10, const std::string & Method = "DY",
    const bool & Strong = true, const bool & Warning = true,
    const int & MaxIteration = 1000, const double & Precision = 1e-15, const double & MinStepLength = 1e-15,
    const double & WolfeConst1 = 1e-4, double WolfeConst2 = 0.9, const double & Increment = 1.05) {
    if (lambda0.empty()) lambda0.assign((size_t)M, 0.0);
    const int32_t s = Strong ? -1 : 0, w = Warning ? -1 : 0;
    if (UnconstrainedSolver == "ConjugateGradient" && WolfeConst2 == 0.9) WolfeConst2 = 0.45;   // hpp:536
    FLGPU_SYM(augmentedlagrangian)(f, fd, c, cd, x, N, M, UnconstrainedSolver.c_str(), lambda0.data(), miu0,
        fdd, cdd, ExactStep, Memory, Method.c_str(), f_fd, s, w, MaxIteration, Precision, MinStepLength,
        WolfeConst1, WolfeConst2, Increment, (int)UnconstrainedSolver.size(), (int)Method.size());
}

// Line-search policy of later calls on this thread (an extension: the reference signatures have no room for it).
// false (default) = the reference's Wolfe / StrongWolfe searchers, trial for trial; true = FLGPU_LS_FAST, which accepts
// the first trial satisfying the Wolfe conditions (include/flgpu.h; not a reference routine, Increment unused).
extern "C" void flgpu_set_line_search(int policy);
inline void FastLineSearch(const bool & on) { flgpu_set_line_search(on ? 1 : 0); }

} }
#endif
