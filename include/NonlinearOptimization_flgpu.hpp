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

#include <cstdint>
#include <string>

namespace FL { namespace NO {

typedef void (*f_t)(double &, const double *, const int &);
typedef void (*fd_t)(double *, const double *, const int &);
typedef int (*f_fd_t)(double &, double *, const double *, const int &);

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

} }
#endif
