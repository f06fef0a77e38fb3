// flgpu_lbfgs_gram.hpp -- layouts shared by K1/K2/K3 and the scalar statement of K2.
//
// The reference's two-loop recursion (f90:589-607) interleaves 2k+2 full-length dot products
// with 2k full-length axpys, each dot depending on the previous axpy.  Here the recursion is
// carried out on the Gram representation instead: every vector it touches lies in
// span{g, s_1..s_k, y_1..y_k}, so with
//     A_j = s_j.g,  B_j = y_j.g,  SY_ij = s_i.y_j (i not newer than j),  YY_ij = y_i.y_j
// the alpha_i / beta_i recurrences need no further pass over the data (this is the
// Byrd-Nocedal-Schnabel compact form evaluated in the reference's own operation order).
// One fused pass (K1) produces A, B, g.g and the new row/column of SY, YY; K2 (this file's
// recurrences, run by a single warp on the device) yields the 2k+1 coefficients; one fused pass
// (K3) forms p = -( gamma (g - sum alpha_i y_i) + sum (alpha_i - beta_i) s_i ).
#pragma once

#ifdef __CUDACC__
#define FLGPU_HD __host__ __device__
#else
#define FLGPU_HD
#endif

namespace flgpu {

// dots produced by K1 for memory m (indices are ring slots):
//   [0,m) A_j = s_j.g1   [m,2m) B_j = y_j.g1   [2m] g1.g1
//   [2m+1,3m+1) s_j.y_new   [3m+1,4m+1) y_j.y_new
FLGPU_HD inline int nd_of(int m) { return 4 * m + 1; }
FLGPU_HD inline int d_A(int, int j) { return j; }
FLGPU_HD inline int d_B(int m, int j) { return m + j; }
FLGPU_HD inline int d_GG(int m) { return 2 * m; }
FLGPU_HD inline int d_SYN(int m, int j) { return 2 * m + 1 + j; }
FLGPU_HD inline int d_YYN(int m, int j) { return 3 * m + 1 + j; }
// coefficients produced by K2: [0] gamma, [1,1+m) alpha_j, [1+m,1+2m) e_j = alpha_j - beta_j
FLGPU_HD inline int nc_of(int m) { return 2 * m + 1; }
// slot of the t-th newest pair (t = 0 is `recent`), valid for t < k (f90:590-597 visiting order)
FLGPU_HD inline int slot_of_age(int recent, int t, int m) { return (recent - t + m) % m; }

// K1 covers the k-1 older columns in passes; thread shape of the pass that still has `rem` columns to cover:
// MT columns per thread group, NG groups per warp (include/flgpu_k1.cuh).  Shared with the test host simulator, which
// models the kernels' summation order.  Measured on B200 (profiles/r01_k1_shapes.md): shapes with 4 or 8 groups per
// warp (128 / 64-byte column runs) or more than 5 columns per thread (> 128 registers, 1 CTA/SM) run at 2.6-5.6 TB/s;
// <5,2> (256-byte runs, 2 CTAs/SM) sustains 6.3-7.0 TB/s even counting the re-read of x, g per pass.
inline void k1_pass_shape(int rem, int &mt, int &ng) {
    if (rem <= 2)      { mt = 2; ng = 1; }
    else if (rem <= 4) { mt = 4; ng = 1; }
    else if (rem <= 5) { mt = 5; ng = 1; }
    else if (rem <= 8) { mt = 4; ng = 2; }
    else               { mt = 5; ng = 2; }   // 10 columns per pass; m = 30 takes 3 passes
}

// Scalar statement of K2.  D: global dots; SY, YY: persistent m x m row-major Gram blocks in
// ring-slot indexing, updated in place with the newest pair (slot `recent`); C: coefficients.
// sq, yq, al: work arrays of m doubles.  Multiplies and adds are separate roundings.
inline void lbfgs_gram_solve(int m, int k, int recent, const double *D, double *SY, double *YY,
                             double *C, double *sq, double *yq, double *al) {
    const int r = recent;
    for (int t = 0; t < k; t++) {
        const int j = slot_of_age(recent, t, m);
        SY[j * m + r] = D[d_SYN(m, j)];
        YY[j * m + r] = D[d_YYN(m, j)];
        YY[r * m + j] = D[d_YYN(m, j)];
        sq[j] = D[d_A(m, j)];
        yq[j] = D[d_B(m, j)];
    }
    for (int j = 0; j < nc_of(m); j++) C[j] = 0.0;
    // first loop, newest -> oldest (f90:590-597)
    for (int t = 0; t < k; t++) {
        const int i = slot_of_age(recent, t, m);
        const double rho = 1.0 / SY[i * m + i];          // f90:623 rho=1/dot_product(y,s)
        const double alpha = rho * sq[i];                // f90:591
        al[i] = alpha;
        for (int u = 0; u < k; u++) {
            const int j = slot_of_age(recent, u, m);
            if (u > t) sq[j] = sq[j] - alpha * SY[j * m + i];   // s_j.(q - alpha y_i), j older than i
            yq[j] = yq[j] - alpha * YY[j * m + i];
        }
    }
    // scaling (f90:598): p/rho(recent)/dot_product(y_recent,y_recent)
    const double rho_r = 1.0 / SY[r * m + r];
    const double gamma = 1.0 / rho_r / YY[r * m + r];
    for (int u = 0; u < k; u++) {
        const int j = slot_of_age(recent, u, m);
        yq[j] = gamma * yq[j];                           // now y_j.r
    }
    // second loop, oldest -> newest (f90:599-606)
    for (int t = k - 1; t >= 0; t--) {
        const int i = slot_of_age(recent, t, m);
        const double rho = 1.0 / SY[i * m + i];
        const double beta = rho * yq[i];                 // f90:600
        const double e = al[i] - beta;
        C[1 + i] = al[i];
        C[1 + m + i] = e;
        for (int u = 0; u < t; u++) {
            const int j = slot_of_age(recent, u, m);     // j newer than i
            yq[j] = yq[j] + e * SY[i * m + j];
        }
    }
    C[0] = gamma;
}

}  // namespace flgpu
