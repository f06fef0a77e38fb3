// flgpu_reduce_geom.h -- geometry of the partition-independent reduction (see flgpu_reduce.cuh): chunk size, chunk
// count and the rank tree.  Plain C++ (no CUDA headers) so that host-only code -- the test host simulator, language
// bindings -- shares the exact definitions the kernels use.
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define FLGPU_RG_HD __host__ __device__
#else
#define FLGPU_RG_HD
#endif

namespace flgpu {
namespace red {

constexpr int kThreads = 256;          // threads of every reducing kernel
constexpr int kWarps = kThreads / 32;
constexpr int kBlockChunks = 4096;     // chunk sums combined by one thread block of the tree kernel (16 per thread)
constexpr int kTopMax = 64;            // block values combined by one warp
constexpr int kMaxRanks = 16;

// Chunk size in elements: a power of two chosen from the GLOBAL dimension only (1024 up to 2^22 elements, 8192 from
// 2^25 on; larger only when needed to keep a 2^31+ vector within kTopMax * kBlockChunks chunks).
FLGPU_RG_HD inline int64_t chunk_elems(int64_t n_global) {
    int64_t c = 1024;
    const int64_t t = n_global >> 12;
    while (c < 8192 && (c << 1) <= t) c <<= 1;
    while ((n_global + c - 1) / c > (int64_t)kTopMax * kBlockChunks) c <<= 1;
    return c;
}
// number of chunks of a shard of n_local elements (an empty shard still has one, empty, chunk: its sum is 0)
FLGPU_RG_HD inline int64_t num_chunks(int64_t n_local, int64_t ch) {
    return n_local > 0 ? (n_local + ch - 1) / ch : 1;
}

// Aligned binary tree over the (at most kMaxRanks) per-rank values v[r * stride], missing ranks = +0.0.
FLGPU_RG_HD inline double rank_tree(const double *v, int G, int64_t stride = 1) {
    double a[kMaxRanks];
    for (int r = 0; r < kMaxRanks; r++) a[r] = r < G ? v[(int64_t)r * stride] : 0.0;
    for (int s = 1; s < kMaxRanks; s <<= 1)
        for (int j = 0; j + s < kMaxRanks; j += 2 * s) a[j] = a[j] + a[j + s];
    return a[0];
}

}  // namespace red
}  // namespace flgpu
