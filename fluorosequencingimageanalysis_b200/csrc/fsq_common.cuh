// Shared helpers for libfsq (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/fsq.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libfsq is written for sm_100a (B200) only"
#endif

namespace fsq {

void set_error(const char* fmt, ...);

#define FSQ_CUDA_CHECK(expr)                                                         \
    do {                                                                             \
        cudaError_t _e = (expr);                                                     \
        if (_e != cudaSuccess) {                                                     \
            fsq::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),   \
                           __FILE__, __LINE__);                                      \
            return FSQ_E_CUDA;                                                       \
        }                                                                            \
    } while (0)

#define FSQ_LAUNCH_CHECK()                                                           \
    do {                                                                             \
        cudaError_t _e = cudaGetLastError();                                         \
        if (_e != cudaSuccess) {                                                     \
            fsq::set_error("kernel launch failed: %s (%s:%d)",                       \
                           cudaGetErrorString(_e), __FILE__, __LINE__);              \
            return FSQ_E_CUDA;                                                       \
        }                                                                            \
    } while (0)

int sm_count();

// ---- sub-warp ("group") collectives: G lanes, G in {8, 32}; mask = the group's lanes ----
template <int G>
__device__ __forceinline__ double group_sum(double v, unsigned mask) {
#pragma unroll
    for (int m = G / 2; m >= 1; m >>= 1) v += __shfl_xor_sync(mask, v, m, G);
    return v;   // xor butterfly: bitwise identical in every lane of the group
}

template <int G>
__device__ __forceinline__ double group_bcast(double v, int src, unsigned mask) {
    return __shfl_sync(mask, v, src, G);
}

template <int G>
__device__ __forceinline__ int group_bcast_i(int v, int src, unsigned mask) {
    return __shfl_sync(mask, v, src, G);
}

// numpy's add.reduce over a contiguous run of n < 128 doubles (pairwise_sum: eight strided accumulators, their
// tree sum, then the remainder in sequence; plain sequence below 8 elements), with no FMA contraction
template <class F>
__device__ __forceinline__ double np_sum(int n, F f) {         // numpy pairwise_sum for n < 128
    if (n < 8) {
        double r = 0.0;
        for (int i = 0; i < n; ++i) r = __dadd_rn(r, f(i));
        return r;
    }
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = f(j);
    int i = 8;
#pragma unroll 4
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], f(i + j));
    }
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                           __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __dadd_rn(res, f(i));
    return res;
}

// pflib.illumina_s_n (pflib.py:261-281) of a size x size integer window, in numpy's own arithmetic so that the value
// is bit-identical to the reference's: (amax(sub) - mean(op)) / std(op) with op = top row, bottom row, then the
// (h, 0), (h, -1) pairs of the middle rows (:278-280); mean(op) is exact (integers), std(op) = sqrt(add.reduce((op -
// mean)^2) / len(op)) with add.reduce in pairwise order (np_sum; len(op) <= 128, i.e. size <= 33).  px(r, c) -> integer
// pixel.  SIZE > 0: compile-time size (everything unrolls), SIZE == 0: run-time `size`.
template <int SIZE, class PX>
__device__ __forceinline__ double illumina_sn(int size_rt, PX px) {
    const int size = SIZE > 0 ? SIZE : size_rt;
    const int ne = 2 * size + 2 * (size > 2 ? size - 2 : 0);
    auto edge = [&](int k) -> long long {
        if (k < size) return px(0, k);
        if (k < 2 * size) return px(size - 1, k - size);
        const int kk = k - 2 * size;
        return px(1 + kk / 2, (kk & 1) ? size - 1 : 0);
    };
    long long es = 0, mx = px(0, 0);
#pragma unroll
    for (int r = 0; r < size; ++r)
#pragma unroll
        for (int c = 0; c < size; ++c) { const long long v = px(r, c); mx = v > mx ? v : mx; }
#pragma unroll
    for (int k = 0; k < ne; ++k) es += edge(k);
    const double mean = __ddiv_rn((double)es, (double)ne);
    const double var_sum = np_sum(ne, [&](int k) { const double d = __dsub_rn((double)edge(k), mean); return __dmul_rn(d, d); });
    return __ddiv_rn(__dsub_rn((double)mx, mean), sqrt(__ddiv_rn(var_sum, (double)ne)));
}

}  // namespace fsq
