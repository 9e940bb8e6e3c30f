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

}  // namespace fsq
