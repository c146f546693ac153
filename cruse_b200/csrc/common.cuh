// common.cuh -- shared helpers for libcruse_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdarg>
#include <cstdint>
#include "../../include/cruse_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "cruse_b200 kernels are written for sm_100a (B200) only"
#endif

namespace cruse {

void set_error(const char* fmt, ...);

#define CRUSE_CHECK_ARG(cond, ...)                 \
    do {                                           \
        if (!(cond)) {                             \
            ::cruse::set_error(__VA_ARGS__);       \
            return -1;                             \
        }                                          \
    } while (0)

#define CRUSE_CUDA_OK(expr)                                                                  \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            ::cruse::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),       \
                               __FILE__, __LINE__);                                          \
            return -2;                                                                       \
        }                                                                                    \
    } while (0)

#define CRUSE_LAUNCH_OK()                                                                    \
    do {                                                                                     \
        cudaError_t _e = cudaGetLastError();                                                 \
        if (_e != cudaSuccess) {                                                             \
            ::cruse::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),   \
                               __FILE__, __LINE__);                                          \
            return -3;                                                                       \
        }                                                                                    \
    } while (0)

int sm_count();

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float apply_act(float v, int act, float alpha) {
    switch (act) {
        case CRUSE_ACT_RELU: return v > 0.f ? v : 0.f;
        case CRUSE_ACT_PRELU: return v > 0.f ? v : alpha * v;
        case CRUSE_ACT_SIGMOID: return sigmoidf_(v);
        default: return v;
    }
}

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace cruse
