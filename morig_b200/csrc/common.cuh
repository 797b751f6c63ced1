// Shared helpers for the morig_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/morig_b200.h"

namespace morig {

void set_error(const char *fmt, ...);
int  sm_count();

#define MORIG_CHECK_ARG(cond, ...)                                  \
    do {                                                            \
        if (!(cond)) {                                              \
            ::morig::set_error(__VA_ARGS__);                        \
            return MORIG_E_BADARG;                                  \
        }                                                           \
    } while (0)

#define MORIG_CUDA(call)                                                                  \
    do {                                                                                  \
        cudaError_t e__ = (call);                                                         \
        if (e__ != cudaSuccess) {                                                         \
            ::morig::set_error("%s failed: %s", #call, cudaGetErrorString(e__));          \
            return (int)e__;                                                              \
        }                                                                                 \
    } while (0)

#define MORIG_LAUNCH_CHECK(name)                                                          \
    do {                                                                                  \
        cudaError_t e__ = cudaGetLastError();                                             \
        if (e__ != cudaSuccess) {                                                         \
            ::morig::set_error("launch of %s failed: %s", name, cudaGetErrorString(e__)); \
            return (int)e__;                                                              \
        }                                                                                 \
    } while (0)

// Order-independent (hence deterministic) float max through integer atomics.
// Destination must have been initialised to -inf (or any value written by this function).
__device__ __forceinline__ void atomic_max_f32(float *addr, float v) {
    v += 0.0f;  // canonicalise -0.0 -> +0.0 so the sign test below is consistent
    if (v >= 0.0f)
        atomicMax(reinterpret_cast<int *>(addr), __float_as_int(v));
    else
        atomicMin(reinterpret_cast<unsigned int *>(addr), __float_as_uint(v));
}

__device__ __forceinline__ float neg_inf() { return __int_as_float(0xff800000); }

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace morig
