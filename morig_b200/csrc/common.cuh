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

// |value| bookkeeping for the fp16-split tensor-core consumers: warp-wide max of the lanes' non-negative `v`, then one
// integer atomic max per warp (bit patterns of non-negative floats order like the floats; exact, order independent).
// Must be reached by all 32 lanes of the warp.
__device__ __forceinline__ void amax_commit(float *amax, float v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, off));
    if (amax && (threadIdx.x & 31) == 0 && v > 0.f) atomicMax(reinterpret_cast<int *>(amax), __float_as_int(v));
}

// the same for a whole thread block (blockDim.x a multiple of 32, at most 1024): ONE atomic per block -- same-address
// atomics serialise, so the memory-bound row kernels must not issue one per warp.  Must be reached by every thread.
__device__ __forceinline__ void amax_commit_block(float *amax, float v) {
    __shared__ float s_am[32];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, off));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    if (lane == 0) s_am[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = lane < nw ? s_am[lane] : 0.f;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, off));
        if (amax && lane == 0 && v > 0.f) atomicMax(reinterpret_cast<int *>(amax), __float_as_int(v));
    }
}

// ---- programmatic dependent launch (PDL) --------------------------------------------------------------------------------
// The forward is ~70 dependent launches in one stream (or one replayed CUDA graph); with the launch attribute below the
// next kernel's CTAs may become resident as soon as every CTA of the current one has executed pdl_trigger() (or exited)
// and SM resources free up, run their prologue (barrier / TMEM set-up, parameter staging) and then block in pdl_wait()
// until the predecessor grid has completed and flushed its memory.  Rule kept by every kernel of this library:
// NO global memory is read or written before pdl_wait() except launch parameters -- so the overlap can never be observed.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

bool pdl_enabled();      // false when MORIG_NO_PDL=1 (A/B switch)

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace morig
