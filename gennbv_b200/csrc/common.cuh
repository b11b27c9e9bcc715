// common.cuh -- shared helpers of libgennbv_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/gennbv_b200.h"

namespace gnbv {

void set_error(const char* fmt, ...);
void stage_mark(int id, cudaStream_t stream);      // no-op unless gnbv_profile_enable(1)

// Kernel-variant switches (api.cu; environment variables read once per process, see gnbv_kernel_mode in the header).
int conv2_tc_mode();      // GNBV_CONV2_TC
int conv1_mma_mode();     // GNBV_CONV1_MMA
int gemm_mma_mode();      // GNBV_GEMM_MMA

#define GNBV_REQUIRE(cond, ...)                 \
    do {                                        \
        if (!(cond)) {                          \
            ::gnbv::set_error(__VA_ARGS__);     \
            return GNBV_E_ARG;                  \
        }                                       \
    } while (0)

#define GNBV_CUDA_CHECK(expr)                                                              \
    do {                                                                                   \
        cudaError_t e__ = (expr);                                                          \
        if (e__ != cudaSuccess) {                                                          \
            ::gnbv::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__),     \
                              __FILE__, __LINE__);                                         \
            return GNBV_E_CUDA;                                                            \
        }                                                                                  \
    } while (0)

#define GNBV_LAUNCH_CHECK(name)                                                            \
    do {                                                                                   \
        cudaError_t e__ = cudaGetLastError();                                              \
        if (e__ != cudaSuccess) {                                                          \
            ::gnbv::set_error("launch of %s failed: %s", name, cudaGetErrorString(e__));   \
            return GNBV_E_CUDA;                                                            \
        }                                                                                  \
    } while (0)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device) instead of before every launch: the driver
// call costs microseconds on the host, and a minibatch update issues dozens of launches.  Grows monotonically.
int ensure_dyn_smem_impl(const void* func, size_t bytes);
template <typename F>
static inline int ensure_dyn_smem(F* func, size_t bytes) { return ensure_dyn_smem_impl(reinterpret_cast<const void*>(func), bytes); }

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// streaming (read-once) 128-bit load / store: keep the grids out of L1
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
// same, for buffers the kernel also writes (in/out state): no .nc
__device__ __forceinline__ float4 ld_stream4(const float* p) {
    float4 v;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p)
                 : "memory");
    return v;
}
__device__ __forceinline__ void stg_stream4(float* p, float4 v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

}  // namespace gnbv
