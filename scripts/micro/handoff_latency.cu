// Micro-benchmark of the hand-off latencies in a TS-form (A in tensor memory) tcgen05 pipeline on one CTA:
//   (1) tcgen05.st of 96 columns (6 x .x16) by a full warp + tcgen05.wait::st
//   (2) 18 tcgen05.mma (M128 N16 K8, A in TMEM) + tcgen05.commit -> mbarrier completion seen by the issuing warp
//   (3) mbarrier arrive by one warp -> try_wait returning true in another warp (ping-pong through shared memory)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../gennbv_b200/csrc/tc.cuh"
using namespace gnbv;

__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void st16(uint32_t taddr, uint32_t v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(v) : "memory");
}

__global__ void __launch_bounds__(128, 1) lat_kernel(long long* out, int reps) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar, ping, pong;
    __shared__ uint32_t tbase;
    const int warp = tc::uniform_warp_index(), lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 32 * 1024 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 1.0f;
    if (threadIdx.x == 0) { tc::mbar_init(tc::smem_u32(&bar), 1); tc::mbar_init(tc::smem_u32(&ping), 1); tc::mbar_init(tc::smem_u32(&pong), 1); tc::fence_mbar_init(); }
    if (warp == 0) tc::tmem_alloc(tc::smem_u32(&tbase), 512);
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tm = __shfl_sync(0xffffffffu, tbase, 0);
    if (warp == 0) {
        // (1) store latency
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            for (int k = 0; k < 6; ++k) st16(tm + k * 16, r);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        long long t1 = clock64();
        // (2) MMA chunk + commit latency
        const uint32_t idesc = tc::make_idesc_tf32(128, 16);
        const uint64_t bdesc = tc::make_smem_desc(tc::smem_u32(smem), 128, 256);
        tc::tc_fence_before();
        __syncwarp();
        long long t2 = clock64();
        for (int r = 0; r < reps; ++r) {
            tc::tc_fence_after();
            if (tc::elect_one()) {
                for (int k = 0; k < 18; ++k) mma_ts(tm + 256, tm + (k % 6) * 8, bdesc, idesc, 1u);
                tc::mma_commit(tc::smem_u32(&bar));
            }
            __syncwarp();
            tc::mbar_wait(tc::smem_u32(&bar), r & 1, 0);
        }
        long long t3 = clock64();
        if (lane == 0 && blockIdx.x == 0) { out[0] = (t1 - t0) / reps; out[1] = (t3 - t2) / reps; }
    }
    // (3) mbarrier ping-pong between warp 1 and warp 2
    if (warp == 1 || warp == 2) {
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            if (warp == 1) {
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(&ping)) : "memory");
                tc::mbar_wait(tc::smem_u32(&pong), r & 1, 0);
            } else {
                tc::mbar_wait(tc::smem_u32(&ping), r & 1, 0);
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(&pong)) : "memory");
            }
        }
        long long t1 = clock64();
        if (warp == 1 && lane == 0 && blockIdx.x == 0) out[2] = (t1 - t0) / reps;
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tbase, 512);
}

int main() {
    long long* d_out;
    cudaMalloc(&d_out, 64);
    cudaFuncSetAttribute(lat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int it = 0; it < 2; ++it) lat_kernel<<<148, 128, 64 * 1024>>>(d_out, 2000);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[3];
    cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("tcgen05.st 96 cols + wait::st        : %lld cycles\n18 MMA (N=16) + commit -> barrier seen : %lld cycles\nmbarrier round trip (2 hand-offs)     : %lld cycles   (%s)\n",
           h[0], h[1], h[2], cudaGetErrorString(e));
    return 0;
}
