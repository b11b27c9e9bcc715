// Micro-benchmark: issue rate of tcgen05.mma.kind::tf32 (cta_group::1, M = 128) as a function of N, with the A operand in
// tensor memory (TS) or in shared memory (SS).  One CTA per SM, one thread issues REPS MMAs back to back, then commits and
// waits; reports cycles per MMA.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../gennbv_b200/csrc/tc.cuh"
using namespace gnbv;

__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

template <int N, bool TS>
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* out, int reps) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tbase;
    const int warp = tc::uniform_warp_index();
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 1.0f;
    if (threadIdx.x == 0) { tc::mbar_init(tc::smem_u32(&bar), 1); tc::fence_mbar_init(); }
    if (warp == 0) tc::tmem_alloc(tc::smem_u32(&tbase), 512);
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    if (warp == 0) {
        const uint32_t tm = __shfl_sync(0xffffffffu, tbase, 0);
        const uint32_t idesc = tc::make_idesc_tf32(128, N);
        const uint64_t bdesc = tc::make_smem_desc(tc::smem_u32(smem), 128, 256);             // B: N rows x 8 k, SBO = 256 B
        const uint64_t adesc = tc::make_smem_desc(tc::smem_u32(smem) + 16384, 128, 256);     // A (SS): 128 rows x 8 k
        long long t0 = 0, t1 = 0;
        if (tc::elect_one()) {
            t0 = clock64();
            for (int r = 0; r < reps; ++r) {
                if (TS) mma_ts(tm + 256, tm + (r & 7) * 8, bdesc, idesc, r ? 1u : 0u);
                else tc::mma_tf32(tm + 256, adesc, bdesc, idesc, r ? 1u : 0u);
            }
            tc::mma_commit(tc::smem_u32(&bar));
        }
        __syncwarp();
        tc::mbar_wait(tc::smem_u32(&bar), 0);
        t1 = clock64();
        if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1;
        long long tt0 = __shfl_sync(0xffffffffu, t0, 0);
        // the elected lane may not be lane 0: reduce max of t0 over the warp
        for (int o = 16; o > 0; o >>= 1) { long long v = __shfl_xor_sync(0xffffffffu, t0, o); t0 = t0 > v ? t0 : v; }
        (void)tt0;
        if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = t1 - t0;
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tbase, 512);
}

template <int N, bool TS>
void run(const char* name, long long* d_out) {
    const int reps = 4096;
    cudaFuncSetAttribute(rate_kernel<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int it = 0; it < 2; ++it) rate_kernel<N, TS><<<148, 128, 64 * 1024>>>(d_out, reps);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2];
    cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-8s N=%3d  %s  cycles/MMA = %.1f   (%s)\n", name, N, TS ? "A in TMEM" : "A in smem", (double)h[1] / reps, cudaGetErrorString(e));
}

int main() {
    long long* d_out;
    cudaMalloc(&d_out, 16);
    run<16, true>("TS", d_out); run<32, true>("TS", d_out); run<64, true>("TS", d_out); run<128, true>("TS", d_out); run<256, true>("TS", d_out);
    run<16, false>("SS", d_out); run<32, false>("SS", d_out); run<64, false>("SS", d_out); run<128, false>("SS", d_out); run<256, false>("SS", d_out);
    return 0;
}
