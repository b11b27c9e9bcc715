// gae.cu -- TensorRolloutBuffer_Grid_Obs.compute_returns_and_advantage
// (stable_baselines3/common/buffers.py:706-724) as one launch: one thread per env walks the T steps
// backwards; loads are coalesced across envs ([T,N] layout).  Replaces a Python loop of T x ~8 kernels.
// Compiled with -fmad=false: one rounding per torch op, in torch's evaluation order.
#include "common.cuh"

namespace gnbv {

__global__ void gae_kernel(const float* __restrict__ rewards, const float* __restrict__ values,
                           const uint8_t* __restrict__ episode_starts, const float* __restrict__ last_values,
                           const uint8_t* __restrict__ dones, float gamma, float gamma_lambda, int T, int N,
                           float* __restrict__ advantages, float* __restrict__ returns) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float last = 0.0f;
    float next_value = last_values[n];
    float nnt = 1.0f - (dones[n] ? 1.0f : 0.0f);
    for (int t = T - 1; t >= 0; --t) {
        size_t i = (size_t)t * N + n;
        float v = values[i], r = rewards[i];
        // delta = r + gamma * next_values * next_non_terminal - v       (buffers.py:720)
        float delta = __fsub_rn(__fadd_rn(r, __fmul_rn(__fmul_rn(gamma, next_value), nnt)), v);
        // last = delta + gamma * lambda * next_non_terminal * last        (buffers.py:721)
        last = __fadd_rn(delta, __fmul_rn(__fmul_rn(gamma_lambda, nnt), last));
        advantages[i] = last;
        returns[i] = __fadd_rn(last, v);                                  // buffers.py:724
        next_value = v;
        nnt = 1.0f - (float)episode_starts[i];                            // next_non_terminal of step t-1
    }
}

}  // namespace gnbv

extern "C" int gnbv_gae(const float* rewards, const float* values, const uint8_t* episode_starts,
                        const float* last_values, const uint8_t* dones, double gamma, double gae_lambda, int T, int N,
                        float* advantages, float* returns, void* stream) {
    GNBV_REQUIRE(rewards && values && episode_starts && last_values && dones && advantages && returns,
                 "gnbv_gae: null pointer argument");
    GNBV_REQUIRE(T > 0 && N > 0, "gnbv_gae: T and N must be positive");
    // python: `self.gamma * tensor` rounds the double to fp32 at the multiply; `self.gamma * self.gae_lambda`
    // is a double product rounded once when it meets the tensor
    gnbv::gae_kernel<<<(unsigned)gnbv::ceil_div(N, 128), 128, 0, (cudaStream_t)stream>>>(
        rewards, values, episode_starts, last_values, dones, (float)gamma, (float)(gamma * gae_lambda), T, N, advantages,
        returns);
    GNBV_LAUNCH_CHECK("gae_kernel");
    return GNBV_OK;
}
