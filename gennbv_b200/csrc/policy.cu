// policy.cu -- actor / critic heads, MultiCategorical distribution, PPO loss, gradient clipping + Adam.
//
//   heads               ActorCriticPolicy_Train_Eval: action_net Linear(256,240), value_net Linear(256,1)
//                       (stable_baselines3/common/policies.py:984,994,1007-1011), stored adjacently as [A+1,256]
//   multicategorical    MultiCategoricalDistribution.log_prob / entropy / sample / mode
//                       (stable_baselines3/common/distributions.py:299-352)
//   ppo loss            PPO_Grid_Obs.train, the loss lines (stable_baselines3/ppo/ppo_grid_obs.py:213-262)
//   clip + Adam         clip_grad_norm_ + th.optim.Adam(eps=1e-5) step (ppo_grid_obs.py:271-275; policies.py:855)
#include "gemm.cuh"

#include <curand_kernel.h>
#include <float.h>
#include <math.h>

#include <algorithm>

namespace gnbv {

constexpr int MAX_SUBSPACES = 16;
struct SubSpaces {
    int n;
    int off[MAX_SUBSPACES + 1];
};

// one warp per row
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// log_prob = sum_k (l[a_k] - lse_k); entropy = sum_k -(sum_j p_j * logp_j)   (distributions.py:330-337)
__global__ void multicat_eval_kernel(const float* __restrict__ logits, int64_t ld, const int64_t* __restrict__ actions,
                                     SubSpaces sp, float* __restrict__ log_prob, float* __restrict__ entropy, int B) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= B) return;
    const float* l = logits + (int64_t)row * ld;
    float lp = 0.f, ent = 0.f;
    for (int k = 0; k < sp.n; ++k) {
        const int o = sp.off[k], n = sp.off[k + 1] - o;
        float m = -FLT_MAX;
        for (int j = lane; j < n; j += 32) m = fmaxf(m, l[o + j]);
        m = warp_max(m);
        float s = 0.f;
        for (int j = lane; j < n; j += 32) s += expf(l[o + j] - m);
        s = warp_sum(s);
        const float lse = m + logf(s);
        float h = 0.f;
        for (int j = lane; j < n; j += 32) {
            float lq = l[o + j] - lse;
            h -= expf(lq) * fmaxf(lq, -FLT_MAX);
        }
        ent += warp_sum(h);
        if (actions) {
            int a = (int)actions[(int64_t)row * sp.n + k];
            a = min(max(a, 0), n - 1);
            lp += l[o + a] - lse;
        }
    }
    if (lane == 0) {
        if (log_prob) log_prob[row] = lp;
        if (entropy) entropy[row] = ent;
    }
}

// sample (Gumbel-max with Philox) or mode (argmax); also returns the log-prob of the drawn action
__global__ void multicat_sample_kernel(const float* __restrict__ logits, int64_t ld, SubSpaces sp, uint64_t seed,
                                       uint64_t offset, int deterministic, int64_t* __restrict__ actions,
                                       float* __restrict__ log_prob, int B) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= B) return;
    const float* l = logits + (int64_t)row * ld;
    curandStatePhilox4_32_10_t st;
    if (!deterministic) curand_init(seed, (uint64_t)row * 32 + lane, offset, &st);
    float lp = 0.f;
    for (int k = 0; k < sp.n; ++k) {
        const int o = sp.off[k], n = sp.off[k + 1] - o;
        float best = -FLT_MAX, m = -FLT_MAX;
        int besti = 0x7fffffff;
        for (int j = lane; j < n; j += 32) {
            float v = l[o + j];
            m = fmaxf(m, v);
            if (!deterministic) {
                float u = curand_uniform(&st);                  // (0,1]
                v += -logf(-logf(fminf(u, 0.99999994f)));
            }
            if (v > best) { best = v; besti = j; }
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            float ob = __shfl_xor_sync(0xffffffffu, best, s);
            int oi = __shfl_xor_sync(0xffffffffu, besti, s);
            if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
        }
        m = warp_max(m);
        float ssum = 0.f;
        for (int j = lane; j < n; j += 32) ssum += expf(l[o + j] - m);
        ssum = warp_sum(ssum);
        if (besti == 0x7fffffff) besti = 0;
        lp += l[o + besti] - (m + logf(ssum));
        if (lane == 0) actions[(int64_t)row * sp.n + k] = besti;
    }
    if (lane == 0 && log_prob) log_prob[row] = lp;
}

// d logits from per-row upstream gradients g_lp (w.r.t. log_prob) and g_ent (w.r.t. entropy):
//   d logp / d l_j = [j == a] - p_j ;  d H / d l_j = -p_j (logp_j + H)
__global__ void multicat_backward_kernel(const float* __restrict__ logits, int64_t ld, const int64_t* __restrict__ actions,
                                         SubSpaces sp, const float* __restrict__ g_lp, const float* __restrict__ g_ent,
                                         float* __restrict__ dlogits, int64_t ldd, int B) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= B) return;
    const float* l = logits + (int64_t)row * ld;
    float* d = dlogits + (int64_t)row * ldd;
    const float glp = g_lp[row], ge = g_ent[row];
    for (int k = 0; k < sp.n; ++k) {
        const int o = sp.off[k], n = sp.off[k + 1] - o;
        float m = -FLT_MAX;
        for (int j = lane; j < n; j += 32) m = fmaxf(m, l[o + j]);
        m = warp_max(m);
        float s = 0.f;
        for (int j = lane; j < n; j += 32) s += expf(l[o + j] - m);
        s = warp_sum(s);
        const float lse = m + logf(s);
        float h = 0.f;
        for (int j = lane; j < n; j += 32) { float lq = l[o + j] - lse; h -= expf(lq) * lq; }
        h = warp_sum(h);
        int a = (int)actions[(int64_t)row * sp.n + k];
        a = min(max(a, 0), n - 1);
        for (int j = lane; j < n; j += 32) {
            float lq = l[o + j] - lse, pj = expf(lq);
            d[o + j] = glp * ((j == a ? 1.f : 0.f) - pj) - ge * pj * (lq + h);
        }
    }
}

// ------------------------------------------------------------------------------------------------ PPO loss
struct PpoHyper {
    float clip_range, clip_range_vf, ent_coef, vf_coef, pg_coef;
    int use_clip_vf, normalize_adv;
};
// scalars out (float[8]): loss, policy_loss, value_loss, entropy_loss, approx_kl, clip_fraction, adv_mean, adv_std
template <int NT>
__device__ __forceinline__ float block_sum(float v, float* sh) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) t += sh[w];        // same order in every thread: deterministic
    return t;
}

__global__ void __launch_bounds__(1024)
ppo_loss_kernel(const float* __restrict__ log_prob, const float* __restrict__ entropy, const float* __restrict__ values,
                const float* __restrict__ old_values, const float* __restrict__ old_log_prob,
                const float* __restrict__ advantages, const float* __restrict__ returns, PpoHyper hp, int B,
                float* __restrict__ scalars, float* __restrict__ g_lp, float* __restrict__ g_ent, float* __restrict__ g_v) {
    __shared__ float sh[32];
    const int tid = threadIdx.x;
    // advantage normalisation over the minibatch: (a - mean) / (std_unbiased + 1e-8)   (ppo_grid_obs.py:215-216)
    float s = 0.f;
    for (int i = tid; i < B; i += 1024) s += advantages[i];
    const float mean = block_sum<1024>(s, sh) / (float)B;
    float q = 0.f;
    for (int i = tid; i < B; i += 1024) { float d = advantages[i] - mean; q += d * d; }
    const float var = block_sum<1024>(q, sh) / (float)(B > 1 ? B - 1 : 1);
    const float stdv = sqrtf(var);
    float pl = 0.f, vl = 0.f, el = 0.f, kl = 0.f, cf = 0.f;
    const float invB = 1.0f / (float)B;
    for (int i = tid; i < B; i += 1024) {
        float adv = advantages[i];
        if (hp.normalize_adv) adv = (adv - mean) / (stdv + 1e-8f);
        const float log_ratio = log_prob[i] - old_log_prob[i];
        const float ratio = expf(log_ratio);
        const float lo = 1.f - hp.clip_range, hi = 1.f + hp.clip_range;
        const float rc = fminf(fmaxf(ratio, lo), hi);
        const float p1 = adv * ratio, p2 = adv * rc;
        pl += fminf(p1, p2);
        cf += (fabsf(ratio - 1.f) > hp.clip_range) ? 1.f : 0.f;
        kl += (ratio - 1.f) - log_ratio;
        // d(-mean(min(p1,p2)))/d logp ; torch.min splits the gradient evenly on ties, clamp passes it on [lo, hi]
        const bool in_range = ratio >= lo && ratio <= hi;
        const float d1 = adv * ratio, d2 = in_range ? adv * ratio : 0.f;
        const float dmin = p1 < p2 ? d1 : (p1 > p2 ? d2 : 0.5f * (d1 + d2));
        g_lp[i] = -hp.pg_coef * invB * dmin;
        // value loss (ppo_grid_obs.py:231-242)
        const float v = values[i], ov = old_values[i];
        float vp = v, dvp = 1.f;
        if (hp.use_clip_vf) {
            const float dv = v - ov;
            vp = ov + fminf(fmaxf(dv, -hp.clip_range_vf), hp.clip_range_vf);
            dvp = (dv >= -hp.clip_range_vf && dv <= hp.clip_range_vf) ? 1.f : 0.f;
        }
        const float err = vp - returns[i];
        vl += err * err;
        g_v[i] = hp.vf_coef * 2.f * invB * err * dvp;
        el += entropy[i];
        g_ent[i] = -hp.ent_coef * invB;
    }
    pl = block_sum<1024>(pl, sh); vl = block_sum<1024>(vl, sh); el = block_sum<1024>(el, sh);
    kl = block_sum<1024>(kl, sh); cf = block_sum<1024>(cf, sh);
    if (tid == 0) {
        const float policy_loss = -pl * invB, value_loss = vl * invB, entropy_loss = -el * invB;
        scalars[0] = policy_loss * hp.pg_coef + hp.ent_coef * entropy_loss + hp.vf_coef * value_loss;   // :253
        scalars[1] = policy_loss; scalars[2] = value_loss; scalars[3] = entropy_loss;
        scalars[4] = kl * invB; scalars[5] = cf * invB; scalars[6] = mean; scalars[7] = stdv;
    }
}

// ------------------------------------------------------------------------------------------------ clip + Adam
// sum of squares of the flat gradient in a fixed order: per-block partials, then one block.
__global__ void __launch_bounds__(256)
sumsq_partial_kernel(const float* __restrict__ g, int64_t n, float* __restrict__ partial) {
    __shared__ float sh[8];
    float s = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) s = fmaf(g[i], g[i], s);
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += sh[w];
        partial[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(256)
sumsq_final_kernel(const float* __restrict__ partial, int nb, float max_norm, float* __restrict__ out /*[2]: norm, coef*/) {
    __shared__ double sh[8];
    double s = 0.0;
    for (int i = threadIdx.x; i < nb; i += 256) s += (double)partial[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += sh[w];
        float norm = (float)sqrt(t);
        out[0] = norm;
        out[1] = fminf(max_norm / (norm + 1e-6f), 1.0f);       // clip_grad_norm_: clamp(max_norm / (norm + 1e-6), max = 1)
    }
}

// th.optim.Adam (no amsgrad, no weight decay): m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
// p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps), with g pre-multiplied by the clip coefficient.
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
            const float* __restrict__ clip /*[2]*/, float lr, float b1, float b2, float eps, float bc1, float bc2_sqrt,
            float grad_scale) {
    const float coef = (clip ? clip[1] : 1.0f) * grad_scale;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const float gi = g[i] * coef;
        const float mi = b1 * m[i] + (1.f - b1) * gi;
        const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi; v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] -= (lr / bc1) * (mi / denom);
    }
}

static int make_subspaces(const int* nvec, int nsub, SubSpaces& sp, int& total) {
    if (!nvec || nsub <= 0 || nsub > MAX_SUBSPACES) { set_error("multicategorical: 1..%d sub-spaces supported", MAX_SUBSPACES); return GNBV_E_ARG; }
    sp.n = nsub; sp.off[0] = 0;
    for (int k = 0; k < nsub; ++k) {
        if (nvec[k] <= 0) { set_error("multicategorical: nvec[%d] = %d", k, nvec[k]); return GNBV_E_ARG; }
        sp.off[k + 1] = sp.off[k] + nvec[k];
    }
    total = sp.off[nsub];
    return GNBV_OK;
}

}  // namespace gnbv

using namespace gnbv;

extern "C" int gnbv_policy_heads_forward(const float* features, const float* head_w, const float* head_b, float* out,
                                         int batch, int feat_dim, int num_out, void* stream) {
    GNBV_REQUIRE(features && head_w && head_b && out && batch > 0 && feat_dim > 0 && num_out > 0,
                 "gnbv_policy_heads_forward: bad arguments");
    GemmEpilogue ep;
    ep.bias = head_b;
    return launch_gemm(features, feat_dim, 1, head_w, 1, feat_dim, out, num_out, batch, num_out, feat_dim, ep, nullptr,
                       (cudaStream_t)stream);
}

extern "C" int gnbv_multicategorical_evaluate(const float* logits, int64_t logits_row_stride, const int* nvec, int num_sub,
                                              const int64_t* actions, float* log_prob, float* entropy, int batch,
                                              void* stream) {
    GNBV_REQUIRE(logits && batch > 0 && (log_prob || entropy), "gnbv_multicategorical_evaluate: bad arguments");
    GNBV_REQUIRE(!log_prob || actions, "gnbv_multicategorical_evaluate: log_prob needs actions");
    SubSpaces sp; int total;
    int rc = make_subspaces(nvec, num_sub, sp, total);
    if (rc) return rc;
    GNBV_REQUIRE(logits_row_stride >= total, "gnbv_multicategorical_evaluate: row stride < sum(nvec)");
    multicat_eval_kernel<<<(unsigned)ceil_div(batch, 8), 256, 0, (cudaStream_t)stream>>>(logits, logits_row_stride, actions, sp,
                                                                                         log_prob, entropy, batch);
    GNBV_LAUNCH_CHECK("multicat_eval_kernel");
    return GNBV_OK;
}

extern "C" int gnbv_multicategorical_sample(const float* logits, int64_t logits_row_stride, const int* nvec, int num_sub,
                                            uint64_t seed, uint64_t offset, int deterministic, int64_t* actions,
                                            float* log_prob, int batch, void* stream) {
    GNBV_REQUIRE(logits && actions && batch > 0, "gnbv_multicategorical_sample: bad arguments");
    SubSpaces sp; int total;
    int rc = make_subspaces(nvec, num_sub, sp, total);
    if (rc) return rc;
    GNBV_REQUIRE(logits_row_stride >= total, "gnbv_multicategorical_sample: row stride < sum(nvec)");
    multicat_sample_kernel<<<(unsigned)ceil_div(batch, 8), 256, 0, (cudaStream_t)stream>>>(logits, logits_row_stride, sp, seed,
                                                                                           offset, deterministic, actions,
                                                                                           log_prob, batch);
    GNBV_LAUNCH_CHECK("multicat_sample_kernel");
    return GNBV_OK;
}

extern "C" int gnbv_multicategorical_backward(const float* logits, int64_t logits_row_stride, const int* nvec, int num_sub,
                                              const int64_t* actions, const float* grad_log_prob, const float* grad_entropy,
                                              float* dlogits, int64_t dlogits_row_stride, int batch, void* stream) {
    GNBV_REQUIRE(logits && actions && grad_log_prob && grad_entropy && dlogits && batch > 0,
                 "gnbv_multicategorical_backward: bad arguments");
    SubSpaces sp; int total;
    int rc = make_subspaces(nvec, num_sub, sp, total);
    if (rc) return rc;
    GNBV_REQUIRE(logits_row_stride >= total && dlogits_row_stride >= total, "gnbv_multicategorical_backward: row stride < sum(nvec)");
    multicat_backward_kernel<<<(unsigned)ceil_div(batch, 8), 256, 0, (cudaStream_t)stream>>>(
        logits, logits_row_stride, actions, sp, grad_log_prob, grad_entropy, dlogits, dlogits_row_stride, batch);
    GNBV_LAUNCH_CHECK("multicat_backward_kernel");
    return GNBV_OK;
}

extern "C" int gnbv_ppo_loss(const float* log_prob, const float* entropy, const float* values, const float* old_values,
                             const float* old_log_prob, const float* advantages, const float* returns, int batch,
                             double clip_range, double clip_range_vf, double ent_coef, double vf_coef, double pg_coef,
                             int normalize_advantage, float* scalars, float* grad_log_prob, float* grad_entropy,
                             float* grad_values, void* stream) {
    GNBV_REQUIRE(log_prob && entropy && values && old_values && old_log_prob && advantages && returns && scalars &&
                     grad_log_prob && grad_entropy && grad_values && batch > 0,
                 "gnbv_ppo_loss: bad arguments");
    PpoHyper hp;
    hp.clip_range = (float)clip_range; hp.use_clip_vf = clip_range_vf >= 0; hp.clip_range_vf = (float)clip_range_vf;
    hp.ent_coef = (float)ent_coef; hp.vf_coef = (float)vf_coef; hp.pg_coef = (float)pg_coef;
    hp.normalize_adv = normalize_advantage;
    ppo_loss_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(log_prob, entropy, values, old_values, old_log_prob, advantages,
                                                          returns, hp, batch, scalars, grad_log_prob, grad_entropy, grad_values);
    GNBV_LAUNCH_CHECK("ppo_loss_kernel");
    return GNBV_OK;
}

constexpr int SUMSQ_BLOCKS = 592;      // 4 x 148

extern "C" size_t gnbv_clip_adam_workspace_bytes(void) { return (SUMSQ_BLOCKS + 2) * sizeof(float); }

extern "C" int gnbv_grad_norm(const float* grads, int64_t n, double max_norm, float* workspace, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GNBV_REQUIRE(grads && workspace && n > 0, "gnbv_grad_norm: bad arguments");
    sumsq_partial_kernel<<<SUMSQ_BLOCKS, 256, 0, stream>>>(grads, n, workspace + 2);
    GNBV_LAUNCH_CHECK("sumsq_partial_kernel");
    sumsq_final_kernel<<<1, 256, 0, stream>>>(workspace + 2, SUMSQ_BLOCKS, (float)max_norm, workspace);
    GNBV_LAUNCH_CHECK("sumsq_final_kernel");
    return GNBV_OK;
}

extern "C" int gnbv_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                              const float* clip_workspace, double lr, double beta1, double beta2, double eps, int64_t step,
                              double grad_scale, void* stream) {
    GNBV_REQUIRE(params && grads && exp_avg && exp_avg_sq && n > 0 && step >= 1, "gnbv_adam_step: bad arguments");
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    int blocks = (int)std::min<int64_t>(ceil_div(n, 256), 148 * 8);
    adam_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, n, clip_workspace, (float)lr,
                                                          (float)beta1, (float)beta2, (float)eps, (float)bc1,
                                                          (float)sqrt(bc2), (float)grad_scale);
    GNBV_LAUNCH_CHECK("adam_kernel");
    return GNBV_OK;
}
