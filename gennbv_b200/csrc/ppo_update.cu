// ppo_update.cu -- one PPO minibatch update (stable_baselines3/ppo/ppo_grid_obs.py:201-275) as a fixed, graph-capturable
// sequence of launches with NO host-side decision inside it:
//
//   gather rows/columns of the minibatch -> encoder forward (batch-stat BN) -> heads -> MultiCategorical -> PPO loss fwd+bwd
//   -> MultiCategorical bwd -> heads bwd -> encoder bwd            (gnbv_ppo_minibatch_grads, optionally in two phases)
//   [data-parallel callers all-reduce the flat gradient arena + the KL vote here]
//   -> global-norm clip -> Adam                                      (gnbv_ppo_minibatch_apply)
//
// What the reference decides on the host is decided on the device:
//   * which rows form the minibatch: a device cursor `ctl[0]` counts the minibatches of the epoch; the gather kernel reads
//     storage_rows[cursor*B .. +B) (one permutation per rollout, buffers.py:673,750);
//   * the KL early stop (ppo_grid_obs.py:259-268): the loss kernel's approx_kl is compared with 1.5*target_kl on the
//     device and the result written as a vote (1.0 / 0.0) into a float slot that the caller appends to the gradient
//     bucket, so that ONE all-reduce (sum) carries both the gradients and the "any rank wants to stop" decision;
//     gnbv_ppo_minibatch_apply turns vote > 0 into the sticky flag ctl[1].  While the flag is set, Adam is a no-op, the
//     Adam step counter and the BatchNorm running statistics stay untouched and nothing is logged: minibatches that were
//     already enqueued change no state, which is what the reference's `break` achieves.  The host reads ctl once per epoch;
//   * the Adam step count (bias corrections) lives in ctl[2].
//   ctl (int64[8]): 0 cursor, 1 stop flag, 2 adam step, 3 cursor value at which the stop was raised (-1), 4 logged rows.
#include "encoder.cuh"
#include "gemm.cuh"

#include <math.h>

namespace gnbv {
namespace {

struct MbWs {            // float offsets into the minibatch scratch
    size_t rows, actions, old_v, old_lp, adv, ret, feats, out, lp, ent, g_lp, g_ent, g_v, dout, dfeat, scalars, gemm, total;
};

MbWs make_mb_ws(int B, int A1, int nsub, int F) {
    MbWs w;
    size_t o = 0;
    auto take = [&](size_t n) { size_t r = o; o += (n + 63) & ~(size_t)63; return r; };
    w.rows = take((size_t)B * 2);                    // int64
    w.actions = take((size_t)B * nsub * 2);          // int64
    w.old_v = take(B); w.old_lp = take(B); w.adv = take(B); w.ret = take(B);
    w.feats = take((size_t)B * F); w.out = take((size_t)B * A1);
    w.lp = take(B); w.ent = take(B); w.g_lp = take(B); w.g_ent = take(B); w.g_v = take(B);
    w.dout = take((size_t)B * A1); w.dfeat = take((size_t)B * F);
    w.scalars = take(64);
    size_t g = std::max(gemm_workspace_floats(B, A1, F), std::max(gemm_workspace_floats(B, F, A1), gemm_workspace_floats(A1, F, B)));
    w.gemm = take(g);
    w.total = o;
    return w;
}

// rows of this minibatch + its scalar columns (buffers.py:753-762 without the observation gather)
__global__ void mb_gather_kernel(const int64_t* __restrict__ storage_rows, int64_t rows_base, const int64_t* __restrict__ ctl, int B, int nsub,
                                 const float* __restrict__ actions, const float* __restrict__ values,
                                 const float* __restrict__ log_probs, const float* __restrict__ advantages,
                                 const float* __restrict__ returns, int64_t* __restrict__ rows_out,
                                 int64_t* __restrict__ actions_out, float* __restrict__ old_v, float* __restrict__ old_lp,
                                 float* __restrict__ adv, float* __restrict__ ret) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const int64_t r = storage_rows[rows_base + ctl[0] * B + i];
    rows_out[i] = r;
    for (int k = 0; k < nsub; ++k) actions_out[(int64_t)i * nsub + k] = (int64_t)actions[r * nsub + k];    // `.long()`
    old_v[i] = values[r]; old_lp[i] = log_probs[r]; adv[i] = advantages[r]; ret[i] = returns[r];
}

// values column of `out` -> contiguous (ppo_loss reads it), and later g_v -> the value column of dout
__global__ void col_copy_kernel(const float* __restrict__ src, int64_t lds, float* __restrict__ dst, int64_t ldd, int B) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) dst[(int64_t)i * ldd] = src[(int64_t)i * lds];
}

__global__ void colsum_rows_kernel(const float* __restrict__ x, int64_t ld, int rows, int cols, float* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    float s = 0.f;
    for (int r = 0; r < rows; ++r) s += x[(int64_t)r * ld + c];
    out[c] = s;
}

// vote = (approx_kl > 1.5 target_kl) as a float the gradient all-reduce sums over ranks
__global__ void kl_vote_kernel(const float* __restrict__ scalars, float threshold, int enabled, float* __restrict__ vote) {
    *vote = (enabled && scalars[4] > threshold) ? 1.0f : 0.0f;
}

// decision + bookkeeping, one thread: runs after the (optional) all-reduce and BEFORE the clip / Adam kernels of the same call
__global__ void ppo_ctl_kernel(int64_t* __restrict__ ctl, const float* __restrict__ vote, const float* __restrict__ scalars,
                               float* __restrict__ log, int64_t log_capacity, float* __restrict__ adam_coef, double b1, double b2) {
    const bool was_stopped = ctl[1] != 0;
    const bool stop_now = !was_stopped && vote && *vote > 0.f;
    if (!was_stopped) {                                  // the stopping minibatch itself IS logged (ppo_grid_obs.py:227-262)
        const int64_t k = ctl[4];
        if (log && k < log_capacity)
            for (int j = 0; j < 8; ++j) log[k * 8 + j] = scalars[j];
        ctl[4] = k + 1;
    }
    if (stop_now) { ctl[1] = 1; ctl[3] = ctl[0]; }
    const bool live = !was_stopped && !stop_now;
    if (live) ctl[2] += 1;
    const double t = (double)ctl[2];
    adam_coef[0] = live ? 1.f : 0.f;
    adam_coef[1] = (float)(1.0 - pow(b1, t));            // bias corrections of th.optim.Adam
    adam_coef[2] = (float)sqrt(1.0 - pow(b2, t));
    ctl[0] += 1;
}

__global__ void __launch_bounds__(256)
adam_ctl_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
                const float* __restrict__ clip /*[2]*/, const float* __restrict__ adam_coef /*[3]*/, float lr, float b1, float b2,
                float eps, float grad_scale) {
    if (adam_coef[0] == 0.f) return;                     // KL stop raised: no state change
    const float coef = clip[1] * grad_scale, bc1 = adam_coef[1], bc2_sqrt = adam_coef[2];
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const float gi = g[i] * coef;
        const float mi = b1 * m[i] + (1.f - b1) * gi;
        const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi; v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] -= (lr / bc1) * (mi / denom);
    }
}

}  // namespace
}  // namespace gnbv

using namespace gnbv;

extern "C" size_t gnbv_ppo_minibatch_workspace_bytes(int batch, int num_logits, int num_sub, int feat_dim) {
    if (batch <= 0 || num_logits <= 0 || num_sub <= 0 || feat_dim <= 0) return 0;
    return make_mb_ws(batch, num_logits + 1, num_sub, feat_dim).total * 4 + 256;
}

extern "C" int gnbv_ppo_minibatch_grads(const gnbv_ppo_minibatch* a, int phases, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GNBV_REQUIRE(a && a->enc && a->enc_grads && a->head_w && a->head_b && a->head_w_grad && a->head_b_grad && a->nvec &&
                     a->observations && a->actions && a->values && a->log_probs && a->advantages && a->returns &&
                     a->storage_rows && a->ctl && a->vote && a->enc_workspace && a->mb_workspace,
                 "gnbv_ppo_minibatch_grads: null pointer argument");
    GNBV_REQUIRE(a->batch > 0 && a->num_sub > 0 && a->feat_dim == 256, "gnbv_ppo_minibatch_grads: bad sizes");
    GNBV_REQUIRE((phases & GNBV_BWD_ALL) != 0, "gnbv_ppo_minibatch_grads: empty phase mask");
    int A = 0;
    for (int k = 0; k < a->num_sub; ++k) A += a->nvec[k];
    const int B = a->batch, F = a->feat_dim, A1 = A + 1;
    MbWs w = make_mb_ws(B, A1, a->num_sub, F);
    GNBV_REQUIRE(a->mb_workspace_bytes >= w.total * 4, "gnbv_ppo_minibatch_grads: minibatch workspace %zu B < %zu B",
                 a->mb_workspace_bytes, w.total * 4);
    GNBV_REQUIRE(((uintptr_t)a->mb_workspace & 255) == 0, "gnbv_ppo_minibatch_grads: workspace must be 256 B aligned");
    float* ws = reinterpret_cast<float*>(a->mb_workspace);
    int64_t* rows = reinterpret_cast<int64_t*>(ws + w.rows);
    int64_t* acts = reinterpret_cast<int64_t*>(ws + w.actions);
    const unsigned nb = (unsigned)ceil_div(B, 128);
    int rc;
    if (phases & GNBV_BWD_LINEAR) {
        mb_gather_kernel<<<nb, 128, 0, stream>>>(a->storage_rows, a->rows_base, a->ctl, B, a->num_sub, a->actions, a->values, a->log_probs,
                                                 a->advantages, a->returns, rows, acts, ws + w.old_v, ws + w.old_lp, ws + w.adv,
                                                 ws + w.ret);
        GNBV_LAUNCH_CHECK("mb_gather_kernel");
        rc = encoder_forward_impl(a->enc, a->observations, a->obs_row_stride, rows, B, a->grid_size, a->state_dim, 1, ws + w.feats,
                                  a->enc_workspace, a->enc_workspace_bytes, a->ctl + 1, stream);
        if (rc) return rc;
        rc = gnbv_policy_heads_forward(ws + w.feats, a->head_w, a->head_b, ws + w.out, B, F, A1, stream);
        if (rc) return rc;
        rc = gnbv_multicategorical_evaluate(ws + w.out, A1, a->nvec, a->num_sub, acts, ws + w.lp, ws + w.ent, B, stream);
        if (rc) return rc;
        // value column -> contiguous (the dfeat scratch is free until the heads' backward GEMM writes it)
        col_copy_kernel<<<nb, 128, 0, stream>>>(ws + w.out + A, A1, ws + w.dfeat, 1, B);
        rc = gnbv_ppo_loss(ws + w.lp, ws + w.ent, ws + w.dfeat, ws + w.old_v, ws + w.old_lp, ws + w.adv, ws + w.ret, B, a->clip_range,
                           a->clip_range_vf, a->ent_coef, a->vf_coef, a->pg_coef, a->normalize_advantage, ws + w.scalars,
                           ws + w.g_lp, ws + w.g_ent, ws + w.g_v, stream);
        if (rc) return rc;
        kl_vote_kernel<<<1, 1, 0, stream>>>(ws + w.scalars, (float)(1.5 * a->target_kl), a->target_kl >= 0 ? 1 : 0, a->vote);
        rc = gnbv_multicategorical_backward(ws + w.out, A1, a->nvec, a->num_sub, acts, ws + w.g_lp, ws + w.g_ent, ws + w.dout, A1, B,
                                            stream);
        if (rc) return rc;
        col_copy_kernel<<<nb, 128, 0, stream>>>(ws + w.g_v, 1, ws + w.dout + A, A1, B);
        GNBV_LAUNCH_CHECK("col_copy_kernel");
        GemmEpilogue none;
        rc = launch_gemm(ws + w.dout, A1, 1, a->head_w, F, 1, ws + w.dfeat, F, B, F, A1, none, ws + w.gemm, stream);           // dfeat = dout W
        if (rc) return rc;
        rc = launch_gemm(ws + w.dout, 1, A1, ws + w.feats, F, 1, a->head_w_grad, F, A1, F, B, none, ws + w.gemm, stream);    // dW = dout^T feats
        if (rc) return rc;
        colsum_rows_kernel<<<(unsigned)ceil_div(A1, 128), 128, 0, stream>>>(ws + w.dout, A1, B, A1, a->head_b_grad);
        GNBV_LAUNCH_CHECK("colsum_rows_kernel");
    }
    return gnbv_encoder_backward_phase(a->enc, a->observations, a->obs_row_stride, rows, B, a->grid_size, a->state_dim, 1,
                                       ws + w.feats, ws + w.dfeat, a->enc_grads, a->enc_workspace, a->enc_workspace_bytes, phases,
                                       stream);
}

extern "C" int gnbv_ppo_minibatch_scalars(const gnbv_ppo_minibatch* a, const float** scalars) {
    GNBV_REQUIRE(a && scalars && a->mb_workspace && a->nvec, "gnbv_ppo_minibatch_scalars: null pointer argument");
    int A = 0;
    for (int k = 0; k < a->num_sub; ++k) A += a->nvec[k];
    MbWs w = make_mb_ws(a->batch, A + 1, a->num_sub, a->feat_dim);
    *scalars = reinterpret_cast<const float*>(a->mb_workspace) + w.scalars;
    return GNBV_OK;
}

extern "C" int gnbv_ppo_minibatch_apply(const gnbv_ppo_minibatch* a, float* params, const float* grads, float* exp_avg,
                                        float* exp_avg_sq, int64_t n, double max_grad_norm, double lr, double beta1, double beta2,
                                        double eps, double grad_scale, float* clip_workspace, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GNBV_REQUIRE(a && params && grads && exp_avg && exp_avg_sq && clip_workspace && a->ctl && a->vote && n > 0,
                 "gnbv_ppo_minibatch_apply: bad arguments");
    const float* scalars;
    int rc = gnbv_ppo_minibatch_scalars(a, &scalars);
    if (rc) return rc;
    float* adam_coef = clip_workspace + gnbv_clip_adam_workspace_bytes() / 4;          // 3 floats behind the clip scratch
    ppo_ctl_kernel<<<1, 1, 0, stream>>>(a->ctl, a->vote, scalars, a->log, a->log_capacity, adam_coef, beta1, beta2);
    GNBV_LAUNCH_CHECK("ppo_ctl_kernel");
    rc = gnbv_grad_norm(grads, n, max_grad_norm, clip_workspace, stream);
    if (rc) return rc;
    int blocks = (int)std::min<int64_t>(ceil_div(n, 256), 148 * 8);
    adam_ctl_kernel<<<blocks, 256, 0, stream>>>(params, grads, exp_avg, exp_avg_sq, n, clip_workspace, adam_coef, (float)lr,
                                                (float)beta1, (float)beta2, (float)eps, (float)grad_scale);
    GNBV_LAUNCH_CHECK("adam_ctl_kernel");
    return GNBV_OK;
}

extern "C" size_t gnbv_ppo_apply_workspace_bytes(void) { return gnbv_clip_adam_workspace_bytes() + 16 * sizeof(float); }
