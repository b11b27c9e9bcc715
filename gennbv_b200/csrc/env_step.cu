// env_step.cu -- the scalar / bookkeeping part of Env_Train_GenNBV.step() for all envs, on device,
// without a single host read-back (the reference syncs ~10x per env per step).
//
//   actions_to_poses_kernel      step(): clip, forced init_action, idx*unit+low      env_train_gennbv.py:246-255
//   obs_update_kernel            post_process_camera_tensor (rgb part), update_obs_buf, obs dict assembly
//                                env_train_base.py:514-518; env_train_gennbv.py:273-275,359-366; wrapper :27-56
//   reward_termination_kernel    compute_reward, _reward_*, check_termination, update_extra_episode_info,
//                                the episode statistics of reset_idx          env_train_base.py:377-398,629-639;
//                                                                              env_train_gennbv.py:424-457,535-556
//   reset_envs_kernel            reset_idx's buffer resets                      env_train_gennbv.py:395-421
//
// Compiled with -fmad=false (one rounding per torch op).
#include "common.cuh"

namespace gnbv {

// ---------------------------------------------------------------------------------------------------------------
__global__ void actions_to_poses_kernel(const int64_t* __restrict__ actions_in, const int64_t* __restrict__ episode_length,
                                        const int64_t* __restrict__ idx_low, const int64_t* __restrict__ idx_up,
                                        const int64_t* __restrict__ init_action, const float* __restrict__ unit,
                                        const float* __restrict__ low_world, int64_t* __restrict__ actions_out,
                                        float* __restrict__ poses, int N, int A) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * A) return;
    int n = i / A, a = i - n * A;
    int64_t v = actions_in[i];
    v = v < idx_low[a] ? idx_low[a] : (v > idx_up[a] ? idx_up[a] : v);       // torch.clip (:247)
    if (episode_length[n] == 0) v = init_action[a];                           // :249-253
    actions_out[i] = v;
    poses[i] = __fadd_rn(__fmul_rn((float)v, unit[a]), low_world[a]);         // env_train_base.py:665-667
}

// ---------------------------------------------------------------------------------------------------------------
// torchvision rgb_to_grayscale on a uint8 image: (0.2989 r + 0.587 g + 0.114 b) in fp32, truncated to uint8,
// then .to(float32) (env_train_base.py:516).  F.interpolate(mode="nearest"): src = min(floor(dst * in/out), in-1).
__device__ __forceinline__ float gray_u8(uchar4 p) {
    float l = __fadd_rn(__fadd_rn(__fmul_rn(0.2989f, (float)p.x), __fmul_rn(0.587f, (float)p.y)), __fmul_rn(0.114f, (float)p.z));
    return (float)(unsigned char)(int)l;
}

__global__ void __launch_bounds__(256)
obs_update_kernel(const uchar4* __restrict__ rgba, const float* __restrict__ poses, float* __restrict__ pose_hist,
                  float* __restrict__ rgb_hist, float* __restrict__ obs, int64_t obs_stride, int64_t state_off,
                  int64_t rgb_off, int H, int W, int B, int A, int K, int RH, int RW, float scale_h, float scale_w) {
    const int n = blockIdx.x, tid = threadIdx.x;
    extern __shared__ float sh[];                         // pose history staging: B*A floats
    // pose history: drop the oldest entry, append the new pose (deque.extend, :273-274); obs["state"] = stack(dim=1)
    float* ph = pose_hist + (size_t)n * B * A;
    float* orow = obs + (size_t)n * obs_stride;
    const int BA = B * A;
    for (int i = tid; i < BA; i += blockDim.x) sh[i] = (i + A < BA) ? ph[i + A] : poses[n * A + (i + A - BA)];
    __syncthreads();
    for (int i = tid; i < BA; i += blockDim.x) { ph[i] = sh[i]; orow[state_off + i] = sh[i]; }
    // rgb history: shift frames, append the new grayscale frame; obs["state_rgb"] = cat(dim=1)
    const int F = RH * RW;
    float* rh = rgb_hist + (size_t)n * K * F;
    const uchar4* img = rgba ? rgba + (size_t)n * H * W : nullptr;
    for (int i = tid; i < F; i += blockDim.x) {
        int y = i / RW, x = i - y * RW;
        float g = 0.0f;
        if (img) {
            int sy = min((int)floorf(__fmul_rn((float)y, scale_h)), H - 1);
            int sx = min((int)floorf(__fmul_rn((float)x, scale_w)), W - 1);
            g = gray_u8(img[sy * W + sx]);
        }
        for (int k = 0; k < K; ++k) {                    // each thread owns pixel i of every frame: no hazard
            float v = (k + 1 < K) ? rh[(k + 1) * F + i] : g;
            rh[k * F + i] = v;
            orow[rgb_off + k * F + i] = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
struct RewardParams {
    float scale_cov, scale_short, scale_term;   // reward_scales[...] (already x dt), cast to fp32 like torch does
    int has_term, only_positive, max_step_done;
    int accumulate_reset;                       // eval env: `reset_buf |= ...` (env_eval_gennbv.py:327-333) instead of `=`
    int64_t max_episode_length;
    float max_episode_length_s;
    float ratio_threshold;
};

// stats layout (float64): [0] ring write position, [1] ring count, [2..2+100) reward ring, [102..202) length ring,
// [202] mean reward, [203] mean length, [204..207) rew_<name> means (cov, short, term)
constexpr int STAT_RING = 100;
constexpr int STAT_DOUBLES = 2 + 2 * STAT_RING + 2 + 3;

__global__ void __launch_bounds__(1024)
reward_termination_kernel(const float* __restrict__ cov_sum, const float* __restrict__ num_valid,
                          float* __restrict__ ratio_prev, int64_t* __restrict__ episode_length,
                          const uint8_t* __restrict__ collision, float* __restrict__ rew_buf,
                          uint8_t* __restrict__ reset_buf, uint8_t* __restrict__ time_out_buf,
                          uint8_t* __restrict__ dones_out, float* __restrict__ episode_sums /*[3,N]*/,
                          float* __restrict__ cur_reward_sum, float* __restrict__ cur_episode_length,
                          double* __restrict__ stats, uint8_t* __restrict__ time_outs_extra, RewardParams p, int N) {
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        int64_t len = episode_length[n] + 1;                               // post_physics_step :336
        episode_length[n] = len;
        // _reward_surface_coverage (:535-539)
        float ratio = __fdiv_rn(cov_sum[n], num_valid[n]);
        float r_cov = __fmul_rn(__fsub_rn(ratio, ratio_prev[n]), p.scale_cov);
        ratio_prev[n] = ratio;
        // _reward_short_path (:541-545): -(clip(len - 30, 0, 2)) as int64, times the scale
        int64_t extra = len - 30;
        extra = extra < 0 ? 0 : (extra > 2 ? 2 : extra);
        float r_short = __fmul_rn((float)(-extra), p.scale_short);
        float rew = __fadd_rn(__fadd_rn(0.0f, r_cov), r_short);            // compute_reward (:382-387)
        episode_sums[0 * N + n] = __fadd_rn(episode_sums[0 * N + n], r_cov);
        episode_sums[1 * N + n] = __fadd_rn(episode_sums[1 * N + n], r_short);
        if (p.only_positive) rew = fmaxf(rew, 0.0f);
        // check_termination (:438-457)
        bool col = collision ? collision[n] != 0 : false;
        bool tout = p.max_step_done ? (len >= p.max_episode_length) : (time_out_buf[n] != 0);
        bool rst = col || (p.max_step_done && tout) || (ratio > p.ratio_threshold) || (p.accumulate_reset && reset_buf[n] != 0);
        if (p.has_term) {                                                  // _reward_termination (:555-556)
            float r_term = __fmul_rn((rst && !tout) ? 1.0f : 0.0f, p.scale_term);
            rew = __fadd_rn(rew, r_term);
            episode_sums[2 * N + n] = __fadd_rn(episode_sums[2 * N + n], r_term);
        }
        rew_buf[n] = rew;
        reset_buf[n] = rst;
        dones_out[n] = rst;
        time_out_buf[n] = tout;
        // update_extra_episode_info (env_train_base.py:629-634) -- uses the reward before the algorithm's bootstrap
        cur_reward_sum[n] = __fadd_rn(cur_reward_sum[n], rew);
        cur_episode_length[n] = __fadd_rn(cur_episode_length[n], 1.0f);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // serial, in env order, exactly like the Python lists: deque(maxlen=100) of finished-episode stats
        int pos = (int)stats[0], cnt = (int)stats[1];
        double s_cov = 0, s_short = 0, s_term = 0;
        int n_reset = 0;
        for (int n = 0; n < N; ++n) {
            if (!reset_buf[n]) continue;
            ++n_reset;
            s_cov += episode_sums[0 * N + n]; s_short += episode_sums[1 * N + n]; s_term += episode_sums[2 * N + n];
            stats[2 + pos] = (double)cur_reward_sum[n];
            stats[2 + STAT_RING + pos] = (double)cur_episode_length[n];
            pos = (pos + 1) % STAT_RING;
            cnt = min(cnt + 1, STAT_RING);
            cur_reward_sum[n] = 0.0f;
            cur_episode_length[n] = 0.0f;
        }
        stats[0] = pos; stats[1] = cnt;
        double mr = 0, ml = 0;
        for (int i = 0; i < cnt; ++i) { mr += stats[2 + i]; ml += stats[2 + STAT_RING + i]; }
        stats[2 + 2 * STAT_RING] = cnt ? mr / cnt : 0.0;
        stats[2 + 2 * STAT_RING + 1] = cnt ? ml / cnt : 0.0;
        if (n_reset) {                                                     // reset_idx (:424-428): only when some env resets
            stats[2 + 2 * STAT_RING + 2] = s_cov / n_reset / p.max_episode_length_s;
            stats[2 + 2 * STAT_RING + 3] = s_short / n_reset / p.max_episode_length_s;
            stats[2 + 2 * STAT_RING + 4] = s_term / n_reset / p.max_episode_length_s;
            // extras["time_outs"] is (re)bound only inside reset_idx (:435-436) and therefore stays stale between
            // resets: reproduce by refreshing the exported copy only on steps where some env resets
            for (int n = 0; n < N; ++n) time_outs_extra[n] = time_out_buf[n];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
reset_envs_kernel(uint8_t* __restrict__ reset_buf, float* __restrict__ prob, float* __restrict__ scan,
                  float* __restrict__ pose_hist, float* __restrict__ rgb_hist, float* __restrict__ ratio_prev,
                  int64_t* __restrict__ actions, int64_t* __restrict__ episode_length, float* __restrict__ episode_sums,
                  const float* __restrict__ init_pose, const int64_t* __restrict__ init_action, int N, int V, int BA,
                  int A, int KF, int vec_ok) {
    const int n = blockIdx.y;
    if (!reset_buf[n]) return;
    const size_t base = (size_t)n * V;
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    if (vec_ok) {
        for (int v = gtid * 4; v < V; v += gsz * 4) {
            stg_stream4(prob + base + v, make_float4(0.f, 0.f, 0.f, 0.f));
            stg_stream4(scan + base + v, make_float4(0.f, 0.f, 0.f, 0.f));
        }
    } else {
        for (int v = gtid; v < V; v += gsz) { prob[base + v] = 0.f; scan[base + v] = 0.f; }
    }
    for (int i = gtid; i < BA; i += gsz) pose_hist[(size_t)n * BA + i] = init_pose[i % A];
    for (int i = gtid; i < KF; i += gsz) rgb_hist[(size_t)n * KF + i] = 0.0f;
    if (gtid < A) actions[n * A + gtid] = init_action[gtid];
    if (gtid < 3) episode_sums[gtid * N + n] = 0.0f;
    if (gtid == 0) { ratio_prev[n] = 0.0f; episode_length[n] = 0; }
    // reset_buf[env_ids] = 0 (env_train_gennbv.py:373) is applied afterwards by clear_flags_kernel: `dones_out` of
    // gnbv_reward_termination keeps the done flags for the caller
}

__global__ void clear_flags_kernel(uint8_t* flags, int N) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) flags[i] = 0;
}

}  // namespace gnbv

using namespace gnbv;

extern "C" int gnbv_actions_to_poses(const int64_t* actions_in, const int64_t* episode_length, const int64_t* idx_low,
                                     const int64_t* idx_up, const int64_t* init_action, const float* unit,
                                     const float* low_world, int64_t* actions_out, float* poses, int N, int A,
                                     void* stream) {
    GNBV_REQUIRE(actions_in && episode_length && idx_low && idx_up && init_action && unit && low_world && actions_out && poses,
                 "gnbv_actions_to_poses: null pointer argument");
    GNBV_REQUIRE(N > 0 && A > 0, "gnbv_actions_to_poses: N, A must be positive");
    actions_to_poses_kernel<<<(unsigned)ceil_div((int64_t)N * A, 256), 256, 0, (cudaStream_t)stream>>>(
        actions_in, episode_length, idx_low, idx_up, init_action, unit, low_world, actions_out, poses, N, A);
    GNBV_LAUNCH_CHECK("actions_to_poses_kernel");
    return GNBV_OK;
}

extern "C" int gnbv_obs_update(const uint8_t* rgba, const float* poses, float* pose_hist, float* rgb_hist, float* obs,
                               int64_t obs_row_stride, int64_t state_off, int64_t rgb_off, int N, int H, int W,
                               int hist_len, int pose_dim, int rgb_frames, int rgb_h, int rgb_w, void* stream) {
    GNBV_REQUIRE(poses && pose_hist && rgb_hist && obs, "gnbv_obs_update: null pointer argument");
    GNBV_REQUIRE(N > 0 && hist_len > 0 && pose_dim > 0 && rgb_frames > 0 && rgb_h > 0 && rgb_w > 0,
                 "gnbv_obs_update: sizes must be positive");
    GNBV_REQUIRE(rgba == nullptr || (H > 0 && W > 0), "gnbv_obs_update: image size must be positive");
    GNBV_REQUIRE(((uintptr_t)rgba & 3) == 0, "gnbv_obs_update: rgba must be 4-byte aligned");
    const size_t smem = (size_t)hist_len * pose_dim * 4;
    GNBV_REQUIRE(smem <= 48 * 1024, "gnbv_obs_update: pose history too long for shared memory");
    // F.interpolate's nearest scale = (float)in / out  (computed in fp32 by ATen's area_pixel_compute_scale)
    const float sh = (float)H / (float)rgb_h, sw = (float)W / (float)rgb_w;
    obs_update_kernel<<<N, 256, smem, (cudaStream_t)stream>>>(reinterpret_cast<const uchar4*>(rgba), poses, pose_hist,
                                                              rgb_hist, obs, obs_row_stride, state_off, rgb_off, H, W,
                                                              hist_len, pose_dim, rgb_frames, rgb_h, rgb_w, sh, sw);
    GNBV_LAUNCH_CHECK("obs_update_kernel");
    return GNBV_OK;
}

extern "C" size_t gnbv_episode_stats_doubles(void) { return STAT_DOUBLES; }

extern "C" int gnbv_reward_termination(const float* cov_sum, const float* num_valid, float* ratio_prev,
                                       int64_t* episode_length, const uint8_t* collision, float* rew_buf,
                                       uint8_t* reset_buf, uint8_t* time_out_buf, uint8_t* dones_out, float* episode_sums,
                                       float* cur_reward_sum, float* cur_episode_length, double* stats,
                                       uint8_t* time_outs_extra, double scale_cov, double scale_short, double scale_term,
                                       int has_termination_reward, int only_positive_rewards, int max_step_done,
                                       int64_t max_episode_length, double max_episode_length_s, double ratio_threshold,
                                       int accumulate_reset, int N, void* stream) {
    GNBV_REQUIRE(cov_sum && num_valid && ratio_prev && episode_length && rew_buf && reset_buf && time_out_buf && dones_out &&
                     episode_sums && cur_reward_sum && cur_episode_length && stats && time_outs_extra,
                 "gnbv_reward_termination: null pointer argument");
    GNBV_REQUIRE(N > 0, "gnbv_reward_termination: N must be positive");
    RewardParams p;
    p.scale_cov = (float)scale_cov; p.scale_short = (float)scale_short; p.scale_term = (float)scale_term;
    p.has_term = has_termination_reward; p.only_positive = only_positive_rewards; p.max_step_done = max_step_done;
    p.max_episode_length = max_episode_length; p.max_episode_length_s = (float)max_episode_length_s;
    p.ratio_threshold = (float)ratio_threshold;
    p.accumulate_reset = accumulate_reset;
    reward_termination_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(cov_sum, num_valid, ratio_prev, episode_length, collision,
                                                                    rew_buf, reset_buf, time_out_buf, dones_out, episode_sums,
                                                                    cur_reward_sum, cur_episode_length, stats,
                                                                    time_outs_extra, p, N);
    GNBV_LAUNCH_CHECK("reward_termination_kernel");
    return GNBV_OK;
}

extern "C" int gnbv_reset_envs(uint8_t* reset_buf, float* prob_grid, float* scanned_gt, float* pose_hist, float* rgb_hist,
                               float* ratio_prev, int64_t* actions, int64_t* episode_length, float* episode_sums,
                               const float* init_pose, const int64_t* init_action, int N, int G, int hist_len,
                               int pose_dim, int rgb_frames, int rgb_h, int rgb_w, int clear_reset_buf, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GNBV_REQUIRE(reset_buf && prob_grid && scanned_gt && pose_hist && rgb_hist && ratio_prev && actions && episode_length &&
                     episode_sums && init_pose && init_action,
                 "gnbv_reset_envs: null pointer argument");
    GNBV_REQUIRE(N > 0 && G > 0, "gnbv_reset_envs: N, G must be positive");
    const int64_t V = (int64_t)G * G * G;
    const int vec_ok = (V % 4 == 0) && (((uintptr_t)prob_grid | (uintptr_t)scanned_gt) & 15) == 0;
    dim3 grid((unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(V, 1024 * 4), 64)), (unsigned)N);
    reset_envs_kernel<<<grid, 256, 0, stream>>>(reset_buf, prob_grid, scanned_gt, pose_hist, rgb_hist, ratio_prev, actions,
                                                episode_length, episode_sums, init_pose, init_action, N, (int)V,
                                                hist_len * pose_dim, pose_dim, rgb_frames * rgb_h * rgb_w, vec_ok);
    GNBV_LAUNCH_CHECK("reset_envs_kernel");
    if (clear_reset_buf) {
        clear_flags_kernel<<<(unsigned)ceil_div(N, 256), 256, 0, stream>>>(reset_buf, N);
        GNBV_LAUNCH_CHECK("clear_flags_kernel");
    }
    return GNBV_OK;
}
