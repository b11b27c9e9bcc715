/* gennbv_b200.h -- C ABI of the B200-native GenNBV hot path (libgennbv_b200.so).
 *
 * The reference (zjwzcx/GenNBV @ c373f76) has no FFI layer: its seams are Python call
 * signatures.  Each entry point below names the reference lines it
 * replaces; gennbv_b200/*.py keeps the reference's Python names on top of these and
 * INTEGRATION.md shows the ctypes binding a maintainer of the reference would add.
 *
 * Conventions (all entry points):
 *   - every pointer is a DEVICE pointer on the current CUDA device unless marked "host";
 *     the caller owns all memory, the library never allocates and never synchronises;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - tensors are dense, row-major, in the reference's own dtypes and layouts;
 *   - return value 0 = OK, <0 = error (GNBV_E_*); gnbv_last_error() returns a
 *     thread-local, human-readable message for the last failing call.
 */
#ifndef GENNBV_B200_H
#define GENNBV_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GNBV_OK 0
#define GNBV_E_ARG (-1)      /* bad argument (null pointer, unsupported size, misalignment) */
#define GNBV_E_CUDA (-2)     /* a CUDA runtime call / kernel launch failed */
#define GNBV_E_WORKSPACE (-3) /* workspace too small */

#define GNBV_ABI_VERSION 3

int gnbv_abi_version(void);
/* Kernel variants in effect for this process: which = 0 -> GNBV_CONV2_TC, 1 -> GNBV_CONV1_MMA, 2 -> GNBV_GEMM_MMA
 * (environment variables read once; defaults and bit meanings in gennbv_b200/csrc/api.cu). */
int gnbv_kernel_mode(int which);
const char* gnbv_last_error(void);

/* ---- optional stage timing (measurement aid, not on the reference's API surface) ----
 * After gnbv_profile_enable(1) the multi-kernel entry points record a CUDA event on their stream at every stage
 * boundary; gnbv_profile_elapsed_ms(a, b) synchronises on event b and returns the device time between stages a and b.
 * Encoder stage ids: forward GNBV_ST_FWD_BEGIN .. GNBV_ST_FWD_END, backward GNBV_ST_BWD_BEGIN .. GNBV_ST_BWD_END
 * (each id marks the START of the named stage). */
#define GNBV_MAX_STAGES 48
enum {
    GNBV_ST_FWD_BEGIN = 0, GNBV_ST_FWD_ACTION_MLP = 0, GNBV_ST_FWD_CONV1 = 1, GNBV_ST_FWD_BN1 = 2, GNBV_ST_FWD_CONV2 = 3,
    GNBV_ST_FWD_BN2 = 4, GNBV_ST_FWD_GRID_FC = 5, GNBV_ST_FWD_OUT_FC = 6, GNBV_ST_FWD_END = 7,
    GNBV_ST_BWD_BEGIN = 16, GNBV_ST_BWD_LINEAR = 16, GNBV_ST_BWD_GRID_FC = 17, GNBV_ST_BWD_BN2 = 18, GNBV_ST_BWD_CONV2_WGRAD = 19,
    GNBV_ST_BWD_CONV2_DGRAD = 20, GNBV_ST_BWD_BN1 = 21, GNBV_ST_BWD_CONV1_WGRAD = 22, GNBV_ST_BWD_END = 23
};
int gnbv_profile_enable(int on);
int gnbv_profile_elapsed_ms(int stage_from, int stage_to, float* ms);

/* ---- flags for gnbv_voxelize_step ---- */
#define GNBV_RAW_DEPTH 1u   /* depth is the raw sensor image: apply post_process_camera_tensor's chain first */

/* Workspace for gnbv_voxelize_step (bytes): two occupancy bit-masks + coverage partial sums. */
size_t gnbv_voxelize_workspace_bytes(int num_envs, int grid_size);

/* One env.step() worth of state encoding for all N envs, three launches, no host sync.
 * Replaces (per step, for all envs at once):
 *   Env_Train_Base.post_process_camera_tensor   gennbv/env/env_train_base.py:519-523  (depth chain; GNBV_RAW_DEPTH)
 *   Env_Train_GenNBV.back_projection_fg         gennbv/env/env_train_gennbv.py:494-533
 *   scanned_pts_to_idx_3D                       gennbv/utils.py:230-270
 *   pose_coord_to_idx_3D                        gennbv/utils.py:273-306
 *   bresenham3D_pycuda                          gennbv/utils.py:24-227
 *   Env_Train_GenNBV.update_occ_grid            gennbv/env/env_train_gennbv.py:277-326
 *   grid_occupancy_tri_cls                      gennbv/utils.py:309-325
 *   the sum in _reward_surface_coverage         gennbv/env/env_train_gennbv.py:537
 *
 *   depth      [N,H,W] f32   processed depth, or raw sensor depth with GNBV_RAW_DEPTH
 *   seg        [N,H,W] i32   segmentation ids (foreground = seg > 50)
 *   kinv       [3,3]   f32   inverse intrinsics
 *   c2w        [N,4,4] f32   camera-to-world, env-local (env origin already subtracted)
 *   range_gt   [N,6]   f32   x_max,x_min,y_max,y_min,z_max,z_min
 *   voxel_size [N,3]   f32
 *   pose_xyz   [N,3]   f32   camera position (ray source)
 *   grid_gt    [N,G,G,G] f32 GT occupancy {0,1}
 *   prob_grid  [N,G,G,G] f32 in/out
 *   scanned_gt [N,G,G,G] f32 in/out
 *   tri_out    f32 out: row n starts at tri_out + n*tri_row_stride (elements), G^3 values in {-1,0,1};
 *              tri_row_stride = G^3 for a dense [N,G,G,G] tensor, or the flattened observation row
 *              length when writing straight into the obs buffer (env_wrapper_gennbv_train.py:27-56)
 *   cov_sum    [N] f32 out   sum of scanned_gt per env after the update
 *   num_targets[N] i32 out   number of distinct occupied voxels hit this step (may be NULL)
 */
int gnbv_voxelize_step(const float* depth, const int32_t* seg, const float* kinv, const float* c2w,
                       const float* range_gt, const float* voxel_size, const float* pose_xyz,
                       const float* grid_gt, float* prob_grid, float* scanned_gt,
                       float* tri_out, int64_t tri_row_stride, float* cov_sum, int32_t* num_targets,
                       void* workspace, size_t workspace_bytes,
                       int num_envs, int height, int width, int grid_size, uint32_t flags, void* stream);

/* The two phases of gnbv_voxelize_step as separate entry points (same arguments, same workspace):
 *   gnbv_scan_raycast : back_projection_fg + scanned_pts_to_idx_3D + pose_coord_to_idx_3D + bresenham3D_pycuda
 *                       -> target / touched bit-masks in the workspace (see gnbv_voxelize_masks)
 *   gnbv_grid_update  : the dense part of update_occ_grid (env_train_gennbv.py:311-326), grid_occupancy_tri_cls
 *                       and the coverage sum, consuming those masks.
 * gnbv_voxelize_step == gnbv_scan_raycast followed by gnbv_grid_update. */
int gnbv_scan_raycast(const float* depth, const int32_t* seg, const float* kinv, const float* c2w,
                      const float* range_gt, const float* voxel_size, const float* pose_xyz,
                      int32_t* num_targets, void* workspace, size_t workspace_bytes,
                      int num_envs, int height, int width, int grid_size, uint32_t flags, void* stream);
int gnbv_grid_update(const float* grid_gt, float* prob_grid, float* scanned_gt,
                     float* tri_out, int64_t tri_row_stride, float* cov_sum,
                     void* workspace, size_t workspace_bytes, int num_envs, int grid_size, void* stream);

/* gnbv_grid_update moving only the bytes that can change: 16-byte groups no ray touched keep their prob / scanned values and cost
 * 8 B per voxel (prob read, tri write) instead of 24.  cov_sum is IN/OUT here: it must hold the env's coverage sum before the
 * step (0 after a reset) and receives the exact increment.  Needs 16-byte aligned grids, G^3 and tri_row_stride % 4 == 0. */
int gnbv_grid_update_sparse(const float* grid_gt, float* prob_grid, float* scanned_gt,
                            float* tri_out, int64_t tri_row_stride, float* cov_sum,
                            void* workspace, size_t workspace_bytes, int num_envs, int grid_size, void* stream);

/* Debug/compat view of the step's intermediate sets, valid after gnbv_voxelize_step on the same
 * workspace: bit v of row n (v = (x*G+y)*G+z, little-endian within u32 words) of
 *   target mask  = voxels returned by scanned_pts_to_idx_3D (+unique) for env n  (utils.py:230-270)
 *   touched mask = union of bresenham3D_pycuda's in-bounds path voxels for env n  (utils.py:24-227)
 * words_per_env receives the row pitch in u32 words. Returned pointers alias `workspace`. */
int gnbv_voxelize_masks(void* workspace, int num_envs, int grid_size,
                        const uint32_t** target_mask, const uint32_t** touched_mask, int64_t* words_per_env);

/* Kernels behind the reference's free functions, for callers that use them outside the env (gennbv/utils.py):
 *   gnbv_points_to_voxel_mask : scanned_pts_to_idx_3D (utils.py:230-270) for ONE env from explicit world points
 *       [num_points,3] f32: ORs bit (x*G+y)*G+z of `mask` (ceil(G^3/32) u32 words, zeroed by the caller) for every point
 *       strictly inside the grid volume; the set bits in increasing order are the unique, clamped, sorted index rows.
 *   gnbv_bresenham_rays : bresenham3D_pycuda (utils.py:24-227).  Call once with counts != NULL (out = NULL) to get the
 *       per-ray number of in-bounds voxels, exclusive-scan them into offsets, call again with out [sum,3] i64: rows
 *       are written in ray order, duplicates kept, exactly the reference's concatenated trajectory. source3 / targets i32. */
int gnbv_points_to_voxel_mask(const float* points, int64_t num_points, const float* range_gt6, const float* voxel_size3,
                              uint32_t* mask, int grid_size, void* stream);
int gnbv_bresenham_rays(const int32_t* source3, const int32_t* targets, int64_t num_rays, int grid_size, int64_t* counts,
                        const int64_t* offsets, int64_t* out, void* stream);

/* reset_idx's grid part (env_train_gennbv.py:413-417): zero prob_grid / scanned_gt rows whose
 * reset flag is non-zero.  reset_flags [N] u8 on device -- no host round trip. */
int gnbv_reset_grids(float* prob_grid, float* scanned_gt, const uint8_t* reset_flags,
                     int num_envs, int grid_size, void* stream);

/* ---- env.step() bookkeeping: everything below runs for all envs on device, no host read-back ---- */

/* Env_Train_GenNBV.step() head (env_train_gennbv.py:246-255 + env_train_base.py:665-667):
 * actions_out = clip(actions_in, idx_low, idx_up); rows with episode_length == 0 are forced to init_action;
 * poses = actions_out * unit + low_world (fp32).  actions [N,A] i64, poses [N,A] f32, per-axis vectors [A]. */
int gnbv_actions_to_poses(const int64_t* actions_in, const int64_t* episode_length, const int64_t* idx_low,
                          const int64_t* idx_up, const int64_t* init_action, const float* unit, const float* low_world,
                          int64_t* actions_out, float* poses, int num_envs, int action_dim, void* stream);

/* rgb part of post_process_camera_tensor (env_train_base.py:514-518: nearest resize to rgb_h x rgb_w, torchvision
 * rgb_to_grayscale on uint8, .float()), update_obs_buf (env_train_gennbv.py:273-275: push pose and frame into the
 * histories) and the "state" / "state_rgb" columns of the flattened observation (env_train_gennbv.py:359-366 +
 * env_wrapper_gennbv_train.py:27-56).
 *   rgba [N,H,W,4] u8 (NULL = black frame); poses [N,pose_dim] f32;
 *   pose_hist [N,hist_len,pose_dim] f32 in/out (oldest first); rgb_hist [N,rgb_frames,rgb_h,rgb_w] f32 in/out;
 *   obs row n starts at obs + n*obs_row_stride; state at +state_off, frames at +rgb_off (elements). */
int gnbv_obs_update(const uint8_t* rgba, const float* poses, float* pose_hist, float* rgb_hist, float* obs,
                    int64_t obs_row_stride, int64_t state_off, int64_t rgb_off, int num_envs, int height, int width,
                    int hist_len, int pose_dim, int rgb_frames, int rgb_h, int rgb_w, void* stream);

/* Number of doubles in the `stats` block of gnbv_reward_termination. Layout: [0] ring position, [1] ring count,
 * [2,102) finished-episode rewards, [102,202) finished-episode lengths, [202] mean reward, [203] mean length,
 * [204..206] rew_surface_coverage / rew_short_path / rew_termination (infos["episode"], env_train_gennbv.py:424-428). */
size_t gnbv_episode_stats_doubles(void);

/* post_physics_step's scalar part: episode_length += 1 (env_train_gennbv.py:336); compute_reward with
 * _reward_surface_coverage, _reward_short_path, check_termination, _reward_termination
 * (env_train_base.py:377-398; env_train_gennbv.py:438-457,535-556); update_extra_episode_info
 * (env_train_base.py:629-639) and reset_idx's episode statistics (env_train_gennbv.py:424-436).
 * All [N]; flags are u8 {0,1}; episode_sums [3,N] (coverage, short_path, termination); scales are the Python
 * doubles reward_scales[name] (cfg scale x dt). collision may be NULL (no contacts).
 * time_outs_extra reproduces infos["time_outs"], which the reference rebinds only inside reset_idx.
 * ratio_threshold: coverage-ratio termination (0.99 in the train env, +inf = none as in the eval env);
 * accumulate_reset != 0: the eval env's `reset_buf |= ...` (env_eval_gennbv.py:327-333) -- flags already set stay set. */
int gnbv_reward_termination(const float* cov_sum, const float* num_valid, float* ratio_prev, int64_t* episode_length,
                            const uint8_t* collision, float* rew_buf, uint8_t* reset_buf, uint8_t* time_out_buf,
                            uint8_t* dones_out, float* episode_sums, float* cur_reward_sum, float* cur_episode_length,
                            double* stats, uint8_t* time_outs_extra, double scale_cov, double scale_short,
                            double scale_term, int has_termination_reward, int only_positive_rewards, int max_step_done,
                            int64_t max_episode_length, double max_episode_length_s, double ratio_threshold,
                            int accumulate_reset, int num_envs, void* stream);

/* reset_idx's buffer resets for rows with reset_buf != 0 (env_train_gennbv.py:395-421): grids zeroed, pose history
 * <- init_pose, frames <- 0, ratio <- 0, actions <- init_action, episode_length <- 0, episode_sums <- 0;
 * clear_reset_buf != 0 then zeroes reset_buf (env_train_gennbv.py:373). */
int gnbv_reset_envs(uint8_t* reset_buf, float* prob_grid, float* scanned_gt, float* pose_hist, float* rgb_hist,
                    float* ratio_prev, int64_t* actions, int64_t* episode_length, float* episode_sums,
                    const float* init_pose, const int64_t* init_action, int num_envs, int grid_size, int hist_len,
                    int pose_dim, int rgb_frames, int rgb_h, int rgb_w, int clear_reset_buf, void* stream);

/* ---- Hybrid_Encoder (gennbv/network/hybrid_encoder.py:12-91) ---- */

/* Device pointers of the encoder parameters, in the reference's state_dict order / shapes:
 *   features_extractor.naive_encoder_grid.0.{weight [16,1,3,3,3], bias [16]}           conv1_w, conv1_b
 *   features_extractor.naive_encoder_grid.1.{weight, bias, running_mean, running_var [16], num_batches_tracked i64}
 *   features_extractor.naive_encoder_grid.3.{weight [16,16,3,3,3], bias [16]}          conv2_w, conv2_b
 *   features_extractor.naive_encoder_grid.4.{...}                                      bn2_*
 *   features_extractor.output_layer_grid.0.{weight [256, 16*G2^3], bias [256]}         grid_fc_w, grid_fc_b
 *   features_extractor.naive_encoder_action.0.{weight [256, 4*state_dim], bias}        act_fc1_w, act_fc1_b
 *   features_extractor.naive_encoder_action.2.{weight [256,256], bias}                 act_fc2_w, act_fc2_b
 *   features_extractor.output_layer.0.{weight [256,512], bias}                         out_fc_w, out_fc_b
 * Running statistics are updated in place by a training-mode forward (momentum 0.1), as nn.BatchNorm3d does. */
typedef struct gnbv_encoder_params {
    const float *conv1_w, *conv1_b, *bn1_w, *bn1_b;
    float *bn1_rm, *bn1_rv;
    int64_t* bn1_nbt;
    const float *conv2_w, *conv2_b, *bn2_w, *bn2_b;
    float *bn2_rm, *bn2_rv;
    int64_t* bn2_nbt;
    const float *grid_fc_w, *grid_fc_b, *act_fc1_w, *act_fc1_b, *act_fc2_w, *act_fc2_b, *out_fc_w, *out_fc_b;
    /* Optional 2-D semantic branch (all NULL = off, the released architecture): Conv2d(2,16,3,s2)+ReLU ->
     * Conv2d(16,16,3,s2)+ReLU -> Flatten -> Linear(3600,256)+ReLU on the k = 2 grayscale 64x64 frames stored behind the grid
     * columns of the observation row; out_fc_w is then [256, 768] (action | grid | semantic).  Parity unpinned: the reference's
     * forward has no such branch (hybrid_encoder.py:69-91) and the paper gives no layer sizes. */
    const float *rgb_conv1_w, *rgb_conv1_b, *rgb_conv2_w, *rgb_conv2_b, *rgb_fc_w, *rgb_fc_b;
} gnbv_encoder_params;

/* Workspace (bytes) for a forward (with_backward = 0) or forward + backward (1) of `batch` rows. */
size_t gnbv_encoder_workspace_bytes(int batch, int grid_size, int state_dim, int with_backward);

/* Hybrid_Encoder.forward (hybrid_encoder.py:69-91), generalised from the hard-coded 20^3 grid to any G:
 *   obs row n = obs + n*obs_row_stride: [0, state_dim) pose history (state_dim = buffer_size*6), then G^3 tri-class
 *   grid values; later columns (the rgb frames) are not read, as in the reference.
 *   row_index (device, [batch] i64, or NULL): sample b reads row row_index[b] -- a PPO minibatch is consumed straight
 *   from the rollout buffer without the gather copy of buffers.py:753-762;
 *   training != 0: BatchNorm3d uses batch statistics and updates the running ones (policies.py:206-214);
 *   features [batch,256] out.  Activations needed by gnbv_encoder_backward stay in `workspace`. */
int gnbv_encoder_forward(const gnbv_encoder_params* params, const float* obs, int64_t obs_row_stride,
                         const int64_t* row_index, int batch, int grid_size, int state_dim, int training,
                         float* features, void* workspace, size_t workspace_bytes, void* stream);

/* Test / debug aid: where a saved activation of the last gnbv_encoder_forward lives inside `workspace` (float offset + count):
 *   GNBV_WS_Y1    conv1 output before BatchNorm1, channels-last [B, G1^3, 16]
 *   GNBV_WS_STAT1 [4][16]: batch (or running) mean, invstd, a = gamma*invstd, b = beta - mean*a of BatchNorm1; every kernel
 *                 forms the ReLU input as fmaf(a, y, b), so (a, b, y) determine the ReLU masks exactly
 *   GNBV_WS_Y2 / GNBV_WS_STAT2 / GNBV_WS_ACT2: the same for layer 2, channel-major [B, 16, G2^3]. */
#define GNBV_WS_Y1 0
#define GNBV_WS_STAT1 1
#define GNBV_WS_Y2 2
#define GNBV_WS_STAT2 3
#define GNBV_WS_ACT2 4
int gnbv_encoder_workspace_view(int batch, int grid_size, int state_dim, int which, int64_t* offset_floats, int64_t* count);

/* Profiling aid of the tcgen05 TS-form conv2 forward kernel (csrc/conv2_ts.cu): with GNBV_TS_DEBUG bit 32 set, CTA 0 accumulates the
 * cycles each warp role spends waiting / working; this copies the 32 counters to the host (synchronises). */
int gnbv_debug_ts_profile(unsigned long long* out32);

/* Gradient destinations, one per trainable tensor of gnbv_encoder_params (same shapes). */
typedef struct gnbv_encoder_grads {
    float *conv1_w, *conv1_b, *bn1_w, *bn1_b, *conv2_w, *conv2_b, *bn2_w, *bn2_b;
    float *grid_fc_w, *grid_fc_b, *act_fc1_w, *act_fc1_b, *act_fc2_w, *act_fc2_b, *out_fc_w, *out_fc_b;
    float *rgb_conv1_w, *rgb_conv1_b, *rgb_conv2_w, *rgb_conv2_b, *rgb_fc_w, *rgb_fc_b;      /* semantic branch, or NULL */
} gnbv_encoder_grads;

/* Backward of gnbv_encoder_forward on the same `workspace` (allocated with with_backward = 1 and untouched since the
 * forward): dfeatures [batch,256] in, every gradient of `grads` overwritten.  `features` is the forward's output
 * (ReLU mask).  No gradient w.r.t. the observation is produced (it is data).  Replaces autograd through
 * hybrid_encoder.py:69-91 inside PPO_Grid_Obs.train (ppo_grid_obs.py:272). */
int gnbv_encoder_backward(const gnbv_encoder_params* params, const float* obs, int64_t obs_row_stride,
                          const int64_t* row_index, int batch, int grid_size, int state_dim, int training,
                          const float* features, const float* dfeatures, const gnbv_encoder_grads* grads,
                          void* workspace, size_t workspace_bytes, void* stream);

/* The same backward in two phases, so that a data-parallel caller can start the all-reduce of the large gradient slice
 * while the convolution backward still runs (the flat gradient arena of gennbv_b200/policy.py is ordered conv tensors
 * first): GNBV_BWD_LINEAR computes the gradients of out_fc, act_fc1/2 and grid_fc (final after this phase) and the
 * gradient entering the conv stack; GNBV_BWD_CONV (must follow, same workspace) computes bn2, conv2, bn1, conv1.
 * gnbv_encoder_backward == phases GNBV_BWD_ALL. */
#define GNBV_BWD_LINEAR 1
#define GNBV_BWD_CONV 2
#define GNBV_BWD_ALL 3
int gnbv_encoder_backward_phase(const gnbv_encoder_params* params, const float* obs, int64_t obs_row_stride,
                                const int64_t* row_index, int batch, int grid_size, int state_dim, int training,
                                const float* features, const float* dfeatures, const gnbv_encoder_grads* grads,
                                void* workspace, size_t workspace_bytes, int phases, void* stream);

/* ---- actor / critic heads + MultiCategorical distribution + PPO loss + optimizer ---- */

/* action_net Linear(256, sum(nvec)) and value_net Linear(256,1) (stable_baselines3/common/policies.py:984,994,
 * 1007-1011) evaluated as one product: head_w [num_out, feat_dim] holds action_net.weight followed by
 * value_net.weight, head_b [num_out] likewise; out [batch, num_out] = logits | value. */
int gnbv_policy_heads_forward(const float* features, const float* head_w, const float* head_b, float* out,
                              int batch, int feat_dim, int num_out, void* stream);

/* MultiCategoricalDistribution (stable_baselines3/common/distributions.py:299-352) over logits rows split by
 * nvec (host array, num_sub entries):
 *   evaluate : log_prob [B] = sum_k log softmax_k[a_k], entropy [B] = sum_k H_k  (either output may be NULL)
 *   sample   : actions [B,num_sub] i64 ~ Categorical (Gumbel-max, Philox(seed, offset)) or the mode if deterministic;
 *              log_prob [B] of the drawn actions (may be NULL)
 *   backward : dlogits [B, sum(nvec)] from per-row gradients w.r.t. log_prob and entropy. */
int gnbv_multicategorical_evaluate(const float* logits, int64_t logits_row_stride, const int* nvec, int num_sub,
                                   const int64_t* actions, float* log_prob, float* entropy, int batch, void* stream);
int gnbv_multicategorical_sample(const float* logits, int64_t logits_row_stride, const int* nvec, int num_sub,
                                 uint64_t seed, uint64_t offset, int deterministic, int64_t* actions, float* log_prob,
                                 int batch, void* stream);
int gnbv_multicategorical_backward(const float* logits, int64_t logits_row_stride, const int* nvec, int num_sub,
                                   const int64_t* actions, const float* grad_log_prob, const float* grad_entropy,
                                   float* dlogits, int64_t dlogits_row_stride, int batch, void* stream);

/* The loss lines of PPO_Grid_Obs.train (stable_baselines3/ppo/ppo_grid_obs.py:213-262) on one minibatch, forward and
 * backward in one launch: advantage normalisation (unbiased std + 1e-8), clipped surrogate, clipped value loss
 * (clip_range_vf < 0 disables the clipping), entropy bonus, loss = pg_coef*pg + ent_coef*ent + vf_coef*vf
 * (pg_coef = 10 in the reference, :253), approx_kl and clip_fraction.
 *   scalars [8] out: loss, policy_loss, value_loss, entropy_loss, approx_kl, clip_fraction, adv_mean, adv_std
 *   grad_log_prob / grad_entropy / grad_values [B] out: d loss / d (log_prob, entropy, values). */
int gnbv_ppo_loss(const float* log_prob, const float* entropy, const float* values, const float* old_values,
                  const float* old_log_prob, const float* advantages, const float* returns, int batch,
                  double clip_range, double clip_range_vf, double ent_coef, double vf_coef, double pg_coef,
                  int normalize_advantage, float* scalars, float* grad_log_prob, float* grad_entropy,
                  float* grad_values, void* stream);

/* th.nn.utils.clip_grad_norm_ + th.optim.Adam step (ppo_grid_obs.py:271-275; Adam eps 1e-5, policies.py:855) on flat
 * fp32 buffers.  gnbv_grad_norm writes workspace[0] = ||g||_2 and workspace[1] = min(1, max_norm / (norm + 1e-6));
 * gnbv_adam_step multiplies the gradient by workspace[1] (if clip_workspace != NULL) and by grad_scale. */
size_t gnbv_clip_adam_workspace_bytes(void);
int gnbv_grad_norm(const float* grads, int64_t n, double max_norm, float* workspace, void* stream);
int gnbv_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                   const float* clip_workspace, double lr, double beta1, double beta2, double eps, int64_t step,
                   double grad_scale, void* stream);

/* ---- one PPO minibatch update as a fixed launch sequence (CUDA-graph capturable, no host decision inside) ----
 * Replaces the body of the minibatch loop of PPO_Grid_Obs.train (stable_baselines3/ppo/ppo_grid_obs.py:201-275) together
 * with RolloutBuffer._get_samples (stable_baselines3/common/buffers.py:753-762):
 *   gnbv_ppo_minibatch_grads : gather the minibatch rows (device cursor ctl[0] into storage_rows), evaluate_actions in
 *       training mode, loss forward + backward, all parameter gradients (enc_grads, head_*_grad), scalars, KL vote;
 *       `phases` as in gnbv_encoder_backward_phase (GNBV_BWD_LINEAR also does everything up to the loss);
 *   gnbv_ppo_minibatch_apply : sticky KL-stop decision from the (all-reduced) vote, log row, clip_grad_norm_ + Adam on the
 *       flat arenas, cursor++.  While the stop flag ctl[1] is set nothing changes state (Adam, BN running statistics, log).
 * ctl (device int64[8], zeroed by the caller at the start of an epoch except ctl[2] and ctl[3] = -1 at the start of train()):
 *   [0] minibatch cursor of the epoch, [1] stop flag, [2] Adam step count, [3] cursor at which the stop was raised,
 *   [4] number of log rows written.
 * vote: device float the caller places right behind the flat gradient bucket so that ONE all-reduce(sum) carries gradients
 *   and the stop decision of all ranks (1.0 = this rank's approx_kl > 1.5 target_kl; target_kl < 0 disables the test).
 * log: [log_capacity, 8] floats, row k = scalars of the k-th executed minibatch (layout of gnbv_ppo_loss). */
typedef struct gnbv_ppo_minibatch {
    const gnbv_encoder_params* enc;          /* host struct of device pointers */
    const gnbv_encoder_grads* enc_grads;
    const float *head_w, *head_b;            /* [A+1, feat_dim], [A+1]: action_net | value_net */
    float *head_w_grad, *head_b_grad;
    const int* nvec;                         /* host */
    int num_sub, feat_dim;
    const float* observations;               /* rollout storage, row = t*N + n */
    int64_t obs_row_stride;
    const float *actions, *values, *log_probs, *advantages, *returns;   /* [rows, num_sub] f32, [rows] f32 x4 */
    const int64_t* storage_rows;             /* the rollout's permutation mapped to storage rows */
    int64_t rows_base;                       /* minibatch k reads storage_rows[rows_base + k*batch .. + batch), k = ctl[0] */
    int batch, grid_size, state_dim, normalize_advantage;
    double clip_range, clip_range_vf, ent_coef, vf_coef, pg_coef, target_kl;
    int64_t* ctl;
    float* vote;
    float* log;
    int64_t log_capacity;
    void* enc_workspace;                     /* gnbv_encoder_workspace_bytes(batch, grid, state, 1) */
    size_t enc_workspace_bytes;
    void* mb_workspace;                      /* gnbv_ppo_minibatch_workspace_bytes(...) */
    size_t mb_workspace_bytes;
} gnbv_ppo_minibatch;
size_t gnbv_ppo_minibatch_workspace_bytes(int batch, int num_logits, int num_sub, int feat_dim);
size_t gnbv_ppo_apply_workspace_bytes(void);
int gnbv_ppo_minibatch_grads(const gnbv_ppo_minibatch* args, int phases, void* stream);
int gnbv_ppo_minibatch_scalars(const gnbv_ppo_minibatch* args, const float** scalars);   /* device pointer to float[8] */
int gnbv_ppo_minibatch_apply(const gnbv_ppo_minibatch* args, float* params, const float* grads, float* exp_avg,
                             float* exp_avg_sq, int64_t n, double max_grad_norm, double lr, double beta1, double beta2,
                             double eps, double grad_scale, float* apply_workspace, void* stream);

/* Bare fp32 GEMM on device pointers: C[M,N] = relu?(A*B + bias), element strides (see gennbv_b200/csrc/gemm.cuh);
 * replaces the cuBLAS calls behind nn.Linear (hybrid_encoder.py:39-54; policies.py:984,994). */
size_t gnbv_sgemm_workspace_bytes(int M, int N, int K);
int gnbv_sgemm(const float* A, int64_t sa_m, int64_t sa_k, const float* B, int64_t sb_k, int64_t sb_n, float* C,
               int64_t ldc, int M, int N, int K, const float* bias, int relu, float* workspace, size_t workspace_bytes,
               void* stream);

/* Same contract as gnbv_sgemm on the tcgen05 tensor cores with split-precision (3xTF32) operands: fp32-grade accuracy
 * (~1e-6 relative) at tensor-core rate.  workspace: gnbv_tc_gemm_workspace_bytes(M,N,K) bytes, zero-initialised once
 * (float slot 0 is a sticky error flag). */
size_t gnbv_tc_gemm_workspace_bytes(int M, int N, int K);
int gnbv_tc_gemm(const float* A, int64_t sa_m, int64_t sa_k, const float* B, int64_t sb_k, int64_t sb_n, float* C,
                 int64_t ldc, int M, int N, int K, const float* bias, int relu, float* workspace, size_t workspace_bytes,
                 void* stream);

/* ---- eval accuracy ("next" row): exact bidirectional 1-NN / chamfer ----
 * Replaces pytorch3d.loss.chamfer_distance as called by gennbv/env/env_eval_gennbv.py:253-261 (pytorch3d is a third-party
 * dependency, "0.7.8 works" per the reference README, not vendored; defaults: squared L2, mean over points, both
 * directions).  Clouds are packed: x [sum P1_e, 3] f32 with offsets [E+1] i64, same for y.
 *   cham_x[e] = mean_i min_j |x_i - y_j|^2,  cham_y[e] = mean_j min_i |x_i - y_j|^2;  the loss is their sum. */
size_t gnbv_chamfer_workspace_bytes(int num_clouds);
int gnbv_chamfer(const float* x, const int64_t* x_offsets, const float* y, const int64_t* y_offsets, int num_clouds,
                 float* cham_x, float* cham_y, float* workspace, size_t workspace_bytes, void* stream);
/* Same result through an exact uniform-grid nearest-neighbour search (per direction: bounding box, per-cell counts,
 * exclusive scan, counting-sort fill, shell search -- gennbv_b200/csrc/nn_grid.cuh) instead of the P1*P2 scan.
 *   total_x / total_y = x_offsets[E] / y_offsets[E] (known to the host; they size the workspace);
 *   cells_per_axis in [1,160]: cells along the longest bounding-box axis of every cloud (cubic cells);
 *   min_x [total_x] / min_y [total_y] f32: optional per-point minima (NULL to skip); equal to the brute-force kernel's
 *   bit for bit.  The workspace must be 256-byte aligned. */
size_t gnbv_chamfer_grid_workspace_bytes(int num_clouds, int64_t total_x, int64_t total_y, int cells_per_axis);
int gnbv_chamfer_grid(const float* x, const int64_t* x_offsets, const float* y, const int64_t* y_offsets, int num_clouds,
                      int64_t total_x, int64_t total_y, int cells_per_axis, float* cham_x, float* cham_y,
                      float* min_x, float* min_y, void* workspace, size_t workspace_bytes, void* stream);
/* Per-point minima min_j |q_i - r_j|^2 of the brute-force scan (cross-check entry; workspace as gnbv_chamfer). */
int gnbv_nn_sqdist_brute(const float* q, const int64_t* q_offsets, const float* r, const int64_t* r_offsets, int num_clouds,
                         float* min_out, float* workspace, size_t workspace_bytes, void* stream);

/* TensorRolloutBuffer_Grid_Obs.compute_returns_and_advantage  (stable_baselines3/common/buffers.py:706-724)
 *   rewards, values [T,N] f32; episode_starts [T,N] u8; last_values [N] f32; dones [N] u8
 *   advantages, returns [T,N] f32 out.  gamma / gae_lambda are the Python doubles of the buffer. */
int gnbv_gae(const float* rewards, const float* values, const uint8_t* episode_starts,
             const float* last_values, const uint8_t* dones, double gamma, double gae_lambda,
             int n_steps, int num_envs, float* advantages, float* returns, void* stream);

/* ---- eval env: scanned-point history (gennbv/env/env_eval_gennbv.py:160-164, 253-257) ----
 * gnbv_scan_points: back_projection_fg (env_train_gennbv.py:494-526) for every foreground pixel (seg > 50), same fp32
 * chain as the voxelize path, appended to env n's history as the packed 1 cm lattice key
 *   key = (kx + 2^17) << 36 | (ky + 2^17) << 18 | (kz + 2^17),  k = nearbyint(p * 100.f)   (== torch.round(p, decimals=2) * 100),
 *   |k| clamped to 2^17 - 1 (1.31 km); bits 54..62 are free for a caller-side env index (gnbv_keys_to_points ignores them)
 * keys [N, capacity] i64, counts [N] i32 in/out (number of keys held), overflow [1] i32 set to 1 if a history is full.
 * Ascending key order is the row order of torch.unique(dim=0) on the rounded points.  flags: GNBV_RAW_DEPTH as above.
 * gnbv_keys_to_points: keys [n] -> points [n,3] f32 = k / 100.f, the rows torch.round(decimals=2) produces. */
int gnbv_scan_points(const float* depth, const int32_t* seg, const float* kinv, const float* c2w, int64_t* keys,
                     int32_t* counts, int32_t* overflow, int num_envs, int height, int width, int64_t capacity,
                     uint32_t flags, void* stream);
int gnbv_keys_to_points(const int64_t* keys, int64_t num_keys, float* points, void* stream);
/* Inverse direction for callers that hold float points (`torch.round(pts, decimals=2)`, env_eval_gennbv.py:254): packed keys. */
int gnbv_points_to_keys(const float* points, int64_t num_points, int64_t* keys, void* stream);
/* Histories [num_envs, capacity] of the listed envs -> one array: env_rows[r]'s first offsets[r+1]-offsets[r] keys at
 * out + offsets[r], each OR-ed with r << tag_shift (tag_shift >= 54), so that ONE sort de-duplicates every env. */
int gnbv_pack_env_keys(const int64_t* keys, int64_t capacity, const int64_t* env_rows, const int64_t* offsets, int num_rows,
                       int tag_shift, int64_t* out, void* stream);

/* `torch.unique` on 64-bit keys (env_eval_gennbv.py:254-257 after the key packing), in-tree: stable LSD radix sort over the
 * low `key_bits` bits (8 bits per pass) followed by a compaction of the run heads.  `keys` [n] is used as a sort buffer
 * (destroyed); unique_out [>= n] receives the distinct keys in ascending order, *count_out (device) their number. */
size_t gnbv_sort_unique_workspace_bytes(int64_t n);
int gnbv_sort_unique_u64(uint64_t* keys, int64_t n, int key_bits, uint64_t* unique_out, int64_t* count_out, void* workspace,
                         size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GENNBV_B200_H */
