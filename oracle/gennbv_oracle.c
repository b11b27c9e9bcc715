/* gennbv_oracle.c -- CPU restatement of GenNBV's per-environment state-encoding path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it.  The product package (gennbv_b200) never links, imports or calls it and
 * fails loudly when its CUDA library is missing.
 *
 * Every function cites the reference lines (zjwzcx/GenNBV @ c373f76) it follows.
 * Parity status: PINNED in the build container against the reference's own Python
 * (imported through oracle/ref_loader.py) -- see oracle/gen_golden.py and
 * tests/test_oracle_vs_reference.py; the PyCUDA Bresenham kernel cannot execute on
 * CPU, it is restated from its kernel text (gennbv/utils.py:43-197) and pinned by
 * hand-derived known-answer rays (tests/test_oracle_kat.py).
 *
 * Floating point: the reference computes the back-projection with three torch
 * einsums that lower to CPU GEMMs.  Measured here (oracle/gen_golden.py, 100 % of
 * 3.1 M points): each output is the sequential chain
 *      acc = fl(a0*b0); acc = fma(a1,b1,acc); acc = fma(a2,b2,acc); ...
 * in increasing k.  This file states exactly that with fmaf(); compile with
 * -ffp-contract=off so nothing else is fused.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

/* env_train_base.py:520-523 : nan_to_num(neginf=0) -> clamp(min=-50) -> abs */
float gnbv_oracle_depth_post(float d) {
    if (isnan(d)) d = 0.0f;
    else if (isinf(d)) d = d > 0 ? FLT_MAX : 0.0f;
    if (d < -50.0f) d = -50.0f;
    return fabsf(d);
}

void gnbv_oracle_post_process_depth(const float* raw, float* out, int64_t n) {
    for (int64_t i = 0; i < n; ++i) out[i] = gnbv_oracle_depth_post(raw[i]);
}

/* env_train_gennbv.py:503-526 : world coordinate of pixel p of one env.
 * depth is the processed (non-negative) depth; (u,v) integer pixel coordinates (:172-181). */
static void back_project_one(float d, float u, float v, const float* kinv, const float* c2w, float* w) {
    float px = d * u, py = d * v, pz = d * 1.0f;              /* einsum 'ij,jk->ijk' : outer product */
    float cam[4];
    for (int r = 0; r < 3; ++r) {                             /* einsum 'ij,nkj->nki' */
        float acc = kinv[r * 3 + 0] * px;
        acc = fmaf(kinv[r * 3 + 1], py, acc);
        acc = fmaf(kinv[r * 3 + 2], pz, acc);
        cam[r] = acc;
    }
    cam[3] = 1.0f;
    for (int i = 0; i < 3; ++i) {                             /* einsum 'nij,nkj->nki' */
        float acc = c2w[i * 4 + 0] * cam[0];
        acc = fmaf(c2w[i * 4 + 1], cam[1], acc);
        acc = fmaf(c2w[i * 4 + 2], cam[2], acc);
        acc = fmaf(c2w[i * 4 + 3], cam[3], acc);
        w[i] = acc;
    }
}

/* Dense version of back_projection_fg for one batch: world [N,P,3]; fg [N,P] (seg > 50, :504).
 * depth[~fg] = 0 (:509) is applied, as in the reference, before the products. */
void gnbv_oracle_back_projection(const float* depth, const int32_t* seg, const float* kinv,
                                 const float* c2w, int N, int H, int W, float* world, uint8_t* fg) {
    int P = H * W;
    for (int n = 0; n < N; ++n)
        for (int p = 0; p < P; ++p) {
            int64_t i = (int64_t)n * P + p;
            uint8_t f = seg[i] > 50;
            float d = f ? depth[i] : 0.0f;
            back_project_one(d, (float)(p % W), (float)(p / W), kinv, c2w + n * 16, world + i * 3);
            fg[i] = f;
        }
}

/* gennbv/utils.py:242-243 */
static void voxel_bounds(const float* range6, const float* vs3, float* lo, float* hi) {
    for (int a = 0; a < 3; ++a) {
        hi[a] = range6[2 * a] + 0.5f * vs3[a];
        lo[a] = range6[2 * a + 1] - 0.5f * vs3[a];
    }
}

/* gennbv/utils.py:251-267 for one point: returns 1 and the clamped index if strictly inside. */
static int point_to_idx(const float* w, const float* lo, const float* hi, const float* vs, int G, int* idx) {
    for (int a = 0; a < 3; ++a)
        if (!(hi[a] > w[a] && w[a] > lo[a])) return 0;
    for (int a = 0; a < 3; ++a) {
        long v = (long)floorf((w[a] - lo[a]) / vs[a]);
        if (v < 0) v = 0;
        if (v > G - 1) v = G - 1;
        idx[a] = (int)v;
    }
    return 1;
}

/* gennbv/utils.py:273-306 (if_col=False): unclamped camera voxel. */
void gnbv_oracle_pose_to_idx(const float* pose_xyz, const float* range6, const float* vs3, int N, int64_t* out) {
    for (int n = 0; n < N; ++n) {
        float lo[3], hi[3];
        voxel_bounds(range6 + n * 6, vs3 + n * 3, lo, hi);
        for (int a = 0; a < 3; ++a)
            out[n * 3 + a] = (int64_t)floorf((pose_xyz[n * 3 + a] - lo[a]) / vs3[n * 3 + a]);
    }
}

/* Emit callback style Bresenham -- follows the kernel text gennbv/utils.py:48-167:
 * dominant axis tie-break dx -> dy -> dz; start voxel emitted if in bounds; every step
 * emitted if in bounds; loop also stops when `idx` (number emitted) reaches max_pts. */
typedef void (*emit_fn)(void* ctx, int x, int y, int z);

static int in_map(int x, int y, int z, int G) {
    return x >= 0 && x < G && y >= 0 && y < G && z >= 0 && z < G;
}

static int bresenham_line(int x0, int y0, int z0, int x1, int y1, int z1, int G, int max_pts,
                          emit_fn emit, void* ctx) {
    int dx = abs(x1 - x0), dy = abs(y1 - y0), dz = abs(z1 - z0);
    int sx = x0 < x1 ? 1 : -1, sy = y0 < y1 ? 1 : -1, sz = z0 < z1 ? 1 : -1;
    int dm = dx > dy ? dx : dy;
    if (dz > dm) dm = dz;
    int x = x0, y = y0, z = z0, idx = 0;
    if (in_map(x, y, z, G)) { emit(ctx, x, y, z); idx++; }
    if (dm == dx) {
        int p1 = 2 * dy - dx, p2 = 2 * dz - dx;
        for (int i = 0; i < dx && idx < max_pts; ++i) {
            if (p1 >= 0) { y += sy; p1 -= 2 * dx; }
            if (p2 >= 0) { z += sz; p2 -= 2 * dx; }
            x += sx; p1 += 2 * dy; p2 += 2 * dz;
            if (in_map(x, y, z, G)) { emit(ctx, x, y, z); idx++; }
        }
    } else if (dm == dy) {
        int p1 = 2 * dx - dy, p2 = 2 * dz - dy;
        for (int i = 0; i < dy && idx < max_pts; ++i) {
            if (p1 >= 0) { x += sx; p1 -= 2 * dy; }
            if (p2 >= 0) { z += sz; p2 -= 2 * dy; }
            y += sy; p1 += 2 * dx; p2 += 2 * dz;
            if (in_map(x, y, z, G)) { emit(ctx, x, y, z); idx++; }
        }
    } else {
        int p1 = 2 * dx - dz, p2 = 2 * dy - dz;
        for (int i = 0; i < dz && idx < max_pts; ++i) {
            if (p1 >= 0) { x += sx; p1 -= 2 * dz; }
            if (p2 >= 0) { y += sy; p2 -= 2 * dz; }
            z += sz; p1 += 2 * dx; p2 += 2 * dy;
            if (in_map(x, y, z, G)) { emit(ctx, x, y, z); idx++; }
        }
    }
    return idx;
}

typedef struct { int64_t* out; int64_t n; int64_t cap; } list_ctx;
static void emit_list(void* c, int x, int y, int z) {
    list_ctx* l = (list_ctx*)c;
    if (l->n < l->cap) { l->out[l->n * 3] = x; l->out[l->n * 3 + 1] = y; l->out[l->n * 3 + 2] = z; }
    l->n++;
}

/* bresenham3D_pycuda (gennbv/utils.py:24-227): concatenated in-bounds voxels of every ray, in
 * ray order, duplicates kept.  Returns the number of rows (call with cap = 0 to size). */
int64_t gnbv_oracle_bresenham3d(const int64_t* src, const int64_t* tgt, int64_t num_rays, int G,
                                int64_t* out, int64_t cap) {
    list_ctx l = { out, 0, cap };
    for (int64_t r = 0; r < num_rays; ++r)
        bresenham_line((int)src[0], (int)src[1], (int)src[2], (int)tgt[r * 3], (int)tgt[r * 3 + 1],
                       (int)tgt[r * 3 + 2], G, 3 * G, emit_list, &l);
    return l.n;
}

typedef struct { uint8_t* m; int G; } mark_ctx;
static void emit_mark(void* c, int x, int y, int z) {
    mark_ctx* k = (mark_ctx*)c;
    k->m[((int64_t)x * k->G + y) * k->G + z] = 1;
}

/* One env.step() worth of state encoding for N envs on dense fp32 grids:
 *   back_projection_fg      env_train_gennbv.py:494-533
 *   scanned_pts_to_idx_3D   gennbv/utils.py:230-270
 *   pose_coord_to_idx_3D    gennbv/utils.py:273-306
 *   update_occ_grid         env_train_gennbv.py:277-326   (incl. bresenham3D_pycuda)
 *   grid_occupancy_tri_cls  gennbv/utils.py:309-325
 *   coverage count          env_train_gennbv.py:537 (the sum; the division is done by the caller)
 * In/out: prob_grid, scanned_gt [N,G^3] f32.  Out: tri [N,G^3] f32, cov_sum [N] f32,
 * num_targets [N] i32, optional target/touched masks [N,G^3] u8 (may be NULL).
 * raw_depth != 0 applies post_process_camera_tensor's depth chain first. */
void gnbv_oracle_voxelize_step(const float* depth, const int32_t* seg, const float* kinv,
                               const float* c2w, const float* range6, const float* vs3,
                               const float* pose_xyz, const float* grid_gt, float* prob_grid,
                               float* scanned_gt, float* tri, float* cov_sum, int32_t* num_targets,
                               uint8_t* target_mask_out, uint8_t* touched_mask_out,
                               int N, int H, int W, int G, int raw_depth) {
    int P = H * W;
    int64_t V = (int64_t)G * G * G;
    uint8_t* tmask = (uint8_t*)malloc(V);
    uint8_t* rmask = (uint8_t*)malloc(V);
    for (int n = 0; n < N; ++n) {
        float lo[3], hi[3];
        const float* vs = vs3 + n * 3;
        voxel_bounds(range6 + n * 6, vs, lo, hi);
        memset(tmask, 0, V);
        memset(rmask, 0, V);
        int nt = 0;
        for (int p = 0; p < P; ++p) {
            int64_t i = (int64_t)n * P + p;
            if (!(seg[i] > 50)) continue;                        /* only fg points are gathered (:527) */
            float d = raw_depth ? gnbv_oracle_depth_post(depth[i]) : depth[i];
            float w[3];
            int idx[3];
            back_project_one(d, (float)(p % W), (float)(p / W), kinv, c2w + n * 16, w);
            if (!point_to_idx(w, lo, hi, vs, G, idx)) continue;
            int64_t lin = ((int64_t)idx[0] * G + idx[1]) * G + idx[2];
            if (!tmask[lin]) { tmask[lin] = 1; nt++; }             /* unique (:266, :301) */
        }
        num_targets[n] = nt;
        float* prob = prob_grid + n * V;
        float* scan = scanned_gt + n * V;
        const float* gt = grid_gt + n * V;
        if (nt > 0) {                                              /* `continue` at :298-299 */
            int src[3];
            for (int a = 0; a < 3; ++a)
                src[a] = (int)(int64_t)floorf((pose_xyz[n * 3 + a] - lo[a]) / vs[a]);
            mark_ctx mk = { rmask, G };
            for (int x = 0; x < G; ++x)
                for (int y = 0; y < G; ++y)
                    for (int z = 0; z < G; ++z)
                        if (tmask[((int64_t)x * G + y) * G + z])
                            bresenham_line(src[0], src[1], src[2], x, y, z, G, 3 * G, emit_mark, &mk);
            /* :311-314 : `prob[path] -= 0.05` is a non-accumulating indexed RMW (each distinct
             * voxel once), then every target is overwritten with 1.0 */
            for (int64_t v = 0; v < V; ++v) {
                if (rmask[v]) prob[v] = prob[v] - 0.05f;
                if (tmask[v]) prob[v] = 1.0f;
            }
        }
        float s = 0.0f;
        for (int64_t v = 0; v < V; ++v) {
            float pr = prob[v];
            tri[n * V + v] = (pr > 0.5f ? 1.0f : 0.0f) - (pr < 0.0f ? 1.0f : 0.0f);   /* utils.py:318-321 */
            float occ = tmask[v] ? 1.0f : 0.0f;
            float sg = scan[v] + occ * gt[v];                                           /* :323-326 */
            sg = sg < 0.0f ? 0.0f : (sg > 1.0f ? 1.0f : sg);
            scan[v] = sg;
            s += sg;                      /* exact while the values are {0,1} and the sum < 2^24 */
        }
        cov_sum[n] = s;
        if (target_mask_out) memcpy(target_mask_out + n * V, tmask, V);
        if (touched_mask_out) memcpy(touched_mask_out + n * V, rmask, V);
    }
    free(tmask);
    free(rmask);
}

/* TensorRolloutBuffer_Grid_Obs.compute_returns_and_advantage (buffers.py:706-724), fp32. */
void gnbv_oracle_gae(const float* rewards, const float* values, const uint8_t* episode_starts,
                     const float* last_values, const uint8_t* dones, double gamma_d, double lam_d,
                     int T, int N, float* advantages, float* returns) {
    /* `self.gamma * tensor` casts the python double to fp32; `self.gamma * self.gae_lambda` is a
     * python double product that is cast to fp32 only when it meets the tensor (buffers.py:720-721) */
    const float gamma = (float)gamma_d, gl = (float)(gamma_d * lam_d);
    for (int n = 0; n < N; ++n) {
        float last = 0.0f;
        for (int t = T - 1; t >= 0; --t) {
            float nnt, nv;
            if (t == T - 1) { nnt = 1.0f - (float)(dones[n] != 0); nv = last_values[n]; }
            else { nnt = 1.0f - (float)episode_starts[(t + 1) * N + n]; nv = values[(t + 1) * N + n]; }
            /* torch evaluates left to right with one rounding per op */
            float delta = (rewards[t * N + n] + (gamma * nv) * nnt) - values[t * N + n];
            last = delta + (gl * nnt) * last;
            advantages[t * N + n] = last;
            returns[t * N + n] = last + values[t * N + n];
        }
    }
}
