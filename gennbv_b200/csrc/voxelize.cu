// voxelize.cu -- depth -> point cloud -> occupancy voxels -> ray-cast free space -> prob / tri-class
// grids + coverage sum, for all environments per launch (sm_100a).
//
// Replaces the per-env Python loops of Env_Train_GenNBV.update_occ_grid
// (gennbv/env/env_train_gennbv.py:277-326) and the helpers it calls; see include/gennbv_b200.h.
//
// Three launches per step, all on the caller's stream, no host synchronisation:
//   K1 scan_raycast_kernel : one CTA per env.  Occupancy bit-masks live in shared memory
//        (G^3 bits each: 32 KB at G = 64).  Phase 1 unprojects every foreground pixel with the
//        reference's exact fp32 operation order and ORs its voxel bit (shared-memory atomics,
//        warp-level dedup of equal words).  Phase 2 compacts the distinct target voxels and walks
//        the reference's integer 3-D Bresenham for each one, entering the line analytically at the
//        first in-grid step, ORing the "touched" mask.  Phase 3 stores both masks (64 KB/env) for K2.
//   K2 grid_update_kernel  : dense, HBM-bound pass over [N, G^3]: prob -= 0.05 on touched,
//        prob = 1 on targets, tri-class, scanned_gt |= target & gt, per-chunk coverage partial sums.
//        128-bit streaming loads/stores, 12 independent 16 B loads in flight per thread.
//   K3 coverage_finalize_kernel : fixed-order sum of the partials (deterministic).
//
// Compiled with -fmad=false: every fp32 operation below is rounded exactly as written
// (the fused steps of the reference's einsum chain are explicit fmaf calls).
#include "common.cuh"

#include <float.h>

#include <algorithm>

namespace gnbv {

constexpr int K1_THREADS = 1024;
constexpr int K1_WARPS = K1_THREADS / 32;
constexpr int LIST_CAP = 8192;           // compacted targets per ray-cast round (32 KB)
constexpr int K2_THREADS = 256;
constexpr int K2_UNROLL = 4;             // float4 items per thread
constexpr int K2_CHUNK = K2_THREADS * K2_UNROLL * 4;   // voxels per CTA (4096)

struct VoxelizeLayout {
    int64_t words;       // u32 words per env per mask (multiple of 4)
    int64_t chunks;      // K2 CTAs per env
    size_t off_tmask, off_rmask, off_partial, total;
};

static VoxelizeLayout make_layout(int N, int G) {
    VoxelizeLayout L;
    int64_t V = (int64_t)G * G * G;
    L.words = ceil_div(ceil_div(V, 32), 4) * 4;
    L.chunks = ceil_div(V, K2_CHUNK);
    L.off_tmask = 0;
    L.off_rmask = L.off_tmask + (size_t)N * L.words * 4;
    L.off_partial = L.off_rmask + (size_t)N * L.words * 4;
    L.total = L.off_partial + (size_t)N * L.chunks * 4;
    L.total = (L.total + 255) & ~(size_t)255;
    return L;
}

// env_train_base.py:520-523 : nan_to_num(neginf=0) -> clamp(min=-50) -> abs
__device__ __forceinline__ float depth_post(float d) {
    if (d != d) d = 0.0f;
    else if (d == -INFINITY) d = 0.0f;
    else if (d == INFINITY) d = FLT_MAX;
    d = fmaxf(d, -50.0f);
    return fabsf(d);
}

struct EnvGeom {
    float kinv[9];
    float c2w[12];
    float lo[3], hi[3], vs[3];
};

// env_train_gennbv.py:519-526 + gennbv/utils.py:251-267 for one pixel; returns the linear voxel
// index or -1.  Operation order == oracle/gennbv_oracle.c::back_project_one (sequential fma chain).
__device__ __forceinline__ int pixel_to_voxel(const EnvGeom& g, float d, float u, float v, int G) {
    float px = __fmul_rn(d, u), py = __fmul_rn(d, v), pz = d;
    float cam[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        float acc = __fmul_rn(g.kinv[r * 3 + 0], px);
        acc = __fmaf_rn(g.kinv[r * 3 + 1], py, acc);
        acc = __fmaf_rn(g.kinv[r * 3 + 2], pz, acc);
        cam[r] = acc;
    }
    int idx[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        float acc = __fmul_rn(g.c2w[i * 4 + 0], cam[0]);
        acc = __fmaf_rn(g.c2w[i * 4 + 1], cam[1], acc);
        acc = __fmaf_rn(g.c2w[i * 4 + 2], cam[2], acc);
        acc = __fmaf_rn(g.c2w[i * 4 + 3], 1.0f, acc);
        if (!(g.hi[i] > acc && acc > g.lo[i])) return -1;          // strict, on the float point
        float q = floorf(__fdiv_rn(__fsub_rn(acc, g.lo[i]), g.vs[i]));
        int k = (int)q;
        k = max(0, min(G - 1, k));                                  // clamp after the mask (utils.py:267)
        idx[i] = k;
    }
    return (idx[0] * G + idx[1]) * G + idx[2];
}

// number of secondary-axis increments after n Bresenham iterations (see raycast())
__device__ __forceinline__ int bres_steps(int d_minor, int d_major, int n) {
    const long long num = 2LL * d_minor * n + d_major;
    if (num < 0x7fffffffLL && d_major < 0x3fffffff) return (int)((unsigned)num / (unsigned)(2 * d_major));   // the usual case: 32-bit divide
    return (int)(num / (2LL * d_major));
}

// gennbv/utils.py:48-167 for one ray, marking the touched mask instead of writing a trajectory.
// The reference walks i = 0..d_major from the (usually out-of-grid) camera voxel and keeps only the
// in-bounds voxels.  After n iterations the state is closed-form,
//     major = m0 + n*s,   minor_k = k0 + s_k * floor((2*d_k*n + d_major) / (2*d_major)),
// (p_k starts at 2*d_k - d_major, gains 2*d_k per iteration and loses 2*d_major per minor step), so
// the walk starts at the first n whose major coordinate is inside [0,G) and stops at the last one.
// The per-ray cap of 3G emitted voxels (utils.py:37) can never bind: emitted voxels have distinct
// major coordinates, hence at most G of them.
__device__ __forceinline__ void raycast(uint32_t* rmask, int G, int x0, int y0, int z0, int x1, int y1, int z1) {
    int dx = abs(x1 - x0), dy = abs(y1 - y0), dz = abs(z1 - z0);
    int sx = x0 < x1 ? 1 : -1, sy = y0 < y1 ? 1 : -1, sz = z0 < z1 ? 1 : -1;
    int dm = max(dx, max(dy, dz));
    // permute so that "a" is the dominant axis with the reference's tie-break dx -> dy -> dz
    int a0, b0, c0, da, db, dc, sa, sb, sc, stride_a, stride_b, stride_c;
    if (dm == dx)      { a0 = x0; b0 = y0; c0 = z0; da = dx; db = dy; dc = dz; sa = sx; sb = sy; sc = sz; stride_a = G * G; stride_b = G; stride_c = 1; }
    else if (dm == dy) { a0 = y0; b0 = x0; c0 = z0; da = dy; db = dx; dc = dz; sa = sy; sb = sx; sc = sz; stride_a = G; stride_b = G * G; stride_c = 1; }
    else               { a0 = z0; b0 = x0; c0 = y0; da = dz; db = dx; dc = dy; sa = sz; sb = sx; sc = sy; stride_a = 1; stride_b = G * G; stride_c = G; }
    // iterations n in [0, da] with 0 <= a0 + n*sa < G
    int n_lo, n_hi;
    if (sa > 0) { n_lo = max(0, -a0); n_hi = min(da, G - 1 - a0); }
    else        { n_lo = max(0, a0 - (G - 1)); n_hi = min(da, a0); }
    if (n_lo > n_hi) return;
    int a = a0 + n_lo * sa, b = b0, c = c0;
    int p1 = 2 * db - da, p2 = 2 * dc - da;
    if (n_lo > 0) {
        int kb = bres_steps(db, da, n_lo), kc = bres_steps(dc, da, n_lo);
        b += sb * kb;
        c += sc * kc;
        p1 = (int)(2LL * db - da + 2LL * n_lo * db - 2LL * kb * da);   // result is within (-2*da, 2*db]
        p2 = (int)(2LL * dc - da + 2LL * n_lo * dc - 2LL * kc * da);
    }
    for (int n = n_lo;; ++n) {
        if ((unsigned)b < (unsigned)G && (unsigned)c < (unsigned)G) {
            int lin = a * stride_a + b * stride_b + c * stride_c;
            uint32_t bit = 1u << (lin & 31);
            if (!(rmask[lin >> 5] & bit)) atomicOr(&rmask[lin >> 5], bit);
        }
        if (n == n_hi) break;
        if (p1 >= 0) { b += sb; p1 -= 2 * da; }
        if (p2 >= 0) { c += sc; p2 -= 2 * da; }
        a += sa;
        p1 += 2 * db;
        p2 += 2 * dc;
    }
}

__global__ void __launch_bounds__(K1_THREADS, 1)
scan_raycast_kernel(const float* __restrict__ depth, const int32_t* __restrict__ seg,
                    const float* __restrict__ kinv, const float* __restrict__ c2w,
                    const float* __restrict__ range_gt, const float* __restrict__ voxel_size,
                    const float* __restrict__ pose_xyz, uint32_t* __restrict__ tmask_out,
                    uint32_t* __restrict__ rmask_out, int32_t* __restrict__ num_targets,
                    int P, int W, int G, int words, uint32_t flags, int vec_ok) {
    extern __shared__ __align__(16) uint32_t smem[];
    uint32_t* tmask = smem;                    // [words]
    uint32_t* rmask = smem + words;            // [words]
    uint32_t* list = rmask + words;            // [LIST_CAP]
    __shared__ EnvGeom geom;
    __shared__ int src[3];
    __shared__ int warp_tot[K1_WARPS];
    __shared__ int total_targets;

    const int n = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 2 * words; i += K1_THREADS) smem[i] = 0u;
    if (tid < 9) geom.kinv[tid] = kinv[tid];
    if (tid < 12) geom.c2w[tid] = c2w[n * 16 + tid];
    if (tid < 3) {
        // gennbv/utils.py:242-243 (same for pose_coord_to_idx_3D :286-295)
        float vs = voxel_size[n * 3 + tid];
        float hi = __fadd_rn(range_gt[n * 6 + 2 * tid], __fmul_rn(0.5f, vs));
        float lo = __fsub_rn(range_gt[n * 6 + 2 * tid + 1], __fmul_rn(0.5f, vs));
        geom.vs[tid] = vs; geom.hi[tid] = hi; geom.lo[tid] = lo;
        // pose_coord_to_idx_3D (utils.py:297): unclamped camera voxel; .int() in bresenham3D_pycuda
        float q = floorf(__fdiv_rn(__fsub_rn(pose_xyz[n * 3 + tid], lo), vs));
        q = fminf(fmaxf(q, -1048576.0f), 1048576.0f);   // keep the integer walk overflow-free (|coord| <= 2^20)
        src[tid] = (int)q;
    }
    __syncthreads();

    // ---- phase 1: unproject + voxel scatter ------------------------------------------------
    const float* dn = depth + (size_t)n * P;
    const int32_t* sn = seg + (size_t)n * P;
    const bool raw = flags & GNBV_RAW_DEPTH;
    const EnvGeom g = geom;
    auto scatter = [&](int p, float d, int s) {
        int lin = -1;
        if (s > 50) {                                           // fg (env_train_gennbv.py:504)
            if (raw) d = depth_post(d);
            int vrow = p / W;
            lin = pixel_to_voxel(g, d, (float)(p - vrow * W), (float)vrow, G);
        }
        if (lin >= 0) {
            uint32_t bit = 1u << (lin & 31);
            uint32_t* wp = &tmask[lin >> 5];
            if (!(*wp & bit)) atomicOr(wp, bit);
        }
    };
    if (vec_ok) {
        const float4* d4 = reinterpret_cast<const float4*>(dn);
        const int4* s4 = reinterpret_cast<const int4*>(sn);
        for (int i = tid; i < P / 4; i += K1_THREADS) {
            float4 d = __ldg(d4 + i);
            int4 s = __ldg(s4 + i);
            scatter(4 * i + 0, d.x, s.x);
            scatter(4 * i + 1, d.y, s.y);
            scatter(4 * i + 2, d.z, s.z);
            scatter(4 * i + 3, d.w, s.w);
        }
    } else {
        for (int p = tid; p < P; p += K1_THREADS) scatter(p, __ldg(dn + p), __ldg(sn + p));
    }
    __syncthreads();

    // ---- phase 2: compact targets (ascending voxel index), ray-cast --------------------------------
    // Warp-cooperative compaction: warp w owns the words [w*wpw, (w+1)*wpw); lanes read 32 consecutive words at a time.  Every
    // non-empty word is expanded by the whole warp at once (lane j owns bit j), so the cost follows the number of non-empty
    // words, not the longest run of set bits a single thread happens to own.  Neighbouring list entries are neighbouring
    // voxels: the 32 rays of a warp walk nearly the same cells, which keeps the shared-memory reads of the walk conflict-free.
    const int wpw = (words + K1_WARPS - 1) / K1_WARPS;
    const int w_begin = min(warp * wpw, words), w_end = min(w_begin + wpw, words);
    int cnt = 0;
    for (int w = w_begin + lane; w < w_end; w += 32) cnt += __popc(tmask[w]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) warp_tot[warp] = cnt;
    __syncthreads();
    if (warp == 0) {
        int t = lane < K1_WARPS ? warp_tot[lane] : 0;
        int s = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int q = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += q;
        }
        if (lane < K1_WARPS) warp_tot[lane] = s - t;      // exclusive
        if (lane == 31) total_targets = s;
    }
    __syncthreads();
    const int my_base = warp_tot[warp];
    const int total = total_targets;
    if (tid == 0 && num_targets) num_targets[n] = total;

    const int sx0 = src[0], sy0 = src[1], sz0 = src[2];
    for (int round = 0; round * LIST_CAP < total; ++round) {
        const int win_lo = round * LIST_CAP, win_hi = min(total, win_lo + LIST_CAP);
        if (my_base < win_hi && my_base + cnt > win_lo) {
            int rank = my_base;
            for (int w0 = w_begin; w0 < w_end; w0 += 32) {
                const int w = w0 + lane;
                const uint32_t mine = w < w_end ? tmask[w] : 0u;
                unsigned nz = __ballot_sync(0xffffffffu, mine != 0u);
                while (nz) {
                    const int srcl = __ffs(nz) - 1;
                    nz &= nz - 1;
                    const uint32_t wd = __shfl_sync(0xffffffffu, mine, srcl);
                    if ((wd >> lane) & 1u) {
                        const int pos = rank + __popc(wd & ((1u << lane) - 1u));
                        if (pos >= win_lo && pos < win_hi) list[pos - win_lo] = (uint32_t)((w0 + srcl) * 32 + lane);
                    }
                    rank += __popc(wd);
                }
            }
        }
        __syncthreads();
        for (int i = tid; i < win_hi - win_lo; i += K1_THREADS) {
            int lin = (int)list[i];
            int z = lin % G, t = lin / G;
            int y = t % G, x = t / G;
            raycast(rmask, G, sx0, sy0, sz0, x, y, z);
        }
        __syncthreads();
    }

    // ---- phase 3: publish the masks ------------------------------------------------------------
    uint4* to = reinterpret_cast<uint4*>(tmask_out + (size_t)n * words);
    uint4* ro = reinterpret_cast<uint4*>(rmask_out + (size_t)n * words);
    const uint4* ts = reinterpret_cast<const uint4*>(tmask);
    const uint4* rs = reinterpret_cast<const uint4*>(rmask);
    for (int i = tid; i < words / 4; i += K1_THREADS) { to[i] = ts[i]; ro[i] = rs[i]; }
}

// update_occ_grid's dense part (env_train_gennbv.py:311-326) + grid_occupancy_tri_cls (utils.py:318-321)
// for 4 voxels, m = 4 mask bits each of touched / target.
__device__ __forceinline__ void update4(float4& p, float4& s, const float4& gt, float4& tri, uint32_t tb, uint32_t rb,
                                        float& cov) {
    float* pp = &p.x; float* ss = &s.x; const float* gg = &gt.x; float* tt = &tri.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float pr = pp[k];
        if (rb & (1u << k)) pr = __fsub_rn(pr, 0.05f);
        if (tb & (1u << k)) pr = 1.0f;
        pp[k] = pr;
        tt[k] = (pr > 0.5f ? 1.0f : 0.0f) - (pr < 0.0f ? 1.0f : 0.0f);
        float occ = (tb & (1u << k)) ? 1.0f : 0.0f;
        float sg = __fadd_rn(ss[k], __fmul_rn(occ, gg[k]));
        sg = fminf(fmaxf(sg, 0.0f), 1.0f);
        ss[k] = sg;
        cov += sg;
    }
}

template <bool VEC>
__global__ void __launch_bounds__(K2_THREADS)
grid_update_kernel(const uint32_t* __restrict__ tmask, const uint32_t* __restrict__ rmask,
                   const float* __restrict__ grid_gt, float* __restrict__ prob, float* __restrict__ scan,
                   float* __restrict__ tri, int64_t tri_stride, float* __restrict__ partial,
                   int V, int words, int chunks) {
    const int n = blockIdx.y, chunk = blockIdx.x, tid = threadIdx.x;
    const size_t base = (size_t)n * V;
    const uint32_t* tm = tmask + (size_t)n * words;
    const uint32_t* rm = rmask + (size_t)n * words;
    float* trin = tri + (size_t)n * tri_stride;
    float cov = 0.0f;
    const int v0 = chunk * K2_CHUNK;
    if (VEC) {
        float4 p[K2_UNROLL], s[K2_UNROLL], g[K2_UNROLL];
        uint32_t tb[K2_UNROLL], rb[K2_UNROLL];
        bool ok[K2_UNROLL];
#pragma unroll
        for (int k = 0; k < K2_UNROLL; ++k) {
            int v = v0 + (k * K2_THREADS + tid) * 4;
            ok[k] = v < V;
            if (ok[k]) {
                p[k] = ld_stream4(prob + base + v);
                s[k] = ld_stream4(scan + base + v);
                g[k] = ldg_stream4(grid_gt + base + v);
                tb[k] = (__ldg(tm + (v >> 5)) >> (v & 31)) & 15u;
                rb[k] = (__ldg(rm + (v >> 5)) >> (v & 31)) & 15u;
            }
        }
#pragma unroll
        for (int k = 0; k < K2_UNROLL; ++k) {
            if (ok[k]) {
                int v = v0 + (k * K2_THREADS + tid) * 4;
                float4 t;
                update4(p[k], s[k], g[k], t, tb[k], rb[k], cov);
                stg_stream4(prob + base + v, p[k]);
                stg_stream4(scan + base + v, s[k]);
                stg_stream4(trin + v, t);
            }
        }
    } else {
        for (int v = v0 + tid; v < min(V, v0 + K2_CHUNK); v += K2_THREADS) {
            uint32_t tbit = (tm[v >> 5] >> (v & 31)) & 1u, rbit = (rm[v >> 5] >> (v & 31)) & 1u;
            float pr = prob[base + v];
            if (rbit) pr = __fsub_rn(pr, 0.05f);
            if (tbit) pr = 1.0f;
            prob[base + v] = pr;
            trin[v] = (pr > 0.5f ? 1.0f : 0.0f) - (pr < 0.0f ? 1.0f : 0.0f);
            float sg = __fadd_rn(scan[base + v], __fmul_rn(tbit ? 1.0f : 0.0f, grid_gt[base + v]));
            sg = fminf(fmaxf(sg, 0.0f), 1.0f);
            scan[base + v] = sg;
            cov += sg;
        }
    }
    // block reduction in a fixed order (values are {0,1}: exact in fp32 for any order anyway)
    __shared__ float wsum[K2_THREADS / 32];
    cov = warp_sum(cov);
    if ((tid & 31) == 0) wsum[tid >> 5] = cov;
    __syncthreads();
    if (tid == 0) {
        float t = 0.0f;
#pragma unroll
        for (int w = 0; w < K2_THREADS / 32; ++w) t += wsum[w];
        partial[(size_t)n * chunks + chunk] = t;
    }
}

// The same update moving only the bytes that can change.  A 16-byte group (4 voxels) no ray touched this step keeps its
// prob / scanned values, so for it the kernel reads prob (to form the tri-class value, which must be written to this step's
// observation row regardless) and writes tri: 8 B per voxel instead of 24.  Touched groups take the full path.  The coverage
// sum is kept INCREMENTALLY: scanned_gt only changes at target voxels, so partial = sum over touched groups of
// (scanned_new - scanned_old) -- exact small integers in fp32 -- and cov_sum[n] += the env's partials (coverage_add_kernel).
__global__ void __launch_bounds__(K2_THREADS)
grid_update_sparse_kernel(const uint32_t* __restrict__ tmask, const uint32_t* __restrict__ rmask,
                          const float* __restrict__ grid_gt, float* __restrict__ prob, float* __restrict__ scan,
                          float* __restrict__ tri, int64_t tri_stride, float* __restrict__ partial,
                          int V, int words, int chunks) {
    const int n = blockIdx.y, chunk = blockIdx.x, tid = threadIdx.x;
    const size_t base = (size_t)n * V;
    const uint32_t* tm = tmask + (size_t)n * words;
    const uint32_t* rm = rmask + (size_t)n * words;
    float* trin = tri + (size_t)n * tri_stride;
    float dcov = 0.0f;
    const int v0 = chunk * K2_CHUNK;
    float4 p[K2_UNROLL], s[K2_UNROLL], g[K2_UNROLL];
    uint32_t tb[K2_UNROLL], rb[K2_UNROLL];
    bool ok[K2_UNROLL];
#pragma unroll
    for (int k = 0; k < K2_UNROLL; ++k) {
        const int v = v0 + (k * K2_THREADS + tid) * 4;
        ok[k] = v < V;
        tb[k] = rb[k] = 0;
        if (ok[k]) {                                        // mask words and prob of all groups first: nothing here depends on a load
            tb[k] = __ldg(tm + (v >> 5));
            rb[k] = __ldg(rm + (v >> 5));
            p[k] = ld_stream4(prob + base + v);
        }
    }
#pragma unroll
    for (int k = 0; k < K2_UNROLL; ++k) {
        const int v = v0 + (k * K2_THREADS + tid) * 4;
        tb[k] = (tb[k] >> (v & 31)) & 15u;
        rb[k] = (rb[k] >> (v & 31)) & 15u;
        if (ok[k] && tb[k]) {                               // scanned_gt can only change where a target voxel is
            s[k] = ld_stream4(scan + base + v);
            g[k] = ldg_stream4(grid_gt + base + v);
        }
    }
#pragma unroll
    for (int k = 0; k < K2_UNROLL; ++k) {
        if (!ok[k]) continue;
        const int v = v0 + (k * K2_THREADS + tid) * 4;
        float4 t;
        if (tb[k]) {
            float before = (s[k].x + s[k].y) + (s[k].z + s[k].w), after = 0.0f;
            update4(p[k], s[k], g[k], t, tb[k], rb[k], after);
            dcov += after - before;
            stg_stream4(prob + base + v, p[k]);
            stg_stream4(scan + base + v, s[k]);
        } else {
            float* pp = &p[k].x; float* tt = &t.x;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float pr = pp[e];
                if (rb[k] & (1u << e)) pr = __fsub_rn(pr, 0.05f);
                pp[e] = pr;
                tt[e] = (pr > 0.5f ? 1.0f : 0.0f) - (pr < 0.0f ? 1.0f : 0.0f);
            }
            if (rb[k]) stg_stream4(prob + base + v, p[k]);
        }
        stg_stream4(trin + v, t);
    }
    __shared__ float wsum[K2_THREADS / 32];
    dcov = warp_sum(dcov);
    if ((tid & 31) == 0) wsum[tid >> 5] = dcov;
    __syncthreads();
    if (tid == 0) {
        float t = 0.0f;
#pragma unroll
        for (int w = 0; w < K2_THREADS / 32; ++w) t += wsum[w];
        partial[(size_t)n * chunks + chunk] = t;
    }
}

__global__ void coverage_add_kernel(const float* __restrict__ partial, float* __restrict__ cov_sum, int N, int chunks) {
    int n = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (n >= N) return;
    float t = 0.0f;
    for (int c = lane; c < chunks; c += 32) t += partial[(size_t)n * chunks + c];
    t = warp_sum(t);
    if (lane == 0) cov_sum[n] += t;
}

__global__ void coverage_finalize_kernel(const float* __restrict__ partial, float* __restrict__ cov_sum, int N, int chunks) {
    int n = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (n >= N) return;
    float t = 0.0f;
    for (int c = lane; c < chunks; c += 32) t += partial[(size_t)n * chunks + c];
    t = warp_sum(t);
    if (lane == 0) cov_sum[n] = t;
}

__global__ void __launch_bounds__(256)
reset_grids_kernel(float* __restrict__ prob, float* __restrict__ scan, const uint8_t* __restrict__ flags, int V, int vec_ok) {
    const int n = blockIdx.y;
    if (!flags[n]) return;
    const size_t base = (size_t)n * V;
    for (int v = (blockIdx.x * 256 + threadIdx.x) * 4; v < V; v += gridDim.x * 256 * 4) {
        if (vec_ok && v + 3 < V) {
            stg_stream4(prob + base + v, make_float4(0.f, 0.f, 0.f, 0.f));
            stg_stream4(scan + base + v, make_float4(0.f, 0.f, 0.f, 0.f));
        } else {
            for (int k = v; k < min(v + 4, V); ++k) { prob[base + k] = 0.f; scan[base + k] = 0.f; }
        }
    }
}


// ---- compatibility kernels behind the reference's free functions (gennbv/utils.py) -------------------------------
// scanned_pts_to_idx_3D (utils.py:230-270) for one env given explicit world points: OR the voxel bit of every point
// strictly inside the grid volume (clamped index), mask[words] must be zeroed by the caller.
__global__ void points_to_mask_kernel(const float* __restrict__ pts, int64_t n, const float* __restrict__ range6,
                                      const float* __restrict__ vs3, uint32_t* __restrict__ mask, int G) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int idx[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float vs = vs3[a];
        const float hi = __fadd_rn(range6[2 * a], __fmul_rn(0.5f, vs)), lo = __fsub_rn(range6[2 * a + 1], __fmul_rn(0.5f, vs));
        const float w = pts[i * 3 + a];
        if (!(hi > w && w > lo)) return;
        int k = (int)floorf(__fdiv_rn(__fsub_rn(w, lo), vs));
        idx[a] = max(0, min(G - 1, k));
    }
    const int lin = (idx[0] * G + idx[1]) * G + idx[2];
    atomicOr(&mask[lin >> 5], 1u << (lin & 31));
}

// bresenham3D_pycuda (utils.py:24-227), two passes so that the concatenated output keeps the reference's layout
// (ray order, duplicates kept): pass 0 counts the in-bounds voxels of each ray, pass 1 writes them at offsets[ray].
__global__ void bresenham_rays_kernel(const int32_t* __restrict__ src, const int32_t* __restrict__ tgt, int64_t num_rays,
                                      int G, int64_t* __restrict__ counts, const int64_t* __restrict__ offsets,
                                      int64_t* __restrict__ out) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= num_rays) return;
    const int x0 = src[0], y0 = src[1], z0 = src[2], x1 = tgt[r * 3], y1 = tgt[r * 3 + 1], z1 = tgt[r * 3 + 2];
    const int dx = abs(x1 - x0), dy = abs(y1 - y0), dz = abs(z1 - z0);
    const int sx = x0 < x1 ? 1 : -1, sy = y0 < y1 ? 1 : -1, sz = z0 < z1 ? 1 : -1;
    const int dm = max(dx, max(dy, dz));
    int x = x0, y = y0, z = z0, p1, p2, dmaj, d1, d2;
    int axis = dm == dx ? 0 : (dm == dy ? 1 : 2);                      // tie-break dx -> dy -> dz (utils.py:69,102,133)
    if (axis == 0) { dmaj = dx; d1 = dy; d2 = dz; } else if (axis == 1) { dmaj = dy; d1 = dx; d2 = dz; } else { dmaj = dz; d1 = dx; d2 = dy; }
    p1 = 2 * d1 - dmaj; p2 = 2 * d2 - dmaj;
    int64_t n = 0;
    int64_t* o = out ? out + offsets[r] * 3 : nullptr;
    const int cap = 3 * G;
    auto emit = [&]() {
        if ((unsigned)x < (unsigned)G && (unsigned)y < (unsigned)G && (unsigned)z < (unsigned)G) {
            if (o) { o[n * 3] = x; o[n * 3 + 1] = y; o[n * 3 + 2] = z; }
            ++n;
        }
    };
    emit();
    for (int i = 0; i < dmaj && n < cap; ++i) {
        if (axis == 0) { if (p1 >= 0) { y += sy; p1 -= 2 * dmaj; } if (p2 >= 0) { z += sz; p2 -= 2 * dmaj; } x += sx; }
        else if (axis == 1) { if (p1 >= 0) { x += sx; p1 -= 2 * dmaj; } if (p2 >= 0) { z += sz; p2 -= 2 * dmaj; } y += sy; }
        else { if (p1 >= 0) { x += sx; p1 -= 2 * dmaj; } if (p2 >= 0) { y += sy; p2 -= 2 * dmaj; } z += sz; }
        p1 += 2 * d1; p2 += 2 * d2;
        emit();
    }
    if (counts) counts[r] = n;
}

}  // namespace gnbv

using namespace gnbv;

extern "C" size_t gnbv_voxelize_workspace_bytes(int num_envs, int grid_size) {
    if (num_envs <= 0 || grid_size <= 0) return 0;
    return make_layout(num_envs, grid_size).total;
}

extern "C" int gnbv_voxelize_masks(void* workspace, int num_envs, int grid_size, const uint32_t** target_mask,
                                   const uint32_t** touched_mask, int64_t* words_per_env) {
    GNBV_REQUIRE(workspace && num_envs > 0 && grid_size > 0, "gnbv_voxelize_masks: bad arguments");
    VoxelizeLayout L = make_layout(num_envs, grid_size);
    if (target_mask) *target_mask = reinterpret_cast<const uint32_t*>((char*)workspace + L.off_tmask);
    if (touched_mask) *touched_mask = reinterpret_cast<const uint32_t*>((char*)workspace + L.off_rmask);
    if (words_per_env) *words_per_env = L.words;
    return GNBV_OK;
}

static int check_workspace(const char* who, void* workspace, size_t workspace_bytes, const VoxelizeLayout& L) {
    if (!workspace) { set_error("%s: null workspace", who); return GNBV_E_ARG; }
    if (workspace_bytes < L.total) {
        set_error("%s: workspace %zu B < required %zu B", who, workspace_bytes, L.total);
        return GNBV_E_WORKSPACE;
    }
    if ((uintptr_t)workspace & 15) { set_error("%s: workspace must be 16-byte aligned", who); return GNBV_E_ARG; }
    return GNBV_OK;
}

extern "C" int gnbv_scan_raycast(const float* depth, const int32_t* seg, const float* kinv, const float* c2w,
                                 const float* range_gt, const float* voxel_size, const float* pose_xyz,
                                 int32_t* num_targets, void* workspace, size_t workspace_bytes, int N, int H, int W,
                                 int G, uint32_t flags, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GNBV_REQUIRE(depth && seg && kinv && c2w && range_gt && voxel_size && pose_xyz,
                 "gnbv_scan_raycast: null pointer argument");
    GNBV_REQUIRE(N > 0 && H > 0 && W > 0 && G > 0, "gnbv_scan_raycast: N, H, W, G must be positive");
    GNBV_REQUIRE((int64_t)H * W < (1LL << 30), "gnbv_scan_raycast: image too large");
    VoxelizeLayout L = make_layout(N, G);
    int rc = check_workspace("gnbv_scan_raycast", workspace, workspace_bytes, L);
    if (rc) return rc;
    const size_t smem = (size_t)(2 * L.words + LIST_CAP) * 4;
    GNBV_REQUIRE(smem <= 200 * 1024, "gnbv_scan_raycast: grid_size %d needs %zu B of shared memory (max G = 92)", G, smem);
    { int rc_ = ensure_dyn_smem(scan_raycast_kernel, smem); if (rc_) return rc_; }
    uint32_t* tmask = reinterpret_cast<uint32_t*>((char*)workspace + L.off_tmask);
    uint32_t* rmask = reinterpret_cast<uint32_t*>((char*)workspace + L.off_rmask);
    const int vec_in = ((H * W) % 4 == 0) && (((uintptr_t)depth | (uintptr_t)seg) & 15) == 0;
    scan_raycast_kernel<<<N, K1_THREADS, smem, stream>>>(depth, seg, kinv, c2w, range_gt, voxel_size, pose_xyz, tmask,
                                                         rmask, num_targets, H * W, W, G, (int)L.words, flags, vec_in);
    GNBV_LAUNCH_CHECK("scan_raycast_kernel");
    return GNBV_OK;
}

static int grid_update_impl(const float* grid_gt, float* prob_grid, float* scanned_gt, float* tri_out, int64_t tri_row_stride,
                            float* cov_sum, void* workspace, size_t workspace_bytes, int N, int G, bool sparse, void* stream_);

extern "C" int gnbv_grid_update(const float* grid_gt, float* prob_grid, float* scanned_gt, float* tri_out,
                                int64_t tri_row_stride, float* cov_sum, void* workspace, size_t workspace_bytes, int N,
                                int G, void* stream) {
    return grid_update_impl(grid_gt, prob_grid, scanned_gt, tri_out, tri_row_stride, cov_sum, workspace, workspace_bytes, N, G,
                            false, stream);
}

extern "C" int gnbv_grid_update_sparse(const float* grid_gt, float* prob_grid, float* scanned_gt, float* tri_out,
                                       int64_t tri_row_stride, float* cov_sum, void* workspace, size_t workspace_bytes, int N,
                                       int G, void* stream) {
    return grid_update_impl(grid_gt, prob_grid, scanned_gt, tri_out, tri_row_stride, cov_sum, workspace, workspace_bytes, N, G,
                            true, stream);
}

static int grid_update_impl(const float* grid_gt, float* prob_grid, float* scanned_gt, float* tri_out, int64_t tri_row_stride,
                            float* cov_sum, void* workspace, size_t workspace_bytes, int N, int G, bool sparse, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GNBV_REQUIRE(grid_gt && prob_grid && scanned_gt && tri_out && cov_sum, "gnbv_grid_update: null pointer argument");
    GNBV_REQUIRE(N > 0 && G > 0, "gnbv_grid_update: N, G must be positive");
    const int64_t V = (int64_t)G * G * G;
    GNBV_REQUIRE(tri_row_stride >= V, "gnbv_grid_update: tri_row_stride (%lld) < G^3", (long long)tri_row_stride);
    VoxelizeLayout L = make_layout(N, G);
    int rc = check_workspace("gnbv_grid_update", workspace, workspace_bytes, L);
    if (rc) return rc;
    uint32_t* tmask = reinterpret_cast<uint32_t*>((char*)workspace + L.off_tmask);
    uint32_t* rmask = reinterpret_cast<uint32_t*>((char*)workspace + L.off_rmask);
    float* partial = reinterpret_cast<float*>((char*)workspace + L.off_partial);
    const bool vec = (V % 4 == 0) && (tri_row_stride % 4 == 0) && (((uintptr_t)grid_gt | (uintptr_t)prob_grid |
                                                                     (uintptr_t)scanned_gt | (uintptr_t)tri_out) & 15) == 0;
    dim3 grid2((unsigned)L.chunks, (unsigned)N);
    if (sparse && vec) {
        grid_update_sparse_kernel<<<grid2, K2_THREADS, 0, stream>>>(tmask, rmask, grid_gt, prob_grid, scanned_gt, tri_out,
                                                                    tri_row_stride, partial, (int)V, (int)L.words, (int)L.chunks);
        GNBV_LAUNCH_CHECK("grid_update_sparse_kernel");
        coverage_add_kernel<<<(unsigned)ceil_div(N, 8), 256, 0, stream>>>(partial, cov_sum, N, (int)L.chunks);
        GNBV_LAUNCH_CHECK("coverage_add_kernel");
        return GNBV_OK;
    }
    GNBV_REQUIRE(!sparse, "gnbv_grid_update_sparse: needs 16-byte aligned grids and G^3, tri_row_stride multiples of 4");
    if (vec)
        grid_update_kernel<true><<<grid2, K2_THREADS, 0, stream>>>(tmask, rmask, grid_gt, prob_grid, scanned_gt, tri_out,
                                                                   tri_row_stride, partial, (int)V, (int)L.words, (int)L.chunks);
    else
        grid_update_kernel<false><<<grid2, K2_THREADS, 0, stream>>>(tmask, rmask, grid_gt, prob_grid, scanned_gt, tri_out,
                                                                    tri_row_stride, partial, (int)V, (int)L.words, (int)L.chunks);
    GNBV_LAUNCH_CHECK("grid_update_kernel");
    coverage_finalize_kernel<<<(unsigned)ceil_div(N, 8), 256, 0, stream>>>(partial, cov_sum, N, (int)L.chunks);
    GNBV_LAUNCH_CHECK("coverage_finalize_kernel");
    return GNBV_OK;
}

extern "C" int gnbv_voxelize_step(const float* depth, const int32_t* seg, const float* kinv, const float* c2w,
                                  const float* range_gt, const float* voxel_size, const float* pose_xyz,
                                  const float* grid_gt, float* prob_grid, float* scanned_gt, float* tri_out,
                                  int64_t tri_row_stride, float* cov_sum, int32_t* num_targets, void* workspace,
                                  size_t workspace_bytes, int N, int H, int W, int G, uint32_t flags, void* stream) {
    int rc = gnbv_scan_raycast(depth, seg, kinv, c2w, range_gt, voxel_size, pose_xyz, num_targets, workspace,
                               workspace_bytes, N, H, W, G, flags, stream);
    if (rc) return rc;
    return gnbv_grid_update(grid_gt, prob_grid, scanned_gt, tri_out, tri_row_stride, cov_sum, workspace, workspace_bytes,
                            N, G, stream);
}

extern "C" int gnbv_reset_grids(float* prob_grid, float* scanned_gt, const uint8_t* reset_flags, int N, int G,
                                void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GNBV_REQUIRE(prob_grid && scanned_gt && reset_flags && N > 0 && G > 0, "gnbv_reset_grids: bad arguments");
    const int64_t V = (int64_t)G * G * G;
    dim3 grid((unsigned)std::min<int64_t>(ceil_div(V, 1024 * 4), 64), (unsigned)N);
    const int vec_ok = (V % 4 == 0) && (((uintptr_t)prob_grid | (uintptr_t)scanned_gt) & 15) == 0;
    reset_grids_kernel<<<grid, 256, 0, stream>>>(prob_grid, scanned_gt, reset_flags, (int)V, vec_ok);
    GNBV_LAUNCH_CHECK("reset_grids_kernel");
    return GNBV_OK;
}

extern "C" int gnbv_points_to_voxel_mask(const float* points, int64_t num_points, const float* range_gt6,
                                         const float* voxel_size3, uint32_t* mask, int G, void* stream) {
    GNBV_REQUIRE(range_gt6 && voxel_size3 && mask && G > 0 && num_points >= 0 && (points || num_points == 0),
                 "gnbv_points_to_voxel_mask: bad arguments");
    if (num_points == 0) return GNBV_OK;
    points_to_mask_kernel<<<(unsigned)ceil_div(num_points, 256), 256, 0, (cudaStream_t)stream>>>(points, num_points, range_gt6,
                                                                                               voxel_size3, mask, G);
    GNBV_LAUNCH_CHECK("points_to_mask_kernel");
    return GNBV_OK;
}

extern "C" int gnbv_bresenham_rays(const int32_t* source3, const int32_t* targets, int64_t num_rays, int G, int64_t* counts,
                                   const int64_t* offsets, int64_t* out, void* stream) {
    GNBV_REQUIRE(source3 && G > 0 && num_rays >= 0 && (targets || num_rays == 0) && (counts || (offsets && out)),
                 "gnbv_bresenham_rays: bad arguments");
    if (num_rays == 0) return GNBV_OK;
    bresenham_rays_kernel<<<(unsigned)ceil_div(num_rays, 256), 256, 0, (cudaStream_t)stream>>>(source3, targets, num_rays, G,
                                                                                             counts, offsets, out);
    GNBV_LAUNCH_CHECK("bresenham_rays_kernel");
    return GNBV_OK;
}
