// eval_points.cu -- the scanned-point history of the eval env (gennbv/env/env_eval_gennbv.py:160-164, 253-257).
//
// The reference keeps, per env, the concatenation of every step's foreground world points (`pts_target_list`, [n,3] f32,
// unbounded) and at episode end reduces it with `torch.unique(torch.round(pts, decimals=2), dim=0)` to the distinct
// 1 cm lattice points, sorted lexicographically.  torch.round(decimals=2) is nearbyint(x * 100.f) / 100.f in fp32, so a
// point is identified by the integer triple k = nearbyint(p * 100); here every step appends the packed triple
//     key = (kx + 2^17) << 36 | (ky + 2^17) << 18 | (kz + 2^17)         (54 bits; |k| <= 2^17 - 1, i.e. +-1.31 km at 1 cm)
// (8 B per point instead of 12, ascending key order == the row order of torch.unique(dim=0), and the 9 spare bits take an
// env index so that ONE sort dedups the histories of up to 512 envs at episode end) to a per-env history
// of capacity (max_episode_length + 1) * H * W, which cannot overflow.  At episode end the caller sorts / dedups the keys
// and gnbv_keys_to_points turns them back into the fp32 rows k / 100.f the reference would hold.
//
// World points use the reference's exact fp32 operation order (env_train_gennbv.py:519-526; same chain as
// pixel_to_voxel in voxelize.cu and oracle/gennbv_oracle.c::back_project_one); compiled with -fmad=false.
#include "common.cuh"

#include <float.h>

#include <algorithm>

namespace gnbv {

constexpr int PTS_THREADS = 256;
constexpr int KEY_BITS = 18;
constexpr int KEY_BIAS = 1 << (KEY_BITS - 1);
constexpr int KEY_MASK = (1 << KEY_BITS) - 1;
constexpr float KEY_LIM = 131071.0f;       // |k| <= 2^17 - 1 (1.31 km at 1 cm); the reference's depth clamp keeps |p| <~ 60 m

__device__ __forceinline__ float depth_post_eval(float d) {     // env_train_base.py:520-523
    if (d != d) d = 0.0f;
    else if (d == -INFINITY) d = 0.0f;
    else if (d == INFINITY) d = FLT_MAX;
    d = fmaxf(d, -50.0f);
    return fabsf(d);
}

__device__ __forceinline__ int64_t pack_key(float x, float y, float z) {
    const float p[3] = {x, y, z};
    int64_t key = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float k = nearbyintf(__fmul_rn(p[a], 100.0f));          // round-half-even, as std::nearbyint on the CPU
        k = fminf(fmaxf(k, -KEY_LIM), KEY_LIM);
        if (k != k) k = 0.0f;
        key = (key << KEY_BITS) | (int64_t)((int)k + KEY_BIAS);
    }
    return key;
}

// grid (blocks, N): every foreground pixel of env n appends one key to keys[n, counts[n]++] (warp-aggregated atomics)
__global__ void __launch_bounds__(PTS_THREADS)
scan_points_kernel(const float* __restrict__ depth, const int32_t* __restrict__ seg, const float* __restrict__ kinv,
                   const float* __restrict__ c2w, int64_t* __restrict__ keys, int32_t* __restrict__ counts,
                   int P, int W, int64_t cap, uint32_t flags, int32_t* __restrict__ overflow) {
    __shared__ float sk[9], sc[12];
    const int n = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
    if (tid < 9) sk[tid] = kinv[tid];
    if (tid < 12) sc[tid] = c2w[n * 16 + tid];
    __syncthreads();
    const float* dn = depth + (size_t)n * P;
    const int32_t* sn = seg + (size_t)n * P;
    int64_t* kn = keys + (int64_t)n * cap;
    const bool raw = flags & GNBV_RAW_DEPTH;
    const int per_block = (P + gridDim.x - 1) / gridDim.x;
    const int p_begin = blockIdx.x * per_block, p_end = min(P, p_begin + per_block);
    for (int base = p_begin; base < p_end; base += PTS_THREADS) {          // uniform trip count per warp
        const int p = base + tid;
        bool fg = false;
        int64_t key = 0;
        if (p < p_end && __ldg(sn + p) > 50) {                              // fg (env_train_gennbv.py:504)
            fg = true;
            float d = __ldg(dn + p);
            if (raw) d = depth_post_eval(d);
            const int vrow = p / W;
            const float u = (float)(p - vrow * W), v = (float)vrow;
            const float px = __fmul_rn(d, u), py = __fmul_rn(d, v), pz = d;
            float cam[3], w[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                float acc = __fmul_rn(sk[r * 3 + 0], px);
                acc = __fmaf_rn(sk[r * 3 + 1], py, acc);
                acc = __fmaf_rn(sk[r * 3 + 2], pz, acc);
                cam[r] = acc;
            }
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                float acc = __fmul_rn(sc[i * 4 + 0], cam[0]);
                acc = __fmaf_rn(sc[i * 4 + 1], cam[1], acc);
                acc = __fmaf_rn(sc[i * 4 + 2], cam[2], acc);
                acc = __fmaf_rn(sc[i * 4 + 3], 1.0f, acc);
                w[i] = acc;
            }
            key = pack_key(w[0], w[1], w[2]);
        }
        const unsigned ballot = __ballot_sync(0xffffffffu, fg);
        if (ballot) {
            int warp_base = 0;
            if (lane == 0) warp_base = atomicAdd(&counts[n], __popc(ballot));
            warp_base = __shfl_sync(0xffffffffu, warp_base, 0);
            if (fg) {
                const int64_t slot = (int64_t)warp_base + __popc(ballot & ((1u << lane) - 1u));
                if (slot < cap) kn[slot] = key;
                else *overflow = 1;
            }
        }
    }
}

__global__ void keys_to_points_kernel(const int64_t* __restrict__ keys, int64_t n, float* __restrict__ pts) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t key = keys[i];
    // bits above 3 * KEY_BITS (an env index added by the caller for batched dedup) are ignored
    const int kz = (int)(key & KEY_MASK) - KEY_BIAS, ky = (int)((key >> KEY_BITS) & KEY_MASK) - KEY_BIAS,
              kx = (int)((key >> (2 * KEY_BITS)) & KEY_MASK) - KEY_BIAS;
    pts[3 * i + 0] = __fdiv_rn((float)kx, 100.0f);
    pts[3 * i + 1] = __fdiv_rn((float)ky, 100.0f);
    pts[3 * i + 2] = __fdiv_rn((float)kz, 100.0f);
}

// arbitrary world points -> packed 1 cm lattice keys (the same rounding as the history kernel)
__global__ void points_to_keys_kernel(const float* __restrict__ pts, int64_t n, int64_t* __restrict__ keys) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = pack_key(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
}

// histories of a list of envs -> one array, env r's valid prefix at out[offsets[r] ..), tagged with r in the bits above the key
__global__ void pack_env_keys_kernel(const int64_t* __restrict__ keys, int64_t cap, const int64_t* __restrict__ env_rows,
                                     const int64_t* __restrict__ offsets, int tag_shift, int64_t* __restrict__ out) {
    const int r = blockIdx.y;
    const int64_t n = offsets[r + 1] - offsets[r];
    const int64_t* src = keys + env_rows[r] * cap;
    int64_t* dst = out + offsets[r];
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x)
        dst[j] = src[j] | ((int64_t)r << tag_shift);
}

}  // namespace gnbv

using namespace gnbv;

extern "C" int gnbv_points_to_keys(const float* points, int64_t num_points, int64_t* keys, void* stream) {
    GNBV_REQUIRE(points && keys && num_points > 0, "gnbv_points_to_keys: bad arguments");
    points_to_keys_kernel<<<(unsigned)ceil_div(num_points, 256), 256, 0, (cudaStream_t)stream>>>(points, num_points, keys);
    GNBV_LAUNCH_CHECK("points_to_keys_kernel");
    return GNBV_OK;
}

extern "C" int gnbv_pack_env_keys(const int64_t* keys, int64_t capacity, const int64_t* env_rows, const int64_t* offsets,
                                  int num_rows, int tag_shift, int64_t* out, void* stream) {
    GNBV_REQUIRE(keys && env_rows && offsets && out && num_rows > 0 && capacity > 0 && tag_shift >= 54 && tag_shift < 63,
                 "gnbv_pack_env_keys: bad arguments");
    pack_env_keys_kernel<<<dim3(64, num_rows), 256, 0, (cudaStream_t)stream>>>(keys, capacity, env_rows, offsets, tag_shift, out);
    GNBV_LAUNCH_CHECK("pack_env_keys_kernel");
    return GNBV_OK;
}

extern "C" int gnbv_scan_points(const float* depth, const int32_t* seg, const float* kinv, const float* c2w, int64_t* keys,
                                int32_t* counts, int32_t* overflow, int num_envs, int height, int width, int64_t capacity,
                                uint32_t flags, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GNBV_REQUIRE(depth && seg && kinv && c2w && keys && counts && overflow, "gnbv_scan_points: null pointer argument");
    GNBV_REQUIRE(num_envs > 0 && height > 0 && width > 0 && capacity > 0 && capacity < (1LL << 31),
                 "gnbv_scan_points: bad dimensions");
    const int P = height * width;
    const int blocks = (int)std::min<int64_t>(ceil_div(P, PTS_THREADS * 4), 64);
    scan_points_kernel<<<dim3(blocks, num_envs), PTS_THREADS, 0, stream>>>(depth, seg, kinv, c2w, keys, counts, P, width, capacity,
                                                                          flags, overflow);
    GNBV_LAUNCH_CHECK("scan_points_kernel");
    return GNBV_OK;
}

extern "C" int gnbv_keys_to_points(const int64_t* keys, int64_t num_keys, float* points, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GNBV_REQUIRE(num_keys >= 0, "gnbv_keys_to_points: negative count");
    if (num_keys == 0) return GNBV_OK;
    GNBV_REQUIRE(keys && points, "gnbv_keys_to_points: null pointer argument");
    keys_to_points_kernel<<<(unsigned)ceil_div(num_keys, 256), 256, 0, stream>>>(keys, num_keys, points);
    GNBV_LAUNCH_CHECK("keys_to_points_kernel");
    return GNBV_OK;
}
