// conv2_tc.cu -- Conv3d(16,16,3,stride 2) forward as an implicit GEMM on the 5th-generation tensor cores.
//
//   D[128 voxels x 16 co] (TMEM, fp32)  +=  A[128 x K] * W[16 x K]^T,   K = 432 = 27 taps x 16 ci, 3xTF32 split (tc.cuh)
//
// A tile = up to 128 output voxels of one (sample, x2) plane (RT consecutive y2 rows).  K is walked in nine stages
// (i, j) of 48 = 3 z-taps x 16 ci: for one output voxel those 48 inputs are 192 CONTIGUOUS bytes of the channels-last
// conv1 activation, so the loader threads issue coalesced 16-byte loads, apply BatchNorm1 + ReLU, split into (hi, lo) and
// store straight into the canonical UMMA K-major layout (conflict-free: 8 consecutive threads fill one core-matrix
// column).  The whole weight tensor lives in shared memory as (hi, lo) K-major tiles for the lifetime of the persistent
// CTA.  One thread issues 18 tcgen05.mma.kind::tf32 per stage; two A stages + tcgen05.commit -> mbarrier overlap the
// loads of stage s+1 with the MMAs of stage s.  Epilogue: tcgen05.ld -> + bias -> channel-major store + Welford
// statistics for BatchNorm2 (one record per tile).
// Status: parity-green but opt-in (GNBV_CONV2_TC=1): measured 1.43 ms vs 0.55 ms for the CUDA-core kernel at B = 256, because
// each staged A element feeds only N = 16 MACs while its staging costs ~20 instructions.  Design note for a follow-up: a
// space-to-depth shared-memory layout [x-plane][y parity][z parity][ci quad][y/2][z/2] x 16 B makes the A tile of every tap a
// plain UMMA descriptor view (SBO = 128 B between 8-voxel groups, LBO = one ci-quad block), so each activation is
// transformed and stored once instead of 3.4 times (M = 64 tiles of 4 rows x 16 voxels fit 178 KB of shared memory).
#include "conv2_tc.cuh"
#include "tc.cuh"

#include <algorithm>

namespace gnbv {

constexpr int C2 = 16, C2_TAPS = 27;
constexpr int C2T_THREADS = 256;
constexpr int KSTAGE = 48;                                  // 3 z-taps x 16 ci
constexpr uint32_t A_SBO = (KSTAGE / 4) * 128;              // 1536 B between 8-row groups
constexpr uint32_t A_TILE = 128 * KSTAGE * 4;               // 24,576 B (one of hi / lo)
constexpr int KW = C2_TAPS * C2;                            // 432
constexpr uint32_t W_SBO = (KW / 4) * 128;                  // 13,824 B
constexpr uint32_t W_TILE = 2 * W_SBO;                      // 16 rows = 2 row groups: 27,648 B (one of hi / lo)
constexpr size_t C2T_SMEM = 2 * W_TILE + 4 * A_TILE;        // 153,600 B
constexpr int C2_PART_STRIDE = 2 * C2 + 4;

static int rows_per_tile(int G2) { return std::max(1, std::min(G2, 128 / G2)); }
bool conv2_tc_supported(int G1, int G2) { return G2 >= 1 && G2 <= 128 && G1 >= 2 * G2 + 1; }
int conv2_tc_tiles(int B, int G2) {
    const int rt = rows_per_tile(G2);
    return B * G2 * (int)ceil_div(G2, rt);
}

__global__ void __launch_bounds__(C2T_THREADS, 1)
conv2_fwd_tc_kernel(const float* __restrict__ y1, const float* __restrict__ stat1, const float* __restrict__ w,
                    const float* __restrict__ bias, float* __restrict__ y2, float* __restrict__ part, int* __restrict__ err,
                    int G1, int G2, int RT, int total_tiles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* w_hi = smem;
    uint8_t* w_lo = smem + W_TILE;
    uint8_t* a_base = smem + 2 * W_TILE;                    // [2 stages][hi | lo]
    __shared__ __align__(8) uint64_t mbar[3];
    __shared__ uint32_t tmem_base_smem;
    __shared__ float a1s[C2], b1s[C2], bs[C2];
    __shared__ float red[C2T_THREADS / 32][C2_PART_STRIDE];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int P1 = G1 * G1 * G1, P2 = G2 * G2 * G2, NYB = (G2 + RT - 1) / RT;

    // ---- one-time setup: weights -> (hi, lo) K-major tiles with k = tap*16 + ci; barriers; TMEM
    for (int e = tid; e < C2 * KW; e += C2T_THREADS) {
        const int co = e / KW, k = e - co * KW, tap = k >> 4, ci = k & 15;
        float h, l;
        tc::split_tf32(w[(co * C2 + ci) * C2_TAPS + tap], h, l);
        const uint32_t off = (uint32_t)((co >> 3) * W_SBO + (k >> 2) * 128 + (co & 7) * 16 + (k & 3) * 4);
        *reinterpret_cast<float*>(w_hi + off) = h;
        *reinterpret_cast<float*>(w_lo + off) = l;
    }
    if (tid < C2) { a1s[tid] = stat1[2 * C2 + tid]; b1s[tid] = stat1[3 * C2 + tid]; bs[tid] = bias[tid]; }
    if (tid == 0) {
        tc::mbar_init(tc::smem_u32(&mbar[0]), 1);
        tc::mbar_init(tc::smem_u32(&mbar[1]), 1);
        tc::mbar_init(tc::smem_u32(&mbar[2]), 1);
        tc::fence_mbar_init();
    }
    if (warp == 0) tc::tmem_alloc(tc::smem_u32(&tmem_base_smem), 32);
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_d = tmem_base_smem;
    const uint32_t idesc = tc::make_idesc_tf32(128, C2);
    const uint32_t w_hi_a = tc::smem_u32(w_hi), w_lo_a = tc::smem_u32(w_lo);

    uint32_t phase[2] = {0, 0}, phase_acc = 0;
    bool pending[2] = {false, false};
    bool ok = true;
    // per-thread loader slots: 6 float4 per stage; slot q -> f = tid + 256 q -> (m, c)
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int b = tile / (G2 * NYB), rem = tile - b * G2 * NYB, x2 = rem / NYB, y0 = (rem - x2 * NYB) * RT;
        const int nr = min(RT, G2 - y0), nvox = nr * G2;
        const float* yb = y1 + (int64_t)b * P1 * C2;
        for (int s = 0; s < 9; ++s) {
            const int i = s / 3, j = s - 3 * i, buf = s & 1;
            if (pending[buf]) {                               // the MMAs that last read this buffer must be done
                ok = tc::mbar_wait(tc::smem_u32(&mbar[buf]), phase[buf]) && ok;
                phase[buf] ^= 1;
                pending[buf] = false;
            }
            uint8_t* a_hi = a_base + buf * 2 * A_TILE;
            uint8_t* a_lo = a_hi + A_TILE;
            float4 v[6];
#pragma unroll
            for (int q = 0; q < 6; ++q) {
                const int f = tid + q * C2T_THREADS, m = (f / 96) * 8 + (f & 7), c = (f >> 3) % 12;
                v[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (m < nvox) {
                    const int r = m / G2, z2 = m - r * G2;
                    const float* src = yb + (((int64_t)(2 * x2 + i) * G1 + (2 * (y0 + r) + j)) * G1 + 2 * z2) * C2 + c * 4;
                    v[q] = __ldg(reinterpret_cast<const float4*>(src));
                }
            }
#pragma unroll
            for (int q = 0; q < 6; ++q) {
                const int f = tid + q * C2T_THREADS, m = (f / 96) * 8 + (f & 7), c = (f >> 3) % 12;
                float4 h = make_float4(0.f, 0.f, 0.f, 0.f), l = h;
                if (m < nvox) {
                    const int cb = (c & 3) * 4;
                    const float x0 = fmaxf(fmaf(a1s[cb + 0], v[q].x, b1s[cb + 0]), 0.f), x1 = fmaxf(fmaf(a1s[cb + 1], v[q].y, b1s[cb + 1]), 0.f);
                    const float x2v = fmaxf(fmaf(a1s[cb + 2], v[q].z, b1s[cb + 2]), 0.f), x3 = fmaxf(fmaf(a1s[cb + 3], v[q].w, b1s[cb + 3]), 0.f);
                    tc::split_tf32(x0, h.x, l.x); tc::split_tf32(x1, h.y, l.y); tc::split_tf32(x2v, h.z, l.z); tc::split_tf32(x3, h.w, l.w);
                }
                const uint32_t off = (uint32_t)((m >> 3) * A_SBO + c * 128 + (m & 7) * 16);
                *reinterpret_cast<float4*>(a_hi + off) = h;
                *reinterpret_cast<float4*>(a_lo + off) = l;
            }
            tc::fence_async_smem();
            __syncthreads();
            if (tid == 0) {
                tc::tc_fence_after();
                const uint32_t ah = tc::smem_u32(a_hi), al = ah + A_TILE;
#pragma unroll
                for (int ks = 0; ks < KSTAGE / 8; ++ks) {
                    const uint32_t ao = ks * 256, wo = (uint32_t)(s * (KSTAGE / 4) + ks * 2) * 128;
                    const uint64_t dah = tc::make_smem_desc(ah + ao, 128, A_SBO), dal = tc::make_smem_desc(al + ao, 128, A_SBO);
                    const uint64_t dwh = tc::make_smem_desc(w_hi_a + wo, 128, W_SBO), dwl = tc::make_smem_desc(w_lo_a + wo, 128, W_SBO);
                    tc::mma_tf32(tmem_d, dal, dwh, idesc, (s == 0 && ks == 0) ? 0u : 1u);
                    tc::mma_tf32(tmem_d, dah, dwl, idesc, 1u);
                    tc::mma_tf32(tmem_d, dah, dwh, idesc, 1u);
                }
                tc::mma_commit(tc::smem_u32(&mbar[buf]));
                if (s == 8) tc::mma_commit(tc::smem_u32(&mbar[2]));
            }
            pending[buf] = true;
        }
        // ---- epilogue
        ok = tc::mbar_wait(tc::smem_u32(&mbar[2]), phase_acc) && ok;
        phase_acc ^= 1;
        tc::tc_fence_after();
        float acc[1][C2];
        int nvalid = 0;
        if (warp < 4) {
            const int m = warp * 32 + lane;
            tc::tmem_ld16(tmem_d + ((uint32_t)(warp * 32) << 16), acc[0]);
            if (m < nvox) {
                const int r = m / G2, z2 = m - r * G2;
                const int64_t pos = ((int64_t)x2 * G2 + y0 + r) * G2 + z2;
#pragma unroll
                for (int c = 0; c < C2; ++c) {
                    acc[0][c] += bs[c];
                    y2[((int64_t)b * C2 + c) * P2 + pos] = acc[0][c];
                }
                nvalid = 1;
            }
        } else {
#pragma unroll
            for (int c = 0; c < C2; ++c) acc[0][c] = 0.f;
        }
        if (part) {
            // block-level (count, mean, M2): two-pass inside each warp, Chan merge across warps (as encoder.cu::block_stats)
            const int nw = warp_sum_i(nvalid);
#pragma unroll
            for (int c = 0; c < C2; ++c) {
                float sv = nvalid ? acc[0][c] : 0.f;
                sv = warp_sum(sv);
                const float mean = nw > 0 ? sv / (float)nw : 0.f;
                float d2 = nvalid ? (acc[0][c] - mean) * (acc[0][c] - mean) : 0.f;
                d2 = warp_sum(d2);
                if (lane == 0) { red[warp][c] = mean; red[warp][C2 + c] = d2; }
            }
            if (lane == 0) red[warp][2 * C2] = (float)nw;
        }
        tc::tc_fence_before();
        __syncthreads();                                      // TMEM drained + red[] complete before reuse
        if (part && tid < C2) {
            float n = 0.f, mean = 0.f, M2 = 0.f;
#pragma unroll
            for (int wv = 0; wv < C2T_THREADS / 32; ++wv) {
                const float cnt = red[wv][2 * C2];
                if (cnt > 0.f) {
                    const float delta = red[wv][tid] - mean, nt = n + cnt;
                    mean += delta * cnt / nt;
                    M2 += red[wv][C2 + tid] + delta * delta * n * cnt / nt;
                    n = nt;
                }
            }
            float* pr = part + (int64_t)tile * C2_PART_STRIDE;
            pr[tid] = mean; pr[C2 + tid] = M2;
            if (tid == 0) pr[2 * C2] = n;
        }
        tc::tc_fence_after();
        __syncthreads();                                      // red[] consumed before the next tile's epilogue rewrites it
    }
    if (!ok) { if (tid == 0 && err) atomicExch(err, 1); asm volatile("trap;"); }     // a bounded wait expired: fail loudly
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_d, 32);
}

int launch_conv2_fwd_tc(const float* y1, const float* stat1, const float* w, const float* bias, float* y2, float* part,
                        int* err, int B, int G1, int G2, cudaStream_t stream) {
    GNBV_REQUIRE(conv2_tc_supported(G1, G2), "conv2_fwd_tc: unsupported grid (G1=%d G2=%d)", G1, G2);
    const int RT = rows_per_tile(G2), tiles = conv2_tc_tiles(B, G2);
    { int rc_ = ensure_dyn_smem(conv2_fwd_tc_kernel, C2T_SMEM); if (rc_) return rc_; }
    const int grid = std::min(tiles, 148);
    conv2_fwd_tc_kernel<<<grid, C2T_THREADS, C2T_SMEM, stream>>>(y1, stat1, w, bias, y2, part, err, G1, G2, RT, tiles);
    GNBV_LAUNCH_CHECK("conv2_fwd_tc_kernel");
    return GNBV_OK;
}

}  // namespace gnbv
