// conv2_ts.cu -- Conv3d(16,16,3,stride 2) forward on the 5th-generation tensor cores with the A operand in TENSOR MEMORY
// ("TS" form of tcgen05.mma), warp-specialised, persistent.
//
//   D[128 output voxels x 16 co] (TMEM, fp32)  +=  A[128 x K] (TMEM) * W[16 x K]^T (shared memory),  K = 432 = 27 taps x 16 ci
//
// Why A lives in TMEM.  With only N = 16 output channels every A element feeds 16 MACs.  An A operand in shared memory makes
// the MMA read 4 KB of shared memory per 16 K MACs -- 32 cycles of the 128 B/clk shared-memory port for an instruction
// whose math takes 8 -- so the SS form is shared-memory bound (and conv2_tc.cu, which used it, additionally paid an
// im2col copy).  From TMEM the tensor core reads A at its own rate: the floor is the 8 cycles per M128 x N16 x K8 MMA.
// TMEM lanes are GEMM rows, so every producer thread owns one output voxel and, per tap, moves that voxel's 16 input
// activations registers -> TMEM (tcgen05.st): no descriptor regularity is needed and the 3xTF32 split happens in registers.
//
// Pipeline of one CTA (one per SM, 13 warps):
//   warps 0-7   producers.  Stage the BatchNorm1+ReLU-ed input planes of the tile once into shared memory (coalesced 16-byte
//               global loads; layout [line][z parity][channel quad][z/2] so that the per-tap reads below are conflict-free),
//               then, chunk by chunk (chunk = the 3 z-taps of one (dx, dy): K = 48), read the 3 x 64 B each row needs
//               (LDS.128), split hi / lo and tcgen05.st them into one of four TMEM A buffers.  Two groups of four warps
//               (a warp can only touch the TMEM lanes of its own quarter) alternate chunks.
//   warp 12     one thread issues 18 tcgen05.mma.kind::tf32 per chunk (lo*hi + hi*lo + hi*hi for 6 k-steps), A from TMEM,
//               W from the K-major (hi, lo) tiles that stay in shared memory for the lifetime of the CTA;
//               tcgen05.commit -> mbarrier hands the A buffer back / signals the accumulator.
//   warps 8-11  epilogue: tcgen05.ld the accumulator (double-buffered, so the next tile's MMAs run meanwhile), + bias,
//               channel-major store, BatchNorm2 (count, mean, M2) record per tile.
// Work item = (sample, block of 8 output y-rows, group of 4 output x-planes): consecutive x-planes share one input plane,
// which stays staged (ring of 4 plane slots).  M = 8 rows x 16 (z padded from G2): at 64^3 (G2 = 15) 88 % of the rows are real.
#include "conv2_ts.cuh"
#include "tc.cuh"

#include <algorithm>

namespace gnbv {
namespace {

constexpr int C = 16, TAPS = 27;
constexpr int PROD_WARPS = 8, EPI_WARPS = 4;
constexpr int TS_THREADS = (PROD_WARPS + EPI_WARPS + 1) * 32;       // 416
constexpr int ROWS = 8, ZP = 16, LINES = 2 * ROWS + 1, XT = 4;
constexpr int QP = 17 * 16;                                         // bytes of one (line, z parity, quad) strip: 16 + 1 pad slots
constexpr int LINE_B = 2 * 4 * QP;                                  // 2,176
constexpr int PLANE_B = LINES * LINE_B;                             // 36,992
constexpr int NPLANES = 4;
constexpr int KW = TAPS * C;                                        // 432
constexpr uint32_t W_SBO = (KW / 4) * 128;                          // 13,824 B between the two 8-row groups of W
constexpr uint32_t W_TILE = 2 * W_SBO;                              // 27,648 B (one of hi / lo)
constexpr size_t TS_SMEM = 2 * W_TILE + (size_t)NPLANES * PLANE_B;  // 203,264 B
constexpr int NBUF = 4, A_COLS = 96, D_COL0 = NBUF * A_COLS;        // TMEM: 4 x 96 A columns, then 2 x 16 accumulator columns
constexpr int PART_STRIDE = 2 * C + 4;

__device__ __forceinline__ void bar_named(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// 16 consecutive 32-bit columns of this thread's TMEM lane <- registers
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
          "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]^T ; issued by ONE thread
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

struct Item { int b, y0, x_begin, x_end; };
__device__ __forceinline__ Item decode_item(int item, int G2, int NYB, int NXG) {
    Item it;
    it.b = item / (NYB * NXG);
    const int rem = item - it.b * NYB * NXG, yb = rem / NXG, xg = rem - yb * NXG;
    it.y0 = yb * ROWS;
    it.x_begin = xg * XT;
    it.x_end = min(G2, it.x_begin + XT);
    return it;
}

__global__ void __launch_bounds__(TS_THREADS, 1)
conv2_fwd_ts_kernel(const float* __restrict__ y1, const float* __restrict__ stat1, const float* __restrict__ w,
                    const float* __restrict__ bias, float* __restrict__ y2, float* __restrict__ part, int G1, int G2,
                    int total_items) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* w_hi = smem;
    uint8_t* w_lo = smem + W_TILE;
    uint8_t* planes = smem + 2 * W_TILE;
    __shared__ __align__(8) uint64_t bar_full[NBUF], bar_free[NBUF], bar_accfull[2], bar_accfree[2];
    __shared__ uint32_t tmem_base_smem;
    __shared__ float a1s[C], b1s[C], bs[C];
    __shared__ float red[EPI_WARPS][PART_STRIDE];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int P1 = G1 * G1 * G1, P2 = G2 * G2 * G2;
    const int NYB = (G2 + ROWS - 1) / ROWS, NXG = (G2 + XT - 1) / XT;

    // ---- one-time setup: weights -> (hi, lo) K-major tiles with k = tap*16 + ci; barriers; TMEM
    for (int e = tid; e < C * KW; e += TS_THREADS) {
        const int co = e / KW, k = e - co * KW, tap = k >> 4, ci = k & 15;
        float h, l;
        tc::split_tf32(w[(co * C + ci) * TAPS + tap], h, l);
        const uint32_t off = (uint32_t)((co >> 3) * W_SBO + (k >> 2) * 128 + (co & 7) * 16 + (k & 3) * 4);
        *reinterpret_cast<float*>(w_hi + off) = h;
        *reinterpret_cast<float*>(w_lo + off) = l;
    }
    if (tid < C) { a1s[tid] = stat1[2 * C + tid]; b1s[tid] = stat1[3 * C + tid]; bs[tid] = bias[tid]; }
    if (tid == 0) {
        for (int i = 0; i < NBUF; ++i) { tc::mbar_init(tc::smem_u32(&bar_full[i]), 128); tc::mbar_init(tc::smem_u32(&bar_free[i]), 1); }
        for (int i = 0; i < 2; ++i) { tc::mbar_init(tc::smem_u32(&bar_accfull[i]), 1); tc::mbar_init(tc::smem_u32(&bar_accfree[i]), 128); }
        tc::fence_mbar_init();
    }
    if (warp == PROD_WARPS + EPI_WARPS) tc::tmem_alloc(tc::smem_u32(&tmem_base_smem), 512);
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_base_smem;
    bool ok = true;

    if (warp < PROD_WARPS) {
        // =================================================================================== producers
        const int group = warp >> 2, q4 = warp & 3;
        const int m = q4 * 32 + lane, r = m >> 4, z2 = m & 15;
        const uint32_t lane_addr = tmem + ((uint32_t)(q4 * 32) << 16);
        const int sq = tid & 3;                                     // channel quad this thread stages (constant: 4 | 256, 4 | NCH)
        const float4 sa = make_float4(a1s[4 * sq], a1s[4 * sq + 1], a1s[4 * sq + 2], a1s[4 * sq + 3]);
        const float4 sb = make_float4(b1s[4 * sq], b1s[4 * sq + 1], b1s[4 * sq + 2], b1s[4 * sq + 3]);
        const int NCH = G1 * 4;                                     // 16-byte chunks per input line
        uint32_t cc = 0;                                            // chunk counter of this CTA (same in every role)
        for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
            const Item it = decode_item(item, G2, NYB, NXG);
            int staged_hi = 2 * it.x_begin - 1;                     // highest input plane already in the ring
            const int nlines = min(LINES, G1 - 2 * it.y0);
            for (int x2 = it.x_begin; x2 < it.x_end; ++x2) {
                bar_named(1, PROD_WARPS * 32);                      // every producer is done reading the planes about to be replaced
                for (int pl = max(staged_hi + 1, 2 * x2); pl <= 2 * x2 + 2; ++pl) {
                    uint8_t* dst = planes + (pl & (NPLANES - 1)) * PLANE_B;
                    const float* src = y1 + (((int64_t)it.b * G1 + pl) * G1 + 2 * it.y0) * (int64_t)G1 * C;
                    const int total = nlines * NCH;
                    for (int f0 = 0; f0 < total; f0 += 4 * PROD_WARPS * 32) {
                        float4 v[4];
                        int fi[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            fi[u] = f0 + u * PROD_WARPS * 32 + tid;
                            if (fi[u] < total) v[u] = __ldg(reinterpret_cast<const float4*>(src) + fi[u]);   // lines are contiguous
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            if (fi[u] < total) {
                                const int line = fi[u] / NCH, i = fi[u] - line * NCH, z = i >> 2;
                                float4 o;
                                o.x = fmaxf(fmaf(sa.x, v[u].x, sb.x), 0.f); o.y = fmaxf(fmaf(sa.y, v[u].y, sb.y), 0.f);
                                o.z = fmaxf(fmaf(sa.z, v[u].z, sb.z), 0.f); o.w = fmaxf(fmaf(sa.w, v[u].w, sb.w), 0.f);
                                *reinterpret_cast<float4*>(dst + ((line * 2 + (z & 1)) * 4 + sq) * QP + (z >> 1) * 16) = o;
                            }
                        }
                    }
                }
                staged_hi = 2 * x2 + 2;
                bar_named(1, PROD_WARPS * 32);                      // staged planes visible to all producers
                for (int c = 0; c < 9; ++c, ++cc) {
                    if ((int)(cc & 1) != group) continue;
                    const int dx = c / 3, dy = c - 3 * dx, buf = cc & (NBUF - 1);
                    ok = tc::mbar_wait(tc::smem_u32(&bar_free[buf]), ((cc >> 2) & 1) ^ 1) && ok;     // MMAs that read this buffer are done
                    tc::tc_fence_after();
                    const uint8_t* lp = planes + ((2 * x2 + dx) & (NPLANES - 1)) * PLANE_B + (2 * r + dy) * LINE_B;
                    const uint32_t abase = lane_addr + (uint32_t)(buf * A_COLS);
#pragma unroll
                    for (int dz = 0; dz < 3; ++dz) {
                        const uint8_t* sp = lp + (dz & 1) * 4 * QP + (z2 + (dz >> 1)) * 16;
                        uint32_t hi[16], lo[16];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float4 x = *reinterpret_cast<const float4*>(sp + q * QP);
                            const float xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const uint32_t h = __float_as_uint(xv[e]) & 0xffffe000u;
                                hi[4 * q + e] = h;
                                lo[4 * q + e] = __float_as_uint(xv[e] - __uint_as_float(h));
                            }
                        }
                        tmem_st16(abase + dz * 16, hi);
                        tmem_st16(abase + 48 + dz * 16, lo);
                    }
                    tmem_st_wait();
                    tc::tc_fence_before();
                    mbar_arrive(tc::smem_u32(&bar_full[buf]));
                }
            }
        }
    } else if (warp == PROD_WARPS + EPI_WARPS) {
        // =================================================================================== MMA issuer
        const uint32_t idesc = tc::make_idesc_tf32(128, C);
        const uint32_t w_hi_a = tc::smem_u32(w_hi), w_lo_a = tc::smem_u32(w_lo);
        uint32_t cc = 0, tt = 0;
        for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
            const Item it = decode_item(item, G2, NYB, NXG);
            for (int x2 = it.x_begin; x2 < it.x_end; ++x2, ++tt) {
                const uint32_t acc = tt & 1, tmem_d = tmem + D_COL0 + acc * C;
                ok = tc::mbar_wait(tc::smem_u32(&bar_accfree[acc]), ((tt >> 1) & 1) ^ 1) && ok;     // epilogue drained this accumulator
                for (int c = 0; c < 9; ++c, ++cc) {
                    const int buf = cc & (NBUF - 1);
                    ok = tc::mbar_wait(tc::smem_u32(&bar_full[buf]), (cc >> 2) & 1) && ok;
                    tc::tc_fence_after();
                    if (lane == 0) {
                        const uint32_t a_hi = tmem + (uint32_t)(buf * A_COLS), a_lo = a_hi + 48;
#pragma unroll
                        for (int ks = 0; ks < 6; ++ks) {
                            const uint32_t wo = (uint32_t)(c * 12 + ks * 2) * 128;               // k0 = c*48 + ks*8 -> (k0/4)*128 B
                            const uint64_t dwh = tc::make_smem_desc(w_hi_a + wo, 128, W_SBO), dwl = tc::make_smem_desc(w_lo_a + wo, 128, W_SBO);
                            mma_tf32_ts(tmem_d, a_lo + ks * 8, dwh, idesc, (c == 0 && ks == 0) ? 0u : 1u);      // small terms first
                            mma_tf32_ts(tmem_d, a_hi + ks * 8, dwl, idesc, 1u);
                            mma_tf32_ts(tmem_d, a_hi + ks * 8, dwh, idesc, 1u);
                        }
                        tc::mma_commit(tc::smem_u32(&bar_free[buf]));
                        if (c == 8) tc::mma_commit(tc::smem_u32(&bar_accfull[acc]));
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // =================================================================================== epilogue
        const int ew = warp - PROD_WARPS;                            // == warp % 4: the TMEM lane quarter this warp may read
        const int m = ew * 32 + lane, r = m >> 4, z2 = m & 15;
        uint32_t tt = 0;
        for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
            const Item it = decode_item(item, G2, NYB, NXG);
            for (int x2 = it.x_begin; x2 < it.x_end; ++x2, ++tt) {
                const uint32_t acc = tt & 1;
                ok = tc::mbar_wait(tc::smem_u32(&bar_accfull[acc]), (tt >> 1) & 1) && ok;
                tc::tc_fence_after();
                float v[C];
                tc::tmem_ld16(tmem + D_COL0 + acc * C + ((uint32_t)(ew * 32) << 16), v);
                tc::tc_fence_before();
                mbar_arrive(tc::smem_u32(&bar_accfree[acc]));        // the next tile but one may overwrite this accumulator
                const int y2r = it.y0 + r;
                const bool valid = y2r < G2 && z2 < G2;
                if (valid) {
                    const int64_t pos = ((int64_t)x2 * G2 + y2r) * G2 + z2;
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        v[c] += bs[c];
                        y2[((int64_t)it.b * C + c) * P2 + pos] = v[c];
                    }
                }
                if (part) {
                    // (count, mean, M2) of the tile: two-pass inside each warp, Chan merge across the four warps
                    const int nvalid = valid ? 1 : 0, nw = warp_sum_i(nvalid);
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        float sv = warp_sum(valid ? v[c] : 0.f);
                        const float mean = nw > 0 ? sv / (float)nw : 0.f;
                        float d2 = warp_sum(valid ? (v[c] - mean) * (v[c] - mean) : 0.f);
                        if (lane == 0) { red[ew][c] = mean; red[ew][C + c] = d2; }
                    }
                    if (lane == 0) red[ew][2 * C] = (float)nw;
                    bar_named(2, EPI_WARPS * 32);
                    if (m < C) {
                        float n = 0.f, mean = 0.f, M2 = 0.f;
#pragma unroll
                        for (int wv = 0; wv < EPI_WARPS; ++wv) {
                            const float cnt = red[wv][2 * C];
                            if (cnt > 0.f) {
                                const float delta = red[wv][m] - mean, nt = n + cnt;
                                mean += delta * cnt / nt;
                                M2 += red[wv][C + m] + delta * delta * n * cnt / nt;
                                n = nt;
                            }
                        }
                        const int tile = ((it.b * NYB + it.y0 / ROWS) * G2 + x2);
                        float* pr = part + (int64_t)tile * PART_STRIDE;
                        pr[m] = mean; pr[C + m] = M2;
                        if (m == 0) pr[2 * C] = n;
                    }
                    bar_named(2, EPI_WARPS * 32);                    // red[] consumed before the next tile rewrites it
                }
            }
        }
    }
    if (!ok) { asm volatile("trap;"); }                              // a bounded wait expired: fail loudly
    tc::tc_fence_before();
    __syncthreads();
    if (warp == PROD_WARPS + EPI_WARPS) tc::tmem_dealloc(tmem, 512);
}

}  // namespace

bool conv2_ts_supported(int G1, int G2) { return G2 >= 1 && G2 <= ZP && G1 >= 2 * G2 + 1 && G1 <= 2 * ZP; }
int conv2_ts_tiles(int B, int G2) { return B * G2 * (int)ceil_div(G2, ROWS); }

int launch_conv2_fwd_ts(const float* y1, const float* stat1, const float* w, const float* bias, float* y2, float* part, int B,
                        int G1, int G2, cudaStream_t stream) {
    GNBV_REQUIRE(conv2_ts_supported(G1, G2), "conv2_fwd_ts: unsupported grid (G1=%d G2=%d)", G1, G2);
    { int rc_ = ensure_dyn_smem(conv2_fwd_ts_kernel, TS_SMEM); if (rc_) return rc_; }
    const int items = B * (int)ceil_div(G2, ROWS) * (int)ceil_div(G2, XT);
    const int grid = std::min(items, 148);
    conv2_fwd_ts_kernel<<<grid, TS_THREADS, TS_SMEM, stream>>>(y1, stat1, w, bias, y2, part, G1, G2, items);
    GNBV_LAUNCH_CHECK("conv2_fwd_ts_kernel");
    return GNBV_OK;
}

}  // namespace gnbv
