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
#include "tma.cuh"

#include <algorithm>
#include <cstdlib>

namespace gnbv {
namespace {

constexpr int C = 16, TAPS = 27;
constexpr int PROD_WARPS = 8, EPI_WARPS = 4, LOAD_WARPS = 4;
constexpr int TS_THREADS = (PROD_WARPS + EPI_WARPS + 1) * 32;       // 416 (data-gradient kernel: no loader warps)
constexpr int FWD_THREADS = TS_THREADS + LOAD_WARPS * 32;           // 544
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
// lane 0 arrives for the whole warp; the __syncwarp orders every lane's earlier writes / completed tcgen05 ops before it
__device__ __forceinline__ void warp_arrive(uint32_t bar) {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
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
// 16-byte asynchronous global -> shared copy (LDGSTS): arbitrary per-thread destinations, no registers held while in flight
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]^T ; issued by ONE thread
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// ---- optional in-kernel time accounting (GNBV_TS_DEBUG bit 32): cycles spent in each wait / work section, per role, for CTA 0
__device__ unsigned long long g_ts_prof[32];
#define TS_T0() const long long t0_ = clock64()
#define TS_ACC(slot) do { if (dbg & 32) acc_[slot] += (unsigned long long)(clock64() - t0_); } while (0)

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

__global__ void __launch_bounds__(FWD_THREADS, 1)
conv2_fwd_ts_kernel(const float* __restrict__ y1, const float* __restrict__ stat1, const float* __restrict__ w,
                    const float* __restrict__ bias, float* __restrict__ y2, float* __restrict__ part, int G1, int G2,
                    int total_items, int dbg) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* w_hi = smem;
    uint8_t* w_lo = smem + W_TILE;
    uint8_t* planes = smem + 2 * W_TILE;
    __shared__ __align__(8) uint64_t bar_full[NBUF], bar_free[NBUF], bar_accfull[2], bar_accfree[2];
    __shared__ __align__(8) uint64_t bar_pfull[NPLANES], bar_pfree[NPLANES], bar_praw[NPLANES];
    __shared__ uint32_t tmem_base_smem;
    __shared__ float a1s[C], b1s[C], bs[C];
    __shared__ float ot[C][ROWS * ZP + 1];                           // epilogue transpose buffer [channel][row]
    const int tid = threadIdx.x, warp = tc::uniform_warp_index(), lane = tid & 31;
    const int P2 = G2 * G2 * G2;
    const int NYB = (G2 + ROWS - 1) / ROWS, NXG = (G2 + XT - 1) / XT;

    // ---- one-time setup: weights -> (hi, lo) K-major tiles with k = tap*16 + ci; barriers; TMEM
    for (int e = tid; e < C * KW; e += FWD_THREADS) {
        const int co = e / KW, k = e - co * KW, tap = k >> 4, ci = k & 15;
        float h, l;
        tc::split_tf32(w[(co * C + ci) * TAPS + tap], h, l);
        const uint32_t off = (uint32_t)((co >> 3) * W_SBO + (k >> 2) * 128 + (co & 7) * 16 + (k & 3) * 4);
        *reinterpret_cast<float*>(w_hi + off) = h;
        *reinterpret_cast<float*>(w_lo + off) = l;
    }
    if (tid < C) { a1s[tid] = stat1[2 * C + tid]; b1s[tid] = stat1[3 * C + tid]; bs[tid] = bias[tid]; }
    if (tid == 0) {
        // one arrival per WARP (lane 0 after __syncwarp): 128 per-thread arrivals on one shared-memory word serialise
        for (int i = 0; i < NBUF; ++i) { tc::mbar_init(tc::smem_u32(&bar_full[i]), 4); tc::mbar_init(tc::smem_u32(&bar_free[i]), 1); }
        for (int i = 0; i < 2; ++i) { tc::mbar_init(tc::smem_u32(&bar_accfull[i]), 1); tc::mbar_init(tc::smem_u32(&bar_accfree[i]), EPI_WARPS); }
        for (int i = 0; i < NPLANES; ++i) {
            tc::mbar_init(tc::smem_u32(&bar_pfull[i]), LOAD_WARPS);
            tc::mbar_init(tc::smem_u32(&bar_pfree[i]), PROD_WARPS);
            tc::mbar_init(tc::smem_u32(&bar_praw[i]), 1);
        }
        tc::fence_mbar_init();
    }
    if (warp == PROD_WARPS + EPI_WARPS) tc::tmem_alloc(tc::smem_u32(&tmem_base_smem), 512);
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_base_smem;
    bool ok = true;

    if (warp > PROD_WARPS + EPI_WARPS) {
        // =================================================================================== plane loaders
        // Input planes of the work item, in order, into a ring of NPLANES slots, running ahead of the producers.  Per plane:
        //   1. one thread issues a TMA bulk copy (cp.async.bulk, 1-D) per input line straight into the line's slot, raw
        //      channels-last order, completion counted in bytes on an mbarrier; two planes are in flight.  (16-byte
        //      cp.async copies were tried first: the LSU sustained only ~15 GB/s per SM with them, 2.2 TB/s for the GPU.)
        //   2. the four loader warps permute every line IN PLACE into the conflict-free layout
        //      [z parity][channel quad][z/2] and apply BatchNorm1 + ReLU on the way: thread lt owns 16-byte chunk lt of every
        //      line, reads four lines' worth into registers, a named barrier separates the reads of those lines from the
        //      writes to them, then the permuted stores;
        //   3. the mbarrier that publishes the plane to the producers.
        const int lt = tid - (PROD_WARPS + EPI_WARPS + 1) * 32, sq = lane & 3, lw = lt >> 5, nch = G1 * 4;
        const float4 sa = make_float4(a1s[4 * sq], a1s[4 * sq + 1], a1s[4 * sq + 2], a1s[4 * sq + 3]);
        const float4 sb = make_float4(b1s[4 * sq], b1s[4 * sq + 1], b1s[4 * sq + 2], b1s[4 * sq + 3]);
        const uint32_t line_bytes = (uint32_t)G1 * C * 4;
        uint32_t ps = 0;                                            // plane sequence number of this CTA (same in the producers)
        unsigned long long acc_[4] = {0, 0, 0, 0};
        for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
            const Item it = decode_item(item, G2, NYB, NXG);
            const int nplanes = 2 * (it.x_end - it.x_begin) + 1, nlines = min(LINES, G1 - 2 * it.y0);
            auto issue = [&](int k, uint32_t seq) {                 // thread 0 of the loaders only
                const int slot = seq & (NPLANES - 1);
                { TS_T0();
                ok = tc::mbar_wait(tc::smem_u32(&bar_pfree[slot]), ((seq >> 2) & 1) ^ 1) && ok;        // producers are done with the old plane
                TS_ACC(0); }
                tc::fence_async_smem();                             // their generic-proxy reads precede the async-proxy writes
                const uint32_t bar = tc::smem_u32(&bar_praw[slot]);
                mbar_expect_tx(bar, (dbg & 16) ? 0u : (uint32_t)nlines * line_bytes);
                const float* src = y1 + (((int64_t)it.b * G1 + (2 * it.x_begin + k)) * G1 + 2 * it.y0) * (int64_t)G1 * C;
                const uint32_t dst = tc::smem_u32(planes + slot * PLANE_B);
                if (!(dbg & 16))
                for (int line = 0; line < nlines; ++line)
                    bulk_g2s(dst + (uint32_t)(line * LINE_B), src + (int64_t)line * G1 * C, line_bytes, bar);
            };
            if (lt == 0) issue(0, ps);
            for (int k = 0; k < nplanes; ++k, ++ps) {
                if (lt == 0 && k + 1 < nplanes) issue(k + 1, ps + 1);
                const int slot = ps & (NPLANES - 1);
                { TS_T0();
                ok = tc::mbar_wait(tc::smem_u32(&bar_praw[slot]), (ps >> 2) & 1) && ok;             // the raw lines have landed
                TS_ACC(1); }
                TS_T0();
                uint8_t* pl = planes + slot * PLANE_B;
                // warp w permutes lines w, w+4, ... on its own: lane owns chunks lane, lane+32, lane+64, lane+96 of the line
                // (same channel quad, z = lane/4 + 8j), reads them, __syncwarp, writes them to their permuted places
                for (int line = lw; line < nlines; line += LOAD_WARPS) {
                    uint8_t* ln = pl + line * LINE_B;
                    float4 v[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (lane + 32 * j < nch) v[j] = *reinterpret_cast<const float4*>(ln + (lane + 32 * j) * 16);
                    __syncwarp();                                   // every lane holds its chunks of the line: safe to overwrite it
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (lane + 32 * j < nch) {
                            float4 o = v[j];
                            if (!(dbg & 8)) {
                                o.x = fmaxf(fmaf(sa.x, o.x, sb.x), 0.f); o.y = fmaxf(fmaf(sa.y, o.y, sb.y), 0.f);
                                o.z = fmaxf(fmaf(sa.z, o.z, sb.z), 0.f); o.w = fmaxf(fmaf(sa.w, o.w, sb.w), 0.f);
                            }
                            const int zz = (lane >> 2) + 8 * j;
                            *reinterpret_cast<float4*>(ln + ((zz & 1) * 4 + sq) * QP + (zz >> 1) * 16) = o;
                        }
                }
                warp_arrive(tc::smem_u32(&bar_pfull[slot]));
                TS_ACC(2);
            }
        }
        if ((dbg & 32) && blockIdx.x == 0 && lt == 0) { g_ts_prof[0] = acc_[0]; g_ts_prof[1] = acc_[1]; g_ts_prof[2] = acc_[2]; }
    } else if (warp < PROD_WARPS) {
        // =================================================================================== producers
        const int group = warp >> 2, q4 = warp & 3;
        const int m = q4 * 32 + lane, r = m >> 4, z2 = m & 15;
        const uint32_t lane_addr = tmem + ((uint32_t)(q4 * 32) << 16);
        uint32_t cc = 0, ps0 = 0;                                   // chunk counter; plane sequence number of the item's first plane
        unsigned long long acc_[4] = {0, 0, 0, 0};
        for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
            const Item it = decode_item(item, G2, NYB, NXG);
            const int ntiles = it.x_end - it.x_begin;
            for (int x2 = it.x_begin; x2 < it.x_end; ++x2) {
                const int t = x2 - it.x_begin;
                for (int c = 0; c < 9; ++c, ++cc) {
                    const int dx = c / 3, dy = c - 3 * dx, buf = cc & (NBUF - 1);
                    const uint32_t pseq = ps0 + 2 * t + dx;
                    const int slot = pseq & (NPLANES - 1);
                    if ((int)(cc & 1) == group) {
                    { TS_T0();
                    ok = tc::mbar_wait(tc::smem_u32(&bar_pfull[slot]), (pseq >> 2) & 1) && ok;         // the plane is staged
                    TS_ACC(0); }
                    { TS_T0();
                    ok = tc::mbar_wait(tc::smem_u32(&bar_free[buf]), ((cc >> 2) & 1) ^ 1) && ok;     // MMAs that read this buffer are done
                    TS_ACC(1); }
                    TS_T0();
                    tc::tc_fence_after();
                    const uint8_t* lp = planes + slot * PLANE_B + (2 * r + dy) * LINE_B;
                    const uint32_t abase = lane_addr + (uint32_t)(buf * A_COLS);
                    if (!(dbg & 1))
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
                    warp_arrive(tc::smem_u32(&bar_full[buf]));
                    TS_ACC(2);
                    }
                    // hand a plane back to the loaders after its last use (in program order this thread's own chunks of that
                    // dx are done): planes 2t and 2t+1 die with tile t, plane 2t+2 only with the last tile of the item
                    if (dy == 2 && (dx < 2 || t == ntiles - 1)) warp_arrive(tc::smem_u32(&bar_pfree[slot]));
                }
            }
            ps0 += 2 * ntiles + 1;
        }
        if ((dbg & 32) && blockIdx.x == 0 && tid == 0) { g_ts_prof[4] = acc_[0]; g_ts_prof[5] = acc_[1]; g_ts_prof[6] = acc_[2]; }
    } else if (warp == PROD_WARPS + EPI_WARPS) {
        // =================================================================================== MMA issuer
        const uint32_t idesc = tc::make_idesc_tf32(128, C);
        const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);       // provably warp-uniform copies of every MMA operand source
        const uint64_t dwh0 = tc::make_smem_desc(tc::smem_u32(w_hi), 128, W_SBO), dwl0 = tc::make_smem_desc(tc::smem_u32(w_lo), 128, W_SBO);
        uint32_t cc = 0, tt = 0;
        unsigned long long acc_[4] = {0, 0, 0, 0};
        for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
            const Item it = decode_item(item, G2, NYB, NXG);
            for (int x2 = it.x_begin; x2 < it.x_end; ++x2, ++tt) {
                const uint32_t acc = tt & 1, tmem_d = tm + D_COL0 + acc * C;
                { TS_T0();
                ok = tc::mbar_wait(tc::smem_u32(&bar_accfree[acc]), ((tt >> 1) & 1) ^ 1) && ok;     // epilogue drained this accumulator
                TS_ACC(0); }
                for (int c = 0; c < 9; ++c, ++cc) {
                    const int buf = cc & (NBUF - 1);
                    { TS_T0();
                    ok = tc::mbar_wait(tc::smem_u32(&bar_full[buf]), (cc >> 2) & 1) && ok;
                    TS_ACC(1); }
                    TS_T0();
                    tc::tc_fence_after();
                    if (tc::elect_one()) {
                        const uint32_t a_hi = tm + (uint32_t)(buf * A_COLS), a_lo = a_hi + 48;
                        if (!(dbg & 2))
#pragma unroll
                        for (int ks = 0; ks < 6; ++ks) {
                            // k0 = c*48 + ks*8 -> start address + (k0/4)*128 B; the descriptor holds address >> 4 in its low bits
                            const uint64_t wo = (uint64_t)((c * 12 + ks * 2) * 8);
                            const uint64_t dwh = dwh0 + wo, dwl = dwl0 + wo;
                            mma_tf32_ts(tmem_d, a_lo + ks * 8, dwh, idesc, (c == 0 && ks == 0) ? 0u : 1u);      // small terms first
                            mma_tf32_ts(tmem_d, a_hi + ks * 8, dwl, idesc, 1u);
                            mma_tf32_ts(tmem_d, a_hi + ks * 8, dwh, idesc, 1u);
                        }
                        tc::mma_commit(tc::smem_u32(&bar_free[buf]));
                        if (c == 8) tc::mma_commit(tc::smem_u32(&bar_accfull[acc]));
                    }
                    __syncwarp();
                    TS_ACC(2);
                }
            }
        }
        if ((dbg & 32) && blockIdx.x == 0 && lane == 0) { g_ts_prof[8] = acc_[0]; g_ts_prof[9] = acc_[1]; g_ts_prof[10] = acc_[2]; }
    } else {
        // =================================================================================== epilogue
        const int ew = warp - PROD_WARPS;                            // == warp % 4: the TMEM lane quarter this warp may read
        const int m = ew * 32 + lane;
        uint32_t tt = 0;
        unsigned long long acc_[4] = {0, 0, 0, 0};
        const long long tstart = clock64();
        for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
            const Item it = decode_item(item, G2, NYB, NXG);
            for (int x2 = it.x_begin; x2 < it.x_end; ++x2, ++tt) {
                const uint32_t acc = tt & 1;
                { TS_T0();
                ok = tc::mbar_wait(tc::smem_u32(&bar_accfull[acc]), (tt >> 1) & 1) && ok;
                TS_ACC(0); }
                TS_T0();
                tc::tc_fence_after();
                float v[C];
                tc::tmem_ld16(tmem + D_COL0 + acc * C + ((uint32_t)(ew * 32) << 16), v);
                tc::tc_fence_before();
                warp_arrive(tc::smem_u32(&bar_accfree[acc]));        // the next tile but one may overwrite this accumulator
                // Through shared memory, transposed to [channel][position in the run]: for one channel the tile's valid voxels
                // (nrows x G2) are ONE contiguous run of y2 (pos = (x2*G2 + y2)*G2 + z2), so the stores below are coalesced
                // 128-byte lines instead of sixteen strided 60-byte pieces per warp, and the BatchNorm2 record of the tile
                // costs 10 shuffles per channel instead of a 160-shuffle reduction per warp.
                const int rr = m >> 4, zz = m & 15;
                const int nrows = min(ROWS, G2 - it.y0), nval = nrows * G2;
                if (rr < nrows && zz < G2) {
                    const int j = rr * G2 + zz;                      // position inside the tile's contiguous run
#pragma unroll
                    for (int c = 0; c < C; ++c) ot[c][j] = v[c] + bs[c];
                }
                bar_named(2, EPI_WARPS * 32);
                const int64_t pos0 = ((int64_t)x2 * G2 + it.y0) * G2;
                if (!(dbg & 4)) {
                    // warp ew owns channels 4 ew .. 4 ew + 3; its 32 lanes sweep a channel's run: 128-byte coalesced stores
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int c = 4 * ew + k;
                        float* dstc = y2 + ((int64_t)it.b * C + c) * P2 + pos0;
                        float sv = 0.f;
                        for (int j = lane; j < nval; j += 32) {
                            const float x = ot[c][j];
                            dstc[j] = x;
                            sv += x;
                        }
                        if (part) {
                            sv = warp_sum(sv);
                            const float mean = sv / (float)nval;
                            float d2 = 0.f;
                            for (int j = lane; j < nval; j += 32) {
                                const float d = ot[c][j] - mean;
                                d2 = fmaf(d, d, d2);
                            }
                            d2 = warp_sum(d2);
                            if (lane == 0) {
                                const int tile = ((it.b * NYB + it.y0 / ROWS) * G2 + x2);
                                float* pr = part + (int64_t)tile * PART_STRIDE;
                                pr[c] = mean; pr[C + c] = d2;
                                if (c == 0) pr[2 * C] = (float)nval;
                            }
                        }
                    }
                }
                bar_named(2, EPI_WARPS * 32);                        // ot[] consumed before the next tile rewrites it
                TS_ACC(1);
            }
        }
        if ((dbg & 32) && blockIdx.x == 0 && m == 0) { g_ts_prof[12] = acc_[0]; g_ts_prof[13] = acc_[1]; g_ts_prof[14] = (unsigned long long)(clock64() - tstart); g_ts_prof[15] = tt; }
    }
    if (!ok) { asm volatile("trap;"); }                              // a bounded wait expired: fail loudly
    tc::tc_fence_before();
    __syncthreads();
    if (warp == PROD_WARPS + EPI_WARPS) tc::tmem_dealloc(tmem, 512);
}


// ------------------------------------------------------------------------------------------------ data gradient
// dact1[b, v, ci] = sum over taps t with (v - t) even and p = (v - t)/2 inside the output grid of sum_co dy2[b, p, co] * W[co, ci, t],
// then g1 = dact1 * [bn1(y1) > 0] (channels-last) and the BatchNorm1-backward sums (sum g1, sum g1 * xhat1).
// The 8 parity classes of the conv1-output grid have fixed tap sets (1 / 2 / 4 / 8 taps): a tile = 128 voxels of ONE class, so
// all rows share the MMA sequence; per tap a producer thread moves its row's 16 dy2 channels (64 contiguous bytes of the
// channels-last dy2) registers -> TMEM, zeros where p falls outside the grid.  Same warp roles and barriers as the forward
// kernel (chunk = one tap, K = 16; eight 32-column A buffers).  The epilogue warps read the y1 row for the ReLU mask, write
// the g1 row and keep the BatchNorm sums in registers for the whole CTA: one record per CTA, merged in a fixed order later.
constexpr int DG_NBUF = 8, DG_ACOLS = 32, DG_DCOL0 = DG_NBUF * DG_ACOLS;

struct DgClass { int px, py, pz, nx, ny, nz, n, tiles, tile0, ntaps; };
struct DgTable { DgClass c[8]; int tiles_per_sample; };

__host__ __device__ inline DgTable make_dg_table(int G1) {
    DgTable t;
    int acc = 0;
    const int ne = (G1 + 1) / 2, no = G1 / 2;
    for (int k = 0; k < 8; ++k) {
        DgClass& c = t.c[k];
        c.px = (k >> 2) & 1; c.py = (k >> 1) & 1; c.pz = k & 1;
        c.nx = c.px ? no : ne; c.ny = c.py ? no : ne; c.nz = c.pz ? no : ne;
        c.n = c.nx * c.ny * c.nz;
        c.tiles = (c.n + 127) / 128;
        c.tile0 = acc;
        acc += c.tiles;
        c.ntaps = (c.px ? 1 : 2) * (c.py ? 1 : 2) * (c.pz ? 1 : 2);
    }
    t.tiles_per_sample = acc;
    return t;
}

struct DgItem { int b, cls, tile; };
__device__ __forceinline__ DgItem dg_decode(int item, const DgTable& t) {
    DgItem it;
    it.b = item / t.tiles_per_sample;
    const int r = item - it.b * t.tiles_per_sample;
    it.cls = 0;
#pragma unroll
    for (int k = 1; k < 8; ++k) if (r >= t.c[k].tile0) it.cls = k;
    it.tile = r - t.c[it.cls].tile0;
    return it;
}
// tap number q (0 .. ntaps-1) of a class -> per-axis kernel offsets: an odd coordinate uses offset 1, an even one 0 and 2
__device__ __forceinline__ void dg_tap(const DgClass& c, int q, int& i, int& j, int& l) {
    const int nl = c.pz ? 1 : 2, nj = c.py ? 1 : 2;
    const int ql = q % nl, qj = (q / nl) % nj, qi = q / (nl * nj);
    i = c.px ? 1 : 2 * qi; j = c.py ? 1 : 2 * qj; l = c.pz ? 1 : 2 * ql;
}

__global__ void __launch_bounds__(TS_THREADS, 1)
conv2_dgrad_ts_kernel(const float* __restrict__ dy2cl, const float* __restrict__ w, const float* __restrict__ y1,
                      const float* __restrict__ stat1, float* __restrict__ g1, float* __restrict__ bpart, int G1, int G2,
                      int total_items) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* w_hi = smem;
    uint8_t* w_lo = smem + W_TILE;
    __shared__ __align__(8) uint64_t bar_full[DG_NBUF], bar_free[DG_NBUF], bar_accfull[2], bar_accfree[2];
    __shared__ uint32_t tmem_base_smem;
    __shared__ float st[4 * C];
    __shared__ float red[EPI_WARPS][2 * C];
    __shared__ DgTable tab;
    const int tid = threadIdx.x, warp = tc::uniform_warp_index(), lane = tid & 31;
    const int P1 = G1 * G1 * G1, P2 = G2 * G2 * G2;

    // B operand = W^T per tap: rows n = ci, k = tap*16 + co
    for (int e = tid; e < C * KW; e += TS_THREADS) {
        const int ci = e / KW, k = e - ci * KW, tap = k >> 4, co = k & 15;
        float h, l;
        tc::split_tf32(w[(co * C + ci) * TAPS + tap], h, l);
        const uint32_t off = (uint32_t)((ci >> 3) * W_SBO + (k >> 2) * 128 + (ci & 7) * 16 + (k & 3) * 4);
        *reinterpret_cast<float*>(w_hi + off) = h;
        *reinterpret_cast<float*>(w_lo + off) = l;
    }
    if (tid < 4 * C) st[tid] = stat1[tid];
    if (tid == 0) {
        tab = make_dg_table(G1);
        for (int i = 0; i < DG_NBUF; ++i) { tc::mbar_init(tc::smem_u32(&bar_full[i]), 4); tc::mbar_init(tc::smem_u32(&bar_free[i]), 1); }
        for (int i = 0; i < 2; ++i) { tc::mbar_init(tc::smem_u32(&bar_accfull[i]), 1); tc::mbar_init(tc::smem_u32(&bar_accfree[i]), EPI_WARPS); }
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
        const int group = warp >> 2, q4 = warp & 3, m = q4 * 32 + lane;
        const uint32_t lane_addr = tmem + ((uint32_t)(q4 * 32) << 16);
        // The taps of the CTA's tiles form one stream; this group owns every second one (chunk counter parity).  The 64 bytes
        // of the NEXT owned tap are requested before the current one is split and stored, so the L2 latency of the dy2 rows
        // overlaps the register -> TMEM work instead of heading every tap.
        struct Cur { int item, q, ntaps; uint32_t cc; bool done; };
        auto settle = [&](Cur& s) {                                  // move to the first owned tap at or after (item, q)
            while (!s.done) {
                if (s.q >= s.ntaps) {
                    s.item += gridDim.x; s.q = 0;
                    if (s.item >= total_items) { s.done = true; break; }
                    s.ntaps = tab.c[dg_decode(s.item, tab).cls].ntaps;
                    continue;
                }
                if ((int)(s.cc & 1) == group) break;
                ++s.q; ++s.cc;
            }
        };
        auto fetch = [&](const Cur& s, float4 (&x)[4]) {
            const DgItem it = dg_decode(s.item, tab);
            const DgClass c = tab.c[it.cls];
            const int v = it.tile * 128 + m;
            const bool vrow = v < c.n;
            const int vc = vrow ? v : 0;
            const int zi = vc % c.nz, t2 = vc / c.nz, yi = t2 % c.ny, xi = t2 / c.ny;
            int i, j, l;
            dg_tap(c, s.q, i, j, l);
            // p = (v - tap) / 2 per axis: odd coordinate (offset 1) -> index itself; even: offset 0 -> index, offset 2 -> index - 1
            const int px = xi - (i >> 1), py = yi - (j >> 1), pz = zi - (l >> 1);
            const bool valid = vrow && px >= 0 && px < G2 && py >= 0 && py < G2 && pz >= 0 && pz < G2;
            if (valid) {
                const float4* src = reinterpret_cast<const float4*>(dy2cl + ((int64_t)it.b * P2 + (int64_t)(px * G2 + py) * G2 + pz) * C);
#pragma unroll
                for (int u = 0; u < 4; ++u) x[u] = __ldg(src + u);
            } else {
#pragma unroll
                for (int u = 0; u < 4; ++u) x[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        Cur cur;
        cur.item = blockIdx.x; cur.q = 0; cur.cc = 0; cur.done = cur.item >= total_items;
        cur.ntaps = cur.done ? 0 : tab.c[dg_decode(cur.item, tab).cls].ntaps;
        settle(cur);
        float4 xa[4], xb[4];
        if (!cur.done) fetch(cur, xa);
        while (!cur.done) {
            Cur nxt = cur;
            ++nxt.q; ++nxt.cc;
            settle(nxt);
            if (!nxt.done) fetch(nxt, xb);
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float xv[4] = {xa[u].x, xa[u].y, xa[u].z, xa[u].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const uint32_t h = __float_as_uint(xv[e]) & 0xffffe000u;
                    hi[4 * u + e] = h;
                    lo[4 * u + e] = __float_as_uint(xv[e] - __uint_as_float(h));
                }
            }
            const int buf = cur.cc & (DG_NBUF - 1);
            ok = tc::mbar_wait(tc::smem_u32(&bar_free[buf]), ((cur.cc >> 3) & 1) ^ 1) && ok;
            tc::tc_fence_after();
            const uint32_t abase = lane_addr + (uint32_t)(buf * DG_ACOLS);
            tmem_st16(abase, hi);
            tmem_st16(abase + 16, lo);
            tmem_st_wait();
            tc::tc_fence_before();
            warp_arrive(tc::smem_u32(&bar_full[buf]));
            cur = nxt;
#pragma unroll
            for (int u = 0; u < 4; ++u) xa[u] = xb[u];
        }
    } else if (warp == PROD_WARPS + EPI_WARPS) {
        // =================================================================================== MMA issuer
        const uint32_t idesc = tc::make_idesc_tf32(128, C);
        const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
        const uint64_t dwh0 = tc::make_smem_desc(tc::smem_u32(w_hi), 128, W_SBO), dwl0 = tc::make_smem_desc(tc::smem_u32(w_lo), 128, W_SBO);
        uint32_t cc = 0, tt = 0;
        for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++tt) {
            const DgItem it = dg_decode(item, tab);
            const DgClass c = tab.c[it.cls];
            const uint32_t acc = tt & 1, tmem_d = tm + DG_DCOL0 + acc * C;
            ok = tc::mbar_wait(tc::smem_u32(&bar_accfree[acc]), ((tt >> 1) & 1) ^ 1) && ok;
            for (int q = 0; q < c.ntaps; ++q, ++cc) {
                const int buf = cc & (DG_NBUF - 1);
                int i, j, l;
                dg_tap(c, q, i, j, l);
                const int tap = (i * 3 + j) * 3 + l;
                ok = tc::mbar_wait(tc::smem_u32(&bar_full[buf]), (cc >> 3) & 1) && ok;
                tc::tc_fence_after();
                if (tc::elect_one()) {
                    const uint32_t a_hi = tm + (uint32_t)(buf * DG_ACOLS), a_lo = a_hi + 16;
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) {
                        const uint64_t wo = (uint64_t)((tap * 4 + ks * 2) * 8);                  // k0 = tap*16 + ks*8
                        const uint64_t dwh = dwh0 + wo, dwl = dwl0 + wo;
                        mma_tf32_ts(tmem_d, a_lo + ks * 8, dwh, idesc, (q == 0 && ks == 0) ? 0u : 1u);
                        mma_tf32_ts(tmem_d, a_hi + ks * 8, dwl, idesc, 1u);
                        mma_tf32_ts(tmem_d, a_hi + ks * 8, dwh, idesc, 1u);
                    }
                    tc::mma_commit(tc::smem_u32(&bar_free[buf]));
                    if (q == c.ntaps - 1) tc::mma_commit(tc::smem_u32(&bar_accfull[acc]));
                }
                __syncwarp();
            }
        }
    } else {
        // =================================================================================== epilogue
        const int ew = warp - PROD_WARPS, m = ew * 32 + lane;
        float s1[C], s2[C];
#pragma unroll
        for (int k = 0; k < C; ++k) { s1[k] = 0.f; s2[k] = 0.f; }
        // element offset of this thread's row in y1 / g1 for a work item, or -1 past the end of the class / of the items
        auto row_of = [&](int item) -> int64_t {
            if (item >= total_items) return -1;
            const DgItem it = dg_decode(item, tab);
            const DgClass c = tab.c[it.cls];
            const int v = it.tile * 128 + m;
            if (v >= c.n) return -1;
            const int zi = v % c.nz, t2 = v / c.nz, yi = t2 % c.ny, xi = t2 / c.ny;
            return ((int64_t)it.b * P1 + ((int64_t)(2 * xi + c.px) * G1 + (2 * yi + c.py)) * G1 + (2 * zi + c.pz)) * C;
        };
        // y1 comes from HBM (it does not fit the L2): rows of the tile four items ahead are prefetched into the L2, rows of the
        // next item into registers, so the ReLU mask never waits for DRAM once the accumulator is ready
        const int step = gridDim.x;
        for (int a = 1; a <= 3; ++a) {
            const int64_t r = row_of(blockIdx.x + a * step);
            if (r >= 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(y1 + r));
        }
        float4 ya[4], yb[4];
        int64_t row = row_of(blockIdx.x);
        if (row >= 0) {
#pragma unroll
            for (int u = 0; u < 4; ++u) ya[u] = __ldg(reinterpret_cast<const float4*>(y1 + row) + u);
        }
        uint32_t tt = 0;
        for (int item = blockIdx.x; item < total_items; item += step, ++tt) {
            const int64_t far = row_of(item + 4 * step), nrow = row_of(item + step);
            if (far >= 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(y1 + far));
            if (nrow >= 0) {
#pragma unroll
                for (int u = 0; u < 4; ++u) yb[u] = __ldg(reinterpret_cast<const float4*>(y1 + nrow) + u);
            }
            const uint32_t acc = tt & 1;
            ok = tc::mbar_wait(tc::smem_u32(&bar_accfull[acc]), (tt >> 1) & 1) && ok;
            tc::tc_fence_after();
            float a[C];
            tc::tmem_ld16(tmem + DG_DCOL0 + acc * C + ((uint32_t)(ew * 32) << 16), a);
            tc::tc_fence_before();
            warp_arrive(tc::smem_u32(&bar_accfree[acc]));
            if (row >= 0) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float y4[4] = {ya[u].x, ya[u].y, ya[u].z, ya[u].w};
                    float o4[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int ch = 4 * u + e;
                        const float pre = fmaf(st[2 * C + ch], y4[e], st[3 * C + ch]);
                        const float gv = pre > 0.f ? a[ch] : 0.f;
                        o4[e] = gv;
                        s1[ch] += gv;
                        s2[ch] = fmaf(gv, (y4[e] - st[ch]) * st[C + ch], s2[ch]);
                    }
                    reinterpret_cast<float4*>(g1 + row)[u] = make_float4(o4[0], o4[1], o4[2], o4[3]);
                }
            }
            row = nrow;
#pragma unroll
            for (int u = 0; u < 4; ++u) ya[u] = yb[u];
        }
        // one (sum g1, sum g1 * xhat1) record per CTA: lanes -> warps -> block, fixed order
#pragma unroll
        for (int k = 0; k < C; ++k) {
            const float r1 = warp_sum(s1[k]), r2 = warp_sum(s2[k]);
            if (lane == 0) { red[ew][k] = r1; red[ew][C + k] = r2; }
        }
        bar_named(2, EPI_WARPS * 32);
        if (m < 2 * C) bpart[(int64_t)blockIdx.x * 2 * C + m] = (red[0][m] + red[1][m]) + (red[2][m] + red[3][m]);
    }
    if (!ok) { asm volatile("trap;"); }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == PROD_WARPS + EPI_WARPS) tc::tmem_dealloc(tmem, 512);
}

}  // namespace

extern "C" int gnbv_debug_ts_profile(unsigned long long* out32) {
    return cudaMemcpyFromSymbol(out32, g_ts_prof, sizeof(unsigned long long) * 32) == cudaSuccess ? 0 : -2;
}

bool conv2_ts_supported(int G1, int G2) { return G2 >= 1 && G2 <= ZP && G1 >= 2 * G2 + 1 && G1 <= 2 * ZP; }
int conv2_ts_tiles(int B, int G2) { return B * G2 * (int)ceil_div(G2, ROWS); }

int launch_conv2_fwd_ts(const float* y1, const float* stat1, const float* w, const float* bias, float* y2, float* part, int B,
                        int G1, int G2, cudaStream_t stream) {
    GNBV_REQUIRE(conv2_ts_supported(G1, G2), "conv2_fwd_ts: unsupported grid (G1=%d G2=%d)", G1, G2);
    { int rc_ = ensure_dyn_smem(conv2_fwd_ts_kernel, TS_SMEM); if (rc_) return rc_; }
    const int items = B * (int)ceil_div(G2, ROWS) * (int)ceil_div(G2, XT);
    const int grid = std::min(items, 148);
    static const int dbg = getenv("GNBV_TS_DEBUG") ? atoi(getenv("GNBV_TS_DEBUG")) : 0;      // ablation switches (profiling only)
    conv2_fwd_ts_kernel<<<grid, FWD_THREADS, TS_SMEM, stream>>>(y1, stat1, w, bias, y2, part, G1, G2, items, dbg);
    GNBV_LAUNCH_CHECK("conv2_fwd_ts_kernel");
    return GNBV_OK;
}

int conv2_dgrad_ts_records(int B, int G1) { return std::min(B * make_dg_table(G1).tiles_per_sample, 148); }

int launch_conv2_dgrad_ts(const float* dy2cl, const float* w, const float* y1, const float* stat1, float* g1, float* bpart, int B,
                          int G1, int G2, cudaStream_t stream) {
    GNBV_REQUIRE(conv2_ts_supported(G1, G2), "conv2_dgrad_ts: unsupported grid (G1=%d G2=%d)", G1, G2);
    const size_t smem = 2 * W_TILE;
    { int rc_ = ensure_dyn_smem(conv2_dgrad_ts_kernel, smem); if (rc_) return rc_; }
    const int items = B * make_dg_table(G1).tiles_per_sample;
    const int grid = conv2_dgrad_ts_records(B, G1);
    conv2_dgrad_ts_kernel<<<grid, TS_THREADS, smem, stream>>>(dy2cl, w, y1, stat1, g1, bpart, G1, G2, items);
    GNBV_LAUNCH_CHECK("conv2_dgrad_ts_kernel");
    return GNBV_OK;
}

}  // namespace gnbv
