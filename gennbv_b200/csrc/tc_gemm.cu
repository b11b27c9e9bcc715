// tc_gemm.cu -- split-precision (3xTF32) tcgen05 GEMM: C[M,N] = epilogue(A * B) with fp32-grade accuracy on the
// 5th-generation tensor cores (see tc.cuh for the numerics and the shared-memory layout).
//
// One CTA = one 128 x BN output tile over one K split.  256 threads stage A and B tiles (arbitrary strides, either
// dimension contiguous) from global memory into the canonical UMMA layout, splitting every value into (hi, lo) on the
// way; one elected thread issues 3 x BK/8 tcgen05.mma per stage into a TMEM accumulator (128 lanes x BN columns);
// two shared-memory stages overlap the loads of stage i+1 with the MMAs of stage i (tcgen05.commit -> mbarrier);
// warps 0-3 read the accumulator back with tcgen05.ld and store either the final tile (bias / ReLU fused) or a split-K
// partial that splitk_reduce_kernel (gemm.cu) sums in a fixed order.
#include "gemm.cuh"
#include "tc.cuh"

#include <algorithm>

namespace gnbv {

constexpr int TC_BM = 128;

struct TcGemmArgs {
    const float* A; int64_t sa_m, sa_k;
    const float* B; int64_t sb_k, sb_n;
    float* C; int64_t ldc;
    float* ws;
    int M, N, K, Kc, splits;
    const float* bias; int relu;
    int vecA, vecB;
    int* err;
};

// Operand staging: an [R x BK] tile (rows r0.., K range k0..) goes to shared memory as (hi, lo) K-major UMMA tiles.
// MODE 0: K contiguous in global (element (r,k) at base[r*ld_r + k]); MODE 1: rows contiguous (base[r + k*ld_k]).

// ---- warp-specialised pipeline --------------------------------------------------------------------------------------
// warps 0-7   producers: per K stage (BK = 32) request the stage's A and B elements from global memory into REGISTERS first
//             (12 x 16 B per thread in flight, issued before the shared-memory slot is known to be free), then wait for the
//             slot, split every value into (hi, lo) and store both into the canonical K-major UMMA layout; one mbarrier
//             arrival per warp.
// warp 8      one elected thread issues the 12 tcgen05.mma of a stage (3xTF32) and tcgen05.commit's the slot back.
// warps 9-12  epilogue: TMEM -> registers -> shared-memory transpose -> coalesced global stores, on the accumulator buffer the MMA
//             warp is not writing (two TMEM accumulators of BN columns each).
// The loads of stage s+1 overlap the MMAs of stage s; the load latency of a stage itself is exposed to the producers (DESIGN.md
// section 10, item 1).
constexpr int WS_PROD = 256, WS_EPI = 128, WS_THREADS = WS_PROD + 32 + WS_EPI;

template <int R>
struct TileRegs { float4 v[R * (tc::BK / 4) / WS_PROD]; };

template <int R, int MODE>
__device__ __forceinline__ void tile_fetch(const float* __restrict__ base, int64_t ld, int r0, int rows, int k0, int k_end, bool vec,
                                           TileRegs<R>& t) {
    constexpr int NIT = R * (tc::BK / 4) / WS_PROD;
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
        const int u = threadIdx.x + it * WS_PROD;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (MODE == 0) {
            const int r_low = u & 7, k4 = (u >> 3) & 7, r = (u >> 6) * 8 + r_low;
            const int gr = r0 + r, gk = k0 + k4 * 4;
            if (gr < rows) {
                const float* p = base + (int64_t)gr * ld + gk;
                if (vec && gk + 3 < k_end) v = __ldg(reinterpret_cast<const float4*>(p));
                else {
                    if (gk < k_end) v.x = __ldg(p);
                    if (gk + 1 < k_end) v.y = __ldg(p + 1);
                    if (gk + 2 < k_end) v.z = __ldg(p + 2);
                    if (gk + 3 < k_end) v.w = __ldg(p + 3);
                }
            }
        } else {
            // rows contiguous in global memory: a warp covers ONE core matrix per step (lane -> row rl = lane / 4 of the
            // 8-row group, k = lane % 4 of the 4-k chunk), four scalars per float4 slot.  The 32 stores of a step then hit 32
            // different banks (a core matrix is 128 contiguous bytes) and the loads are four fully used 32-byte sectors;
            // float4 loads along the rows would put 16 lanes on one bank at store time (measured 9x slower).
            const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, rl = lane >> 2, kl = lane & 3;
            float x[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int cm = (it * 4 + e) * (WS_PROD / 32) + w, rg = cm >> 3, kc = cm & 7;
                const int gr = r0 + rg * 8 + rl, gk = k0 + kc * 4 + kl;
                x[e] = (gk < k_end && gr < rows) ? __ldg(base + (int64_t)gk * ld + gr) : 0.f;
            }
            v = make_float4(x[0], x[1], x[2], x[3]);
        }
        t.v[it] = v;
    }
}

template <int R, int MODE>
__device__ __forceinline__ void tile_store(const TileRegs<R>& t, uint8_t* hi, uint8_t* lo) {
    constexpr int NIT = R * (tc::BK / 4) / WS_PROD;
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
        const int u = threadIdx.x + it * WS_PROD;
        const float4 v = t.v[it];
        if (MODE == 0) {
            const int r_low = u & 7, k4 = (u >> 3) & 7, r = (u >> 6) * 8 + r_low;
            float4 h, l;
            tc::split_tf32(v.x, h.x, l.x); tc::split_tf32(v.y, h.y, l.y); tc::split_tf32(v.z, h.z, l.z); tc::split_tf32(v.w, h.w, l.w);
            const uint32_t off = tc::tile_offset(r, k4 * 4);
            *reinterpret_cast<float4*>(hi + off) = h;
            *reinterpret_cast<float4*>(lo + off) = l;
        } else {
            const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, rl = lane >> 2, kl = lane & 3;
            const float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int cm = (it * 4 + e) * (WS_PROD / 32) + w, rg = cm >> 3, kc = cm & 7;
                float h, l;
                tc::split_tf32(x[e], h, l);
                const uint32_t off = tc::tile_offset(rg * 8 + rl, kc * 4 + kl);
                *reinterpret_cast<float*>(hi + off) = h;
                *reinterpret_cast<float*>(lo + off) = l;
            }
        }
    }
}

template <int BN, int A_MODE, int B_MODE>   // A_MODE 0: k contiguous, 1: m contiguous.  B_MODE 0: n contiguous, 1: k contiguous
__global__ void __launch_bounds__(WS_THREADS, 1)
tc_gemm_kernel(TcGemmArgs g, int tiles_n, int tiles_mn, int total_items) {
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr uint32_t A_TILE = TC_BM * tc::BK * 4, B_TILE = BN * tc::BK * 4, STAGE = 2 * A_TILE + 2 * B_TILE;
    constexpr int NST = BN >= 256 ? 2 : 4;                // 96 KB / 48 KB per stage
    __shared__ __align__(8) uint64_t bar_full[NST], bar_empty[NST], bar_accfull[2], bar_accfree[2];
    __shared__ uint32_t tmem_base_smem;
    __shared__ float tr[WS_EPI / 32][32][17];              // epilogue transpose: 32 rows x 16 columns per warp (+1 pad)
    const int tid = threadIdx.x, warp = tc::uniform_warp_index(), lane = tid & 31;
    constexpr uint32_t TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;             // two accumulators
    constexpr int MMA_WARP = WS_PROD / 32, EPI_WARP0 = MMA_WARP + 1;

    if (tid == 0) {
        for (int i = 0; i < NST; ++i) { tc::mbar_init(tc::smem_u32(&bar_full[i]), WS_PROD / 32); tc::mbar_init(tc::smem_u32(&bar_empty[i]), 1); }
        for (int i = 0; i < 2; ++i) { tc::mbar_init(tc::smem_u32(&bar_accfull[i]), 1); tc::mbar_init(tc::smem_u32(&bar_accfree[i]), WS_EPI / 32); }
        tc::fence_mbar_init();
    }
    if (warp == MMA_WARP) tc::tmem_alloc(tc::smem_u32(&tmem_base_smem), TMEM_COLS);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_d = tmem_base_smem;
    bool ok = true;
    // work item -> (K split z, m tile, n tile); every role walks the same list (persistent CTA, round robin)
    auto decode = [&](int item, int& m0, int& n0, int& z, int& k_begin, int& k_end, int& nstages) {
        z = item / tiles_mn;
        const int r = item - z * tiles_mn, mt = r / tiles_n;
        m0 = mt * TC_BM; n0 = (r - mt * tiles_n) * BN;
        k_begin = z * g.Kc; k_end = min(g.K, k_begin + g.Kc);
        nstages = (k_end - k_begin + tc::BK - 1) / tc::BK;
    };

    if (warp < MMA_WARP) {
        // ------------------------------------------------------------------------------------------- producers
        TileRegs<TC_BM> ra;
        TileRegs<BN> rb;
        uint32_t sc = 0;                                       // stage counter of this CTA (same in the MMA warp)
        for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
            int m0, n0, z, k_begin, k_end, nstages;
            decode(item, m0, n0, z, k_begin, k_end, nstages);
            for (int s = 0; s < nstages; ++s, ++sc) {
                const int buf = sc % NST, k0 = k_begin + s * tc::BK;
                if (A_MODE == 0) tile_fetch<TC_BM, 0>(g.A, g.sa_m, m0, g.M, k0, k_end, g.vecA, ra);
                else             tile_fetch<TC_BM, 1>(g.A, g.sa_k, m0, g.M, k0, k_end, g.vecA, ra);
                if (B_MODE == 1) tile_fetch<BN, 0>(g.B, g.sb_n, n0, g.N, k0, k_end, g.vecB, rb);
                else             tile_fetch<BN, 1>(g.B, g.sb_k, n0, g.N, k0, k_end, g.vecB, rb);
                ok = tc::mbar_wait(tc::smem_u32(&bar_empty[buf]), ((sc / NST) & 1) ^ 1) && ok;      // the MMAs that read this slot are done
                uint8_t* st = smem + buf * STAGE;
                tile_store<TC_BM, A_MODE == 0 ? 0 : 1>(ra, st, st + A_TILE);
                tile_store<BN, B_MODE == 1 ? 0 : 1>(rb, st + 2 * A_TILE, st + 2 * A_TILE + B_TILE);
                tc::fence_async_smem();                 // generic-proxy smem writes -> visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(&bar_full[buf])) : "memory");
            }
        }
    } else if (warp == MMA_WARP) {
        // ------------------------------------------------------------------------------------------- MMA issuer
        const uint32_t idesc = tc::make_idesc_tf32(TC_BM, BN);
        const uint32_t td = __shfl_sync(0xffffffffu, tmem_d, 0);
        uint32_t sc = 0, tt = 0;
        for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++tt) {
            int m0, n0, z, k_begin, k_end, nstages;
            decode(item, m0, n0, z, k_begin, k_end, nstages);
            const uint32_t acc = tt & 1;
            ok = tc::mbar_wait(tc::smem_u32(&bar_accfree[acc]), ((tt >> 1) & 1) ^ 1) && ok;          // epilogue drained this accumulator
            for (int s = 0; s < nstages; ++s, ++sc) {
                const int buf = sc % NST;
                ok = tc::mbar_wait(tc::smem_u32(&bar_full[buf]), (sc / NST) & 1) && ok;
                tc::tc_fence_after();
                if (tc::elect_one()) {
                    const uint32_t a_hi = tc::smem_u32(smem + buf * STAGE), a_lo = a_hi + A_TILE, b_hi = a_hi + 2 * A_TILE, b_lo = b_hi + B_TILE;
                    tc::mma_stage_3xtf32(td + acc * BN, a_hi, a_lo, b_hi, b_lo, idesc, s == 0);
                    tc::mma_commit(tc::smem_u32(&bar_empty[buf]));
                    if (s == nstages - 1) tc::mma_commit(tc::smem_u32(&bar_accfull[acc]));
                }
                __syncwarp();
            }
        }
    } else {
        // ------------------------------------------------------------------------------------------- epilogue
        // each epilogue warp reads its TMEM lane quarter (warp % 4), 16 columns at a time, transposes them through shared memory so that
        // four lanes write one row's 64 contiguous bytes (full 32-byte sectors, 8 rows per store instruction)
        const int e = warp - EPI_WARP0, q = warp & 3;          // q: the TMEM lane quarter the hardware lets this warp read
        uint32_t tt = 0;
        for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++tt) {
            int m0, n0, z, k_begin, k_end, nstages;
            decode(item, m0, n0, z, k_begin, k_end, nstages);
            const uint32_t acc = tt & 1;
            ok = tc::mbar_wait(tc::smem_u32(&bar_accfull[acc]), (tt >> 1) & 1) && ok;
            tc::tc_fence_after();
            const uint32_t trow = tmem_d + acc * BN + ((uint32_t)(q * 32) << 16);
            float* out = g.splits == 1 ? g.C : g.ws + (int64_t)z * g.M * g.N;
            const int64_t ldo = g.splits == 1 ? g.ldc : g.N;
            const bool vec = (ldo % 4 == 0) && (((uintptr_t)out & 15) == 0);
#pragma unroll 1
            for (int c = 0; c < BN; c += 16) {
                float v[16];
                tc::tmem_ld16(trow + c, v);
                if (c + 16 >= BN) {                          // last read of this accumulator: hand it back before the stores
                    tc::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(&bar_accfree[acc])) : "memory");
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) tr[e][lane][j] = v[j];
                __syncwarp();
                const int cq = lane & 3, n = n0 + c + cq * 4;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int rl = (lane >> 2) + 8 * i, m = m0 + q * 32 + rl;
                    if (m < g.M && n < g.N) {
                        float x[4] = {tr[e][rl][cq * 4], tr[e][rl][cq * 4 + 1], tr[e][rl][cq * 4 + 2], tr[e][rl][cq * 4 + 3]};
                        if (g.splits == 1) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                if (g.bias && n + q < g.N) x[q] += __ldg(g.bias + n + q);
                                if (g.relu) x[q] = fmaxf(x[q], 0.f);
                            }
                        }
                        float* p = out + (int64_t)m * ldo + n;
                        if (vec && n + 3 < g.N) *reinterpret_cast<float4*>(p) = make_float4(x[0], x[1], x[2], x[3]);
                        else
#pragma unroll
                            for (int q = 0; q < 4; ++q) if (n + q < g.N) p[q] = x[q];
                    }
                }
                __syncwarp();
            }
        }
    }
    if (!ok) { if (lane == 0 && g.err) atomicExch(g.err, 1); asm volatile("trap;"); }      // a bounded wait expired: fail loudly
    tc::tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) tc::tmem_dealloc(tmem_d, TMEM_COLS);
}

// split-K reduction kernel lives in gemm.cu
void launch_splitk_reduce(const float* ws, float* C, int64_t ldc, int M, int N, int splits, const float* bias, int relu,
                          cudaStream_t stream);

static void tc_pick_splits(int M, int N, int K, int BN, int& splits, int& Kc) {
    int64_t tiles = ceil_div(M, TC_BM) * ceil_div(N, BN);
    int64_t want = std::max<int64_t>(1, ceil_div(148, tiles));
    int64_t max_splits = std::max<int64_t>(1, K / 256);          // at least 8 stages per split
    splits = (int)std::min(want, max_splits);
    Kc = (int)(ceil_div(ceil_div(K, splits), tc::BK) * tc::BK);
    splits = (int)ceil_div(K, Kc);
}

size_t tc_gemm_workspace_floats(int M, int N, int K) {
    int s, kc;
    tc_pick_splits(M, N, K, N >= 256 ? 256 : 64, s, kc);
    return (s > 1 ? (size_t)s * M * N : 0) + 64;       // + error flag slot
}

template <int BN>
static int launch_tc(const TcGemmArgs& g, int a_mode, int b_mode, dim3 grid, cudaStream_t stream) {
    constexpr size_t smem = (BN >= 256 ? 2 : 4) * (2 * TC_BM * tc::BK * 4 + 2 * BN * tc::BK * 4);
    const int items = (int)(grid.x * grid.y * grid.z);
    const int pgrid = std::min(items, 148);                 // persistent CTAs, one per SM
#define GNBV_TC_LAUNCH(AM, BMODE)                                                                                        \
    do {                                                                                                                 \
        GNBV_CUDA_CHECK(cudaFuncSetAttribute(tc_gemm_kernel<BN, AM, BMODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                             (int)smem));                                                                \
        tc_gemm_kernel<BN, AM, BMODE><<<pgrid, WS_THREADS, smem, stream>>>(g, (int)grid.x, (int)(grid.x * grid.y), items);                                            \
    } while (0)
    if (a_mode == 0 && b_mode == 0) GNBV_TC_LAUNCH(0, 0);
    else if (a_mode == 0 && b_mode == 1) GNBV_TC_LAUNCH(0, 1);
    else if (a_mode == 1 && b_mode == 0) GNBV_TC_LAUNCH(1, 0);
    else GNBV_TC_LAUNCH(1, 1);
#undef GNBV_TC_LAUNCH
    GNBV_LAUNCH_CHECK("tc_gemm_kernel");
    return GNBV_OK;
}

int launch_tc_gemm(const float* A, int64_t sa_m, int64_t sa_k, const float* B, int64_t sb_k, int64_t sb_n, float* C,
                   int64_t ldc, int M, int N, int K, const GemmEpilogue& ep, float* workspace, cudaStream_t stream) {
    GNBV_REQUIRE(A && B && C && workspace && M > 0 && N > 0 && K > 0, "tc_gemm: bad arguments (M=%d N=%d K=%d)", M, N, K);
    GNBV_REQUIRE(sa_k == 1 || sa_m == 1, "tc_gemm: A must be contiguous along m or k");
    GNBV_REQUIRE(sb_n == 1 || sb_k == 1, "tc_gemm: B must be contiguous along n or k");
    TcGemmArgs g;
    g.A = A; g.sa_m = sa_m; g.sa_k = sa_k; g.B = B; g.sb_k = sb_k; g.sb_n = sb_n; g.C = C; g.ldc = ldc;
    g.M = M; g.N = N; g.K = K; g.bias = ep.bias; g.relu = ep.relu;
    const int BN = N >= 256 ? 256 : 64;
    tc_pick_splits(M, N, K, BN, g.splits, g.Kc);
    g.err = reinterpret_cast<int*>(workspace);           // slot 0..63 of the workspace: sticky error flag (bounded waits)
    g.ws = workspace + 64;
    const int a_mode = (sa_k == 1) ? 0 : 1, b_mode = (sb_n == 1) ? 0 : 1;
    const int64_t lda = a_mode == 0 ? sa_m : sa_k, ldb = b_mode == 0 ? sb_k : sb_n;
    g.vecA = ((uintptr_t)A % 16 == 0) && (lda % 4 == 0);
    g.vecB = ((uintptr_t)B % 16 == 0) && (ldb % 4 == 0);
    dim3 grid((unsigned)ceil_div(N, BN), (unsigned)ceil_div(M, TC_BM), (unsigned)g.splits);
    int rc = BN == 256 ? launch_tc<256>(g, a_mode, b_mode, grid, stream) : launch_tc<64>(g, a_mode, b_mode, grid, stream);
    if (rc) return rc;
    if (g.splits > 1) launch_splitk_reduce(g.ws, C, ldc, M, N, g.splits, ep.bias, ep.relu, stream);
    return GNBV_OK;
}

}  // namespace gnbv

extern "C" size_t gnbv_tc_gemm_workspace_bytes(int M, int N, int K) { return gnbv::tc_gemm_workspace_floats(M, N, K) * 4; }

extern "C" int gnbv_tc_gemm(const float* A, int64_t sa_m, int64_t sa_k, const float* B, int64_t sb_k, int64_t sb_n,
                            float* C, int64_t ldc, int M, int N, int K, const float* bias, int relu, float* workspace,
                            size_t workspace_bytes, void* stream) {
    if (workspace_bytes < gnbv::tc_gemm_workspace_floats(M, N, K) * 4) {
        gnbv::set_error("gnbv_tc_gemm: workspace %zu B too small", workspace_bytes);
        return GNBV_E_WORKSPACE;
    }
    gnbv::GemmEpilogue ep;
    ep.bias = bias; ep.relu = relu;
    return gnbv::launch_tc_gemm(A, sa_m, sa_k, B, sb_k, sb_n, C, ldc, M, N, K, ep, workspace, (cudaStream_t)stream);
}
