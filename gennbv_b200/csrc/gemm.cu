// gemm.cu -- see gemm.cuh.
#include "gemm.cuh"
#include "mma.cuh"

#include <algorithm>
#include <cstdlib>

namespace gnbv {

constexpr int BM = 128, BN = 64, BK = 16, TM = 8, TN = 4, GEMM_THREADS = 256;
constexpr int AS_LD = BM + 4, BS_LD = BN + 4;

struct GemmArgs {
    const float* A; int64_t sa_m, sa_k;
    const float* B; int64_t sb_k, sb_n;
    float* C; int64_t ldc;
    float* ws;
    int M, N, K, Kc, splits;
    const float* bias; int relu;
    int vecA, vecB;
};

__device__ __forceinline__ float4 ld4_guard(const float* base, int64_t off, int64_t stride, int valid, bool vec) {
    // loads 4 elements base[off + j*stride], j < valid (others 0); vec => stride == 1 and 16 B aligned
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid >= 4 && vec) return __ldg(reinterpret_cast<const float4*>(base + off));
    if (valid > 0) v.x = __ldg(base + off);
    if (valid > 1) v.y = __ldg(base + off + stride);
    if (valid > 2) v.z = __ldg(base + off + 2 * stride);
    if (valid > 3) v.w = __ldg(base + off + 3 * stride);
    return v;
}

template <int A_MODE, int B_MODE>   // A_MODE 0: k contiguous, 1: m contiguous.  B_MODE 0: n contiguous, 1: k contiguous
__global__ void __launch_bounds__(GEMM_THREADS)
sgemm_kernel(GemmArgs g) {
    __shared__ __align__(16) float As[2][BK][AS_LD];
    __shared__ __align__(16) float Bs[2][BK][BS_LD];
    const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN, z = blockIdx.z;
    const int k_begin = z * g.Kc, k_end = min(g.K, k_begin + g.Kc);
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    float4 ra[2], rb;
    auto load_tiles = [&](int k0) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            int f = t + i * GEMM_THREADS;
            if (A_MODE == 0) {
                int row = f >> 2, kq = f & 3, m = m0 + row, k = k0 + kq * 4;
                int valid = (m < g.M) ? max(0, min(4, k_end - k)) : 0;
                ra[i] = ld4_guard(g.A, (int64_t)m * g.sa_m + (int64_t)k * g.sa_k, g.sa_k, valid, g.vecA);
            } else {
                int kk = f >> 5, mq = f & 31, m = m0 + mq * 4, k = k0 + kk;
                int valid = (k < k_end) ? max(0, min(4, g.M - m)) : 0;
                ra[i] = ld4_guard(g.A, (int64_t)m * g.sa_m + (int64_t)k * g.sa_k, g.sa_m, valid, g.vecA);
            }
        }
        if (B_MODE == 0) {
            int kk = t >> 4, nq = t & 15, n = n0 + nq * 4, k = k0 + kk;
            int valid = (k < k_end) ? max(0, min(4, g.N - n)) : 0;
            rb = ld4_guard(g.B, (int64_t)k * g.sb_k + (int64_t)n * g.sb_n, g.sb_n, valid, g.vecB);
        } else {
            int nn = t >> 2, kq = t & 3, n = n0 + nn, k = k0 + kq * 4;
            int valid = (n < g.N) ? max(0, min(4, k_end - k)) : 0;
            rb = ld4_guard(g.B, (int64_t)k * g.sb_k + (int64_t)n * g.sb_n, g.sb_k, valid, g.vecB);
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            int f = t + i * GEMM_THREADS;
            if (A_MODE == 0) {
                int row = f >> 2, kq = f & 3;
                As[buf][kq * 4 + 0][row] = ra[i].x; As[buf][kq * 4 + 1][row] = ra[i].y;
                As[buf][kq * 4 + 2][row] = ra[i].z; As[buf][kq * 4 + 3][row] = ra[i].w;
            } else {
                int kk = f >> 5, mq = f & 31;
                *reinterpret_cast<float4*>(&As[buf][kk][mq * 4]) = ra[i];
            }
        }
        if (B_MODE == 0) {
            int kk = t >> 4, nq = t & 15;
            *reinterpret_cast<float4*>(&Bs[buf][kk][nq * 4]) = rb;
        } else {
            int nn = t >> 2, kq = t & 3;
            Bs[buf][kq * 4 + 0][nn] = rb.x; Bs[buf][kq * 4 + 1][nn] = rb.y;
            Bs[buf][kq * 4 + 2][nn] = rb.z; Bs[buf][kq * 4 + 3][nn] = rb.w;
        }
    };

    int buf = 0;
    if (k_begin < k_end) {
        load_tiles(k_begin);
        store_tiles(0);
    }
    __syncthreads();
    for (int k0 = k_begin; k0 < k_end; k0 += BK) {
        const bool more = k0 + BK < k_end;
        if (more) load_tiles(k0 + BK);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * TM]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * TM + 4]);
            float4 b = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * TN]);
            const float a[TM] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bb[TN] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        if (more) store_tiles(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }

#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int m = m0 + ty * TM + i;
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int n = n0 + tx * TN + j;
            if (n >= g.N) continue;
            float v = acc[i][j];
            if (g.splits == 1) {
                if (g.bias) v += __ldg(g.bias + n);
                if (g.relu) v = fmaxf(v, 0.f);
                g.C[(int64_t)m * g.ldc + n] = v;
            } else {
                g.ws[((int64_t)z * g.M + m) * g.N + n] = v;
            }
        }
    }
}

// Tensor-core version of sgemm_kernel: same tiling, loads and split-K contract, inner product by mma.sync m16n8k8 with
// split-precision (3xTF32) operands -- a*b ~= a_lo*b_hi + a_hi*b_lo + a_hi*b_hi with fp32 accumulation, error ~2^-21 per
// product, i.e. fp32-grade (plain TF32 would break the 1e-4 parity budget).  Warp w owns a 32x32 sub-tile (2 m-tiles x 4
// n-tiles = 32 accumulators); fragments are read from the [k][m] / [k][n] shared tiles, whose leading dimensions are
// 8 (mod 32) so that the lanes (g, t) of a fragment load (address k = t, m = g) fall on 32 different banks.
constexpr int MAS_LD = BM + 8, MBS_LD = BN + 8;

template <int A_MODE, int B_MODE>
__global__ void __launch_bounds__(GEMM_THREADS)
sgemm_mma_kernel(GemmArgs g) {
    __shared__ __align__(16) float As[2][BK][MAS_LD];
    __shared__ __align__(16) float Bs[2][BK][MBS_LD];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5, fg = lane >> 2, ft = lane & 3;
    const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN, z = blockIdx.z;
    const int k_begin = z * g.Kc, k_end = min(g.K, k_begin + g.Kc);
    float acc[2][4][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.f;

    float4 ra[2], rb;
    auto load_tiles = [&](int k0) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            int f = t + i * GEMM_THREADS;
            if (A_MODE == 0) {
                int row = f >> 2, kq = f & 3, m = m0 + row, k = k0 + kq * 4;
                int valid = (m < g.M) ? max(0, min(4, k_end - k)) : 0;
                ra[i] = ld4_guard(g.A, (int64_t)m * g.sa_m + (int64_t)k * g.sa_k, g.sa_k, valid, g.vecA);
            } else {
                int kk = f >> 5, mq = f & 31, m = m0 + mq * 4, k = k0 + kk;
                int valid = (k < k_end) ? max(0, min(4, g.M - m)) : 0;
                ra[i] = ld4_guard(g.A, (int64_t)m * g.sa_m + (int64_t)k * g.sa_k, g.sa_m, valid, g.vecA);
            }
        }
        if (B_MODE == 0) {
            int kk = t >> 4, nq = t & 15, n = n0 + nq * 4, k = k0 + kk;
            int valid = (k < k_end) ? max(0, min(4, g.N - n)) : 0;
            rb = ld4_guard(g.B, (int64_t)k * g.sb_k + (int64_t)n * g.sb_n, g.sb_n, valid, g.vecB);
        } else {
            int nn = t >> 2, kq = t & 3, n = n0 + nn, k = k0 + kq * 4;
            int valid = (n < g.N) ? max(0, min(4, k_end - k)) : 0;
            rb = ld4_guard(g.B, (int64_t)k * g.sb_k + (int64_t)n * g.sb_n, g.sb_k, valid, g.vecB);
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            int f = t + i * GEMM_THREADS;
            if (A_MODE == 0) {
                int row = f >> 2, kq = f & 3;
                As[buf][kq * 4 + 0][row] = ra[i].x; As[buf][kq * 4 + 1][row] = ra[i].y;
                As[buf][kq * 4 + 2][row] = ra[i].z; As[buf][kq * 4 + 3][row] = ra[i].w;
            } else {
                int kk = f >> 5, mq = f & 31;
                *reinterpret_cast<float4*>(&As[buf][kk][mq * 4]) = ra[i];
            }
        }
        if (B_MODE == 0) {
            int kk = t >> 4, nq = t & 15;
            *reinterpret_cast<float4*>(&Bs[buf][kk][nq * 4]) = rb;
        } else {
            int nn = t >> 2, kq = t & 3;
            Bs[buf][kq * 4 + 0][nn] = rb.x; Bs[buf][kq * 4 + 1][nn] = rb.y;
            Bs[buf][kq * 4 + 2][nn] = rb.z; Bs[buf][kq * 4 + 3][nn] = rb.w;
        }
    };

    int buf = 0;
    if (k_begin < k_end) {
        load_tiles(k_begin);
        store_tiles(0);
    }
    __syncthreads();
    for (int k0 = k_begin; k0 < k_end; k0 += BK) {
        const bool more = k0 + BK < k_end;
        if (more) load_tiles(k0 + BK);
#pragma unroll
        for (int ks = 0; ks < BK / 8; ++ks) {
            // A fragments of the warp's two m-tiles: a0 (m g, k t), a1 (m g+8, k t), a2 (m g, k t+4), a3 (m g+8, k t+4)
            uint32_t ah[2][4], al[2][4];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const float* ap = &As[buf][8 * ks + ft][wm + 16 * i + fg];
                split_tf32(ap[0], ah[i][0], al[i][0]);
                split_tf32(ap[8], ah[i][1], al[i][1]);
                split_tf32(ap[4 * MAS_LD], ah[i][2], al[i][2]);
                split_tf32(ap[4 * MAS_LD + 8], ah[i][3], al[i][3]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                // B fragment of n-tile j: b0 (k t, n g), b1 (k t+4, n g)
                const float* bp = &Bs[buf][8 * ks + ft][wn + 8 * j + fg];
                uint32_t bh0, bl0, bh1, bl1;
                split_tf32(bp[0], bh0, bl0);
                split_tf32(bp[4 * MBS_LD], bh1, bl1);
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    mma_tf32(acc[i][j], al[i][0], al[i][1], al[i][2], al[i][3], bh0, bh1);
                    mma_tf32(acc[i][j], ah[i][0], ah[i][1], ah[i][2], ah[i][3], bl0, bl1);
                    mma_tf32(acc[i][j], ah[i][0], ah[i][1], ah[i][2], ah[i][3], bh0, bh1);
                }
            }
        }
        if (more) store_tiles(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }

    // accumulators: c0 (m g, n 2t), c1 (m g, n 2t+1), c2 (m g+8, n 2t), c3 (m g+8, n 2t+1)
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int m = m0 + wm + 16 * i + fg + 8 * (e >> 1), n = n0 + wn + 8 * j + 2 * ft + (e & 1);
                if (m >= g.M || n >= g.N) continue;
                float v = acc[i][j][e];
                if (g.splits == 1) {
                    if (g.bias) v += __ldg(g.bias + n);
                    if (g.relu) v = fmaxf(v, 0.f);
                    g.C[(int64_t)m * g.ldc + n] = v;
                } else {
                    g.ws[((int64_t)z * g.M + m) * g.N + n] = v;
                }
            }
}

__global__ void splitk_reduce_kernel(const float* __restrict__ ws, float* __restrict__ C, int64_t ldc, int M, int N,
                                     int splits, const float* __restrict__ bias, int relu) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)M * N) return;
    int m = (int)(idx / N), n = (int)(idx - (int64_t)m * N);
    float v = 0.f;
    for (int z = 0; z < splits; ++z) v += ws[(int64_t)z * M * N + idx];      // fixed order: deterministic
    if (bias) v += __ldg(bias + n);
    if (relu) v = fmaxf(v, 0.f);
    C[(int64_t)m * ldc + n] = v;
}

void launch_splitk_reduce(const float* ws, float* C, int64_t ldc, int M, int N, int splits, const float* bias, int relu,
                          cudaStream_t stream) {
    int64_t total = (int64_t)M * N;
    splitk_reduce_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, stream>>>(ws, C, ldc, M, N, splits, bias, relu);
}

static void pick_splits(int M, int N, int K, int& splits, int& Kc) {
    int64_t tiles = ceil_div(M, BM) * ceil_div(N, BN);
    int64_t want = std::max<int64_t>(1, ceil_div(2 * 148, tiles));
    int64_t max_splits = std::max<int64_t>(1, K / 128);          // at least 128 of K per split
    splits = (int)std::min(want, max_splits);
    Kc = (int)(ceil_div(ceil_div(K, splits), BK) * BK);
    splits = (int)ceil_div(K, Kc);
}

int gemm_splits(int M, int N, int K) {
    int s, kc;
    pick_splits(M, N, K, s, kc);
    return s;
}

// The large contractions (the grid Linear layer: 7 GFLOP forward at B = 256, twice that backward) go to the tcgen05 pipeline
// of tc_gemm.cu when GNBV_GEMM_MMA has bit 2 set (default): measured 0.082 / 0.077 / 0.078 ms against 0.154 / 0.152 / 0.148 ms
// for the mma.sync kernel (forward / dX / dW at B = 256).  Small GEMMs stay on mma.sync (launch + TMEM set-up dominate there).
static bool use_tcgen05(int M, int N, int K) {
    return (gemm_mma_mode() & 2) && N >= 256 && 2.0 * M * N * K >= 1.0e9;
}

size_t gemm_workspace_floats(int M, int N, int K) {
    int s = gemm_splits(M, N, K);
    const size_t a = s > 1 ? (size_t)s * M * N : 0;
    return use_tcgen05(M, N, K) ? std::max(a, tc_gemm_workspace_floats(M, N, K)) : a;
}

int launch_gemm(const float* A, int64_t sa_m, int64_t sa_k, const float* B, int64_t sb_k, int64_t sb_n, float* C,
                int64_t ldc, int M, int N, int K, const GemmEpilogue& ep, float* workspace, cudaStream_t stream) {
    GNBV_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0, "gemm: bad arguments (M=%d N=%d K=%d)", M, N, K);
    GNBV_REQUIRE(sa_k == 1 || sa_m == 1, "gemm: A must be contiguous along m or k");
    GNBV_REQUIRE(sb_n == 1 || sb_k == 1, "gemm: B must be contiguous along n or k");
    if (workspace && use_tcgen05(M, N, K)) return launch_tc_gemm(A, sa_m, sa_k, B, sb_k, sb_n, C, ldc, M, N, K, ep, workspace, stream);
    GemmArgs g;
    g.A = A; g.sa_m = sa_m; g.sa_k = sa_k; g.B = B; g.sb_k = sb_k; g.sb_n = sb_n; g.C = C; g.ldc = ldc;
    g.M = M; g.N = N; g.K = K; g.bias = ep.bias; g.relu = ep.relu;
    pick_splits(M, N, K, g.splits, g.Kc);
    if (!workspace) {                      // no workspace: single pass over K
        g.splits = 1;
        g.Kc = (int)(ceil_div(K, BK) * BK);
    }
    g.ws = workspace;
    const int a_mode = (sa_k == 1) ? 0 : 1, b_mode = (sb_n == 1) ? 0 : 1;
    const int64_t lda = a_mode == 0 ? sa_m : sa_k, ldb = b_mode == 0 ? sb_k : sb_n;
    g.vecA = ((uintptr_t)A % 16 == 0) && (lda % 4 == 0);
    g.vecB = ((uintptr_t)B % 16 == 0) && (ldb % 4 == 0);
    dim3 grid((unsigned)ceil_div(N, BN), (unsigned)ceil_div(M, BM), (unsigned)g.splits);
    if (gemm_mma_mode() & 1) {
        if (a_mode == 0 && b_mode == 0) sgemm_mma_kernel<0, 0><<<grid, GEMM_THREADS, 0, stream>>>(g);
        else if (a_mode == 0 && b_mode == 1) sgemm_mma_kernel<0, 1><<<grid, GEMM_THREADS, 0, stream>>>(g);
        else if (a_mode == 1 && b_mode == 0) sgemm_mma_kernel<1, 0><<<grid, GEMM_THREADS, 0, stream>>>(g);
        else sgemm_mma_kernel<1, 1><<<grid, GEMM_THREADS, 0, stream>>>(g);
    } else if (a_mode == 0 && b_mode == 0) sgemm_kernel<0, 0><<<grid, GEMM_THREADS, 0, stream>>>(g);
    else if (a_mode == 0 && b_mode == 1) sgemm_kernel<0, 1><<<grid, GEMM_THREADS, 0, stream>>>(g);
    else if (a_mode == 1 && b_mode == 0) sgemm_kernel<1, 0><<<grid, GEMM_THREADS, 0, stream>>>(g);
    else sgemm_kernel<1, 1><<<grid, GEMM_THREADS, 0, stream>>>(g);
    GNBV_LAUNCH_CHECK("sgemm_kernel");
    if (g.splits > 1) {
        int64_t total = (int64_t)M * N;
        splitk_reduce_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, stream>>>(workspace, C, ldc, M, N, g.splits, ep.bias,
                                                                                 ep.relu);
        GNBV_LAUNCH_CHECK("splitk_reduce_kernel");
    }
    return GNBV_OK;
}

}  // namespace gnbv

// Exposed for tests and for host code that wants a bare fp32 GEMM on device pointers.
extern "C" size_t gnbv_sgemm_workspace_bytes(int M, int N, int K) { return gnbv::gemm_workspace_floats(M, N, K) * 4; }

extern "C" int gnbv_sgemm(const float* A, int64_t sa_m, int64_t sa_k, const float* B, int64_t sb_k, int64_t sb_n,
                          float* C, int64_t ldc, int M, int N, int K, const float* bias, int relu, float* workspace,
                          size_t workspace_bytes, void* stream) {
    if (workspace_bytes < gnbv::gemm_workspace_floats(M, N, K) * 4) {
        gnbv::set_error("gnbv_sgemm: workspace %zu B too small", workspace_bytes);
        return GNBV_E_WORKSPACE;
    }
    gnbv::GemmEpilogue ep;
    ep.bias = bias; ep.relu = relu;
    return gnbv::launch_gemm(A, sa_m, sa_k, B, sb_k, sb_n, C, ldc, M, N, K, ep, workspace, (cudaStream_t)stream);
}
