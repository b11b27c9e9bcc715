// conv2_mma.cu -- Conv3d(16,16,3,stride 2) forward of the Hybrid_Encoder (hybrid_encoder.py:35-37) as an implicit GEMM on
// the tensor cores through warp-level mma.sync (m16n8k8, TF32 operands, fp32 accumulation) with split-precision (3xTF32)
// operands: x = x_hi + x_lo, w = w_hi + w_lo, acc += x_lo*w_hi + x_hi*w_lo + x_hi*w_hi.  The dropped lo*lo term is
// ~2^-22 relative, i.e. fp32-level, which keeps the <= 1e-4 parity budget with the reference's fp32 (TF32-off) cuDNN path
// that plain TF32 operands would not (tests/test_tc_gemm_gpu.py::test_plain_tf32_would_not_meet_the_budget).
//
// Why mma.sync here and not tcgen05 (conv2_tc.cu): with N = 16 output channels each activation feeds only 16 MACs, and
// the activation needs BN1 + ReLU + the hi/lo split on its way in.  UMMA wants its A operand in a canonical shared-memory
// layout, so every activation was loaded, transformed, split and re-stored (~20 instructions per element, measured 1.43 ms
// vs 0.55 ms for the CUDA-core kernel).  mma.sync takes A from REGISTERS: a thread loads the 4 consecutive channels it
// owns as one 128-bit load, applies BN + ReLU + split in place and issues the MMAs -- no im2col copy at all.
//
// Mapping.  GEMM rows = output voxels, K = 27 taps x 16 input channels, N = 16.  A warp owns MT = 4 row tiles of 16
// voxels.  Within a tap the 16 input channels are two k-steps of 8; the k order inside a k-step is permuted so that
// thread t (= lane & 3) supplies channels 4t..4t+3: k-step A uses (4t, 4t+1) for the fragment's k slots (t, t+4), k-step B
// uses (4t+2, 4t+3).  The weight fragments follow the same permutation and sit pre-split (hi / lo) in shared memory in
// exactly the order the threads read them: [tap][ntile*2 + part][lane] as float4 -> conflict-free LDS.128.
#include "common.cuh"
#include "conv2_mma.cuh"
#include "mma.cuh"
#include "tma.cuh"

#include <algorithm>

namespace gnbv {

namespace {

constexpr int C = 16;
constexpr int NTAPS = 27;
constexpr int MMA_THREADS = 128;
constexpr int MMA_WARPS = MMA_THREADS / 32;
constexpr int MT = 4;                               // m16 tiles per warp
constexpr int ROWS_PER_WARP = 16 * MT;              // 64 output voxels
constexpr int ROWS_PER_ITEM = ROWS_PER_WARP * MMA_WARPS;   // 256 output voxels per work item
constexpr int WSM_FLOAT4 = NTAPS * 4 * 32;          // 3456 float4 = 55,296 B
constexpr int MMA_PART_STRIDE = 2 * C + 4;          // == PART_STRIDE of encoder.cu: mean[16], M2[16], count, pad
constexpr size_t MMA_SMEM = (size_t)WSM_FLOAT4 * 16 + (size_t)(2 * MMA_WARPS * C + C) * 4;

struct Split4 { uint32_t hi[4], lo[4]; };

// v / n for 0 <= v < 2^22 and small n with inv = 1.f / n: (v + 0.5) * inv is at least 0.5 / n away from every integer, far
// more than its fp32 rounding error, so the floor is exact -- 3 instructions instead of the ~20 of an integer division.
__device__ __forceinline__ int fast_div(int v, float inv) { return __float2int_rd(((float)v + 0.5f) * inv); }

// relu(a * y + b) for the thread's 4 channels, split into tf32 hi / lo.
// `cvt.rna.tf32.f32` is not a native instruction on sm_100a (ptxas expands it to ~4 instructions with an infinity check), so
// the split is spelled out: hi = round-to-nearest of the 13 dropped mantissa bits (add half an ulp, mask); lo = v - hi is
// exact in fp32 and is handed over as is -- the tensor core reads only its upper 19 bits, an error below 2^-21 |v|.
// (Activations are finite and far from FLT_MAX, the only place where the unchecked add could overflow.)
__device__ __forceinline__ Split4 bn_relu_split(float4 y, const float (&sc)[4], const float (&sh)[4]) {
    const float v[4] = {fmaxf(fmaf(sc[0], y.x, sh[0]), 0.f), fmaxf(fmaf(sc[1], y.y, sh[1]), 0.f),
                        fmaxf(fmaf(sc[2], y.z, sh[2]), 0.f), fmaxf(fmaf(sc[3], y.w, sh[3]), 0.f)};
    Split4 s;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        s.hi[k] = (__float_as_uint(v[k]) + 0x1000u) & 0xffffe000u;
        s.lo[k] = __float_as_uint(v[k] - __uint_as_float(s.hi[k]));
    }
    return s;
}

__device__ __forceinline__ Split4 split4(float4 y) {
    const float v[4] = {y.x, y.y, y.z, y.w};
    Split4 s;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        s.hi[k] = (__float_as_uint(v[k]) + 0x1000u) & 0xffffe000u;
        s.lo[k] = __float_as_uint(v[k] - __uint_as_float(s.hi[k]));
    }
    return s;
}

// y1 [B, G1^3, 16] channels-last (pre-BN conv1 output); stat1[2*16 + c] = BN1 scale, stat1[3*16 + c] = BN1 shift;
// w [16,16,3,3,3]; y2 [B,16,G2^3] channel-major pre-BN (bias added); part: one (mean[16], M2[16], count) record per item.
__global__ void __launch_bounds__(MMA_THREADS, 3)
conv2_fwd_mma_kernel(const float* __restrict__ y1, const float* __restrict__ stat1, const float* __restrict__ w,
                     const float* __restrict__ bias, float* __restrict__ y2, float* __restrict__ part, int G1, int G2,
                     int chunks, int items) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* wsm = reinterpret_cast<float4*>(smem_raw);                       // [tap][q][lane]
    float* red = reinterpret_cast<float*>(smem_raw + (size_t)WSM_FLOAT4 * 16);   // [2][warps][16] partial sums
    float* bmean = red + 2 * MMA_WARPS * C;                                  // [16]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;

    // ---- stage the weight fragments, pre-split, in the order the threads consume them
    for (int idx = tid; idx < WSM_FLOAT4; idx += MMA_THREADS) {
        const int tap = idx >> 7, q = (idx >> 5) & 3, ln = idx & 31;
        const int gg = ln >> 2, tt = ln & 3, j = q >> 1, lo_part = q & 1;
        const int co = 8 * j + gg;
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float wv = __ldg(w + ((int64_t)co * C + (4 * tt + e)) * NTAPS + tap);
            const float hi = __uint_as_float(to_tf32(wv));
            v[e] = lo_part ? __uint_as_float(to_tf32(wv - hi)) : hi;
        }
        wsm[idx] = make_float4(v[0], v[1], v[2], v[3]);
    }
    float sc[4], sh[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) { sc[e] = stat1[2 * C + 4 * t + e]; sh[e] = stat1[3 * C + 4 * t + e]; }
    float bias_r[2][2];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) bias_r[j][e] = bias[8 * j + 2 * t + e];
    __syncthreads();

    const int P1 = G1 * G1 * G1, P2 = G2 * G2 * G2;
    const float inv_g2 = 1.0f / (float)G2;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int b = item / chunks, chunk = item - b * chunks;
        const int row0 = chunk * ROWS_PER_ITEM + warp * ROWS_PER_WARP;
        const float* in_b = y1 + (int64_t)b * P1 * C + 4 * t;
        // input offsets (floats) of the thread's rows g and g+8 of every m-tile; rows past P2 alias the last voxel
        int off[MT][2];
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int p = min(row0 + m * 16 + g + 8 * h, P2 - 1);
                const int r = fast_div(p, inv_g2), z2 = p - r * G2, x2 = fast_div(r, inv_g2), yy2 = r - x2 * G2;
                off[m][h] = (((2 * x2) * G1 + 2 * yy2) * G1 + 2 * z2) * C;
            }
        float acc[MT][2][4];
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int j = 0; j < 2; ++j) { acc[m][j][0] = acc[m][j][2] = bias_r[j][0]; acc[m][j][1] = acc[m][j][3] = bias_r[j][1]; }

        // (explicit register double-buffering of the tap loads was measured slower here: 0.291 vs 0.275 ms -- the unrolled
        // l loop already keeps three taps of loads in flight and the extra index arithmetic costs more than it hides)
        for (int ij = 0; ij < 9; ++ij) {
            const int i = ij / 3, jy = ij - 3 * i;
            const int tap_off = ((i * G1 + jy) * G1) * C;
#pragma unroll
            for (int l = 0; l < 3; ++l) {
                const int tap = ij * 3 + l;
                const float4* wt = wsm + tap * 128 + lane;
                const float4 w0h = wt[0], w0l = wt[32], w1h = wt[64], w1l = wt[96];       // ntile 0 hi/lo, ntile 1 hi/lo
                const uint32_t bh[2][4] = {{__float_as_uint(w0h.x), __float_as_uint(w0h.y), __float_as_uint(w0h.z), __float_as_uint(w0h.w)},
                                           {__float_as_uint(w1h.x), __float_as_uint(w1h.y), __float_as_uint(w1h.z), __float_as_uint(w1h.w)}};
                const uint32_t bl[2][4] = {{__float_as_uint(w0l.x), __float_as_uint(w0l.y), __float_as_uint(w0l.z), __float_as_uint(w0l.w)},
                                           {__float_as_uint(w1l.x), __float_as_uint(w1l.y), __float_as_uint(w1l.z), __float_as_uint(w1l.w)}};
                float4 raw[MT][2];
#pragma unroll
                for (int m = 0; m < MT; ++m)
#pragma unroll
                    for (int h = 0; h < 2; ++h)
                        raw[m][h] = __ldg(reinterpret_cast<const float4*>(in_b + off[m][h] + tap_off + l * C));
#pragma unroll
                for (int m = 0; m < MT; ++m) {
                    const Split4 r0 = bn_relu_split(raw[m][0], sc, sh), r1 = bn_relu_split(raw[m][1], sc, sh);
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) {             // k-step A: channels (4t, 4t+1); k-step B: (4t+2, 4t+3)
                        const int e0 = 2 * ks, e1 = 2 * ks + 1;
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            mma_tf32(acc[m][j], r0.lo[e0], r1.lo[e0], r0.lo[e1], r1.lo[e1], bh[j][e0], bh[j][e1]);
                            mma_tf32(acc[m][j], r0.hi[e0], r1.hi[e0], r0.hi[e1], r1.hi[e1], bl[j][e0], bl[j][e1]);
                            mma_tf32(acc[m][j], r0.hi[e0], r1.hi[e0], r0.hi[e1], r1.hi[e1], bh[j][e0], bh[j][e1]);
                        }
                    }
                }
            }
        }

        // ---- store y2 (channel-major) and the block's BN statistics record
        const int nrows_item = min(ROWS_PER_ITEM, P2 - chunk * ROWS_PER_ITEM);   // > 0 by construction of `chunks`
        float s[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int p = row0 + m * 16 + g + 8 * h;
                if (p < P2) {
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const float v = acc[m][j][2 * h + e];
                            y2[((int64_t)b * C + (8 * j + 2 * t + e)) * P2 + p] = v;
                            s[j][e] += v;
                        }
                }
            }
        if (part) {                                    // uniform branch
            // pass 1: block mean per channel
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    float v = s[j][e];
                    v += __shfl_xor_sync(0xffffffffu, v, 4); v += __shfl_xor_sync(0xffffffffu, v, 8); v += __shfl_xor_sync(0xffffffffu, v, 16);
                    if (g == 0) red[warp * C + 8 * j + 2 * t + e] = v;
                }
            __syncthreads();
            if (tid < C) {
                float tot = 0.f;
#pragma unroll
                for (int wv = 0; wv < MMA_WARPS; ++wv) tot += red[wv * C + tid];
                bmean[tid] = tot / (float)nrows_item;
            }
            __syncthreads();
            // pass 2: centred sum of squares
            float q2[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
#pragma unroll
            for (int m = 0; m < MT; ++m)
#pragma unroll
                for (int h = 0; h < 2; ++h)
                    if (row0 + m * 16 + g + 8 * h < P2) {
#pragma unroll
                        for (int j = 0; j < 2; ++j)
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const float dlt = acc[m][j][2 * h + e] - bmean[8 * j + 2 * t + e];
                                q2[j][e] = fmaf(dlt, dlt, q2[j][e]);
                            }
                    }
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    float v = q2[j][e];
                    v += __shfl_xor_sync(0xffffffffu, v, 4); v += __shfl_xor_sync(0xffffffffu, v, 8); v += __shfl_xor_sync(0xffffffffu, v, 16);
                    if (g == 0) red[(MMA_WARPS + warp) * C + 8 * j + 2 * t + e] = v;
                }
            __syncthreads();
            if (tid < C) {
                float m2 = 0.f;
#pragma unroll
                for (int wv = 0; wv < MMA_WARPS; ++wv) m2 += red[(MMA_WARPS + wv) * C + tid];
                float* pr = part + (int64_t)item * MMA_PART_STRIDE;
                pr[tid] = bmean[tid];
                pr[C + tid] = m2;
                if (tid == 0) pr[2 * C] = (float)nrows_item;
            }
            __syncthreads();                           // red / bmean are reused by the next item
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// conv2 data gradient + ReLU / BN1-backward statistics (same contract as conv2_dgrad_kernel in encoder.cu):
//   dact1[b,v,ci] = sum_{taps with (v - tap) even and in range} sum_co dy2[b,(v - tap)/2,co] * W2[co,ci,tap]
//   g1 = dact1 * [bn1(y1) > 0]  (channels-last);  bpart[item][0..16) = sum g1, [16..32) = sum g1 * xhat1.
// The 8 parity classes (x&1, y&1, z&1) of the conv1-output grid each have a fixed tap set (2 or 1 taps per axis), so a
// work item = 256 voxels of ONE class of one sample is a dense GEMM: rows = voxels, K = (taps of the class) x 16 co, N = 16 ci.
// A rows are contiguous 16-channel dy2 vectors (one LDG.128 per thread and row, zeros outside the output grid), split
// hi/lo in registers; B = W2[tap][co][ci] fragments pre-split in shared memory like the forward kernel's.
struct DgradClass { int cx, cy, cz, nx, ny, nz, nvox, chunks; };

__device__ __host__ inline DgradClass dgrad_class(int cls, int G1) {
    const int NE = (G1 + 1) / 2, NO = G1 / 2;
    DgradClass c;
    c.cx = (cls >> 2) & 1; c.cy = (cls >> 1) & 1; c.cz = cls & 1;
    c.nx = c.cx ? NO : NE; c.ny = c.cy ? NO : NE; c.nz = c.cz ? NO : NE;
    c.nvox = c.nx * c.ny * c.nz;
    c.chunks = (c.nvox + ROWS_PER_ITEM - 1) / ROWS_PER_ITEM;
    return c;
}

__global__ void __launch_bounds__(MMA_THREADS, 3)
conv2_dgrad_mma_kernel(const float* __restrict__ dy2cl, const float* __restrict__ w, const float* __restrict__ y1,
                       const float* __restrict__ stat1, float* __restrict__ g1, float* __restrict__ bpart, int G1, int G2,
                       int items_per_sample, int total_items) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* wsm = reinterpret_cast<float4*>(smem_raw);                           // [tap][q][lane]
    float* red = reinterpret_cast<float*>(smem_raw + (size_t)WSM_FLOAT4 * 16);   // [warps][32]
    __shared__ int row_src[ROWS_PER_ITEM], row_dst[ROWS_PER_ITEM];
    __shared__ uint32_t row_ok[ROWS_PER_ITEM];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;

    // B[k = co][n = ci] = W2[co][ci][tap]; thread (g,t) of n-tile j reads co = 4t..4t+3 for ci = 8j + g
    for (int idx = tid; idx < WSM_FLOAT4; idx += MMA_THREADS) {
        const int tap = idx >> 7, q = (idx >> 5) & 3, ln = idx & 31;
        const int gg = ln >> 2, tt = ln & 3, j = q >> 1, lo_part = q & 1;
        const int ci = 8 * j + gg;
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float wv = __ldg(w + ((int64_t)(4 * tt + e) * C + ci) * NTAPS + tap);
            const float hi = __uint_as_float(to_tf32(wv));
            v[e] = lo_part ? __uint_as_float(to_tf32(wv - hi)) : hi;
        }
        wsm[idx] = make_float4(v[0], v[1], v[2], v[3]);
    }
    __syncthreads();

    const int P1 = G1 * G1 * G1, P2 = G2 * G2 * G2;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const int b = item / items_per_sample;
        int r = item - b * items_per_sample;
        DgradClass cl = dgrad_class(0, G1);
        for (int cls = 0; cls < 8; ++cls) {                       // block-uniform
            cl = dgrad_class(cls, G1);
            if (r < cl.chunks) break;
            r -= cl.chunks;
        }
        const float inv_nz = 1.0f / (float)cl.nz, inv_ny = 1.0f / (float)cl.ny;
        // per row: dy2 offset of the (di,dj,dl) = (0,0,0) source voxel, output voxel index, validity bits.  Decoding a row
        // costs ~40 instructions; instead of every thread decoding its own 8 rows, the block decodes the item's 256 rows
        // once (2 per thread) into shared memory and every thread picks its rows up from there.
        for (int rr = tid; rr < ROWS_PER_ITEM; rr += MMA_THREADS) {
            const int v = r * ROWS_PER_ITEM + rr;
            const bool in = v < cl.nvox;
            const int vv = in ? v : 0;
            const int q = fast_div(vv, inv_nz), zi = vv - q * cl.nz, xi = fast_div(q, inv_ny), yi = q - xi * cl.ny;
            row_src[rr] = ((xi * G2 + yi) * G2 + zi) * C;
            row_dst[rr] = in ? ((2 * xi + cl.cx) * G1 + (2 * yi + cl.cy)) * G1 + (2 * zi + cl.cz) : -1;
            // bit 2a: source index (coordinate - 0) < G2;  bit 2a+1: (coordinate - 1) >= 0
            row_ok[rr] = in ? ((xi < G2 ? 1u : 0u) | (xi >= 1 ? 2u : 0u) | (yi < G2 ? 4u : 0u) | (yi >= 1 ? 8u : 0u) |
                               (zi < G2 ? 16u : 0u) | (zi >= 1 ? 32u : 0u))
                            : 0u;
        }
        __syncthreads();
        int src[MT][2], dst[MT][2];
        uint32_t okb[MT][2];
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int rr = warp * ROWS_PER_WARP + m * 16 + g + 8 * h;
                src[m][h] = row_src[rr]; dst[m][h] = row_dst[rr]; okb[m][h] = row_ok[rr];
            }
        float acc[MT][2][4];
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[m][j][e] = 0.f;
        const float* dy_b = dy2cl + (int64_t)b * P2 * C + 4 * t;

        // taps of the class: per axis, parity 0 -> kernel offsets {0, 2} (source shift 0, 1); parity 1 -> {1} (shift 0).
        // The 1 / 2 / 4 / 8 taps are walked with the loads of tap k+1 in flight while tap k is multiplied (two register
        // buffers): the A rows come straight from L2 / HBM and their latency is what this kernel has to hide.
        const int nix = cl.cx ? 1 : 2, niy = cl.cy ? 1 : 2, niz = cl.cz ? 1 : 2;
        const int ntap = nix * niy * niz;
        auto tap_of = [&](int k, int& tap, int& delta, uint32_t& need) {
            const int az = niz == 2 ? (k & 1) : 0, k2 = niz == 2 ? (k >> 1) : k;
            const int ay = niy == 2 ? (k2 & 1) : 0, ax = niy == 2 ? (k2 >> 1) : k2;
            const int i = cl.cx ? 1 : 2 * ax, jy = cl.cy ? 1 : 2 * ay, l = cl.cz ? 1 : 2 * az;
            tap = (i * 3 + jy) * 3 + l;
            delta = ((ax * G2 + ay) * G2 + az) * C;               // source shift (i>>1, j>>1, l>>1) = (ax, ay, az)
            need = (1u << ax) | (4u << ay) | (16u << az);
        };
        auto load = [&](float4 (&raw)[MT][2], int delta, uint32_t need) {
#pragma unroll
            for (int m = 0; m < MT; ++m)
#pragma unroll
                for (int h = 0; h < 2; ++h)
                    raw[m][h] = (okb[m][h] & need) == need ? __ldg(reinterpret_cast<const float4*>(dy_b + src[m][h] - delta))
                                                           : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        auto compute = [&](const float4 (&raw)[MT][2], int tap) {
            const float4* wt = wsm + tap * 128 + lane;
            const float4 w0h = wt[0], w0l = wt[32], w1h = wt[64], w1l = wt[96];
            const uint32_t bh[2][4] = {{__float_as_uint(w0h.x), __float_as_uint(w0h.y), __float_as_uint(w0h.z), __float_as_uint(w0h.w)},
                                       {__float_as_uint(w1h.x), __float_as_uint(w1h.y), __float_as_uint(w1h.z), __float_as_uint(w1h.w)}};
            const uint32_t bl[2][4] = {{__float_as_uint(w0l.x), __float_as_uint(w0l.y), __float_as_uint(w0l.z), __float_as_uint(w0l.w)},
                                       {__float_as_uint(w1l.x), __float_as_uint(w1l.y), __float_as_uint(w1l.z), __float_as_uint(w1l.w)}};
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                const Split4 r0 = split4(raw[m][0]), r1 = split4(raw[m][1]);
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    const int e0 = 2 * ks, e1 = 2 * ks + 1;
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        mma_tf32(acc[m][j], r0.lo[e0], r1.lo[e0], r0.lo[e1], r1.lo[e1], bh[j][e0], bh[j][e1]);
                        mma_tf32(acc[m][j], r0.hi[e0], r1.hi[e0], r0.hi[e1], r1.hi[e1], bl[j][e0], bl[j][e1]);
                        mma_tf32(acc[m][j], r0.hi[e0], r1.hi[e0], r0.hi[e1], r1.hi[e1], bh[j][e0], bh[j][e1]);
                    }
                }
            }
        };
        // The epilogue needs y1 at the item's output voxels (streamed from HBM).  Its 16 loads ride in whichever tap buffer is
        // idle during the LAST tap's MMAs -- same register footprint, and the round trip is over when the epilogue starts
        // (issued after the loop they were 12 % of this kernel's stall samples).
        float4 raw_a[MT][2], raw_b[MT][2];
        auto load_y = [&](float4 (&buf)[MT][2]) {
#pragma unroll
            for (int m = 0; m < MT; ++m)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float* yp = y1 + ((int64_t)b * P1 + max(dst[m][h], 0)) * C + 2 * t;
                    const float2 lo2 = __ldg(reinterpret_cast<const float2*>(yp)), hi2 = __ldg(reinterpret_cast<const float2*>(yp + 8));
                    buf[m][h] = make_float4(lo2.x, lo2.y, hi2.x, hi2.y);
                }
        };
        {
            int tap0, d0, tap1 = 0, d1 = 0;
            uint32_t n0, n1 = 0;
            tap_of(0, tap0, d0, n0);
            load(raw_a, d0, n0);
            for (int k = 0; k < ntap; k += 2) {                   // all conditions are block-uniform
                const bool has1 = k + 1 < ntap;
                if (has1) { tap_of(k + 1, tap1, d1, n1); load(raw_b, d1, n1); }
                else load_y(raw_b);                               // single-tap class
                compute(raw_a, tap0);
                if (has1) {
                    if (k + 2 < ntap) { tap_of(k + 2, tap0, d0, n0); load(raw_a, d0, n0); }
                    else load_y(raw_a);                           // last pair of taps
                    compute(raw_b, tap1);
                }
            }
            if (ntap == 1) {
#pragma unroll
                for (int m = 0; m < MT; ++m)
#pragma unroll
                    for (int h = 0; h < 2; ++h) raw_a[m][h] = raw_b[m][h];
            }
        }

        // ---- epilogue: ReLU mask from bn1(y1) (values already in raw_a), store g1 (channels-last), BN1-backward partial sums.
        float s1[2][2] = {{0.f, 0.f}, {0.f, 0.f}}, s2[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
        float k_mean[2][2], k_istd[2][2], k_a[2][2], k_b[2][2];     // BN1 constants of the thread's channels c = 8j + 2t + e
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int c = 8 * j + 2 * t + e;
                k_mean[j][e] = __ldg(stat1 + c); k_istd[j][e] = __ldg(stat1 + C + c);
                k_a[j][e] = __ldg(stat1 + 2 * C + c); k_b[j][e] = __ldg(stat1 + 3 * C + c);
            }
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (dst[m][h] < 0) continue;
                const int64_t base = ((int64_t)b * P1 + dst[m][h]) * C;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const float y2v[2] = {j ? raw_a[m][h].z : raw_a[m][h].x, j ? raw_a[m][h].w : raw_a[m][h].y};
                    float o[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const float pre = fmaf(k_a[j][e], y2v[e], k_b[j][e]);
                        const float gv = pre > 0.f ? acc[m][j][2 * h + e] : 0.f;
                        o[e] = gv;
                        s1[j][e] += gv;
                        s2[j][e] = fmaf(gv, (y2v[e] - k_mean[j][e]) * k_istd[j][e], s2[j][e]);
                    }
                    *reinterpret_cast<float2*>(g1 + base + 8 * j + 2 * t) = make_float2(o[0], o[1]);
                }
            }
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                float a = s1[j][e], q = s2[j][e];
                a += __shfl_xor_sync(0xffffffffu, a, 4); a += __shfl_xor_sync(0xffffffffu, a, 8); a += __shfl_xor_sync(0xffffffffu, a, 16);
                q += __shfl_xor_sync(0xffffffffu, q, 4); q += __shfl_xor_sync(0xffffffffu, q, 8); q += __shfl_xor_sync(0xffffffffu, q, 16);
                if (g == 0) { red[warp * 2 * C + 8 * j + 2 * t + e] = a; red[warp * 2 * C + C + 8 * j + 2 * t + e] = q; }
            }
        __syncthreads();
        if (tid < 2 * C) {
            float tot = 0.f;
#pragma unroll
            for (int wv = 0; wv < MMA_WARPS; ++wv) tot += red[wv * 2 * C + tid];
            bpart[(int64_t)item * 2 * C + tid] = tot;
        }
        __syncthreads();                                          // red[] is reused by the next item
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// conv2 weight gradient (same contract as conv2_wgrad_kernel in encoder.cu):
//   dW2[co,ci,tap] = sum_{b,pos} dy2[b,pos,co] * relu(bn1(y1))[b, 2pos+tap, ci];   db2[co] = sum dy2[b,pos,co]
// GEMM with M = 16 (co), N = 27 taps x 16 ci = 54 n-tiles of 8, K = output positions.  A block walks output z-rows
// (b, x2, y2); a row's G2 positions are ceil(G2/8) k-steps (missing positions contribute zero A).  The four warps of a
// block share the rows and split the taps (warp w owns taps w, w+4, ...: <= 7 taps x 2 ci-halves = 14 n-tiles = 56
// accumulators), so every warp re-loads the small A fragment (dy2, 4 scalars per k-step) and its own B fragments
// (activations: 2 scalars per n-tile and k-step, BN1 + ReLU + hi/lo split in registers).  Per block one record
// part[blk][6912 + 16] in weight layout [co][ci][tap] followed by db2[16], reduced in fixed order by reduce_records_kernel.
constexpr int WG_TAPS_PER_WARP = (NTAPS + MMA_WARPS - 1) / MMA_WARPS;       // 7
constexpr int WG_REC = C * C * NTAPS + C;

__global__ void __launch_bounds__(MMA_THREADS, 4)
conv2_wgrad_mma_kernel(const float* __restrict__ y1, const float* __restrict__ stat1, const float* __restrict__ dy2cl,
                       float* __restrict__ part, int G1, int G2, int total_rows, int rows_per_block) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int P1 = G1 * G1 * G1, P2 = G2 * G2 * G2;
    // BN1 scale / shift of the two input channels this lane supplies as B columns: ci = 8*hf + g
    const float sc0 = stat1[2 * C + g], sh0 = stat1[3 * C + g], sc1 = stat1[2 * C + 8 + g], sh1 = stat1[3 * C + 8 + g];
    float acc[WG_TAPS_PER_WARP][2][4];
#pragma unroll
    for (int k = 0; k < WG_TAPS_PER_WARP; ++k)
#pragma unroll
        for (int hf = 0; hf < 2; ++hf)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[k][hf][e] = 0.f;
    float db_lo = 0.f, db_hi = 0.f;                          // sum of dy2 over positions for co = g and co = g + 8
    const int ksteps = (G2 + 7) / 8;
    const int r0 = blockIdx.x * rows_per_block, r1 = min(total_rows, r0 + rows_per_block);
    for (int row = r0; row < r1; ++row) {
        const int b = row / (G2 * G2), rem = row - b * G2 * G2, x2 = rem / G2, yy2 = rem - x2 * G2;
        const float* dyrow = dy2cl + ((int64_t)b * P2 + (int64_t)(x2 * G2 + yy2) * G2) * C;
        const float* xin = y1 + ((int64_t)b * P1 + ((int64_t)(2 * x2) * G1 + 2 * yy2) * G1) * C;
        for (int ks = 0; ks < ksteps; ++ks) {
            const int za = 8 * ks + t, zb = za + 4;             // positions of the fragment's k slots t and t + 4
            const bool va = za < G2, vb = zb < G2;
            const int zac = min(za, G2 - 1), zbc = min(zb, G2 - 1);
            // A = dy2^T: a0 (co g, pos za), a1 (co g+8, pos za), a2 (co g, pos zb), a3 (co g+8, pos zb)
            const float a0 = va ? __ldg(dyrow + zac * C + g) : 0.f, a1 = va ? __ldg(dyrow + zac * C + 8 + g) : 0.f;
            const float a2 = vb ? __ldg(dyrow + zbc * C + g) : 0.f, a3 = vb ? __ldg(dyrow + zbc * C + 8 + g) : 0.f;
            db_lo += a0 + a2; db_hi += a1 + a3;
            uint32_t ah[4], al[4];
            split_tf32(a0, ah[0], al[0]); split_tf32(a1, ah[1], al[1]); split_tf32(a2, ah[2], al[2]); split_tf32(a3, ah[3], al[3]);
#pragma unroll
            for (int k = 0; k < WG_TAPS_PER_WARP; ++k) {
                const int tap = warp + MMA_WARPS * k;
                if (tap < NTAPS) {                              // warp-uniform
                    const int i = tap / 9, jy = (tap / 3) % 3, l = tap % 3;
                    const float* line = xin + ((int64_t)(i * G1 + jy) * G1 + l) * C;
                    // B (k = position, n = ci): b0 = x[2*za + l][ci], b1 = x[2*zb + l][ci]; clamped positions carry A = 0
                    const float* pa = line + 2 * zac * C + g;
                    const float* pb = line + 2 * zbc * C + g;
                    const float x00 = fmaxf(fmaf(sc0, __ldg(pa), sh0), 0.f), x01 = fmaxf(fmaf(sc0, __ldg(pb), sh0), 0.f);
                    const float x10 = fmaxf(fmaf(sc1, __ldg(pa + 8), sh1), 0.f), x11 = fmaxf(fmaf(sc1, __ldg(pb + 8), sh1), 0.f);
                    uint32_t bh0, bl0, bh1, bl1;
                    split_tf32(x00, bh0, bl0); split_tf32(x01, bh1, bl1);
                    mma_tf32(acc[k][0], al[0], al[1], al[2], al[3], bh0, bh1);
                    mma_tf32(acc[k][0], ah[0], ah[1], ah[2], ah[3], bl0, bl1);
                    mma_tf32(acc[k][0], ah[0], ah[1], ah[2], ah[3], bh0, bh1);
                    split_tf32(x10, bh0, bl0); split_tf32(x11, bh1, bl1);
                    mma_tf32(acc[k][1], al[0], al[1], al[2], al[3], bh0, bh1);
                    mma_tf32(acc[k][1], ah[0], ah[1], ah[2], ah[3], bl0, bl1);
                    mma_tf32(acc[k][1], ah[0], ah[1], ah[2], ah[3], bh0, bh1);
                }
            }
        }
    }
    // ---- the block's record: acc (m = co, n = ci) -> [co][ci][tap]; db2 from warp 0
    float* out = part + (int64_t)blockIdx.x * WG_REC;
#pragma unroll
    for (int k = 0; k < WG_TAPS_PER_WARP; ++k) {
        const int tap = warp + MMA_WARPS * k;
        if (tap < NTAPS) {
#pragma unroll
            for (int hf = 0; hf < 2; ++hf)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int co = g + 8 * (e >> 1), ci = 8 * hf + 2 * t + (e & 1);
                    out[(co * C + ci) * NTAPS + tap] = acc[k][hf][e];
                }
        }
    }
    if (warp == 0) {
        db_lo += __shfl_xor_sync(0xffffffffu, db_lo, 1); db_lo += __shfl_xor_sync(0xffffffffu, db_lo, 2);
        db_hi += __shfl_xor_sync(0xffffffffu, db_hi, 1); db_hi += __shfl_xor_sync(0xffffffffu, db_hi, 2);
        if (t == 0) { out[C * C * NTAPS + g] = db_lo; out[C * C * NTAPS + 8 + g] = db_hi; }
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// conv2 weight gradient, staged version (same contract as conv2_wgrad_mma_kernel, default).  The register-path kernel
// above is latency-bound on its scalar global loads (tensor pipe 26 % busy).  Here a block of 8 warps walks GROUPS of
// four consecutive output rows (b, x2, y2 = 4Yg .. 4Yg+3): the 27 input lines (3 x-planes x 9 y-lines of G1 voxels x 16
// channels) and the 4 dy2 lines of a group are fetched by TMA bulk copies into a double-buffered shared-memory stage while
// the previous group is multiplied.  A k-step = 4 rows x 2 adjacent z positions: fragment slot t <-> (row t, z), slot
// t+4 <-> (row t, z+1).  With that choice the four t-lanes of a fragment load read four DIFFERENT staged lines, and the
// line pitch is padded to 4 (mod 32) floats (dy2 lines: 8 mod 32), so the 32 lanes of every LDS hit 32 different banks
// -- positions of one line alone are always 32 floats apart and would collide four ways.
// The 54 n-tiles (27 taps x 2 channel halves) are dealt round-robin to 18 consumer warps, exactly 3 each.  A 19th warp is the
// producer: its lanes issue the <= 31 bulk copies of a group in parallel and it alone waits for a stage to drain.  Stages are
// handed over through mbarriers in both directions (full: TMA bytes landed; empty: 18 warp arrivals), so there is no
// block-wide barrier in the loop and a warp that finishes a group early starts the next one.
// (Round-2 profile of the previous 16-warp form, where thread 0 issued the copies and 6 warps carried a 4th n-tile: 33 % of all
// warp samples were stall_barrier at the per-group __syncthreads; profiles/r02n_*.)
constexpr int WGS_WARPS = 18;                                        // consumer warps
constexpr int WGS_THREADS = (WGS_WARPS + 1) * 32;                    // + the producer warp
constexpr int WGS_NT = 2 * NTAPS / WGS_WARPS;                        // 3 n-tiles per warp
static_assert(WGS_NT * WGS_WARPS == 2 * NTAPS && WGS_WARPS % 2 == 0, "n-tiles must deal evenly, keeping the channel half per warp");
constexpr int WGS_ROWS = 4;


__host__ __device__ inline int pad_mod32(int n, int r) {             // smallest multiple-of-4 value >= n that is r (mod 32)
    int v = (n / 32) * 32 + r;
    return v >= n ? v : v + 32;
}

__global__ void __launch_bounds__(WGS_THREADS, 1)
conv2_wgrad_staged_kernel(const float* __restrict__ y1, const float* __restrict__ stat1, const float* __restrict__ dy2cl,
                          float* __restrict__ part, int G1, int G2, int total_groups, int groups_per_block) {
    extern __shared__ __align__(128) float dsm[];
    __shared__ __align__(8) uint64_t mbar[2], mbar_empty[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), g = lane >> 2, t = lane & 3;
    const int P1 = G1 * G1 * G1, P2 = G2 * G2 * G2;
    const int LP = pad_mod32(G1 * C, 4), DP = pad_mod32(G2 * C, 8);         // floats per staged y1 / dy2 line
    const int STAGE = 27 * LP + WGS_ROWS * DP;
    const int NYG = (G2 + WGS_ROWS - 1) / WGS_ROWS;
    const uint32_t line_bytes = (uint32_t)G1 * C * 4, dy_bytes = (uint32_t)G2 * C * 4;
    // BN1 scale / shift of the two input channels this lane supplies as B columns: ci = 8*hf + g
    // (n-tile nt = warp + 16k <-> tap nt >> 1, channel half nt & 1 = warp & 1: a warp only ever sees one channel half)
    const int hf = warp & 1;
    const float scv = stat1[2 * C + 8 * hf + g], shv = stat1[3 * C + 8 * hf + g];
    const uint32_t full0 = (uint32_t)__cvta_generic_to_shared(&mbar[0]), empty0 = (uint32_t)__cvta_generic_to_shared(&mbar_empty[0]);
    // acc: the MMA accumulators of ONE group; tot: their running sum over the block's groups, added with ordinary
    // round-to-nearest FADDs.  The tensor core's fp32 accumulation truncates, so a chain of thousands of MMAs into one
    // register drifts (measured 5e-4 relative at B = 128 on the conv1 twin of this kernel); 24 MMAs per chain do not.
    float acc[WGS_NT][4], tot[WGS_NT][4];
#pragma unroll
    for (int k = 0; k < WGS_NT; ++k) { tot[k][0] = tot[k][1] = tot[k][2] = tot[k][3] = 0.f; }
    float db_lo = 0.f, db_hi = 0.f;
    const int g0 = blockIdx.x * groups_per_block, g1 = min(total_groups, g0 + groups_per_block);
    if (tid == 0) {
        mbar_init(&mbar[0], 1); mbar_init(&mbar[1], 1);
        mbar_init(&mbar_empty[0], WGS_WARPS); mbar_init(&mbar_empty[1], WGS_WARPS);
        mbar_init_fence();
    }
    __syncthreads();
    auto decode = [&](int grp, int& b, int& x2, int& y20, int& nrows) {
        b = grp / (G2 * NYG);
        const int rem = grp - b * G2 * NYG;
        x2 = rem / NYG;
        y20 = (rem - x2 * NYG) * WGS_ROWS;
        nrows = min(WGS_ROWS, G2 - y20);
    };
    auto issue = [&](int grp, int stage) {                  // the whole producer warp: one bulk copy per lane
        int b, x2, y20, nrows;
        decode(grp, b, x2, y20, nrows);
        const int nlines = 2 * nrows + 1;                   // input y-lines 2*y20 .. 2*y20 + 2*nrows
        const uint32_t bar = full0 + 8u * stage;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(dsm + stage * STAGE);
        if (lane == 0) mbar_expect_tx(bar, 3 * nlines * line_bytes + nrows * dy_bytes);
        __syncwarp();
        if (lane < 3 * nlines) {
            const int i = lane / nlines, yl = lane - i * nlines;
            bulk_g2s(dst + (uint32_t)((i * 9 + yl) * LP * 4),
                     y1 + ((int64_t)b * P1 + ((int64_t)(2 * x2 + i) * G1 + (2 * y20 + yl)) * G1) * C, line_bytes, bar);
        } else if (lane - 3 * nlines < nrows) {
            const int r = lane - 3 * nlines;
            bulk_g2s(dst + (uint32_t)((27 * LP + r * DP) * 4), dy2cl + ((int64_t)b * P2 + (int64_t)(x2 * G2 + y20 + r) * G2) * C,
                     dy_bytes, bar);
        }
    };
    bool ok = true;
    if (warp == WGS_WARPS) {                                // ---- producer warp ----
        for (int grp = g0; grp < g1; ++grp) {
            const int j = grp - g0, stage = j & 1;
            if (j >= 2) ok = mbar_wait_parity(empty0 + 8u * stage, (uint32_t)(((j >> 1) - 1) & 1)) && ok;   // drained by all 18 warps
            issue(grp, stage);
        }
        if (!ok) { asm volatile("trap;"); }
        return;
    }
    const int nzp = (G2 + 1) / 2;
    for (int grp = g0; grp < g1; ++grp) {
        const int j = grp - g0, stage = j & 1;
        ok = mbar_wait_parity(full0 + 8u * stage, (uint32_t)((j >> 1) & 1)) && ok;
        int b, x2, y20, nrows;
        decode(grp, b, x2, y20, nrows);
        const float* xs = dsm + stage * STAGE;
        const float* dys = xs + 27 * LP;
        const bool vrow = t < nrows;
        const int tc = min(t, nrows - 1);                   // rows past the grid alias the last valid row (their A is 0)
#pragma unroll
        for (int k = 0; k < WGS_NT; ++k) { acc[k][0] = acc[k][1] = acc[k][2] = acc[k][3] = 0.f; }
        for (int zp = 0; zp < nzp; ++zp) {
            const int za = 2 * zp, zb = za + 1;
            const bool vb = zb < G2;
            const int zbc = min(zb, G2 - 1);
            // A = dy2^T: a0 (co g, row t, z za), a1 (co g+8, ...), a2 (co g, row t, z zb), a3 (co g+8, ...)
            const float* da = dys + tc * DP + za * C + g;
            const float* db = dys + tc * DP + zbc * C + g;
            const float a0 = vrow ? da[0] : 0.f, a1 = vrow ? da[8] : 0.f;
            const float a2 = (vrow && vb) ? db[0] : 0.f, a3 = (vrow && vb) ? db[8] : 0.f;
            if (warp == 0) { db_lo += a0 + a2; db_hi += a1 + a3; }
            uint32_t ah[4], al[4];
            split_tf32(a0, ah[0], al[0]); split_tf32(a1, ah[1], al[1]); split_tf32(a2, ah[2], al[2]); split_tf32(a3, ah[3], al[3]);
            // B (k = position, n = ci): b0 = x[row t, 2*za + l][ci], b1 = x[row t, 2*zb + l][ci] for the warp's three n-tiles, then
            // the nine MMAs issued tile-interleaved: the three products of one accumulator (lo*hi, hi*lo, hi*hi, in this order)
            // are a dependent chain, and back to back they left the tensor pipe waiting on its own result (NOP + stall_wait
            // after every HMMA in the round-2 profile).
            // (precomputing the per-n-tile line offsets once per kernel was measured no faster: 3.41 vs 3.33 ms/step)
            uint32_t bh0[WGS_NT], bl0[WGS_NT], bh1[WGS_NT], bl1[WGS_NT];
#pragma unroll
            for (int k = 0; k < WGS_NT; ++k) {
                const int tap = (warp + WGS_WARPS * k) >> 1;
                const int i = tap / 9, r9 = tap - 9 * i, jy = r9 / 3, l = r9 - 3 * jy;
                const float* line = xs + (i * 9 + 2 * tc + jy) * LP + l * C + 8 * hf + g;
                const float x0 = fmaxf(fmaf(scv, line[2 * za * C], shv), 0.f);
                const float x1 = fmaxf(fmaf(scv, line[2 * zbc * C], shv), 0.f);
                split_tf32(x0, bh0[k], bl0[k]); split_tf32(x1, bh1[k], bl1[k]);
            }
#pragma unroll
            for (int k = 0; k < WGS_NT; ++k) mma_tf32(acc[k], al[0], al[1], al[2], al[3], bh0[k], bh1[k]);
#pragma unroll
            for (int k = 0; k < WGS_NT; ++k) mma_tf32(acc[k], ah[0], ah[1], ah[2], ah[3], bl0[k], bl1[k]);
#pragma unroll
            for (int k = 0; k < WGS_NT; ++k) mma_tf32(acc[k], ah[0], ah[1], ah[2], ah[3], bh0[k], bh1[k]);
        }
#pragma unroll
        for (int k = 0; k < WGS_NT; ++k) { tot[k][0] += acc[k][0]; tot[k][1] += acc[k][1]; tot[k][2] += acc[k][2]; tot[k][3] += acc[k][3]; }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + 8u * stage);    // this warp is done with the stage (the producer refills it after all 18)
    }
    if (!ok) { asm volatile("trap;"); }
    float* out = part + (int64_t)blockIdx.x * WG_REC;
#pragma unroll
    for (int k = 0; k < WGS_NT; ++k) {
        const int tap = (warp + WGS_WARPS * k) >> 1;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int co = g + 8 * (e >> 1), ci = 8 * hf + 2 * t + (e & 1);
            out[(co * C + ci) * NTAPS + tap] = tot[k][e];
        }
    }
    if (warp == 0) {
        db_lo += __shfl_xor_sync(0xffffffffu, db_lo, 1); db_lo += __shfl_xor_sync(0xffffffffu, db_lo, 2);
        db_hi += __shfl_xor_sync(0xffffffffu, db_hi, 1); db_hi += __shfl_xor_sync(0xffffffffu, db_hi, 2);
        if (t == 0) { out[C * C * NTAPS + g] = db_lo; out[C * C * NTAPS + 8 + g] = db_hi; }
    }
}

}  // namespace

int conv2_mma_chunks(int G2) { return (int)ceil_div((int64_t)G2 * G2 * G2, ROWS_PER_ITEM); }
int conv2_mma_items(int B, int G2) { return B * conv2_mma_chunks(G2); }

int launch_conv2_fwd_mma(const float* y1, const float* stat1, const float* w, const float* bias, float* y2, float* part,
                         int B, int G1, int G2, cudaStream_t stream) {
    // set on every launch (a host call of about a microsecond): the attribute is per device, and a process may drive several
    { int rc_ = ensure_dyn_smem(conv2_fwd_mma_kernel, MMA_SMEM); if (rc_) return rc_; }
    const int chunks = conv2_mma_chunks(G2), items = B * chunks;
    int sms = 148;
    const int grid = std::min(items, sms * 3);
    conv2_fwd_mma_kernel<<<grid, MMA_THREADS, MMA_SMEM, stream>>>(y1, stat1, w, bias, y2, part, G1, G2, chunks, items);
    GNBV_LAUNCH_CHECK("conv2_fwd_mma_kernel");
    return GNBV_OK;
}

int conv2_dgrad_mma_items_per_sample(int G1) {
    int n = 0;
    for (int cls = 0; cls < 8; ++cls) n += dgrad_class(cls, G1).chunks;
    return n;
}

int launch_conv2_dgrad_mma(const float* dy2cl, const float* w, const float* y1, const float* stat1, float* g1, float* bpart,
                           int B, int G1, int G2, cudaStream_t stream) {
    { int rc_ = ensure_dyn_smem(conv2_dgrad_mma_kernel, MMA_SMEM); if (rc_) return rc_; }
    const int ips = conv2_dgrad_mma_items_per_sample(G1), items = B * ips;
    conv2_dgrad_mma_kernel<<<std::min(items, 148 * 3), MMA_THREADS, MMA_SMEM, stream>>>(dy2cl, w, y1, stat1, g1, bpart, G1, G2, ips, items);
    GNBV_LAUNCH_CHECK("conv2_dgrad_mma_kernel");
    return GNBV_OK;
}

int launch_conv2_wgrad_mma(const float* y1, const float* stat1, const float* dy2cl, float* part, int B, int G1, int G2,
                           int nblocks, int rows_per_block, cudaStream_t stream) {
    conv2_wgrad_mma_kernel<<<nblocks, MMA_THREADS, 0, stream>>>(y1, stat1, dy2cl, part, G1, G2, B * G2 * G2, rows_per_block);
    GNBV_LAUNCH_CHECK("conv2_wgrad_mma_kernel");
    return GNBV_OK;
}

int launch_conv2_wgrad_staged(const float* y1, const float* stat1, const float* dy2cl, float* part, int B, int G1, int G2,
                              int max_blocks, int* nblocks_out, cudaStream_t stream) {
    const int LP = pad_mod32(G1 * C, 4), DP = pad_mod32(G2 * C, 8);
    const size_t smem = (size_t)2 * (27 * LP + WGS_ROWS * DP) * 4;
    GNBV_REQUIRE(smem <= 200 * 1024, "conv2 wgrad: grid too large for the staged kernel's shared-memory lines");
    { int rc_ = ensure_dyn_smem(conv2_wgrad_staged_kernel, smem); if (rc_) return rc_; }
    const int total = B * G2 * (int)ceil_div(G2, WGS_ROWS);
    const int want = std::min(std::min(max_blocks, 148), total);
    const int gpb = (int)ceil_div(total, want), nblk = (int)ceil_div(total, gpb);
    conv2_wgrad_staged_kernel<<<nblk, WGS_THREADS, smem, stream>>>(y1, stat1, dy2cl, part, G1, G2, total, gpb);
    GNBV_LAUNCH_CHECK("conv2_wgrad_staged_kernel");
    *nblocks_out = nblk;
    return GNBV_OK;
}

}  // namespace gnbv
