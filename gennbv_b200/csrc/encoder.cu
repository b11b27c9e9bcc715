// encoder.cu -- Hybrid_Encoder (gennbv/network/hybrid_encoder.py:12-91) + actor / critic heads
// (stable_baselines3/common/policies.py:954-1090) forward and backward, hand-written for sm_100a.
//
// Layout decisions (B200-first, not a translation of the cuDNN/cuBLAS call sequence):
//   * the observation row is consumed in place: `state` columns feed the positional encoding, the tri-class grid
//     columns feed conv1 directly (no reshape / slice copies), the rgb columns are never read (the reference's
//     forward ignores them, hybrid_encoder.py:69-91);
//   * conv1 (Cin = 1, 27 taps) is a register-resident stencil, one thread per output voxel x 16 channels,
//     output kept channels-last (NDHWC) so that conv2 reads 64 B per voxel with 128-bit loads;
//   * BatchNorm3d + ReLU are never materialised for layer 1: conv2 applies a*y+b / max on load; batch statistics
//     are reduced from per-block (sum, sum of squares) partials in double in a fixed order (deterministic);
//   * conv2 (K = 432, N = 16) is an implicit GEMM on CUDA cores: 4 output voxels x 16 channels per thread,
//     weights broadcast from shared memory with LDS.128 (1 shared load per 16 FMA);
//   * Linear layers use the fp32 split-K GEMM of gemm.cu (fp32 everywhere: the 1e-4 parity budget rules out
//     bf16 / tf32 operands, see gemm.cuh).
#include "gemm.cuh"
#include "conv2_tc.cuh"
#include "conv2_mma.cuh"
#include "conv2_ts.cuh"
#include "mma.cuh"
#include "tma.cuh"
#include "encoder.cuh"
#include "sem2d.cuh"

#include <stdlib.h>

#include <algorithm>

namespace gnbv {

constexpr int C1 = 16;            // conv channels (hybrid_encoder.py:32,35)
constexpr int TAPS = 27;
constexpr int CONV1_THREADS = 256;
constexpr int CONV2_THREADS = 128;
constexpr int CONV2_ZT = 4;       // output voxels (along z) per thread


constexpr int PART_STRIDE = 2 * C1 + 4;      // per block: mean[16], M2[16], count, pad

// Block-level (count, mean, M2) of `nv` values per thread and channel (acc[k][c], k < nv valid ones), written to
// part[PART_STRIDE].  Two-pass inside each warp (mean first, then centred squares) and Chan's merge across warps:
// no E[x^2] - mean^2 cancellation anywhere.
template <int NT, int NV>
__device__ __forceinline__ void block_stats(const float (&acc)[NV][C1], int nvalid, float* __restrict__ part,
                                            float (*red)[PART_STRIDE]) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int nw = warp_sum_i(nvalid);
#pragma unroll
    for (int c = 0; c < C1; ++c) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < NV; ++k) if (k < nvalid) s += acc[k][c];
        s = warp_sum(s);
        const float mean = nw > 0 ? s / (float)nw : 0.f;
        float m2 = 0.f;
#pragma unroll
        for (int k = 0; k < NV; ++k) if (k < nvalid) { float dlt = acc[k][c] - mean; m2 = fmaf(dlt, dlt, m2); }
        m2 = warp_sum(m2);
        if (lane == 0) { red[wid][c] = mean; red[wid][C1 + c] = m2; }
    }
    if (lane == 0) red[wid][2 * C1] = (float)nw;
    __syncthreads();
    if (tid < C1) {
        float n = 0.f, mean = 0.f, M2 = 0.f;
#pragma unroll
        for (int w = 0; w < NT / 32; ++w) {
            float cnt = red[w][2 * C1];
            if (cnt > 0.f) {
                float delta = red[w][tid] - mean, nt = n + cnt;
                mean += delta * cnt / nt;
                M2 += red[w][C1 + tid] + delta * delta * n * cnt / nt;
                n = nt;
            }
        }
        part[tid] = mean; part[C1 + tid] = M2;
        if (tid == 0) part[2 * C1] = n;
    }
}

// ------------------------------------------------------------------------------------------------ posenc
// positional_encoding (hybrid_encoder.py:56-67): per pose p[6] -> [sin(p_i * f), ...] then [cos(...)], f in {1,2},
// order (i major, f minor); 24 values per pose.
__global__ void posenc_kernel(const float* __restrict__ obs, int64_t obs_stride, const int64_t* __restrict__ rows,
                              float* __restrict__ pe, int B, int S) {
    // S = buffer_size * 6 state values per row; output [B, 4*S]
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)B * S) return;
    int b = (int)(idx / S), s = (int)(idx - (int64_t)b * S);
    int pose = s / 6, i = s - pose * 6;
    float x = obs[(rows ? rows[b] : (int64_t)b) * obs_stride + s];
    float* o = pe + (int64_t)b * 4 * S + pose * 24;
    float x1 = x * 1.0f, x2 = x * 2.0f;
    o[2 * i] = sinf(x1); o[2 * i + 1] = sinf(x2);
    o[12 + 2 * i] = cosf(x1); o[12 + 2 * i + 1] = cosf(x2);
}

// ------------------------------------------------------------------------------------------------ conv1
// Conv3d(1,16,3,stride 2) on the grid columns of the observation.  y1 [B, G1^3, 16] (pre-BN, channels-last).
// Per-block partial (sum, sumsq) per channel -> part[blk][32] when stats != nullptr.
__global__ void __launch_bounds__(CONV1_THREADS)
conv1_fwd_kernel(const float* __restrict__ obs, int64_t obs_stride, const int64_t* __restrict__ rows, int64_t grid_off,
                 const float* __restrict__ w,
                 const float* __restrict__ bias, float* __restrict__ y1, float* __restrict__ part, int G, int G1) {
    __shared__ __align__(16) float ws[TAPS][C1];
    __shared__ float bs[C1];
    __shared__ float red[CONV1_THREADS / 32][PART_STRIDE];
    const int tid = threadIdx.x, b = blockIdx.y;
    for (int i = tid; i < TAPS * C1; i += CONV1_THREADS) {
        int tap = i / C1, c = i - tap * C1;
        ws[tap][c] = w[c * TAPS + tap];                 // weight [16,1,3,3,3]
    }
    if (tid < C1) bs[tid] = bias[tid];
    __syncthreads();
    const int P1 = G1 * G1 * G1;
    const int p = blockIdx.x * CONV1_THREADS + tid;
    float acc1[1][C1];
    float (&acc)[C1] = acc1[0];
    const bool valid = p < P1;
    if (valid) {
        int z1 = p % G1, t = p / G1;
        int yy = t % G1, xx = t / G1;
        const float* in = obs + (rows ? rows[b] : (int64_t)b) * obs_stride + grid_off + ((int64_t)(2 * xx) * G + 2 * yy) * G + 2 * z1;
#pragma unroll
        for (int c = 0; c < C1; ++c) acc[c] = bs[c];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j)
#pragma unroll
                for (int l = 0; l < 3; ++l) {
                    float v = __ldg(in + ((int64_t)i * G + j) * G + l);
                    const float4* wr = reinterpret_cast<const float4*>(ws[(i * 3 + j) * 3 + l]);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float4 wv = wr[q];
                        acc[4 * q + 0] = fmaf(v, wv.x, acc[4 * q + 0]);
                        acc[4 * q + 1] = fmaf(v, wv.y, acc[4 * q + 1]);
                        acc[4 * q + 2] = fmaf(v, wv.z, acc[4 * q + 2]);
                        acc[4 * q + 3] = fmaf(v, wv.w, acc[4 * q + 3]);
                    }
                }
        float4* o = reinterpret_cast<float4*>(y1 + ((int64_t)b * P1 + p) * C1);
#pragma unroll
        for (int q = 0; q < 4; ++q) o[q] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
    } else {
#pragma unroll
        for (int c = 0; c < C1; ++c) acc[c] = 0.f;
    }
    if (part)
        block_stats<CONV1_THREADS, 1>(acc1, valid ? 1 : 0, part + ((int64_t)b * gridDim.x + blockIdx.x) * PART_STRIDE, red);
}

// ------------------------------------------------------------------------------------------------ BN statistics
// Merges the per-block (count, mean, M2) partials (Chan et al.) in double in a fixed order -- 32 lanes per channel
// stride over the blocks, then a 5-step butterfly -- and produces the affine form y = a*x + b.  In training mode it
// also updates the running statistics like nn.BatchNorm3d (biased variance to normalise, unbiased for running_var,
// momentum 0.1, eps 1e-5).  stat layout [4][16]: mean, invstd, a, b.
__device__ __forceinline__ void chan_merge(double& n, double& mean, double& M2, double nb, double mb, double M2b) {
    if (nb <= 0) return;
    double nt = n + nb, delta = mb - mean;
    mean += delta * nb / nt;
    M2 += M2b + delta * delta * n * nb / nt;
    n = nt;
}

__global__ void __launch_bounds__(32 * C1)
bn_finalize_kernel(const float* __restrict__ part, int nblk, const float* __restrict__ gamma,
                   const float* __restrict__ beta, float* __restrict__ running_mean, float* __restrict__ running_var,
                   int64_t* __restrict__ num_batches_tracked, float* __restrict__ stat, float eps, float momentum,
                   const int64_t* __restrict__ freeze) {
    const int c = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double n = 0, mean = 0, M2 = 0;
    for (int k = lane; k < nblk; k += 32) {
        const float* pr = part + (int64_t)k * PART_STRIDE;
        chan_merge(n, mean, M2, (double)pr[2 * C1], (double)pr[c], (double)pr[C1 + c]);
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        double nb = __shfl_xor_sync(0xffffffffu, n, o), mb = __shfl_xor_sync(0xffffffffu, mean, o),
               M2b = __shfl_xor_sync(0xffffffffu, M2, o);
        // merge in a lane-symmetric way so that both partners obtain identical results
        double n_lo = (lane & o) ? nb : n, m_lo = (lane & o) ? mb : mean, q_lo = (lane & o) ? M2b : M2;
        double n_hi = (lane & o) ? n : nb, m_hi = (lane & o) ? mean : mb, q_hi = (lane & o) ? M2 : M2b;
        chan_merge(n_lo, m_lo, q_lo, n_hi, m_hi, q_hi);
        n = n_lo; mean = m_lo; M2 = q_lo;
    }
    if (lane != 0) return;
    double var = n > 0 ? M2 / n : 0.0;
    float invstd = (float)(1.0 / sqrt(var + (double)eps));
    float a = gamma[c] * invstd;
    stat[0 * C1 + c] = (float)mean;
    stat[1 * C1 + c] = invstd;
    stat[2 * C1 + c] = a;
    stat[3 * C1 + c] = beta[c] - (float)mean * a;
    // `freeze` (device flag, may be null): a set flag leaves the running statistics alone -- the PPO update's sticky
    // KL-stop flag, so that minibatches enqueued after the stop change no state (ppo_update.cu)
    if (running_mean && !(freeze && *freeze != 0)) {
        double unbiased = n > 1 ? M2 / (n - 1) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
        if (c == 0 && num_batches_tracked) *num_batches_tracked += 1;
    }
}

// eval mode: a = gamma / sqrt(running_var + eps), b = beta - running_mean * a
__global__ void bn_eval_affine_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                                      const float* __restrict__ running_mean, const float* __restrict__ running_var,
                                      float* __restrict__ stat, float eps) {
    const int c = threadIdx.x;
    if (c >= C1) return;
    float invstd = 1.0f / sqrtf(running_var[c] + eps);
    float a = gamma[c] * invstd;
    stat[0 * C1 + c] = running_mean[c];
    stat[1 * C1 + c] = invstd;
    stat[2 * C1 + c] = a;
    stat[3 * C1 + c] = beta[c] - running_mean[c] * a;
}

// ------------------------------------------------------------------------------------------------ conv2
// Conv3d(16,16,3,stride 2) as an implicit GEMM; input y1 [B,G1^3,16] with BN1 affine + ReLU applied on load,
// output y2 [B,16,G2^3] (pre-BN, channel-major = the order nn.Flatten feeds the grid Linear).
// (forcing 4 resident blocks/SM via launch bounds spills and measured no gain: the kernel is LSU/FMA co-bound, not occupancy-bound)
__global__ void __launch_bounds__(CONV2_THREADS)
conv2_fwd_kernel(const float* __restrict__ y1, const float* __restrict__ stat1, const float* __restrict__ w,
                 const float* __restrict__ bias, float* __restrict__ y2, float* __restrict__ part, int G1, int G2) {
    __shared__ __align__(16) float ws[TAPS][C1][C1];        // [tap][ci][co]
    __shared__ float a1s[C1], b1s[C1], bs[C1];
    __shared__ float red[CONV2_THREADS / 32][PART_STRIDE];
    const int tid = threadIdx.x, b = blockIdx.y;
    for (int i = tid; i < TAPS * C1 * C1; i += CONV2_THREADS) {
        int tap = i / (C1 * C1), r = i - tap * C1 * C1, ci = r / C1, co = r - ci * C1;
        ws[tap][ci][co] = w[(co * C1 + ci) * TAPS + tap];    // weight [16,16,3,3,3]
    }
    if (tid < C1) { a1s[tid] = stat1[2 * C1 + tid]; b1s[tid] = stat1[3 * C1 + tid]; bs[tid] = bias[tid]; }
    __syncthreads();
    const int ZQ = (G2 + CONV2_ZT - 1) / CONV2_ZT;
    const int items = G2 * G2 * ZQ;
    const int item = blockIdx.x * CONV2_THREADS + tid;
    const int P1 = G1 * G1 * G1, P2 = G2 * G2 * G2;
    float acc[CONV2_ZT][C1];
    const bool active = item < items;
    int x2 = 0, yy2 = 0, z20 = 0;
    if (active) {
        int q = item % ZQ, t = item / ZQ;
        yy2 = t % G2; x2 = t / G2; z20 = q * CONV2_ZT;
#pragma unroll
        for (int s = 0; s < CONV2_ZT; ++s)
#pragma unroll
            for (int c = 0; c < C1; ++c) acc[s][c] = bs[c];
        const float* in_b = y1 + (int64_t)b * P1 * C1;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                const float* row = in_b + (((int64_t)(2 * x2 + i) * G1 + (2 * yy2 + j)) * G1) * C1;
#pragma unroll
                for (int l = 0; l < 3; ++l) {
                    const int tap = (i * 3 + j) * 3 + l;
#pragma unroll
                    for (int cq = 0; cq < 4; ++cq) {             // 4 input channels at a time
                        float4 xin[CONV2_ZT];
                        const float4 av = *reinterpret_cast<const float4*>(&a1s[4 * cq]);
                        const float4 bv = *reinterpret_cast<const float4*>(&b1s[4 * cq]);
#pragma unroll
                        for (int s = 0; s < CONV2_ZT; ++s) {
                            int zi = 2 * (z20 + s) + l;
                            zi = min(zi, G1 - 1);               // out-of-range lanes (z2 >= G2) read a valid voxel, result unused
                            float4 v = __ldg(reinterpret_cast<const float4*>(row + (int64_t)zi * C1) + cq);
                            v.x = fmaxf(fmaf(av.x, v.x, bv.x), 0.f); v.y = fmaxf(fmaf(av.y, v.y, bv.y), 0.f);
                            v.z = fmaxf(fmaf(av.z, v.z, bv.z), 0.f); v.w = fmaxf(fmaf(av.w, v.w, bv.w), 0.f);
                            xin[s] = v;
                        }
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc) {
                            const float4* wr = reinterpret_cast<const float4*>(ws[tap][4 * cq + cc]);
#pragma unroll
                            for (int oq = 0; oq < 4; ++oq) {
                                const float4 wv = wr[oq];
#pragma unroll
                                for (int s = 0; s < CONV2_ZT; ++s) {
                                    const float xv = cc == 0 ? xin[s].x : (cc == 1 ? xin[s].y : (cc == 2 ? xin[s].z : xin[s].w));
                                    acc[s][4 * oq + 0] = fmaf(xv, wv.x, acc[s][4 * oq + 0]);
                                    acc[s][4 * oq + 1] = fmaf(xv, wv.y, acc[s][4 * oq + 1]);
                                    acc[s][4 * oq + 2] = fmaf(xv, wv.z, acc[s][4 * oq + 2]);
                                    acc[s][4 * oq + 3] = fmaf(xv, wv.w, acc[s][4 * oq + 3]);
                                }
                            }
                        }
                    }
                }
            }
        const int pbase = (x2 * G2 + yy2) * G2;
#pragma unroll
        for (int s = 0; s < CONV2_ZT; ++s) {
            int z2 = z20 + s;
            if (z2 < G2) {
#pragma unroll
                for (int c = 0; c < C1; ++c) y2[((int64_t)b * C1 + c) * P2 + pbase + z2] = acc[s][c];
            }
        }
    }
    if (part) {
        const int nvalid = active ? min(CONV2_ZT, G2 - z20) : 0;
        block_stats<CONV2_THREADS, CONV2_ZT>(acc, nvalid, part + ((int64_t)b * gridDim.x + blockIdx.x) * PART_STRIDE, red);
    }
}

// act2 = relu(a2[c] * y2 + b2[c]), [B,16,P2] channel-major
__global__ void bn_relu_apply_kernel(const float* __restrict__ y2, const float* __restrict__ stat2, float* __restrict__ act2,
                                     int64_t total, int P2) {
    // four consecutive elements per thread (the flat [B,16,P2] array is 16-byte aligned and total % 4 == 0 -- 16 * P2 per sample);
    // the channel can change inside a quad only at a row end
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, idx = 4 * q;
    if (idx >= total) return;
    const float4 v = *reinterpret_cast<const float4*>(y2 + idx);
    const int64_t row = idx / P2;
    const int rem = (int)(idx - row * P2);
    const float in[4] = {v.x, v.y, v.z, v.w};
    float out[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int c = (int)((row + (rem + e >= P2 ? 1 : 0)) % C1);
        out[e] = fmaxf(fmaf(stat2[2 * C1 + c], in[e], stat2[3 * C1 + c]), 0.f);
    }
    *reinterpret_cast<float4*>(act2 + idx) = make_float4(out[0], out[1], out[2], out[3]);
}


// =====================================================================================================================
// backward kernels
// =====================================================================================================================

// dy[i,j] = y[i,j] > 0 ? dy[i,j] : 0   (ReLU backward from the stored post-activation), row-strided buffers
__global__ void relu_mask_kernel(const float* __restrict__ dy_in, int64_t ldi, float* __restrict__ dy, int64_t ldd,
                                 const float* __restrict__ y, int64_t ldy, int rows, int cols) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)rows * cols) return;
    int r = (int)(idx / cols), c = (int)(idx - (int64_t)r * cols);
    dy[(int64_t)r * ldd + c] = (y[(int64_t)r * ldy + c] > 0.f) ? dy_in[(int64_t)r * ldi + c] : 0.f;
}

// db[j] = sum_i dy[i,j]  (rows summed in order: deterministic)
// column sums of a [rows, cols] matrix (row stride ld): 32 columns x 8 interleaved row slices per block, the slices combined in a
// fixed order (deterministic), accumulation in double.  Launch with colsum_blocks(cols) blocks of 256 threads.
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ dy, int64_t ld, int rows, int cols, float* __restrict__ db) {
    __shared__ double sh[8][33];
    const int cx = threadIdx.x & 31, sy = threadIdx.x >> 5, j = blockIdx.x * 32 + cx;
    double s = 0.0;
    if (j < cols)
        for (int i = sy; i < rows; i += 8) s += (double)dy[(int64_t)i * ld + j];
    sh[sy][cx] = s;
    __syncthreads();
    if (sy == 0 && j < cols) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += sh[k][cx];
        db[j] = (float)t;
    }
}
static inline unsigned colsum_blocks(int cols) { return (unsigned)((cols + 31) / 32); }

// BN2 backward, pass 1: per (sample, channel) sums of g = dact2 * [act2 > 0] and g * xhat.  part [B][16][2]
__global__ void __launch_bounds__(256)
bn2_bwd_reduce_kernel(const float* __restrict__ dact2, const float* __restrict__ act2, const float* __restrict__ y2,
                      const float* __restrict__ stat2, float* __restrict__ part, int P2) {
    __shared__ float sh[2][8];
    const int c = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const int64_t base = ((int64_t)b * C1 + c) * P2;
    const float mean = stat2[c], invstd = stat2[C1 + c], sc = stat2[2 * C1 + c], sh_ = stat2[3 * C1 + c];
    float s1 = 0.f, s2 = 0.f;
    (void)act2;     // the ReLU mask is recomputed from y2 with the forward's own expression (bn_relu_apply_kernel): 55 MB less to read
    for (int p = tid; p < P2; p += 256) {
        const float yv = y2[base + p];
        float g = fmaf(sc, yv, sh_) > 0.f ? dact2[base + p] : 0.f;
        s1 += g;
        s2 = fmaf(g, (yv - mean) * invstd, s2);
    }
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    if ((tid & 31) == 0) { sh[0][tid >> 5] = s1; sh[1][tid >> 5] = s2; }
    __syncthreads();
    if (tid < 2) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += sh[tid][w];
        part[((int64_t)b * C1 + c) * 2 + tid] = t;
    }
}

// Sums `nrec` records of (S1[16], S2[16]) laid out as rec[k][c*2 + {0,1}] (layout A) or rec[k][{0,1}*16 + c] (layout B)
// in double, fixed order; writes d_gamma = S2, d_beta = S1 and coef[2][16] = S1/n, S2/n.
__global__ void __launch_bounds__(32 * C1)
bn_bwd_finalize_kernel(const float* __restrict__ rec, int nrec, int rec_stride, int layout_b, double count,
                       float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ coef, int zero_coef) {
    const int c = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double s1 = 0, s2 = 0;
    for (int k = lane; k < nrec; k += 32) {
        const float* r = rec + (int64_t)k * rec_stride;
        s1 += layout_b ? r[c] : r[c * 2];
        s2 += layout_b ? r[C1 + c] : r[c * 2 + 1];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    if (lane == 0) {
        dgamma[c] = (float)s2; dbeta[c] = (float)s1;
        // eval-mode BN is a fixed affine map: no batch-statistics terms in dx
        coef[c] = zero_coef ? 0.f : (float)(s1 / count); coef[C1 + c] = zero_coef ? 0.f : (float)(s2 / count);
    }
}

// BN2 backward, pass 2: dy2 = a2 * (g - S1/n - xhat * S2/n), written channels-last [B,P2,16] for the conv2 backward kernels
__global__ void __launch_bounds__(256)
bn2_bwd_apply_kernel(const float* __restrict__ dact2, const float* __restrict__ act2, const float* __restrict__ y2,
                     const float* __restrict__ stat2, const float* __restrict__ coef, float* __restrict__ dy2cl, int P2) {
    const int b = blockIdx.y, p = blockIdx.x * 256 + threadIdx.x;
    if (p >= P2) return;
    float out[C1];
#pragma unroll
    for (int c = 0; c < C1; ++c) {
        const int64_t i = ((int64_t)b * C1 + c) * P2 + p;
        const float yv = y2[i];
        const float g = fmaf(stat2[2 * C1 + c], yv, stat2[3 * C1 + c]) > 0.f ? dact2[i] : 0.f;   // same mask as act2 > 0
        const float xhat = (yv - stat2[c]) * stat2[C1 + c];
        out[c] = stat2[2 * C1 + c] * (g - coef[c] - xhat * coef[C1 + c]);
    }
    float4* o = reinterpret_cast<float4*>(dy2cl + ((int64_t)b * P2 + p) * C1);
#pragma unroll
    for (int q = 0; q < 4; ++q) o[q] = make_float4(out[4 * q], out[4 * q + 1], out[4 * q + 2], out[4 * q + 3]);
}

// conv2 weight gradient.  dW2[co,ci,tap] = sum_{b,pos} dy2[b,pos,co] * relu(bn1(y1))[b, 2pos+tap, ci].
//
// Work unit = one output z-row (b, x2, y2): G2 positions whose 3x3x(2*G2+1) input neighbourhood is nine CONTIGUOUS
// channels-last z-lines of y1 (G1*64 B each) plus one contiguous dy2 line (G2*64 B).  A block walks a range of rows with
// a two-stage pipeline: one thread issues ten TMA bulk copies (cp.async.bulk ... mbarrier::complete_tx) for row i+1 while
// all 256 threads reduce row i out of shared memory.  Thread (g, ci, tg): positions z2 = g mod 4, input channel ci, the 7
// taps [7tg, 7tg+7) x all 16 output channels (112 accumulators): per position 4 LDS.128 (dy2) + 7 LDS.32 feed 112 FMA.
// Partials: part[blk][6912 + 16] in weight layout [co][ci][tap] followed by db2[16].
constexpr int WG2_THREADS = 256;
constexpr int WG2_TT = 7;
constexpr int WG2_MAX_BLOCKS = 592;
constexpr int WG2_REC = C1 * C1 * TAPS + C1;

__global__ void __launch_bounds__(WG2_THREADS)
conv2_wgrad_kernel(const float* __restrict__ y1, const float* __restrict__ stat1, const float* __restrict__ dy2cl,
                   float* __restrict__ part, int G1, int G2, int total_rows, int rows_per_block) {
    extern __shared__ __align__(128) float dsm[];          // [2 stages][9 lines + dy line]  /  later [3][WG2_REC]
    __shared__ __align__(8) uint64_t mbar[2];
    const int tid = threadIdx.x, g = tid >> 6, t64 = tid & 63, ci = t64 >> 2, tg = t64 & 3;
    const float a1 = stat1[2 * C1 + ci], b1 = stat1[3 * C1 + ci];
    const int P1 = G1 * G1 * G1, P2 = G2 * G2 * G2;
    const int LINE = ((G1 * C1 + 31) / 32) * 32;           // floats per staged y1 line (128 B multiple)
    const int DYL = ((G2 * C1 + 31) / 32) * 32;
    const int STAGE = 9 * LINE + DYL;
    const uint32_t line_bytes = (uint32_t)G1 * C1 * 4, dy_bytes = (uint32_t)G2 * C1 * 4;
    int off[WG2_TT];
#pragma unroll
    for (int tt = 0; tt < WG2_TT; ++tt) {
        const int tap = min(WG2_TT * tg + tt, TAPS - 1);
        off[tt] = ((tap / 9) * 3 + (tap / 3) % 3) * LINE + (tap % 3) * C1 + ci;       // line (i,j), voxel offset l
    }
    const int ntap = min(WG2_TT, TAPS - WG2_TT * tg);
    float acc[WG2_TT][C1];
#pragma unroll
    for (int tt = 0; tt < WG2_TT; ++tt)
#pragma unroll
        for (int c = 0; c < C1; ++c) acc[tt][c] = 0.f;
    float dbs[C1];
#pragma unroll
    for (int c = 0; c < C1; ++c) dbs[c] = 0.f;

    const int r0 = blockIdx.x * rows_per_block, r1 = min(total_rows, r0 + rows_per_block);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&mbar[0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&mbar[1])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int row, int stage) {                  // executed by thread 0 only
        const int b = row / (G2 * G2), rem = row - b * G2 * G2, x2 = rem / G2, yy2 = rem - x2 * G2;
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&mbar[stage]);
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(dsm + stage * STAGE);
        mbar_expect_tx(bar, 9 * line_bytes + dy_bytes);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j)
                bulk_g2s(dst + (uint32_t)((i * 3 + j) * LINE * 4),
                         y1 + ((int64_t)b * P1 + ((int64_t)(2 * x2 + i) * G1 + (2 * yy2 + j)) * G1) * C1, line_bytes, bar);
        bulk_g2s(dst + (uint32_t)(9 * LINE * 4), dy2cl + ((int64_t)b * P2 + (int64_t)(x2 * G2 + yy2) * G2) * C1, dy_bytes, bar);
    };
    if (tid == 0 && r0 < r1) issue(r0, 0);
    uint32_t phase[2] = {0, 0};
    bool ok = true;
    for (int row = r0; row < r1; ++row) {
        const int stage = (row - r0) & 1;
        if (tid == 0 && row + 1 < r1) issue(row + 1, stage ^ 1);     // stage^1 was released by the barrier ending row-1
        ok = mbar_wait_parity((uint32_t)__cvta_generic_to_shared(&mbar[stage]), phase[stage]) && ok;
        phase[stage] ^= 1;
        const float* lines = dsm + stage * STAGE;
        const float* dyl = lines + 9 * LINE;
        for (int z2 = g; z2 < G2; z2 += 4) {
            const float4* dyp = reinterpret_cast<const float4*>(dyl + z2 * C1);
            const float4 d0 = dyp[0], d1 = dyp[1], d2 = dyp[2], d3 = dyp[3];
            const float dy[C1] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w, d2.x, d2.y, d2.z, d2.w, d3.x, d3.y, d3.z, d3.w};
            if (t64 == 0) {
#pragma unroll
                for (int c = 0; c < C1; ++c) dbs[c] += dy[c];
            }
            const float* in = lines + 2 * z2 * C1;
#pragma unroll
            for (int tt = 0; tt < WG2_TT; ++tt) {
                if (tt < ntap) {
                    float x = in[off[tt]];
                    x = fmaxf(fmaf(a1, x, b1), 0.f);
#pragma unroll
                    for (int c = 0; c < C1; ++c) acc[tt][c] = fmaf(dy[c], x, acc[tt][c]);
                }
            }
        }
        __syncthreads();                                    // everyone is done with this stage before it is refilled
    }
    if (!ok) { asm volatile("trap;"); }
    // cross-group reduction in a fixed order: groups 1..3 publish, group 0 adds them in order and writes the record
    float* sred = dsm;
    if (g > 0) {
        float* dst = sred + (g - 1) * WG2_REC;
#pragma unroll
        for (int tt = 0; tt < WG2_TT; ++tt)
            if (tt < ntap) {
#pragma unroll
                for (int c = 0; c < C1; ++c) dst[(c * C1 + ci) * TAPS + WG2_TT * tg + tt] = acc[tt][c];
            }
        if (t64 == 0) {
#pragma unroll
            for (int c = 0; c < C1; ++c) dst[C1 * C1 * TAPS + c] = dbs[c];
        }
    }
    __syncthreads();
    if (g == 0) {
        float* out = part + (int64_t)blockIdx.x * WG2_REC;
#pragma unroll
        for (int tt = 0; tt < WG2_TT; ++tt)
            if (tt < ntap) {
#pragma unroll
                for (int c = 0; c < C1; ++c) {
                    const int o = (c * C1 + ci) * TAPS + WG2_TT * tg + tt;
                    out[o] = ((acc[tt][c] + sred[o]) + sred[WG2_REC + o]) + sred[2 * WG2_REC + o];
                }
            }
        if (t64 == 0) {
#pragma unroll
            for (int c = 0; c < C1; ++c) {
                const int o = C1 * C1 * TAPS + c;
                out[o] = ((dbs[c] + sred[o]) + sred[WG2_REC + o]) + sred[2 * WG2_REC + o];
            }
        }
    }
}

// out[j] = sum_k part[k][j] for j < rec (fixed order, double accumulation); out split over two destinations
__global__ void __launch_bounds__(256)
reduce_records_kernel(const float* __restrict__ part, int nrec, int rec, float* __restrict__ out_a, int na,
                      float* __restrict__ out_b) {
    // 32 record columns x 8 interleaved record slices per block, slices combined in a fixed order (deterministic), double
    // accumulation; launch with (rec + 31) / 32 blocks of 256 threads.  (One thread per column walking all the records serially
    // took 52 us per call for 148 x 6928 / 1176 x 448 floats.)
    __shared__ double sh[8][33];
    const int cx = threadIdx.x & 31, sy = threadIdx.x >> 5, j = blockIdx.x * 32 + cx;
    double s = 0.0;
    if (j < rec)
        for (int k = sy; k < nrec; k += 8) s += (double)part[(int64_t)k * rec + j];
    sh[sy][cx] = s;
    __syncthreads();
    if (sy == 0 && j < rec) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += sh[k][cx];
        if (j < na) out_a[j] = (float)t;
        else out_b[j - na] = (float)t;
    }
}

// conv2 data gradient + ReLU/BN1 backward statistics.
//   dact1[b,vi,ci] = sum_{taps with (vi - tap) even and in range} sum_co dy2[b,(vi-tap)/2,co] * W2[co,ci,tap]
//   g1 = dact1 * [bn1(y1) > 0]  -> stored channels-last; per-block (sum g1, sum g1*xhat1) partials -> bpart[blk][32]
// One thread owns 4 conv1-output voxels of equal z parity (z = zpar + 2(4zq + s)) x 16 input channels: the 4 voxels
// share their tap set, so each LDS.128 of weights feeds 16 FMA.
constexpr int DG2_THREADS = 128;
constexpr int DG2_ZT = 4;
__global__ void __launch_bounds__(DG2_THREADS)
conv2_dgrad_kernel(const float* __restrict__ dy2cl, const float* __restrict__ w, const float* __restrict__ y1,
                   const float* __restrict__ stat1, float* __restrict__ g1, float* __restrict__ bpart, int G1, int G2,
                   int vblocks_per_sample, int total_vblocks) {
    __shared__ __align__(16) float ws[TAPS][C1][C1];        // [tap][co][ci]
    __shared__ float red[DG2_THREADS / 32][2 * C1];
    const int tid = threadIdx.x;
    // persistent block: the 27 KB weight tile is staged once, then the block walks virtual blocks (sample, item chunk)
    for (int i = tid; i < TAPS * C1 * C1; i += DG2_THREADS) {
        int tap = i / (C1 * C1), r = i - tap * C1 * C1, co = r / C1, ci = r - co * C1;
        ws[tap][co][ci] = w[(co * C1 + ci) * TAPS + tap];
    }
    __syncthreads();
    const int P1 = G1 * G1 * G1, P2 = G2 * G2 * G2;
    const int ZQ = ((G1 + 1) / 2 + DG2_ZT - 1) / DG2_ZT;
    // Item order: parity class (x&1, y&1, z&1) is the SLOWEST index, so that the 32 threads of a warp (almost always)
    // share one class and therefore one tap set -- the tap loops below are then warp-uniform.
    const int NE = (G1 + 1) / 2, NO = G1 / 2;                 // number of even / odd coordinates
    const int items = G1 * G1 * 2 * ZQ;
    for (int vb = blockIdx.x; vb < total_vblocks; vb += gridDim.x) {
    const int b = vb / vblocks_per_sample, vblk = vb - b * vblocks_per_sample;
    const int item = vblk * DG2_THREADS + tid;
    const bool active = item < items;
    float acc[DG2_ZT][C1];
#pragma unroll
    for (int s = 0; s < DG2_ZT; ++s)
#pragma unroll
        for (int c = 0; c < C1; ++c) acc[s][c] = 0.f;
    float s1[C1], s2[C1];
#pragma unroll
    for (int c = 0; c < C1; ++c) { s1[c] = 0.f; s2[c] = 0.f; }
    if (active) {
        // decode: classes ordered (xpar, ypar, zpar); inside a class: (x index, y index, zq)
        int t = item, xpar = 0, ypar = 0, zpar = 0;
        for (int cls = 0; cls < 8; ++cls) {
            const int cx = (cls >> 2) & 1, cy = (cls >> 1) & 1, cz = cls & 1;
            const int cnt = (cx ? NO : NE) * (cy ? NO : NE) * ZQ;
            if (t < cnt) { xpar = cx; ypar = cy; zpar = cz; break; }
            t -= cnt;
        }
        const int zq = t % ZQ; t /= ZQ;
        const int ny = ypar ? NO : NE;
        const int yi = 2 * (t % ny) + ypar, xi = 2 * (t / ny) + xpar;
        const int zfirst = zpar + 2 * DG2_ZT * zq;                   // z of s = 0; z(s) = zfirst + 2s
        for (int i = 0; i < 3; ++i) {
            const int xr = xi - i;
            if (xr < 0 || (xr & 1) || (xr >> 1) >= G2) continue;
            for (int j = 0; j < 3; ++j) {
                const int yr = yi - j;
                if (yr < 0 || (yr & 1) || (yr >> 1) >= G2) continue;
                const int64_t row2 = ((int64_t)b * P2 + ((int64_t)(xr >> 1) * G2 + (yr >> 1)) * G2) * C1;
                for (int l = 0; l < 3; ++l) {
                    if ((zpar - l) & 1) continue;
                    const int z2first = ((zpar - l) >> 1) + DG2_ZT * zq;   // arithmetic shift: (0-2)>>1 == -1
                    const int tap = (i * 3 + j) * 3 + l;
#pragma unroll
                    for (int oq = 0; oq < 4; ++oq) {
                        float4 dv[DG2_ZT];
#pragma unroll
                        for (int s = 0; s < DG2_ZT; ++s) {
                            const int z2 = z2first + s;
                            const bool ok = z2 >= 0 && z2 < G2 && (zfirst + 2 * s) < G1;
                            dv[s] = ok ? __ldg(reinterpret_cast<const float4*>(dy2cl + row2 + (int64_t)z2 * C1) + oq)
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
#pragma unroll
                        for (int oo = 0; oo < 4; ++oo) {
                            const float4* wr = reinterpret_cast<const float4*>(ws[tap][4 * oq + oo]);
#pragma unroll
                            for (int cq = 0; cq < 4; ++cq) {
                                const float4 wv = wr[cq];
#pragma unroll
                                for (int s = 0; s < DG2_ZT; ++s) {
                                    const float d = oo == 0 ? dv[s].x : (oo == 1 ? dv[s].y : (oo == 2 ? dv[s].z : dv[s].w));
                                    acc[s][4 * cq + 0] = fmaf(d, wv.x, acc[s][4 * cq + 0]);
                                    acc[s][4 * cq + 1] = fmaf(d, wv.y, acc[s][4 * cq + 1]);
                                    acc[s][4 * cq + 2] = fmaf(d, wv.z, acc[s][4 * cq + 2]);
                                    acc[s][4 * cq + 3] = fmaf(d, wv.w, acc[s][4 * cq + 3]);
                                }
                            }
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int s = 0; s < DG2_ZT; ++s) {
            const int zi = zfirst + 2 * s;
            if (zi >= G1) continue;
            const int64_t p = ((int64_t)xi * G1 + yi) * G1 + zi;
            const float4* yp = reinterpret_cast<const float4*>(y1 + ((int64_t)b * P1 + p) * C1);
            float4* gp = reinterpret_cast<float4*>(g1 + ((int64_t)b * P1 + p) * C1);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 yv = __ldg(yp + q);
                const float y4[4] = {yv.x, yv.y, yv.z, yv.w};
                float o4[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int c = 4 * q + e;
                    const float pre = fmaf(stat1[2 * C1 + c], y4[e], stat1[3 * C1 + c]);
                    const float g = pre > 0.f ? acc[s][c] : 0.f;
                    o4[e] = g;
                    s1[c] += g;
                    s2[c] = fmaf(g, (y4[e] - stat1[c]) * stat1[C1 + c], s2[c]);
                }
                gp[q] = make_float4(o4[0], o4[1], o4[2], o4[3]);
            }
        }
    }
#pragma unroll
    for (int c = 0; c < C1; ++c) {
        float a = warp_sum(s1[c]), q = warp_sum(s2[c]);
        if ((tid & 31) == 0) { red[tid >> 5][c] = a; red[tid >> 5][C1 + c] = q; }
    }
    __syncthreads();
    if (tid < 2 * C1) {
        float t = 0.f;
#pragma unroll
        for (int wv = 0; wv < DG2_THREADS / 32; ++wv) t += red[wv][tid];
        bpart[(int64_t)vb * 2 * C1 + tid] = t;
    }
    __syncthreads();                                          // red[] is reused by the next virtual block
    }
}

// conv1 weight gradient with the BN1 backward applied on the fly:
//   dy1 = a1 * (g1 - S1/n - xhat1 * S2/n);  dW1[co,tap] = sum dy1[b,p,co] * tri[b, 2p+tap];  db1[co] = sum dy1
// Groups of 4 threads share a stream of z-pairs (z1 = 2m, 2m+1); thread cg owns co = 4cg..4cg+3 x all 27 taps
// (108 accumulators).  The two voxels of a pair read the same 5-float input rows (one LDG.128 + one LDG.32 per row when
// the grid row is 16 B aligned): 216 FMA per 22 loads.
constexpr int WG1_THREADS = 256;
constexpr int WG1_MAX_BLOCKS = 1184;
constexpr int WG1_REC = C1 * TAPS + C1;
__global__ void __launch_bounds__(WG1_THREADS)
conv1_wgrad_kernel(const float* __restrict__ obs, int64_t obs_stride, const int64_t* __restrict__ rows, int64_t grid_off,
                   const float* __restrict__ g1, const float* __restrict__ y1, const float* __restrict__ stat1,
                   const float* __restrict__ coef, float* __restrict__ part, int G, int G1, int64_t total_items,
                   int items_per_stream, int vec_ok) {
    __shared__ float red[WG1_THREADS / 32][WG1_REC];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int grp = tid >> 2, cg = tid & 3;
    const int P1 = G1 * G1 * G1, M = (G1 + 1) / 2, items_per_sample = G1 * G1 * M;
    float acc[TAPS][4];
#pragma unroll
    for (int t = 0; t < TAPS; ++t) { acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f; }
    float dbs[4] = {0.f, 0.f, 0.f, 0.f};
    float a1[4], mean[4], invstd[4], k1[4], k2[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int c = 4 * cg + q;
        mean[q] = stat1[c]; invstd[q] = stat1[C1 + c]; a1[q] = stat1[2 * C1 + c]; k1[q] = coef[c]; k2[q] = coef[C1 + c];
    }
    const int64_t it0 = ((int64_t)blockIdx.x * (WG1_THREADS / 4) + grp) * items_per_stream;
    const int64_t it1 = min(total_items, it0 + items_per_stream);
    for (int64_t it = it0; it < it1; ++it) {
        const int b = (int)(it / items_per_sample);
        int t = (int)(it - (int64_t)b * items_per_sample);
        const int m = t % M; t /= M;
        const int yy = t % G1, xx = t / G1;
        const int z1 = 2 * m;
        const bool two = z1 + 1 < G1;
        const int64_t gp = (int64_t)b * P1 + ((int64_t)xx * G1 + yy) * G1 + z1;
        float dy[2][4];
#pragma unroll
        for (int v = 0; v < 2; ++v) {
            if (v == 0 || two) {
                const float4 gv = __ldg(reinterpret_cast<const float4*>(g1 + (gp + v) * C1) + cg);
                const float4 yv = __ldg(reinterpret_cast<const float4*>(y1 + (gp + v) * C1) + cg);
                const float g4[4] = {gv.x, gv.y, gv.z, gv.w}, y4[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    dy[v][q] = a1[q] * (g4[q] - k1[q] - ((y4[q] - mean[q]) * invstd[q]) * k2[q]);
                    dbs[q] += dy[v][q];
                }
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) dy[v][q] = 0.f;
            }
        }
        const float* in = obs + (rows ? rows[b] : (int64_t)b) * obs_stride + grid_off + ((int64_t)(2 * xx) * G + 2 * yy) * G + 2 * z1;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const float* r = in + ((int64_t)i * G + j) * G;
                float x[5];
                if (vec_ok) {
                    const float4 v4 = __ldg(reinterpret_cast<const float4*>(r));
                    x[0] = v4.x; x[1] = v4.y; x[2] = v4.z; x[3] = v4.w;
                } else {
                    x[0] = __ldg(r); x[1] = __ldg(r + 1); x[2] = __ldg(r + 2); x[3] = two ? __ldg(r + 3) : 0.f;
                }
                x[4] = two ? __ldg(r + 4) : 0.f;
#pragma unroll
                for (int l = 0; l < 3; ++l) {
                    const int tap = (i * 3 + j) * 3 + l;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        acc[tap][q] = fmaf(dy[0][q], x[l], acc[tap][q]);
                        acc[tap][q] = fmaf(dy[1][q], x[l + 2], acc[tap][q]);
                    }
                }
            }
    }
    // reduce the 8 streams of a warp (lanes differing in bits 2,3,4), then the 8 warps in a fixed order
#pragma unroll
    for (int t = 0; t < TAPS; ++t)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float v = acc[t][q];
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            acc[t][q] = v;
        }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float v = dbs[q];
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        dbs[q] = v;
    }
    if (lane < 4) {
#pragma unroll
        for (int t = 0; t < TAPS; ++t)
#pragma unroll
            for (int q = 0; q < 4; ++q) red[wid][(4 * cg + q) * TAPS + t] = acc[t][q];
#pragma unroll
        for (int q = 0; q < 4; ++q) red[wid][C1 * TAPS + 4 * cg + q] = dbs[q];
    }
    __syncthreads();
    for (int j = tid; j < WG1_REC; j += WG1_THREADS) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < WG1_THREADS / 32; ++w) t += red[w][j];
        part[(int64_t)blockIdx.x * WG1_REC + j] = t;
    }
}

// TMA-staged variant of conv1_wgrad_kernel (used when grid rows are 16 B aligned, i.e. G % 4 == 0).
// Work unit = a block of up to 16 consecutive output rows (b, x1, y1 in [y0, y0+nr)): its input is three contiguous
// slabs of the tri-class grid ((2 nr + 1) lines of G floats from planes 2x1, 2x1+1, 2x1+2) plus the contiguous g1 / y1
// rows -- five bulk copies per stage, double-buffered against the FMA loop.  Thread mapping and accumulation order per
// thread are those of conv1_wgrad_kernel (groups of 4 threads, z-pairs, 108 accumulators).
constexpr int WG1_RB = 16;
__global__ void __launch_bounds__(WG1_THREADS)
conv1_wgrad_tma_kernel(const float* __restrict__ obs, int64_t obs_stride, const int64_t* __restrict__ rows, int64_t grid_off,
                       const float* __restrict__ g1, const float* __restrict__ y1, const float* __restrict__ stat1,
                       const float* __restrict__ coef, float* __restrict__ part, int G, int G1, int total_rb, int rb_per_block) {
    extern __shared__ __align__(128) float dsm[];
    __shared__ float red[WG1_THREADS / 32][WG1_REC];
    __shared__ __align__(8) uint64_t mbar[2];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int grp = tid >> 2, cg = tid & 3;
    const int P1 = G1 * G1 * G1, M = (G1 + 1) / 2, NYB = (G1 + WG1_RB - 1) / WG1_RB;
    const int TL = (2 * WG1_RB + 1) * G;                    // floats per staged tri slab
    const int GL = WG1_RB * G1 * C1;                        // floats per staged g1 / y1 slab
    const int STAGE = 3 * TL + 2 * GL;
    float acc[TAPS][4];
#pragma unroll
    for (int t = 0; t < TAPS; ++t) { acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f; }
    float dbs[4] = {0.f, 0.f, 0.f, 0.f};
    float a1[4], mean[4], invstd[4], k1[4], k2[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int c = 4 * cg + q;
        mean[q] = stat1[c]; invstd[q] = stat1[C1 + c]; a1[q] = stat1[2 * C1 + c]; k1[q] = coef[c]; k2[q] = coef[C1 + c];
    }
    const int rb0 = blockIdx.x * rb_per_block, rb1 = min(total_rb, rb0 + rb_per_block);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&mbar[0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&mbar[1])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto decode = [&](int rb, int& b, int& x1, int& y0, int& nr) {
        b = rb / (G1 * NYB);
        const int rem = rb - b * G1 * NYB;
        x1 = rem / NYB;
        y0 = (rem - x1 * NYB) * WG1_RB;
        nr = min(WG1_RB, G1 - y0);
    };
    auto issue = [&](int rb, int stage) {                   // thread 0 only
        int b, x1, y0, nr;
        decode(rb, b, x1, y0, nr);
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&mbar[stage]);
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(dsm + stage * STAGE);
        const uint32_t tri_bytes = (uint32_t)(2 * nr + 1) * G * 4, g_bytes = (uint32_t)nr * G1 * C1 * 4;
        mbar_expect_tx(bar, 3 * tri_bytes + 2 * g_bytes);
        const float* orow = obs + (rows ? rows[b] : (int64_t)b) * obs_stride + grid_off;
#pragma unroll
        for (int i = 0; i < 3; ++i)
            bulk_g2s(dst + (uint32_t)(i * TL * 4), orow + ((int64_t)(2 * x1 + i) * G + 2 * y0) * G, tri_bytes, bar);
        const int64_t goff = ((int64_t)b * P1 + ((int64_t)x1 * G1 + y0) * G1) * C1;
        bulk_g2s(dst + (uint32_t)(3 * TL * 4), g1 + goff, g_bytes, bar);
        bulk_g2s(dst + (uint32_t)((3 * TL + GL) * 4), y1 + goff, g_bytes, bar);
    };
    if (tid == 0 && rb0 < rb1) issue(rb0, 0);
    uint32_t phase[2] = {0, 0};
    bool ok = true;
    for (int rb = rb0; rb < rb1; ++rb) {
        const int stage = (rb - rb0) & 1;
        if (tid == 0 && rb + 1 < rb1) issue(rb + 1, stage ^ 1);
        ok = mbar_wait_parity((uint32_t)__cvta_generic_to_shared(&mbar[stage]), phase[stage]) && ok;
        phase[stage] ^= 1;
        int b, x1, y0, nr;
        decode(rb, b, x1, y0, nr);
        const float* ts = dsm + stage * STAGE;
        const float* gs = ts + 3 * TL;
        const float* ys = gs + GL;
        for (int item = grp; item < nr * M; item += WG1_THREADS / 4) {
            const int r = item / M, m = item - r * M, z1 = 2 * m;
            const bool two = z1 + 1 < G1;
            float dy[2][4];
#pragma unroll
            for (int v = 0; v < 2; ++v) {
                if (v == 0 || two) {
                    const float4 gv = *reinterpret_cast<const float4*>(gs + ((r * G1 + z1 + v) * C1) + 4 * cg);
                    const float4 yv = *reinterpret_cast<const float4*>(ys + ((r * G1 + z1 + v) * C1) + 4 * cg);
                    const float g4[4] = {gv.x, gv.y, gv.z, gv.w}, y4[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        dy[v][q] = a1[q] * (g4[q] - k1[q] - ((y4[q] - mean[q]) * invstd[q]) * k2[q]);
                        dbs[q] += dy[v][q];
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 4; ++q) dy[v][q] = 0.f;
                }
            }
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const float* rp = ts + i * TL + (2 * r + j) * G + 2 * z1;
                    const float4 v4 = *reinterpret_cast<const float4*>(rp);
                    const float x4 = two ? rp[4] : 0.f;
                    const float x[5] = {v4.x, v4.y, v4.z, v4.w, x4};
#pragma unroll
                    for (int l = 0; l < 3; ++l) {
                        const int tap = (i * 3 + j) * 3 + l;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            acc[tap][q] = fmaf(dy[0][q], x[l], acc[tap][q]);
                            acc[tap][q] = fmaf(dy[1][q], x[l + 2], acc[tap][q]);
                        }
                    }
                }
        }
        __syncthreads();
    }
    if (!ok) { asm volatile("trap;"); }
#pragma unroll
    for (int t = 0; t < TAPS; ++t)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float v = acc[t][q];
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            acc[t][q] = v;
        }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float v = dbs[q];
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        dbs[q] = v;
    }
    if (lane < 4) {
#pragma unroll
        for (int t = 0; t < TAPS; ++t)
#pragma unroll
            for (int q = 0; q < 4; ++q) red[wid][(4 * cg + q) * TAPS + t] = acc[t][q];
#pragma unroll
        for (int q = 0; q < 4; ++q) red[wid][C1 * TAPS + 4 * cg + q] = dbs[q];
    }
    __syncthreads();
    for (int j = tid; j < WG1_REC; j += WG1_THREADS) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < WG1_THREADS / 32; ++w) t += red[w][j];
        part[(int64_t)blockIdx.x * WG1_REC + j] = t;
    }
}

// TMA-staged conv1 forward (grid rows 16 B aligned).  Same work unit as conv1_wgrad_tma_kernel (row blocks of one x1
// plane); one thread per z-quad: the four voxels share their nine 9-float input rows (2 LDS.128 + LDS.32 each), every
// broadcast LDS.128 of weights feeds 16 FMA, each thread stores 256 contiguous bytes of the channels-last output.
__global__ void __launch_bounds__(CONV1_THREADS)
conv1_fwd_tma_kernel(const float* __restrict__ obs, int64_t obs_stride, const int64_t* __restrict__ rows, int64_t grid_off,
                     const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ y1,
                     float* __restrict__ part, int G, int G1, int RBF, int total_rb, int rb_per_block) {
    extern __shared__ __align__(128) float dsm[];
    __shared__ __align__(16) float wsm[TAPS][C1];
    __shared__ float bs[C1];
    __shared__ float red[CONV1_THREADS / 32][PART_STRIDE];
    __shared__ __align__(8) uint64_t mbar[2];
    const int tid = threadIdx.x;
    for (int i = tid; i < TAPS * C1; i += CONV1_THREADS) {
        int tap = i / C1, c = i - tap * C1;
        wsm[tap][c] = w[c * TAPS + tap];
    }
    if (tid < C1) bs[tid] = bias[tid];
    const int P1 = G1 * G1 * G1, M = (G1 + 3) / 4, NYB = (G1 + RBF - 1) / RBF;      // M = z-quads per row
    const int TL = (2 * RBF + 1) * G;
    const int rb0 = blockIdx.x * rb_per_block, rb1 = min(total_rb, rb0 + rb_per_block);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&mbar[0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&mbar[1])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto decode = [&](int rb, int& b, int& x1, int& y0, int& nr) {
        b = rb / (G1 * NYB);
        const int rem = rb - b * G1 * NYB;
        x1 = rem / NYB;
        y0 = (rem - x1 * NYB) * RBF;
        nr = min(RBF, G1 - y0);
    };
    auto issue = [&](int rb, int stage) {
        int b, x1, y0, nr;
        decode(rb, b, x1, y0, nr);
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&mbar[stage]);
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(dsm + stage * 3 * TL);
        const uint32_t tri_bytes = (uint32_t)(2 * nr + 1) * G * 4;
        mbar_expect_tx(bar, 3 * tri_bytes);
        const float* orow = obs + (rows ? rows[b] : (int64_t)b) * obs_stride + grid_off;
#pragma unroll
        for (int i = 0; i < 3; ++i)
            bulk_g2s(dst + (uint32_t)(i * TL * 4), orow + ((int64_t)(2 * x1 + i) * G + 2 * y0) * G, tri_bytes, bar);
    };
    if (tid == 0 && rb0 < rb1) issue(rb0, 0);
    uint32_t phase[2] = {0, 0};
    bool ok = true;
    for (int rb = rb0; rb < rb1; ++rb) {
        const int stage = (rb - rb0) & 1;
        if (tid == 0 && rb + 1 < rb1) issue(rb + 1, stage ^ 1);
        ok = mbar_wait_parity((uint32_t)__cvta_generic_to_shared(&mbar[stage]), phase[stage]) && ok;
        phase[stage] ^= 1;
        int b, x1, y0, nr;
        decode(rb, b, x1, y0, nr);
        const float* ts = dsm + stage * 3 * TL;
        const bool active = tid < nr * M;
        const int r = tid / M, q4 = tid - r * M, z1 = 4 * q4;
        const int nv = active ? min(4, G1 - z1) : 0;                  // valid voxels of this thread's z-quad
        float acc[4][C1];
#pragma unroll
        for (int sv = 0; sv < 4; ++sv)
#pragma unroll
            for (int c = 0; c < C1; ++c) acc[sv][c] = bs[c];
        if (active) {
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const float* rp = ts + i * TL + (2 * r + j) * G + 2 * z1;
                    const float4 va = *reinterpret_cast<const float4*>(rp);
                    const float4 vb = *reinterpret_cast<const float4*>(rp + 4);
                    const float x[9] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w, rp[8]};
#pragma unroll
                    for (int l = 0; l < 3; ++l) {
                        const float4* wr = reinterpret_cast<const float4*>(wsm[(i * 3 + j) * 3 + l]);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float4 wv = wr[q];
#pragma unroll
                            for (int sv = 0; sv < 4; ++sv) {
                                const float xv = x[l + 2 * sv];
                                acc[sv][4 * q + 0] = fmaf(xv, wv.x, acc[sv][4 * q + 0]);
                                acc[sv][4 * q + 1] = fmaf(xv, wv.y, acc[sv][4 * q + 1]);
                                acc[sv][4 * q + 2] = fmaf(xv, wv.z, acc[sv][4 * q + 2]);
                                acc[sv][4 * q + 3] = fmaf(xv, wv.w, acc[sv][4 * q + 3]);
                            }
                        }
                    }
                }
            float4* o = reinterpret_cast<float4*>(y1 + ((int64_t)b * P1 + ((int64_t)x1 * G1 + y0 + r) * G1 + z1) * C1);
#pragma unroll
            for (int sv = 0; sv < 4; ++sv)
                if (sv < nv) {
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        o[4 * sv + q] = make_float4(acc[sv][4 * q], acc[sv][4 * q + 1], acc[sv][4 * q + 2], acc[sv][4 * q + 3]);
                }
        }
        if (part) block_stats<CONV1_THREADS, 4>(acc, nv, part + (int64_t)rb * PART_STRIDE, red);
        __syncthreads();
    }
    if (!ok) { asm volatile("trap;"); }
}

// conv1 weight gradient on the tensor cores, same staging, contract and per-block record as conv1_wgrad_tma_kernel:
//   dy1 = a1 * (g1 - S1/n - xhat1 * S2/n);  dW1[co,tap] = sum dy1[b,p,co] * tri[b, 2p+tap];  db1[co] = sum dy1
// GEMM with M = 16 (co), N = 27 taps padded to 32 (4 n-tiles), K = output positions: a k-step is 8 consecutive z1 of one
// staged y-row.  A = dy1^T is formed in registers from the staged g1 / y1 rows (BN1 backward on the fly) and split hi/lo;
// B = the tri-class input, exact in TF32 (checked like the forward kernel: a row block whose inputs are not TF32 numbers is
// redone with B split too).  The 8 warps of a block take different k-steps and their 16-register accumulators are
// combined through shared memory at the end, exactly like the CUDA-core kernel's record.  Row blocks are RB = 8 output rows
// (45 KB per stage at 64^3) so that two blocks fit an SM: the kernel is a streaming reduction (1.25 GB in, 7 KB out per
// block) and needs copies in flight more than it needs large tiles.
constexpr int WG1M_RB = 8;
constexpr int C1M_THREADS = 256;
constexpr int C1M_WARPS = C1M_THREADS / 32;

template <bool SPLIT_B>
__device__ __forceinline__ uint32_t conv1_wgrad_mma_rowblock(const float* __restrict__ ts, const float* __restrict__ gs,
                                                              const float* __restrict__ ys, int G, int G1, int TL, int nr,
                                                              const float (&kc)[2][5], const int (&boff)[4],
                                                              float (&acc)[4][4], float (&dbs)[2]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int KS = (G1 + 7) / 8, nks = nr * KS;
    uint32_t orbits = 0;
    for (int ks = warp; ks < nks; ks += C1M_WARPS) {
        const int r = ks / KS, z0 = 8 * (ks - r * KS);
        const int za = z0 + t, zb = za + 4;
        const bool va = za < G1, vb = zb < G1;
        const int zac = min(za, G1 - 1), zbc = min(zb, G1 - 1);
        // A = dy1^T: a0 (co g, pos za), a1 (co g+8, pos za), a2 (co g, pos zb), a3 (co g+8, pos zb)
        float av[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int h = q & 1, zc = (q >> 1) ? zbc : zac;
            const bool valid = (q >> 1) ? vb : va;
            const int idx = (r * G1 + zc) * C1 + g + 8 * h;
            const float dy = kc[h][2] * (gs[idx] - kc[h][3] - ((ys[idx] - kc[h][0]) * kc[h][1]) * kc[h][4]);
            av[q] = valid ? dy : 0.f;
        }
        dbs[0] += av[0] + av[2];
        dbs[1] += av[1] + av[3];
        uint32_t ah[4], al[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) split_tf32(av[q], ah[q], al[q]);
        // B (k = position, n = tap 8j + g): input voxel (2r + jj, 2z + l) of slab i
        const float* pa = ts + (2 * r) * G + 2 * zac;
        const float* pb = ts + (2 * r) * G + 2 * zbc;
        if (SPLIT_B) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float x0 = pa[boff[j]], x1 = pb[boff[j]];
                uint32_t h0, l0, h1, l1;
                split_tf32(x0, h0, l0); split_tf32(x1, h1, l1);
                mma_tf32(acc[j], al[0], al[1], al[2], al[3], h0, h1);
                mma_tf32(acc[j], ah[0], ah[1], ah[2], ah[3], l0, l1);
                mma_tf32(acc[j], ah[0], ah[1], ah[2], ah[3], h0, h1);
            }
        } else {
            // the two products of one accumulator are a dependent chain: issue the four n-tiles interleaved
            uint32_t b0[4], b1[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                b0[j] = __float_as_uint(pa[boff[j]]); b1[j] = __float_as_uint(pb[boff[j]]);
                orbits |= b0[j] | b1[j];
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) mma_tf32(acc[j], al[0], al[1], al[2], al[3], b0[j], b1[j]);
#pragma unroll
            for (int j = 0; j < 4; ++j) mma_tf32(acc[j], ah[0], ah[1], ah[2], ah[3], b0[j], b1[j]);
        }
    }
    return orbits & 0x1fffu;
}

// 8 consumer warps + a producer warp (lane 0 issues the five bulk copies of a row block and alone waits for a stage to drain);
// stages are handed over by mbarriers in both directions, so the loop has no block-wide barrier -- the 16-warp-barrier form spent
// 21 % of its warp samples in stall_barrier (profiles/r02n_*).  The TF32-exactness fallback is decided per warp: a warp's
// accumulators depend only on the k-steps it multiplied itself.
constexpr int WG1M_THREADS = C1M_THREADS + 32;
template <int RB>
__global__ void __launch_bounds__(WG1M_THREADS, 2)
conv1_wgrad_mma_kernel(const float* __restrict__ obs, int64_t obs_stride, const int64_t* __restrict__ rows, int64_t grid_off,
                       const float* __restrict__ g1, const float* __restrict__ y1, const float* __restrict__ stat1,
                       const float* __restrict__ coef, float* __restrict__ part, int G, int G1, int total_rb, int rb_per_block) {
    extern __shared__ __align__(128) float dsm[];
    __shared__ float red[C1M_WARPS][WG1_REC];
    __shared__ __align__(8) uint64_t mbar[2], mbar_empty[2];
    const int tid = threadIdx.x, lane = tid & 31, wid = __shfl_sync(0xffffffffu, tid >> 5, 0), g = lane >> 2, t = lane & 3;
    const int P1 = G1 * G1 * G1, NYB = (G1 + RB - 1) / RB;
    const int TL = (2 * RB + 1) * G;                    // floats per staged tri slab
    const int GL = RB * G1 * C1;                        // floats per staged g1 / y1 slab
    const int STAGE = 3 * TL + 2 * GL;
    // acc: MMA accumulators of ONE row block; tot: their running sum over the block's row blocks (round-to-nearest FADDs).
    // The tensor core's fp32 accumulation truncates: hundreds of MMAs chained into one register drifted to 5e-4 relative
    // at B = 128 (tests/test_policy_gpu.py, 64^3 x 128); a dozen per chain do not.
    float acc[4][4], tot[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { tot[j][0] = tot[j][1] = tot[j][2] = tot[j][3] = 0.f; }
    float dbs[2] = {0.f, 0.f};
    float kc[2][5];                                         // channels g, g + 8: mean, invstd, a1, S1/n, S2/n
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int c = g + 8 * h;
        kc[h][0] = stat1[c]; kc[h][1] = stat1[C1 + c]; kc[h][2] = stat1[2 * C1 + c]; kc[h][3] = coef[c]; kc[h][4] = coef[C1 + c];
    }
    int boff[4];                                            // slab offset of tap n = 8j + g (padding taps alias tap 26)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int tc = min(8 * j + g, TAPS - 1);
        boff[j] = (tc / 9) * TL + ((tc / 3) % 3) * G + tc % 3;
    }
    const int rb0 = blockIdx.x * rb_per_block, rb1 = min(total_rb, rb0 + rb_per_block);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&mbar[0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&mbar[1])) : "memory");
        mbar_init(&mbar_empty[0], C1M_WARPS);
        mbar_init(&mbar_empty[1], C1M_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto decode = [&](int rb, int& b, int& x1, int& y0, int& nr) {
        b = rb / (G1 * NYB);
        const int rem = rb - b * G1 * NYB;
        x1 = rem / NYB;
        y0 = (rem - x1 * NYB) * RB;
        nr = min(RB, G1 - y0);
    };
    auto issue = [&](int rb, int stage) {                   // thread 0 only
        int b, x1, y0, nr;
        decode(rb, b, x1, y0, nr);
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&mbar[stage]);
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(dsm + stage * STAGE);
        const uint32_t tri_bytes = (uint32_t)(2 * nr + 1) * G * 4, g_bytes = (uint32_t)nr * G1 * C1 * 4;
        mbar_expect_tx(bar, 3 * tri_bytes + 2 * g_bytes);
        const float* orow = obs + (rows ? rows[b] : (int64_t)b) * obs_stride + grid_off;
#pragma unroll
        for (int i = 0; i < 3; ++i)
            bulk_g2s(dst + (uint32_t)(i * TL * 4), orow + ((int64_t)(2 * x1 + i) * G + 2 * y0) * G, tri_bytes, bar);
        const int64_t goff = ((int64_t)b * P1 + ((int64_t)x1 * G1 + y0) * G1) * C1;
        bulk_g2s(dst + (uint32_t)(3 * TL * 4), g1 + goff, g_bytes, bar);
        bulk_g2s(dst + (uint32_t)((3 * TL + GL) * 4), y1 + goff, g_bytes, bar);
    };
    const uint32_t full0 = (uint32_t)__cvta_generic_to_shared(&mbar[0]), empty0 = (uint32_t)__cvta_generic_to_shared(&mbar_empty[0]);
    bool ok = true;
    if (wid == C1M_WARPS) {                                 // ---- producer warp ----
        if (lane == 0)
            for (int rb = rb0; rb < rb1; ++rb) {
                const int j = rb - rb0, stage = j & 1;
                if (j >= 2) ok = mbar_wait_parity(empty0 + 8u * stage, (uint32_t)(((j >> 1) - 1) & 1)) && ok;
                issue(rb, stage);
            }
        __syncwarp();
    } else {
        for (int rb = rb0; rb < rb1; ++rb) {
            const int j = rb - rb0, stage = j & 1;
            ok = mbar_wait_parity(full0 + 8u * stage, (uint32_t)((j >> 1) & 1)) && ok;
            int b, x1, y0, nr;
            decode(rb, b, x1, y0, nr);
            const float* ts = dsm + stage * STAGE;
            const float* gs = ts + 3 * TL;
            const float* ys = gs + GL;
            const float keep_db[2] = {dbs[0], dbs[1]};
#pragma unroll
            for (int j2 = 0; j2 < 4; ++j2) { acc[j2][0] = acc[j2][1] = acc[j2][2] = acc[j2][3] = 0.f; }
            const uint32_t inexact = conv1_wgrad_mma_rowblock<false>(ts, gs, ys, G, G1, TL, nr, kc, boff, acc, dbs);
            if (__any_sync(0xffffffffu, inexact != 0u)) {       // this warp met an input value that is not a TF32 number: redo with B split
#pragma unroll
                for (int j2 = 0; j2 < 4; ++j2) { acc[j2][0] = acc[j2][1] = acc[j2][2] = acc[j2][3] = 0.f; }
                dbs[0] = keep_db[0]; dbs[1] = keep_db[1];
                conv1_wgrad_mma_rowblock<true>(ts, gs, ys, G, G1, TL, nr, kc, boff, acc, dbs);
            }
#pragma unroll
            for (int j2 = 0; j2 < 4; ++j2) { tot[j2][0] += acc[j2][0]; tot[j2][1] += acc[j2][1]; tot[j2][2] += acc[j2][2]; tot[j2][3] += acc[j2][3]; }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty0 + 8u * stage);
        }
    }
    if (!ok) { asm volatile("trap;"); }
    // accumulators: c0 (co g, tap 8j+2t), c1 (co g, tap 8j+2t+1), c2 (co g+8, tap 8j+2t), c3 (co g+8, tap 8j+2t+1)
    if (wid < C1M_WARPS) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int co = g + 8 * (e >> 1), tap = 8 * j + 2 * t + (e & 1);
            if (tap < TAPS) red[wid][co * TAPS + tap] = tot[j][e];
        }
    {
        float d0 = dbs[0], d1 = dbs[1];
        d0 += __shfl_xor_sync(0xffffffffu, d0, 1); d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
        d1 += __shfl_xor_sync(0xffffffffu, d1, 1); d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
        if (t == 0) { red[wid][C1 * TAPS + g] = d0; red[wid][C1 * TAPS + 8 + g] = d1; }
    }
    }
    __syncthreads();
    for (int j = tid; j < WG1_REC; j += WG1M_THREADS) {
        float tsum = 0.f;
#pragma unroll
        for (int w = 0; w < C1M_WARPS; ++w) tsum += red[w][j];
        part[(int64_t)blockIdx.x * WG1_REC + j] = tsum;
    }
}

// conv1 forward on the tensor cores (mma.sync m16n8k8, TF32 operands, fp32 accumulation), same staging and contract as
// conv1_fwd_tma_kernel.  GEMM rows = output voxels (16 consecutive z1 of one y-row per m-tile), K = 27 taps padded to 32
// (4 k-steps), N = 16 channels.  The weights are split hi/lo once into REGISTER-resident B fragments (w = hi + lo, 2^-22
// relative); the input is the tri-class grid, whose values {-1, 0, 1} are exact in TF32, so one A fragment (plain LDS.32
// from the staged slabs, no conversion) feeds two MMAs per k-step and n-tile: 16 LDS + 16 HMMA per 16 voxels instead of
// 432 FMA per voxel.  Exactness of A is CHECKED, not assumed: the low 13 mantissa bits of everything loaded are OR-ed
// together, and a row block that held anything not representable in TF32 is recomputed with the A operand split as
// well (3 MMAs, SPLIT_A = true) -- arbitrary float grids stay at fp32-level accuracy, ternary grids never take that path.
// BN statistics: per-thread shifted sums (shift = bias, the value of every voxel whose neighbourhood is unknown space)
// -> (n, mean, M2) -> Chan merges across lanes, warps and (bn_merge_kernel) blocks.
__device__ __forceinline__ void chan_merge_f(float& n, float& mean, float& M2, float nb, float mb, float M2b) {
    if (nb > 0.f) {
        const float nt = n + nb, delta = mb - mean, f = nb / nt;
        mean = fmaf(delta, f, mean);
        M2 += M2b + delta * delta * n * f;
        n = nt;
    }
}

template <bool SPLIT_A>
__device__ __forceinline__ uint32_t conv1_mma_rowblock(const float* __restrict__ ts, int G, int G1, int nr, int64_t out_base,
                                                        float* __restrict__ y1, const uint32_t (&bh)[4][2][2],
                                                        const uint32_t (&bl)[4][2][2], const int (&toff)[4][2],
                                                        const float (&bias_r)[2][2], float (&S1)[2][2], float (&S2)[2][2],
                                                        float& cnt) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int ZT = (G1 + 15) / 16, ntiles = nr * ZT;
    uint32_t orbits = 0;
    for (int tile = warp; tile < ntiles; tile += C1M_WARPS) {
        const int r = tile / ZT, z0 = 16 * (tile - r * ZT);
        // corner of the receptive field of rows g / g + 8 in slab i = 0 (rows past G1 alias the last voxel; never stored)
        const float* pa = ts + (2 * r) * G + 2 * min(z0 + g, G1 - 1);
        const float* pb = ts + (2 * r) * G + 2 * min(z0 + g + 8, G1 - 1);
        float acc[2][4];
#pragma unroll
        for (int j = 0; j < 2; ++j) { acc[j][0] = acc[j][2] = bias_r[j][0]; acc[j][1] = acc[j][3] = bias_r[j][1]; }
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const uint32_t a0 = __float_as_uint(pa[toff[s][0]]), a1 = __float_as_uint(pb[toff[s][0]]);
            const uint32_t a2 = __float_as_uint(pa[toff[s][1]]), a3 = __float_as_uint(pb[toff[s][1]]);
            if (SPLIT_A) {
                uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
                split_tf32(__uint_as_float(a0), h0, l0); split_tf32(__uint_as_float(a1), h1, l1);
                split_tf32(__uint_as_float(a2), h2, l2); split_tf32(__uint_as_float(a3), h3, l3);
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    mma_tf32(acc[j], l0, l1, l2, l3, bh[s][j][0], bh[s][j][1]);
                    mma_tf32(acc[j], h0, h1, h2, h3, bl[s][j][0], bl[s][j][1]);
                    mma_tf32(acc[j], h0, h1, h2, h3, bh[s][j][0], bh[s][j][1]);
                }
            } else {
                orbits |= a0 | a1 | a2 | a3;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    mma_tf32(acc[j], a0, a1, a2, a3, bl[s][j][0], bl[s][j][1]);
                    mma_tf32(acc[j], a0, a1, a2, a3, bh[s][j][0], bh[s][j][1]);
                }
            }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int z1 = z0 + g + 8 * h;
            if (z1 < G1) {
                float* o = y1 + (out_base + (int64_t)r * G1 + z1) * C1 + 2 * t;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const float v0 = acc[j][2 * h], v1 = acc[j][2 * h + 1];
                    *reinterpret_cast<float2*>(o + 8 * j) = make_float2(v0, v1);
                    const float d0 = v0 - bias_r[j][0], d1 = v1 - bias_r[j][1];
                    S1[j][0] += d0; S2[j][0] = fmaf(d0, d0, S2[j][0]);
                    S1[j][1] += d1; S2[j][1] = fmaf(d1, d1, S2[j][1]);
                }
                cnt += 1.f;
            }
        }
    }
    return orbits & 0x1fffu;
}

// 8 consumer warps + a producer warp; stages handed over by mbarriers in both directions (no block barrier per row block), the
// exactness fallback decided per warp, and ONE BatchNorm partial record per block (shifted sums kept in registers across the
// block's row blocks) instead of a block-wide merge per row block.
constexpr int C1F_THREADS = C1M_THREADS + 32;
__global__ void __launch_bounds__(C1F_THREADS, 2)
conv1_fwd_mma_kernel(const float* __restrict__ obs, int64_t obs_stride, const int64_t* __restrict__ rows, int64_t grid_off,
                     const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ y1,
                     float* __restrict__ part, int G, int G1, int RBF, int total_rb, int rb_per_block) {
    extern __shared__ __align__(128) float dsm[];
    __shared__ float red[C1M_WARPS][3 * C1];
    __shared__ __align__(8) uint64_t mbar[2], mbar_empty[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), g = lane >> 2, t = lane & 3;
    const int P1 = G1 * G1 * G1, NYB = (G1 + RBF - 1) / RBF;
    const int TL = (2 * RBF + 1) * G;
    // B fragments: k slot q of k-step s <-> tap 8s + q (taps 27..31 are zero padding); n = g <-> channel 8j + g
    uint32_t bh[4][2][2], bl[4][2][2];
    int toff[4][2];
#pragma unroll
    for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int tap = 8 * s + t + 4 * q, tc = min(tap, TAPS - 1);
            toff[s][q] = (tc / 9) * TL + ((tc / 3) % 3) * G + tc % 3;          // slab i, row offset j, voxel offset l
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const float wv = tap < TAPS ? w[(8 * j + g) * TAPS + tap] : 0.f;
                bh[s][j][q] = to_tf32(wv);
                bl[s][j][q] = to_tf32(wv - __uint_as_float(bh[s][j][q]));
            }
        }
    float bias_r[2][2];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) bias_r[j][e] = bias[8 * j + 2 * t + e];
    const int rb0 = blockIdx.x * rb_per_block, rb1 = min(total_rb, rb0 + rb_per_block);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&mbar[0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&mbar[1])) : "memory");
        mbar_init(&mbar_empty[0], C1M_WARPS);
        mbar_init(&mbar_empty[1], C1M_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto decode = [&](int rb, int& b, int& x1, int& y0, int& nr) {
        b = rb / (G1 * NYB);
        const int rem = rb - b * G1 * NYB;
        x1 = rem / NYB;
        y0 = (rem - x1 * NYB) * RBF;
        nr = min(RBF, G1 - y0);
    };
    auto issue = [&](int rb, int stage) {
        int b, x1, y0, nr;
        decode(rb, b, x1, y0, nr);
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&mbar[stage]);
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(dsm + stage * 3 * TL);
        const uint32_t tri_bytes = (uint32_t)(2 * nr + 1) * G * 4;
        mbar_expect_tx(bar, 3 * tri_bytes);
        const float* orow = obs + (rows ? rows[b] : (int64_t)b) * obs_stride + grid_off;
#pragma unroll
        for (int i = 0; i < 3; ++i)
            bulk_g2s(dst + (uint32_t)(i * TL * 4), orow + ((int64_t)(2 * x1 + i) * G + 2 * y0) * G, tri_bytes, bar);
    };
    const uint32_t full0 = (uint32_t)__cvta_generic_to_shared(&mbar[0]), empty0 = (uint32_t)__cvta_generic_to_shared(&mbar_empty[0]);
    bool ok = true;
    // shifted sums (shift = bias) of this thread's outputs over ALL the block's row blocks: <= a few hundred values per slot
    float S1[2][2] = {{0.f, 0.f}, {0.f, 0.f}}, S2[2][2] = {{0.f, 0.f}, {0.f, 0.f}}, cnt = 0.f;
    if (warp == C1M_WARPS) {                                         // ---- producer warp ----
        if (lane == 0)
            for (int rb = rb0; rb < rb1; ++rb) {
                const int j = rb - rb0, stage = j & 1;
                if (j >= 2) ok = mbar_wait_parity(empty0 + 8u * stage, (uint32_t)(((j >> 1) - 1) & 1)) && ok;
                issue(rb, stage);
            }
        __syncwarp();
    } else {
        for (int rb = rb0; rb < rb1; ++rb) {
            const int j = rb - rb0, stage = j & 1;
            ok = mbar_wait_parity(full0 + 8u * stage, (uint32_t)((j >> 1) & 1)) && ok;
            int b, x1, y0, nr;
            decode(rb, b, x1, y0, nr);
            const float* ts = dsm + stage * 3 * TL;
            const int64_t out_base = (int64_t)b * P1 + ((int64_t)x1 * G1 + y0) * G1;
            float K1[2][2], K2[2][2];
            const float kcnt = cnt;
#pragma unroll
            for (int jj = 0; jj < 2; ++jj)
#pragma unroll
                for (int e = 0; e < 2; ++e) { K1[jj][e] = S1[jj][e]; K2[jj][e] = S2[jj][e]; }
            const uint32_t inexact = conv1_mma_rowblock<false>(ts, G, G1, nr, out_base, y1, bh, bl, toff, bias_r, S1, S2, cnt);
            if (__any_sync(0xffffffffu, inexact != 0u)) {            // this warp met an input value that is not a TF32 number
#pragma unroll
                for (int jj = 0; jj < 2; ++jj)
#pragma unroll
                    for (int e = 0; e < 2; ++e) { S1[jj][e] = K1[jj][e]; S2[jj][e] = K2[jj][e]; }
                cnt = kcnt;
                conv1_mma_rowblock<true>(ts, G, G1, nr, out_base, y1, bh, bl, toff, bias_r, S1, S2, cnt);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty0 + 8u * stage);
        }
    }
    if (part) {                                                      // uniform: one (mean, M2, n) record per block
        if (warp < C1M_WARPS) {
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    float n = cnt, mean = bias_r[j][e], M2 = 0.f;
                    if (cnt > 0.f) {
                        const float m = S1[j][e] / cnt;
                        mean += m;
                        M2 = fmaxf(S2[j][e] - S1[j][e] * m, 0.f);
                    }
#pragma unroll
                    for (int o = 4; o < 32; o <<= 1) {               // lanes that share t hold the same channels
                        const float nb = __shfl_xor_sync(0xffffffffu, n, o), mb = __shfl_xor_sync(0xffffffffu, mean, o),
                                    qb = __shfl_xor_sync(0xffffffffu, M2, o);
                        chan_merge_f(n, mean, M2, nb, mb, qb);
                    }
                    if (g == 0) {
                        const int c = 8 * j + 2 * t + e;
                        red[warp][c] = mean; red[warp][C1 + c] = M2; red[warp][2 * C1 + c] = n;
                    }
                }
        }
        __syncthreads();
        if (tid < C1) {
            float n = 0.f, mean = 0.f, M2 = 0.f;
#pragma unroll
            for (int wv = 0; wv < C1M_WARPS; ++wv) chan_merge_f(n, mean, M2, red[wv][2 * C1 + tid], red[wv][tid], red[wv][C1 + tid]);
            float* pr = part + (int64_t)blockIdx.x * PART_STRIDE;
            pr[tid] = mean; pr[C1 + tid] = M2;
            if (tid == 0) pr[2 * C1] = n;
        }
    }
    if (!ok) { asm volatile("trap;"); }
}

// ---- two-level reductions of the per-block statistics (keeps the final single-block kernels short) -----------------
constexpr int MERGE_FAN = 64;
// forward BN partials (mean, M2, count) -> one record per MERGE_FAN input records (same layout)
__global__ void __launch_bounds__(32 * C1)
bn_merge_kernel(const float* __restrict__ part, int nblk, float* __restrict__ out) {
    const int c = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k0 = blockIdx.x * MERGE_FAN, k1 = min(nblk, k0 + MERGE_FAN);
    double n = 0, mean = 0, M2 = 0;
    for (int k = k0 + lane; k < k1; k += 32) {
        const float* pr = part + (int64_t)k * PART_STRIDE;
        chan_merge(n, mean, M2, (double)pr[2 * C1], (double)pr[c], (double)pr[C1 + c]);
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        double nb = __shfl_xor_sync(0xffffffffu, n, o), mb = __shfl_xor_sync(0xffffffffu, mean, o),
               M2b = __shfl_xor_sync(0xffffffffu, M2, o);
        double n_lo = (lane & o) ? nb : n, m_lo = (lane & o) ? mb : mean, q_lo = (lane & o) ? M2b : M2;
        double n_hi = (lane & o) ? n : nb, m_hi = (lane & o) ? mean : mb, q_hi = (lane & o) ? M2 : M2b;
        chan_merge(n_lo, m_lo, q_lo, n_hi, m_hi, q_hi);
        n = n_lo; mean = m_lo; M2 = q_lo;
    }
    if (lane == 0) {
        float* o = out + (int64_t)blockIdx.x * PART_STRIDE;
        o[c] = (float)mean; o[C1 + c] = (float)M2;
        if (c == 0) o[2 * C1] = (float)n;
    }
}

// backward BN partial sums, layout B (S1[16], S2[16]) -> one record per MERGE_FAN input records
__global__ void __launch_bounds__(32 * C1)
sum_merge_kernel(const float* __restrict__ rec, int nrec, float* __restrict__ out) {
    const int c = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k0 = blockIdx.x * MERGE_FAN, k1 = min(nrec, k0 + MERGE_FAN);
    double s1 = 0, s2 = 0;
    for (int k = k0 + lane; k < k1; k += 32) { s1 += rec[(int64_t)k * 2 * C1 + c]; s2 += rec[(int64_t)k * 2 * C1 + C1 + c]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    if (lane == 0) { out[(int64_t)blockIdx.x * 2 * C1 + c] = (float)s1; out[(int64_t)blockIdx.x * 2 * C1 + C1 + c] = (float)s2; }
}

}  // namespace gnbv

using namespace gnbv;

// =====================================================================================================================
// C ABI
// =====================================================================================================================
namespace {

struct EncDims {
    int B, G, G1, G2, P1, P2, S, FEAT, HID;
    int nblk1, nblk2, items2, rbf1, nrb1, nrec1, nrec2;
    int64_t flat2;
};

EncDims make_dims(int B, int G, int state_dim) {
    EncDims d;
    d.B = B; d.G = G;
    d.G1 = (G - 3) / 2 + 1;
    d.G2 = (d.G1 - 3) / 2 + 1;
    d.P1 = d.G1 * d.G1 * d.G1; d.P2 = d.G2 * d.G2 * d.G2;
    d.S = state_dim; d.FEAT = 256; d.HID = 256;
    d.rbf1 = std::max(1, std::min(d.G1, CONV1_THREADS / ((d.G1 + 3) / 4)));
    d.nrb1 = d.G1 * (int)ceil_div(d.G1, d.rbf1);                    // row blocks per sample of the TMA-staged conv1 kernels
    d.nblk1 = (int)ceil_div(d.P1, CONV1_THREADS);
    d.nrec1 = std::max(d.nblk1, d.nrb1);
    d.items2 = d.G2 * d.G2 * (int)ceil_div(d.G2, CONV2_ZT);
    d.nblk2 = (int)ceil_div(d.items2, CONV2_THREADS);
    d.flat2 = (int64_t)C1 * d.P2;
    d.nrec2 = std::max(std::max(d.nblk2, conv2_mma_chunks(d.G2)), conv2_tc_supported(d.G1, d.G2) ? conv2_tc_tiles(1, d.G2) : 0);
    if (conv2_ts_supported(d.G1, d.G2)) d.nrec2 = std::max(d.nrec2, conv2_ts_tiles(1, d.G2));
    return d;
}

// workspace carve-up (floats). Forward activations that backward needs stay here between the two calls.
struct EncWs {
    size_t pe, h1, cat, y1, part1, stat1, y2, part2, stat2, act2, gemm, total;
    size_t s_out1, s_out2, s_feat_pre, s_dflat, s_dy1, s_scratch;       // 2-D semantic branch (used only when its parameters are given)
    // backward-only
    size_t dz, dcat, dh1, dact2, dy2cl, g1, bn2part, coef2, bpart1, coef1, wg2part, wg1part;
    int nblk_dg, nrec_dg, nblk_wg2, nblk_wg1, wg2_pps, wg1_ips;
    int64_t wg1_items;
    size_t merge1, merge2, bmerge1;
};

EncWs make_ws(const EncDims& d, bool backward) {
    EncWs w;
    size_t o = 0;
    auto take = [&](size_t n) { size_t r = o; o += (n + 63) & ~(size_t)63; return r; };
    const size_t B = d.B;
    w.pe = take(B * 4 * d.S);
    w.h1 = take(B * d.HID);
    w.cat = take(B * 3 * d.HID);                               // [action | grid | (semantic)] : row stride 2 or 3 x HID at run time
    w.y1 = take(B * (size_t)d.P1 * C1);
    w.part1 = take(B * (size_t)d.nrec1 * PART_STRIDE);
    w.stat1 = take(4 * C1);
    w.y2 = take(B * (size_t)d.flat2);
    w.part2 = take(B * (size_t)d.nrec2 * PART_STRIDE);
    w.stat2 = take(4 * C1);
    w.merge1 = take((size_t)ceil_div(B * (size_t)d.nrec1, MERGE_FAN) * PART_STRIDE);
    w.merge2 = take((size_t)ceil_div(B * (size_t)d.nrec2, MERGE_FAN) * PART_STRIDE);
    w.act2 = take(B * (size_t)d.flat2);
    w.s_out1 = take(sem2d_out1_floats(d.B));
    w.s_out2 = take(sem2d_out2_floats(d.B));
    size_t g = 0;
    g = std::max(g, gemm_workspace_floats(d.B, d.HID, sem2d_flat()));
    g = std::max(g, gemm_workspace_floats(d.B, d.FEAT, 3 * d.HID));
    g = std::max(g, gemm_workspace_floats(d.B, d.HID, 4 * d.S));
    g = std::max(g, gemm_workspace_floats(d.B, d.HID, d.HID));
    g = std::max(g, gemm_workspace_floats(d.B, d.HID, (int)d.flat2));
    g = std::max(g, gemm_workspace_floats(d.B, d.FEAT, 2 * d.HID));
    w.bmerge1 = 0;
    w.dz = w.dcat = w.dh1 = w.dact2 = w.dy2cl = w.g1 = w.bn2part = w.coef2 = w.bpart1 = w.coef1 = w.wg2part = w.wg1part = 0;
    w.s_dflat = w.s_dy1 = w.s_scratch = 0;
    w.nblk_dg = (int)ceil_div((int64_t)d.G1 * d.G1 * 2 * ceil_div((d.G1 + 1) / 2, DG2_ZT), DG2_THREADS);
    w.nrec_dg = std::max(w.nblk_dg, conv2_dgrad_mma_items_per_sample(d.G1));     // records per sample of either dgrad kernel
    w.wg2_pps = (int)std::max<int64_t>(1, ceil_div((int64_t)d.B * d.G2 * d.G2, (int64_t)WG2_MAX_BLOCKS));   // rows per block
    w.nblk_wg2 = (int)ceil_div((int64_t)d.B * d.G2 * d.G2, (int64_t)w.wg2_pps);
    w.wg1_items = (int64_t)d.B * d.G1 * d.G1 * ((d.G1 + 1) / 2);
    w.wg1_ips = (int)std::max<int64_t>(16, ceil_div(w.wg1_items, (int64_t)WG1_MAX_BLOCKS * (WG1_THREADS / 4)));
    w.nblk_wg1 = (int)ceil_div(w.wg1_items, (int64_t)(WG1_THREADS / 4) * w.wg1_ips);
    if (backward) {
        w.dz = take(B * d.FEAT);
        w.dcat = take(B * 3 * d.HID);
        w.s_dflat = take(sem2d_out2_floats(d.B));
        w.s_dy1 = take(sem2d_out1_floats(d.B));
        w.s_scratch = take(sem2d_scratch_floats(d.B));
        g = std::max(g, gemm_workspace_floats(d.B, 3 * d.HID, d.FEAT));
        g = std::max(g, gemm_workspace_floats(d.FEAT, 3 * d.HID, d.B));
        g = std::max(g, gemm_workspace_floats(d.B, sem2d_flat(), d.HID));
        g = std::max(g, gemm_workspace_floats(d.HID, sem2d_flat(), d.B));
        w.dh1 = take(B * d.HID);
        w.dact2 = take(B * (size_t)d.flat2);
        w.dy2cl = take(B * (size_t)d.flat2);
        w.g1 = take(B * (size_t)d.P1 * C1);
        w.bn2part = take(B * C1 * 2);
        w.coef2 = take(2 * C1);
        w.bpart1 = take(B * (size_t)w.nrec_dg * 2 * C1);
        w.coef1 = take(2 * C1);
        w.bmerge1 = take((size_t)ceil_div(B * (size_t)w.nrec_dg, MERGE_FAN) * 2 * C1);
        w.wg2part = take((size_t)w.nblk_wg2 * WG2_REC);
        w.wg1part = take((size_t)w.nblk_wg1 * WG1_REC);
        g = std::max(g, gemm_workspace_floats(d.B, 2 * d.HID, d.FEAT));
        g = std::max(g, gemm_workspace_floats(d.FEAT, 2 * d.HID, d.B));
        g = std::max(g, gemm_workspace_floats(d.B, (int)d.flat2, d.HID));
        g = std::max(g, gemm_workspace_floats(d.HID, (int)d.flat2, d.B));
        g = std::max(g, gemm_workspace_floats(d.HID, 4 * d.S, d.B));
        g = std::max(g, gemm_workspace_floats(d.HID, d.HID, d.B));
        g = std::max(g, gemm_workspace_floats(d.B, d.HID, d.HID));
    }
    w.gemm = take(g);
    w.total = o;
    return w;
}

}  // namespace

extern "C" size_t gnbv_encoder_workspace_bytes(int batch, int grid_size, int state_dim, int with_backward) {
    if (batch <= 0 || grid_size < 7 || state_dim <= 0) return 0;
    EncDims d = make_dims(batch, grid_size, state_dim);
    return make_ws(d, with_backward != 0).total * 4 + 256;
}

extern "C" int gnbv_encoder_workspace_view(int batch, int grid_size, int state_dim, int which, int64_t* offset_floats,
                                           int64_t* count) {
    GNBV_REQUIRE(offset_floats && count && batch > 0 && grid_size >= 7 && state_dim > 0, "gnbv_encoder_workspace_view: bad arguments");
    EncDims d = make_dims(batch, grid_size, state_dim);
    EncWs w = make_ws(d, false);
    const size_t B = batch;
    switch (which) {
        case GNBV_WS_Y1: *offset_floats = (int64_t)w.y1; *count = (int64_t)(B * d.P1 * C1); break;
        case GNBV_WS_STAT1: *offset_floats = (int64_t)w.stat1; *count = 4 * C1; break;
        case GNBV_WS_Y2: *offset_floats = (int64_t)w.y2; *count = (int64_t)(B * d.flat2); break;
        case GNBV_WS_STAT2: *offset_floats = (int64_t)w.stat2; *count = 4 * C1; break;
        case GNBV_WS_ACT2: *offset_floats = (int64_t)w.act2; *count = (int64_t)(B * d.flat2); break;
        default: GNBV_REQUIRE(false, "gnbv_encoder_workspace_view: unknown view %d", which);
    }
    return GNBV_OK;
}

int gnbv::encoder_forward_impl(const gnbv_encoder_params* p, const float* obs, int64_t obs_row_stride,
                                const int64_t* row_index, int batch, int grid_size, int state_dim, int training,
                                float* features, void* workspace, size_t workspace_bytes, const int64_t* freeze,
                                cudaStream_t stream) {
    GNBV_REQUIRE(p && obs && features && workspace, "gnbv_encoder_forward: null pointer argument");
    GNBV_REQUIRE(batch > 0 && grid_size >= 7 && state_dim > 0 && state_dim % 6 == 0,
                 "gnbv_encoder_forward: bad sizes (batch=%d grid=%d state=%d)", batch, grid_size, state_dim);
    GNBV_REQUIRE(p->conv1_w && p->conv1_b && p->bn1_w && p->bn1_b && p->bn1_rm && p->bn1_rv && p->conv2_w && p->conv2_b &&
                     p->bn2_w && p->bn2_b && p->bn2_rm && p->bn2_rv && p->grid_fc_w && p->grid_fc_b && p->act_fc1_w &&
                     p->act_fc1_b && p->act_fc2_w && p->act_fc2_b && p->out_fc_w && p->out_fc_b,
                 "gnbv_encoder_forward: null parameter pointer");
    EncDims d = make_dims(batch, grid_size, state_dim);
    GNBV_REQUIRE(d.G2 >= 1, "gnbv_encoder_forward: grid too small");
    // the forward-only and forward+backward layouts share every forward offset (backward buffers are appended
    // before the GEMM scratch, which is re-derived per call), so the forward layout is the prefix of both
    EncWs w = make_ws(d, false);
    GNBV_REQUIRE(workspace_bytes >= w.total * 4, "gnbv_encoder_forward: workspace %zu B < %zu B", workspace_bytes, w.total * 4);
    if (workspace_bytes >= make_ws(d, true).total * 4) w = make_ws(d, true);
    GNBV_REQUIRE(((uintptr_t)workspace & 255) == 0, "gnbv_encoder_forward: workspace must be 256 B aligned");
    float* ws = reinterpret_cast<float*>(workspace);
    const int B = batch;
    const bool sem = p->rgb_conv1_w != nullptr;               // 2-D semantic branch: third 256-wide block of `cat`
    GNBV_REQUIRE(!sem || (p->rgb_conv1_b && p->rgb_conv2_w && p->rgb_conv2_b && p->rgb_fc_w && p->rgb_fc_b),
                 "gnbv_encoder_forward: the semantic branch needs all six rgb_* parameters");
    const int CAT = (sem ? 3 : 2) * d.HID;
    GemmEpilogue relu_ep;
    relu_ep.relu = 1;
    int rc;
    stage_mark(GNBV_ST_FWD_ACTION_MLP, stream);
    // ---- action branch: positional encoding -> Linear(4S,256)+ReLU -> Linear(256,256)+ReLU (written into cat[:, :256])
    posenc_kernel<<<(unsigned)ceil_div((int64_t)B * d.S, 256), 256, 0, stream>>>(obs, obs_row_stride, row_index, ws + w.pe, B, d.S);
    GNBV_LAUNCH_CHECK("posenc_kernel");
    relu_ep.bias = p->act_fc1_b;
    rc = launch_gemm(ws + w.pe, 4 * d.S, 1, p->act_fc1_w, 1, 4 * d.S, ws + w.h1, d.HID, B, d.HID, 4 * d.S, relu_ep, ws + w.gemm, stream);
    if (rc) return rc;
    relu_ep.bias = p->act_fc2_b;
    rc = launch_gemm(ws + w.h1, d.HID, 1, p->act_fc2_w, 1, d.HID, ws + w.cat, CAT, B, d.HID, d.HID, relu_ep, ws + w.gemm, stream);
    if (rc) return rc;
    // ---- grid branch
    stage_mark(GNBV_ST_FWD_CONV1, stream);
    float* part1 = training ? ws + w.part1 : nullptr;
    const int vec1f = (d.G % 4 == 0) && (obs_row_stride % 4 == 0) && (state_dim % 4 == 0) && (((uintptr_t)obs & 15) == 0);
    const size_t smem_c1 = ((size_t)2 * 3 * (2 * d.rbf1 + 1) * d.G + 32) * 4;      // + pad: the last quad reads one float past its row
    int nrec1;
    if (vec1f && smem_c1 <= 160 * 1024) {
        const int total_rb = B * d.nrb1;
        const int rbpb = (int)std::max<int64_t>(1, ceil_div(total_rb, 1184));
        if (conv1_mma_mode() & 1) {
            { int rc_ = ensure_dyn_smem(conv1_fwd_mma_kernel, smem_c1); if (rc_) return rc_; }
            conv1_fwd_mma_kernel<<<(unsigned)ceil_div(total_rb, rbpb), C1F_THREADS, smem_c1, stream>>>(
                obs, obs_row_stride, row_index, state_dim, p->conv1_w, p->conv1_b, ws + w.y1, part1, d.G, d.G1, d.rbf1, total_rb, rbpb);
            nrec1 = (int)ceil_div(total_rb, rbpb);                   // one BN record per block
        } else {
            { int rc_ = ensure_dyn_smem(conv1_fwd_tma_kernel, smem_c1); if (rc_) return rc_; }
            conv1_fwd_tma_kernel<<<(unsigned)ceil_div(total_rb, rbpb), CONV1_THREADS, smem_c1, stream>>>(
                obs, obs_row_stride, row_index, state_dim, p->conv1_w, p->conv1_b, ws + w.y1, part1, d.G, d.G1, d.rbf1, total_rb, rbpb);
            nrec1 = total_rb;
        }
    } else {
        conv1_fwd_kernel<<<dim3(d.nblk1, B), CONV1_THREADS, 0, stream>>>(obs, obs_row_stride, row_index, state_dim, p->conv1_w, p->conv1_b,
                                                                          ws + w.y1, part1, d.G, d.G1);
        nrec1 = B * d.nblk1;
    }
    GNBV_LAUNCH_CHECK("conv1_fwd_kernel");
    stage_mark(GNBV_ST_FWD_BN1, stream);
    if (training) {
        const int nm = (int)ceil_div((int64_t)nrec1, MERGE_FAN);
        bn_merge_kernel<<<nm, 32 * C1, 0, stream>>>(part1, nrec1, ws + w.merge1);
        bn_finalize_kernel<<<1, 32 * C1, 0, stream>>>(ws + w.merge1, nm, p->bn1_w, p->bn1_b, p->bn1_rm, p->bn1_rv, p->bn1_nbt,
                                                      ws + w.stat1, 1e-5f, 0.1f, freeze);
    }
    else
        bn_eval_affine_kernel<<<1, 32, 0, stream>>>(p->bn1_w, p->bn1_b, p->bn1_rm, p->bn1_rv, ws + w.stat1, 1e-5f);
    GNBV_LAUNCH_CHECK("bn1 statistics");
    stage_mark(GNBV_ST_FWD_CONV2, stream);
    float* part2 = training ? ws + w.part2 : nullptr;
    // GNBV_CONV2_TC=1 routes conv2 through the tcgen05 tensor cores (3xTF32 implicit GEMM, conv2_tc.cu).  It is parity-green
    // but, with only N = 16 output channels per A element, its im2col + BN + split staging costs about as many
    // instructions as the CUDA-core kernel's FMAs (measured 1.43 ms vs 0.55 ms at B = 256), so it is opt-in this round.
    // GNBV_CONV2_TC bit 2 (value 2) routes it through warp-level mma.sync with the same 3xTF32 split, operands taken from
    // registers (conv2_mma.cu): no im2col staging at all.  Bit 4 does the same for the data gradient (value 6 = both).
    const int tc_mode = conv2_tc_mode();
    const bool use_tc = tc_mode == 1;
    int nrec2;
    if ((tc_mode & 32) && conv2_ts_supported(d.G1, d.G2)) {
        rc = launch_conv2_fwd_ts(ws + w.y1, ws + w.stat1, p->conv2_w, p->conv2_b, ws + w.y2, part2, B, d.G1, d.G2, stream);
        if (rc) return rc;
        nrec2 = conv2_ts_tiles(B, d.G2);
    } else if (tc_mode & 2) {
        rc = launch_conv2_fwd_mma(ws + w.y1, ws + w.stat1, p->conv2_w, p->conv2_b, ws + w.y2, part2, B, d.G1, d.G2, stream);
        if (rc) return rc;
        nrec2 = conv2_mma_items(B, d.G2);
    } else if (use_tc && conv2_tc_supported(d.G1, d.G2)) {
        rc = launch_conv2_fwd_tc(ws + w.y1, ws + w.stat1, p->conv2_w, p->conv2_b, ws + w.y2, part2, nullptr, B, d.G1, d.G2, stream);
        if (rc) return rc;
        nrec2 = conv2_tc_tiles(B, d.G2);
    } else {
        conv2_fwd_kernel<<<dim3(d.nblk2, B), CONV2_THREADS, 0, stream>>>(ws + w.y1, ws + w.stat1, p->conv2_w, p->conv2_b, ws + w.y2,
                                                                          part2, d.G1, d.G2);
        nrec2 = B * d.nblk2;
    }
    GNBV_LAUNCH_CHECK("conv2_fwd_kernel");
    stage_mark(GNBV_ST_FWD_BN2, stream);
    if (training) {
        const int nm = (int)ceil_div((int64_t)nrec2, MERGE_FAN);
        bn_merge_kernel<<<nm, 32 * C1, 0, stream>>>(part2, nrec2, ws + w.merge2);
        bn_finalize_kernel<<<1, 32 * C1, 0, stream>>>(ws + w.merge2, nm, p->bn2_w, p->bn2_b, p->bn2_rm, p->bn2_rv, p->bn2_nbt,
                                                      ws + w.stat2, 1e-5f, 0.1f, freeze);
    }
    else
        bn_eval_affine_kernel<<<1, 32, 0, stream>>>(p->bn2_w, p->bn2_b, p->bn2_rm, p->bn2_rv, ws + w.stat2, 1e-5f);
    GNBV_LAUNCH_CHECK("bn2 statistics");
    const int64_t tot2 = (int64_t)B * d.flat2;
    bn_relu_apply_kernel<<<(unsigned)ceil_div(tot2 / 4, 256), 256, 0, stream>>>(ws + w.y2, ws + w.stat2, ws + w.act2, tot2, d.P2);
    GNBV_LAUNCH_CHECK("bn_relu_apply_kernel");
    // Linear(16*G2^3, 256)+ReLU -> cat[:, 256:512]
    stage_mark(GNBV_ST_FWD_GRID_FC, stream);
    relu_ep.bias = p->grid_fc_b;
    rc = launch_gemm(ws + w.act2, d.flat2, 1, p->grid_fc_w, 1, d.flat2, ws + w.cat + d.HID, CAT, B, d.HID, (int)d.flat2,
                     relu_ep, ws + w.gemm, stream);
    if (rc) return rc;
    if (sem) {
        // frames live behind the grid columns of the observation row (env_wrapper_gennbv_train.py:102-110: state | grid | state_rgb)
        const int64_t rgb_off = (int64_t)state_dim + (int64_t)d.G * d.G * d.G;
        rc = launch_sem2d_forward(obs, obs_row_stride, row_index, rgb_off, p->rgb_conv1_w, p->rgb_conv1_b, p->rgb_conv2_w,
                                  p->rgb_conv2_b, ws + w.s_out1, ws + w.s_out2, B, stream);
        if (rc) return rc;
        relu_ep.bias = p->rgb_fc_b;
        rc = launch_gemm(ws + w.s_out2, sem2d_flat(), 1, p->rgb_fc_w, 1, sem2d_flat(), ws + w.cat + 2 * d.HID, CAT, B, d.HID,
                         sem2d_flat(), relu_ep, ws + w.gemm, stream);
        if (rc) return rc;
    }
    // fuse: Linear(512,256)+ReLU -> features
    stage_mark(GNBV_ST_FWD_OUT_FC, stream);
    relu_ep.bias = p->out_fc_b;
    rc = launch_gemm(ws + w.cat, CAT, 1, p->out_fc_w, 1, CAT, features, d.FEAT, B, d.FEAT, CAT, relu_ep,
                     ws + w.gemm, stream);
    stage_mark(GNBV_ST_FWD_END, stream);
    return rc;
}

extern "C" int gnbv_encoder_forward(const gnbv_encoder_params* p, const float* obs, int64_t obs_row_stride,
                                    const int64_t* row_index, int batch, int grid_size, int state_dim, int training,
                                    float* features, void* workspace, size_t workspace_bytes, void* stream) {
    return encoder_forward_impl(p, obs, obs_row_stride, row_index, batch, grid_size, state_dim, training, features, workspace,
                                workspace_bytes, nullptr, (cudaStream_t)stream);
}

extern "C" int gnbv_encoder_backward(const gnbv_encoder_params* p, const float* obs, int64_t obs_row_stride,
                                     const int64_t* row_index, int batch, int grid_size, int state_dim, int training,
                                     const float* features, const float* dfeatures, const gnbv_encoder_grads* gr,
                                     void* workspace, size_t workspace_bytes, void* stream) {
    return gnbv_encoder_backward_phase(p, obs, obs_row_stride, row_index, batch, grid_size, state_dim, training, features,
                                       dfeatures, gr, workspace, workspace_bytes, GNBV_BWD_ALL, stream);
}

extern "C" int gnbv_encoder_backward_phase(const gnbv_encoder_params* p, const float* obs, int64_t obs_row_stride,
                                           const int64_t* row_index, int batch, int grid_size, int state_dim, int training,
                                           const float* features, const float* dfeatures, const gnbv_encoder_grads* gr,
                                           void* workspace, size_t workspace_bytes, int phases, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GNBV_REQUIRE(phases & GNBV_BWD_ALL, "gnbv_encoder_backward_phase: empty phase mask");
    GNBV_REQUIRE(p && obs && features && dfeatures && gr && workspace, "gnbv_encoder_backward: null pointer argument");
    GNBV_REQUIRE(gr->conv1_w && gr->conv1_b && gr->bn1_w && gr->bn1_b && gr->conv2_w && gr->conv2_b && gr->bn2_w && gr->bn2_b &&
                     gr->grid_fc_w && gr->grid_fc_b && gr->act_fc1_w && gr->act_fc1_b && gr->act_fc2_w && gr->act_fc2_b &&
                     gr->out_fc_w && gr->out_fc_b,
                 "gnbv_encoder_backward: null gradient pointer");
    EncDims d = make_dims(batch, grid_size, state_dim);
    EncWs w = make_ws(d, true);
    GNBV_REQUIRE(workspace_bytes >= w.total * 4, "gnbv_encoder_backward: workspace %zu B < %zu B (allocate with with_backward=1)",
                 workspace_bytes, w.total * 4);
    float* ws = reinterpret_cast<float*>(workspace);
    const int B = batch, H = d.HID;
    const bool sem = p->rgb_conv1_w != nullptr;
    GNBV_REQUIRE(!sem || (gr->rgb_conv1_w && gr->rgb_conv1_b && gr->rgb_conv2_w && gr->rgb_conv2_b && gr->rgb_fc_w && gr->rgb_fc_b),
                 "gnbv_encoder_backward: the semantic branch needs all six rgb_* gradient pointers");
    const int CAT = (sem ? 3 : 2) * H;
    GemmEpilogue none;
    int rc;
    auto blocks = [](int64_t n) { return (unsigned)ceil_div(n, 256); };
    if (phases & GNBV_BWD_LINEAR) {
    stage_mark(GNBV_ST_BWD_LINEAR, stream);
    // ---- fuse layer: features = relu(cat W^T + b)
    relu_mask_kernel<<<blocks((int64_t)B * d.FEAT), 256, 0, stream>>>(dfeatures, d.FEAT, ws + w.dz, d.FEAT, features, d.FEAT, B, d.FEAT);
    GNBV_LAUNCH_CHECK("relu_mask_kernel");
    rc = launch_gemm(ws + w.dz, 1, d.FEAT, ws + w.cat, CAT, 1, gr->out_fc_w, CAT, d.FEAT, CAT, B, none, ws + w.gemm, stream);
    if (rc) return rc;
    colsum_kernel<<<colsum_blocks(d.FEAT), 256, 0, stream>>>(ws + w.dz, d.FEAT, B, d.FEAT, gr->out_fc_b);
    rc = launch_gemm(ws + w.dz, d.FEAT, 1, p->out_fc_w, CAT, 1, ws + w.dcat, CAT, B, CAT, d.FEAT, none, ws + w.gemm, stream);
    if (rc) return rc;
    relu_mask_kernel<<<blocks((int64_t)B * CAT), 256, 0, stream>>>(ws + w.dcat, CAT, ws + w.dcat, CAT, ws + w.cat, CAT, B, CAT);
    // ---- action branch
    rc = launch_gemm(ws + w.dcat, 1, CAT, ws + w.h1, H, 1, gr->act_fc2_w, H, H, H, B, none, ws + w.gemm, stream);
    if (rc) return rc;
    colsum_kernel<<<colsum_blocks(H), 256, 0, stream>>>(ws + w.dcat, CAT, B, H, gr->act_fc2_b);
    rc = launch_gemm(ws + w.dcat, CAT, 1, p->act_fc2_w, H, 1, ws + w.dh1, H, B, H, H, none, ws + w.gemm, stream);
    if (rc) return rc;
    relu_mask_kernel<<<blocks((int64_t)B * H), 256, 0, stream>>>(ws + w.dh1, H, ws + w.dh1, H, ws + w.h1, H, B, H);
    rc = launch_gemm(ws + w.dh1, 1, H, ws + w.pe, 4 * d.S, 1, gr->act_fc1_w, 4 * d.S, H, 4 * d.S, B, none, ws + w.gemm, stream);
    if (rc) return rc;
    colsum_kernel<<<colsum_blocks(H), 256, 0, stream>>>(ws + w.dh1, H, B, H, gr->act_fc1_b);
    // ---- grid branch: Linear
    stage_mark(GNBV_ST_BWD_GRID_FC, stream);
    const float* dcat_g = ws + w.dcat + H;
    rc = launch_gemm(dcat_g, 1, CAT, ws + w.act2, d.flat2, 1, gr->grid_fc_w, d.flat2, H, (int)d.flat2, B, none, ws + w.gemm, stream);
    if (rc) return rc;
    colsum_kernel<<<colsum_blocks(H), 256, 0, stream>>>(dcat_g, CAT, B, H, gr->grid_fc_b);
    rc = launch_gemm(dcat_g, CAT, 1, p->grid_fc_w, d.flat2, 1, ws + w.dact2, d.flat2, B, (int)d.flat2, H, none, ws + w.gemm, stream);
    if (rc) return rc;
    if (sem) {
        // semantic branch: Linear(3600, 256) then the two 2-D convolutions (dcat[:, 2H:3H] already carries the ReLU mask)
        const float* dcat_s = ws + w.dcat + 2 * H;
        const int FS = sem2d_flat();
        rc = launch_gemm(dcat_s, 1, CAT, ws + w.s_out2, FS, 1, gr->rgb_fc_w, FS, H, FS, B, none, ws + w.gemm, stream);
        if (rc) return rc;
        colsum_kernel<<<colsum_blocks(H), 256, 0, stream>>>(dcat_s, CAT, B, H, gr->rgb_fc_b);
        rc = launch_gemm(dcat_s, CAT, 1, p->rgb_fc_w, FS, 1, ws + w.s_dflat, FS, B, FS, H, none, ws + w.gemm, stream);
        if (rc) return rc;
        const int64_t rgb_off = (int64_t)state_dim + (int64_t)d.G * d.G * d.G;
        rc = launch_sem2d_backward(obs, obs_row_stride, row_index, rgb_off, p->rgb_conv2_w, ws + w.s_out1, ws + w.s_out2, ws + w.s_dflat,
                                   ws + w.s_dy1, ws + w.s_scratch, gr->rgb_conv1_w, gr->rgb_conv1_b, gr->rgb_conv2_w, gr->rgb_conv2_b, B,
                                   stream);
        if (rc) return rc;
    }
    GNBV_LAUNCH_CHECK("linear backward");
    }
    if (!(phases & GNBV_BWD_CONV)) return GNBV_OK;
    // ---- BN2 + ReLU backward
    stage_mark(GNBV_ST_BWD_BN2, stream);
    bn2_bwd_reduce_kernel<<<dim3(C1, B), 256, 0, stream>>>(ws + w.dact2, ws + w.act2, ws + w.y2, ws + w.stat2, ws + w.bn2part, d.P2);
    bn_bwd_finalize_kernel<<<1, 32 * C1, 0, stream>>>(ws + w.bn2part, B, 2 * C1, 0, (double)B * d.P2, gr->bn2_w, gr->bn2_b,
                                                      ws + w.coef2, training ? 0 : 1);
    bn2_bwd_apply_kernel<<<dim3((unsigned)ceil_div(d.P2, 256), B), 256, 0, stream>>>(ws + w.dact2, ws + w.act2, ws + w.y2,
                                                                                      ws + w.stat2, ws + w.coef2, ws + w.dy2cl, d.P2);
    GNBV_LAUNCH_CHECK("bn2 backward");
    // ---- conv2 backward
    stage_mark(GNBV_ST_BWD_CONV2_WGRAD, stream);
    const size_t line_f = (size_t)((d.G1 * C1 + 31) / 32) * 32, dyl_f = (size_t)((d.G2 * C1 + 31) / 32) * 32;
    const size_t smem_wg2 = std::max(3 * (size_t)WG2_REC, 2 * (9 * line_f + dyl_f)) * 4;
    GNBV_REQUIRE(smem_wg2 <= 200 * 1024, "gnbv_encoder_backward: grid too large for the conv2 wgrad staging buffers");
    { int rc_ = ensure_dyn_smem(conv2_wgrad_kernel, smem_wg2); if (rc_) return rc_; }
    int nrec_wg2 = w.nblk_wg2;
    if (conv2_tc_mode() & 16) {
        rc = launch_conv2_wgrad_staged(ws + w.y1, ws + w.stat1, ws + w.dy2cl, ws + w.wg2part, B, d.G1, d.G2, w.nblk_wg2, &nrec_wg2, stream);
        if (rc) return rc;
    } else if (conv2_tc_mode() & 8) {
        rc = launch_conv2_wgrad_mma(ws + w.y1, ws + w.stat1, ws + w.dy2cl, ws + w.wg2part, B, d.G1, d.G2, w.nblk_wg2, w.wg2_pps, stream);
        if (rc) return rc;
    } else {
        conv2_wgrad_kernel<<<w.nblk_wg2, WG2_THREADS, smem_wg2, stream>>>(ws + w.y1, ws + w.stat1, ws + w.dy2cl, ws + w.wg2part, d.G1,
                                                                           d.G2, B * d.G2 * d.G2, w.wg2_pps);
    }
    GNBV_LAUNCH_CHECK("conv2_wgrad_kernel");
    reduce_records_kernel<<<colsum_blocks(WG2_REC), 256, 0, stream>>>(ws + w.wg2part, nrec_wg2, WG2_REC, gr->conv2_w, C1 * C1 * TAPS,
                                                                gr->conv2_b);
    stage_mark(GNBV_ST_BWD_CONV2_DGRAD, stream);
    int nrec_dg;
    if ((conv2_tc_mode() & 64) && conv2_ts_supported(d.G1, d.G2)) {
        rc = launch_conv2_dgrad_ts(ws + w.dy2cl, p->conv2_w, ws + w.y1, ws + w.stat1, ws + w.g1, ws + w.bpart1, B, d.G1, d.G2, stream);
        if (rc) return rc;
        nrec_dg = conv2_dgrad_ts_records(B, d.G1);
    } else if (conv2_tc_mode() & 4) {
        rc = launch_conv2_dgrad_mma(ws + w.dy2cl, p->conv2_w, ws + w.y1, ws + w.stat1, ws + w.g1, ws + w.bpart1, B, d.G1, d.G2, stream);
        if (rc) return rc;
        nrec_dg = B * conv2_dgrad_mma_items_per_sample(d.G1);
    } else {
        conv2_dgrad_kernel<<<std::min(B * w.nblk_dg, 148 * 3), DG2_THREADS, 0, stream>>>(ws + w.dy2cl, p->conv2_w, ws + w.y1, ws + w.stat1,
                                                                                          ws + w.g1, ws + w.bpart1, d.G1, d.G2, w.nblk_dg,
                                                                                          B * w.nblk_dg);
        nrec_dg = B * w.nblk_dg;
    }
    GNBV_LAUNCH_CHECK("conv2_dgrad_kernel");
    stage_mark(GNBV_ST_BWD_BN1, stream);
    {
        const int nm = (int)ceil_div((int64_t)nrec_dg, MERGE_FAN);
        sum_merge_kernel<<<nm, 32 * C1, 0, stream>>>(ws + w.bpart1, nrec_dg, ws + w.bmerge1);
        bn_bwd_finalize_kernel<<<1, 32 * C1, 0, stream>>>(ws + w.bmerge1, nm, 2 * C1, 1, (double)B * d.P1, gr->bn1_w, gr->bn1_b,
                                                          ws + w.coef1, training ? 0 : 1);
    }
    // ---- conv1 backward (weights only: the input is data)
    stage_mark(GNBV_ST_BWD_CONV1_WGRAD, stream);
    const int vec1 = (d.G % 4 == 0) && (obs_row_stride % 4 == 0) && (state_dim % 4 == 0) && (((uintptr_t)obs & 15) == 0);
    const size_t smem_wg1 = (size_t)2 * (3 * (2 * WG1_RB + 1) * d.G + 2 * WG1_RB * d.G1 * C1) * 4;
    if (vec1 && smem_wg1 <= 190 * 1024) {
        const int nyb = (int)ceil_div(d.G1, WG1_RB), total_rb = B * d.G1 * nyb;
        const int rbpb = (int)std::max<int64_t>(1, ceil_div(total_rb, w.nblk_wg1));
        int nblk = (int)ceil_div(total_rb, rbpb);                // <= w.nblk_wg1: the partial buffer is large enough
        { int rc_ = ensure_dyn_smem(conv1_wgrad_tma_kernel, smem_wg1); if (rc_) return rc_; }
        if (conv1_mma_mode() & 2) {
            const int nyb_m = (int)ceil_div(d.G1, WG1M_RB), total_m = B * d.G1 * nyb_m;
            const int rbpb_m = (int)std::max<int64_t>(1, ceil_div(total_m, w.nblk_wg1));
            const int nblk_m = (int)ceil_div(total_m, rbpb_m);
            const size_t smem_m = (size_t)2 * (3 * (2 * WG1M_RB + 1) * d.G + 2 * WG1M_RB * d.G1 * C1) * 4;
            { int rc_ = ensure_dyn_smem(conv1_wgrad_mma_kernel<WG1M_RB>, smem_m); if (rc_) return rc_; }
            conv1_wgrad_mma_kernel<WG1M_RB><<<nblk_m, WG1M_THREADS, smem_m, stream>>>(obs, obs_row_stride, row_index, state_dim, ws + w.g1,
                                                                                     ws + w.y1, ws + w.stat1, ws + w.coef1, ws + w.wg1part,
                                                                                     d.G, d.G1, total_m, rbpb_m);
            nblk = nblk_m;
        } else
        conv1_wgrad_tma_kernel<<<nblk, WG1_THREADS, smem_wg1, stream>>>(obs, obs_row_stride, row_index, state_dim, ws + w.g1, ws + w.y1,
                                                                        ws + w.stat1, ws + w.coef1, ws + w.wg1part, d.G, d.G1,
                                                                        total_rb, rbpb);
        GNBV_LAUNCH_CHECK("conv1_wgrad_tma_kernel");
        reduce_records_kernel<<<colsum_blocks(WG1_REC), 256, 0, stream>>>(ws + w.wg1part, nblk, WG1_REC, gr->conv1_w, C1 * TAPS, gr->conv1_b);
        GNBV_LAUNCH_CHECK("reduce_records_kernel");
        stage_mark(GNBV_ST_BWD_END, stream);
        return GNBV_OK;
    }
    conv1_wgrad_kernel<<<w.nblk_wg1, WG1_THREADS, 0, stream>>>(obs, obs_row_stride, row_index, state_dim, ws + w.g1, ws + w.y1, ws + w.stat1,
                                                                ws + w.coef1, ws + w.wg1part, d.G, d.G1, w.wg1_items, w.wg1_ips, vec1);
    GNBV_LAUNCH_CHECK("conv1_wgrad_kernel");
    reduce_records_kernel<<<colsum_blocks(WG1_REC), 256, 0, stream>>>(ws + w.wg1part, w.nblk_wg1, WG1_REC, gr->conv1_w, C1 * TAPS, gr->conv1_b);
    GNBV_LAUNCH_CHECK("reduce_records_kernel");
    stage_mark(GNBV_ST_BWD_END, stream);
    return GNBV_OK;
}
