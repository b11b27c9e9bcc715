// sem2d.cu -- the 2-D semantic branch of the multi-source encoder (GenNBV paper section 3.2: the k most
// recent grayscale frames -> a two-layer convolution -> Linear(Flatten) -> concatenated with the pose and grid embeddings).
// The released reference code carries the frames in the observation (env_train_gennbv.py:363) but its Hybrid_Encoder.forward
// never reads them (hybrid_encoder.py:69-91; the comment `[num_env, 256*3]` at :89 is the only residue), and the paper gives no
// layer sizes.  This build therefore fixes them as
//     Conv2d(k=2, 16, 3, stride 2) + ReLU -> Conv2d(16, 16, 3, stride 2) + ReLU -> Flatten (16 x 15 x 15) -> Linear(3600, 256) + ReLU
// (the 2-D twin of the grid branch, without BatchNorm), OFF by default so that checkpoints keep the reference's keys.
// PARITY UNPINNED: there is no reference forward to compare with; the kernels are checked against torch autograd of the same
// layers (oracle/encoder_ref.py, tests/test_policy_gpu.py).
//
// The branch is 2.7 MFLOP per sample (the grid branch: 100 MFLOP) -- plain fp32 CUDA-core kernels, one thread per output pixel
// with the weights in shared memory; the weight gradients are per-sample partial records reduced in a fixed order in double.
#include "sem2d.cuh"

namespace gnbv {
namespace {

constexpr int K_IN = 2, IMG = 64, C = 16, O1 = 31, O2 = 15, P1 = O1 * O1, P2 = O2 * O2;
constexpr int W1N = C * K_IN * 9, W2N = C * C * 9;                  // 288, 2304

__device__ __forceinline__ const float* frames_of(const float* obs, int64_t stride, const int64_t* rows, int b, int64_t rgb_off) {
    return obs + (rows ? rows[b] : (int64_t)b) * stride + rgb_off;
}

// out1 [B, 31*31, 16] channels-last, post-ReLU
__global__ void __launch_bounds__(128)
sem_conv1_kernel(const float* __restrict__ obs, int64_t stride, const int64_t* __restrict__ rows, int64_t rgb_off,
                 const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ out1) {
    __shared__ float ws[K_IN * 9][C];
    __shared__ float bs[C];
    const int b = blockIdx.y, tid = threadIdx.x;
    for (int i = tid; i < W1N; i += 128) { const int oc = i / (K_IN * 9), t = i - oc * (K_IN * 9); ws[t][oc] = w[i]; }
    if (tid < C) bs[tid] = bias[tid];
    __syncthreads();
    const int p = blockIdx.x * 128 + tid;
    if (p >= P1) return;
    const int oy = p / O1, ox = p - oy * O1;
    const float* in = frames_of(obs, stride, rows, b, rgb_off);
    float acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = bs[c];
    for (int ic = 0; ic < K_IN; ++ic)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const float v = __ldg(in + (ic * IMG + 2 * oy + ky) * IMG + 2 * ox + kx);
                const float* wr = ws[(ic * 3 + ky) * 3 + kx];
#pragma unroll
                for (int c = 0; c < C; ++c) acc[c] = fmaf(v, wr[c], acc[c]);
            }
    float4* o = reinterpret_cast<float4*>(out1 + ((int64_t)b * P1 + p) * C);
#pragma unroll
    for (int q = 0; q < 4; ++q)
        o[q] = make_float4(fmaxf(acc[4 * q], 0.f), fmaxf(acc[4 * q + 1], 0.f), fmaxf(acc[4 * q + 2], 0.f), fmaxf(acc[4 * q + 3], 0.f));
}

// out2 [B, 16*15*15] channel-major (torch Flatten order of [B,16,15,15]), post-ReLU
__global__ void __launch_bounds__(256)
sem_conv2_kernel(const float* __restrict__ out1, const float* __restrict__ w, const float* __restrict__ bias,
                 float* __restrict__ out2) {
    __shared__ float ws[9][C][C];                                    // [tap][ic][oc]
    __shared__ float bs[C];
    const int b = blockIdx.x, tid = threadIdx.x;
    for (int i = tid; i < W2N; i += 256) { const int oc = i / (C * 9), r = i - oc * C * 9, ic = r / 9, t = r - ic * 9; ws[t][ic][oc] = w[i]; }
    if (tid < C) bs[tid] = bias[tid];
    __syncthreads();
    if (tid >= P2) return;
    const int oy = tid / O2, ox = tid - oy * O2;
    float acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = bs[c];
    for (int t = 0; t < 9; ++t) {
        const int ky = t / 3, kx = t - 3 * ky;
        const float4* src = reinterpret_cast<const float4*>(out1 + ((int64_t)b * P1 + (2 * oy + ky) * O1 + 2 * ox + kx) * C);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 v = __ldg(src + q);
            const float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float* wr = ws[t][4 * q + e];
#pragma unroll
                for (int c = 0; c < C; ++c) acc[c] = fmaf(x[e], wr[c], acc[c]);
            }
        }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) out2[(int64_t)b * C * P2 + c * P2 + tid] = fmaxf(acc[c], 0.f);
}

// per sample: dW2 partial [2304] + db2 partial [16];   dy2 = dflat * [out2 > 0]
__global__ void __launch_bounds__(256)
sem_conv2_wgrad_kernel(const float* __restrict__ out1, const float* __restrict__ out2, const float* __restrict__ dflat,
                       float* __restrict__ part /*[B][9][256] then [B][16]*/, int B) {
    __shared__ float dy[C][P2 + 1];
    const int b = blockIdx.x, t = blockIdx.y, tid = threadIdx.x, oc = tid >> 4, ic = tid & 15;
    for (int i = tid; i < C * P2; i += 256) {
        const int c = i / P2, p = i - c * P2;
        const int64_t k = (int64_t)b * C * P2 + i;
        dy[c][p] = out2[k] > 0.f ? dflat[k] : 0.f;
    }
    __syncthreads();
    const int ky = t / 3, kx = t - 3 * ky;
    float acc = 0.f, bsum = 0.f;
    for (int p = 0; p < P2; ++p) {
        const int oy = p / O2, ox = p - oy * O2;
        const float d = dy[oc][p];
        acc = fmaf(d, __ldg(out1 + ((int64_t)b * P1 + (2 * oy + ky) * O1 + 2 * ox + kx) * C + ic), acc);
        bsum += d;
    }
    part[((int64_t)b * 9 + t) * 256 + tid] = acc;                     // [oc][ic] for this tap
    if (t == 0 && ic == 0) part[(int64_t)B * 9 * 256 + (int64_t)b * C + oc] = bsum;
}

// dy1 [B, 961, 16] = (conv2 data gradient) * [out1 > 0]
__global__ void __launch_bounds__(128)
sem_conv2_dgrad_kernel(const float* __restrict__ out1, const float* __restrict__ out2, const float* __restrict__ dflat,
                       const float* __restrict__ w, float* __restrict__ dy1) {
    __shared__ float ws[9][C][C];                                    // [tap][oc][ic]
    const int b = blockIdx.y, tid = threadIdx.x;
    for (int i = tid; i < W2N; i += 128) { const int oc = i / (C * 9), r = i - oc * C * 9, ic = r / 9, t = r - ic * 9; ws[t][oc][ic] = w[i]; }
    __syncthreads();
    const int p = blockIdx.x * 128 + tid;
    if (p >= P1) return;
    const int Y = p / O1, X = p - Y * O1;
    float acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 0.f;
    for (int ky = 0; ky < 3; ++ky) {
        const int yy = Y - ky;
        if (yy < 0 || (yy & 1) || (yy >> 1) >= O2) continue;
        for (int kx = 0; kx < 3; ++kx) {
            const int xx = X - kx;
            if (xx < 0 || (xx & 1) || (xx >> 1) >= O2) continue;
            const int q = (yy >> 1) * O2 + (xx >> 1);
            for (int oc = 0; oc < C; ++oc) {
                const int64_t k = (int64_t)b * C * P2 + oc * P2 + q;
                const float d = __ldg(out2 + k) > 0.f ? __ldg(dflat + k) : 0.f;
                const float* wr = ws[ky * 3 + kx][oc];
#pragma unroll
                for (int c = 0; c < C; ++c) acc[c] = fmaf(d, wr[c], acc[c]);
            }
        }
    }
    const float4* o1 = reinterpret_cast<const float4*>(out1 + ((int64_t)b * P1 + p) * C);
    float4* o = reinterpret_cast<float4*>(dy1 + ((int64_t)b * P1 + p) * C);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 m = __ldg(o1 + q);
        o[q] = make_float4(m.x > 0.f ? acc[4 * q] : 0.f, m.y > 0.f ? acc[4 * q + 1] : 0.f, m.z > 0.f ? acc[4 * q + 2] : 0.f,
                           m.w > 0.f ? acc[4 * q + 3] : 0.f);
    }
}

// per sample: dW1 partial [288] + db1 partial [16]
__global__ void __launch_bounds__(320)
sem_conv1_wgrad_kernel(const float* __restrict__ obs, int64_t stride, const int64_t* __restrict__ rows, int64_t rgb_off,
                       const float* __restrict__ dy1, float* __restrict__ part /*[B][304]*/) {
    const int b = blockIdx.x, tid = threadIdx.x;
    const float* in = frames_of(obs, stride, rows, b, rgb_off);
    const float* d = dy1 + (int64_t)b * P1 * C;
    if (tid < W1N) {
        const int oc = tid / (K_IN * 9), t = tid - oc * (K_IN * 9), ic = t / 9, r = t - ic * 9, ky = r / 3, kx = r - 3 * ky;
        float acc = 0.f;
        for (int p = 0; p < P1; ++p) {
            const int oy = p / O1, ox = p - oy * O1;
            acc = fmaf(__ldg(d + p * C + oc), __ldg(in + (ic * IMG + 2 * oy + ky) * IMG + 2 * ox + kx), acc);
        }
        part[(int64_t)b * (W1N + C) + tid] = acc;
    } else if (tid < W1N + C) {
        const int oc = tid - W1N;
        float acc = 0.f;
        for (int p = 0; p < P1; ++p) acc += __ldg(d + p * C + oc);
        part[(int64_t)b * (W1N + C) + tid] = acc;
    }
}

// out[j] = sum_b part[b*rec + map(j)] in double, fixed order
__global__ void sem_reduce_kernel(const float* __restrict__ part, int B, int64_t rec_stride, int n, float* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    double s = 0.0;
    for (int b = 0; b < B; ++b) s += (double)part[(int64_t)b * rec_stride + j];
    out[j] = (float)s;
}

// dW2 partials are laid out [b][tap][oc][ic]; the parameter is [oc][ic][tap]
__global__ void sem_reduce_w2_kernel(const float* __restrict__ part, int B, float* __restrict__ gw2) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= W2N) return;
    const int oc = j / (C * 9), r = j - oc * C * 9, ic = r / 9, t = r - ic * 9;
    double s = 0.0;
    for (int b = 0; b < B; ++b) s += (double)part[((int64_t)b * 9 + t) * 256 + oc * C + ic];
    gw2[j] = (float)s;
}

}  // namespace

size_t sem2d_out1_floats(int B) { return (size_t)B * P1 * C; }
size_t sem2d_out2_floats(int B) { return (size_t)B * C * P2; }
size_t sem2d_scratch_floats(int B) { return (size_t)B * 9 * 256 + (size_t)B * C + (size_t)B * (W1N + C); }
int sem2d_flat() { return C * P2; }

int launch_sem2d_forward(const float* obs, int64_t stride, const int64_t* rows, int64_t rgb_off, const float* w1, const float* b1,
                         const float* w2, const float* b2, float* out1, float* out2, int B, cudaStream_t stream) {
    sem_conv1_kernel<<<dim3((unsigned)ceil_div(P1, 128), (unsigned)B), 128, 0, stream>>>(obs, stride, rows, rgb_off, w1, b1, out1);
    sem_conv2_kernel<<<B, 256, 0, stream>>>(out1, w2, b2, out2);
    GNBV_LAUNCH_CHECK("sem2d forward");
    return GNBV_OK;
}

int launch_sem2d_backward(const float* obs, int64_t stride, const int64_t* rows, int64_t rgb_off, const float* w2, const float* out1,
                          const float* out2, const float* dflat, float* dy1, float* scratch, float* gw1, float* gb1, float* gw2,
                          float* gb2, int B, cudaStream_t stream) {
    float* part2 = scratch;                                          // [B][9][256] + [B][16]
    float* part1 = scratch + (size_t)B * 9 * 256 + (size_t)B * C;     // [B][304]
    sem_conv2_wgrad_kernel<<<dim3((unsigned)B, 9), 256, 0, stream>>>(out1, out2, dflat, part2, B);
    sem_reduce_w2_kernel<<<(unsigned)ceil_div(W2N, 256), 256, 0, stream>>>(part2, B, gw2);
    sem_reduce_kernel<<<1, 32, 0, stream>>>(part2 + (size_t)B * 9 * 256, B, C, C, gb2);
    sem_conv2_dgrad_kernel<<<dim3((unsigned)ceil_div(P1, 128), (unsigned)B), 128, 0, stream>>>(out1, out2, dflat, w2, dy1);
    sem_conv1_wgrad_kernel<<<B, 320, 0, stream>>>(obs, stride, rows, rgb_off, dy1, part1);
    sem_reduce_kernel<<<(unsigned)ceil_div(W1N, 256), 256, 0, stream>>>(part1, B, W1N + C, W1N, gw1);
    sem_reduce_kernel<<<1, 32, 0, stream>>>(part1 + W1N, B, W1N + C, C, gb1);
    GNBV_LAUNCH_CHECK("sem2d backward");
    return GNBV_OK;
}

}  // namespace gnbv
