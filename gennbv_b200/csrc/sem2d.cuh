// sem2d.cuh -- 2-D semantic branch kernels (sem2d.cu), called from the encoder forward / backward.
#pragma once
#include "common.cuh"

namespace gnbv {

size_t sem2d_out1_floats(int B);      // conv1 output [B, 31*31, 16] channels-last, post-ReLU
size_t sem2d_out2_floats(int B);      // conv2 output [B, 16*15*15] channel-major (Flatten order), post-ReLU
size_t sem2d_scratch_floats(int B);   // per-sample weight-gradient partials
int sem2d_flat();                     // 3600

int launch_sem2d_forward(const float* obs, int64_t stride, const int64_t* rows, int64_t rgb_off, const float* w1, const float* b1,
                         const float* w2, const float* b2, float* out1, float* out2, int B, cudaStream_t stream);
// dflat: gradient w.r.t. out2 (post-ReLU), dy1: scratch [B, 961, 16]
int launch_sem2d_backward(const float* obs, int64_t stride, const int64_t* rows, int64_t rgb_off, const float* w2, const float* out1,
                          const float* out2, const float* dflat, float* dy1, float* scratch, float* gw1, float* gb1, float* gw2,
                          float* gb2, int B, cudaStream_t stream);

}  // namespace gnbv
