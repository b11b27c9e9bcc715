// conv2_ts.cuh -- tcgen05 "TS" (A operand in tensor memory) kernels for Conv3d(16,16,3,stride 2) of the Hybrid_Encoder.
#pragma once
#include "common.cuh"

namespace gnbv {

bool conv2_ts_supported(int G1, int G2);
// Number of (b, y-block, x2) tiles = BN-statistics records the forward kernel writes.
int conv2_ts_tiles(int B, int G2);

// y2 [B,16,G2^3] (pre-BN, channel-major) = conv(relu(a1*y1+b1)) + bias; y1 [B,G1^3,16] channels-last; stat1 [4][16];
// part: per-tile (mean[16], M2[16], count) records (stride 36 floats) or NULL.
int launch_conv2_fwd_ts(const float* y1, const float* stat1, const float* w, const float* bias, float* y2, float* part, int B,
                        int G1, int G2, cudaStream_t stream);

// g1 [B,G1^3,16] = (conv2 data gradient of dy2cl [B,G2^3,16]) * [bn1(y1) > 0]; bpart: conv2_dgrad_ts_records(B, G1) records of
// (sum g1 [16], sum g1 * xhat1 [16]).
int conv2_dgrad_ts_records(int B, int G1);
int launch_conv2_dgrad_ts(const float* dy2cl, const float* w, const float* y1, const float* stat1, float* g1, float* bpart, int B,
                          int G1, int G2, cudaStream_t stream);

}  // namespace gnbv
