// conv2_mma.cuh -- mma.sync (3xTF32) implicit-GEMM kernels for Conv3d(16,16,3,stride 2) of the Hybrid_Encoder.
#pragma once
#include "common.cuh"

namespace gnbv {

// Work items (256 output voxels of one sample) = BN-statistics records the forward kernel writes.
int conv2_mma_chunks(int G2);
int conv2_mma_items(int B, int G2);

// y2 [B,16,G2^3] (pre-BN, channel-major) = conv(relu(a1*y1+b1)) + bias; y1 [B,G1^3,16] channels-last; stat1 [4][16];
// part: per-item (mean[16], M2[16], count) records (stride 36 floats) or NULL.
int launch_conv2_fwd_mma(const float* y1, const float* stat1, const float* w, const float* bias, float* y2, float* part,
                         int B, int G1, int G2, cudaStream_t stream);

// conv2 data gradient with the ReLU mask / BN1-backward sums fused (contract of conv2_dgrad_kernel in encoder.cu):
// dy2cl [B,G2^3,16] channels-last; g1 [B,G1^3,16]; bpart: one [32] record (sum g1, sum g1*xhat1) per work item.
int conv2_dgrad_mma_items_per_sample(int G1);
int launch_conv2_dgrad_mma(const float* dy2cl, const float* w, const float* y1, const float* stat1, float* g1, float* bpart,
                           int B, int G1, int G2, cudaStream_t stream);

// conv2 weight gradient (contract of conv2_wgrad_kernel): block blk reduces output z-rows [blk*rows_per_block, ...) and
// writes part[blk][16*16*27 + 16] ([co][ci][tap] followed by db2[16]).
int launch_conv2_wgrad_mma(const float* y1, const float* stat1, const float* dy2cl, float* part, int B, int G1, int G2,
                           int nblocks, int rows_per_block, cudaStream_t stream);

// Same contract, TMA-staged (groups of four output rows in shared memory, bank-conflict-free fragment loads): launches at
// most min(max_blocks, 148) blocks and reports how many records it wrote.
int launch_conv2_wgrad_staged(const float* y1, const float* stat1, const float* dy2cl, float* part, int B, int G1, int G2,
                              int max_blocks, int* nblocks_out, cudaStream_t stream);

}  // namespace gnbv
