// encoder.cuh -- internal entry points of encoder.cu shared with ppo_update.cu.
#pragma once
#include "common.cuh"

namespace gnbv {

// gnbv_encoder_forward with one more argument: `freeze` (device int64 flag or NULL).  While *freeze != 0 a training-mode
// forward does not update the BatchNorm running statistics (the sticky KL-stop flag of the fused PPO update).
int encoder_forward_impl(const gnbv_encoder_params* p, const float* obs, int64_t obs_row_stride, const int64_t* row_index,
                         int batch, int grid_size, int state_dim, int training, float* features, void* workspace,
                         size_t workspace_bytes, const int64_t* freeze, cudaStream_t stream);

}  // namespace gnbv
