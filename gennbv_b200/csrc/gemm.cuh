// gemm.cuh -- fp32 SIMT GEMM used by the encoder / policy layers (sm_100a).
//
// Why CUDA cores and not tcgen05: BASELINE.json asks for <= 1e-4 relative parity on encoder logits, advantages
// and the policy loss against the reference's fp32 PyTorch path (TF32 off).  bf16 / tf32 tensor-core products
// carry 2^-8 / 2^-11 relative error per operand -- one to two orders above the budget -- so every contraction of
// the policy is computed in fp32 FMA with fp32 accumulation.  (A 3xTF32 / bf16x3 split-precision tcgen05 path is
// the planned replacement for the two large contractions, conv2 and the grid Linear; see DESIGN.md.)
//
// C[M,N] = epilogue( sum_p A(i,p) * B(p,j) ), generic strides so that one kernel serves
//   forward   y  = x W^T      A = x [M,K] row-major,    B(p,j) = W[j*K + p]
//   backward  dx = dy W       A = dy [M,N'] row-major,  B = W [N',K] row-major
//   backward  dW = dy^T x     A(i,p) = dy[p*N' + i],    B = x [M',K] row-major
// Tiles 128x64x16, 256 threads, 8x4 outputs per thread, register-prefetch double buffering, optional
// deterministic split-K (partials in a workspace, fixed-order reduction + epilogue in a second kernel).
#pragma once
#include "common.cuh"

namespace gnbv {

struct GemmEpilogue {
    const float* bias = nullptr;   // [N] added to every row
    int relu = 0;                  // max(x, 0)
    const float* row_scale = nullptr;   // optional [M] factor applied before bias (unused by default)
};

// Launches on `stream`. A element (i,p): A[i*sa_m + p*sa_k]; B element (p,j): B[p*sb_k + j*sb_n]; C[i*ldc + j].
// workspace must hold splits*M*N floats when splits > 1 (gemm_splits() tells how many the launcher will use).
int gemm_splits(int M, int N, int K);
size_t gemm_workspace_floats(int M, int N, int K);
int launch_gemm(const float* A, int64_t sa_m, int64_t sa_k, const float* B, int64_t sb_k, int64_t sb_n, float* C,
                int64_t ldc, int M, int N, int K, const GemmEpilogue& ep, float* workspace, cudaStream_t stream);

// Split-precision (3xTF32) tcgen05 version of the same contract (tc_gemm.cu).  `workspace` must hold
// tc_gemm_workspace_floats(M,N,K) floats and be zero-initialised once (its first 64 floats carry a sticky error flag set
// if a bounded mbarrier wait ever expires).
size_t tc_gemm_workspace_floats(int M, int N, int K);
int launch_tc_gemm(const float* A, int64_t sa_m, int64_t sa_k, const float* B, int64_t sb_k, int64_t sb_n, float* C,
                   int64_t ldc, int M, int N, int K, const GemmEpilogue& ep, float* workspace, cudaStream_t stream);

}  // namespace gnbv
