// chamfer.cu -- exact bidirectional 1-nearest-neighbour squared-L2 distances: the computation behind
// pytorch3d.loss.chamfer_distance as the reference's eval env calls it (gennbv/env/env_eval_gennbv.py:253-261, defaults:
// squared L2, point_reduction = "mean", batch_reduction = "mean", both directions summed).  pytorch3d 0.7.x is a
// third-party dependency that is not vendored in the reference; this restates its published brute-force definition
//     cham(x, y) = mean_i min_j |x_i - y_j|^2 + mean_j min_i |x_i - y_j|^2 .
//
// One launch per direction: a block owns 256 query points (one per thread, in registers) of one cloud and streams the
// reference cloud through shared memory in 2048-point tiles (float4-padded, broadcast LDS.128): 3 SUB + 3 FMA + 1 MIN per
// pair, FP32-ALU bound.  Per-block sums of the minima are written out and reduced in double in a fixed order.
//
// gnbv_chamfer_grid is the same computation with the inner scan replaced by an exact uniform-grid search (nn_grid.cuh):
// per direction, the reference cloud is binned into cubic cells (bounding box -> per-cell counts -> exclusive scan ->
// counting-sort fill, all on the device, all clouds per launch) and every query visits the shells of cells around its
// own cell until the best distance is provably minimal.  Work drops from P1*P2 pairs to ~27 cells x points-per-cell per
// query, which turns the eval stress shape (256 envs x 100 k GT x 819 k scanned points) from seconds into milliseconds;
// the per-point minima are bit-identical to the brute-force kernel's (same fp32 distance expression).
#include "common.cuh"
#include "nn_grid.cuh"

#include <float.h>

#include <algorithm>

namespace gnbv {

constexpr int NN_THREADS = 256;
constexpr int NN_TILE = 2048;

// clouds are packed: points [total,3] f32, offsets [E+1] i64.  partial[e * max_blocks + blk] = sum of min d^2
__global__ void __launch_bounds__(NN_THREADS)
nn_min_sum_kernel(const float* __restrict__ q, const int64_t* __restrict__ q_off, const float* __restrict__ r,
                  const int64_t* __restrict__ r_off, float* __restrict__ partial, float* __restrict__ min_out, int max_blocks) {
    __shared__ float4 tile[NN_TILE];
    __shared__ float wsum[NN_THREADS / 32];
    const int e = blockIdx.y, tid = threadIdx.x;
    const int64_t q0 = q_off[e], nq = q_off[e + 1] - q0, r0 = r_off[e], nr = r_off[e + 1] - r0;
    float block_total = 0.f;
    for (int64_t base = (int64_t)blockIdx.x * NN_THREADS; base < nq; base += (int64_t)gridDim.x * NN_THREADS) {
        const int64_t qi = base + tid;
        const bool active = qi < nq;
        float qx = 0.f, qy = 0.f, qz = 0.f;
        if (active) { qx = q[(q0 + qi) * 3]; qy = q[(q0 + qi) * 3 + 1]; qz = q[(q0 + qi) * 3 + 2]; }
        float best = FLT_MAX;
        for (int64_t t0 = 0; t0 < nr; t0 += NN_TILE) {
            const int n = (int)min((int64_t)NN_TILE, nr - t0);
            __syncthreads();
            for (int i = tid; i < n; i += NN_THREADS) {
                const float* p = r + (r0 + t0 + i) * 3;
                tile[i] = make_float4(p[0], p[1], p[2], 0.f);
            }
            __syncthreads();
#pragma unroll 8
            for (int i = 0; i < n; ++i) {
                const float4 p = tile[i];
                const float dx = qx - p.x, dy = qy - p.y, dz = qz - p.z;
                best = fminf(best, fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
            }
        }
        if (active && min_out) min_out[q0 + qi] = best;
        float s = (active && nr > 0) ? best : 0.f;
        s = warp_sum(s);
        __syncthreads();
        if ((tid & 31) == 0) wsum[tid >> 5] = s;
        __syncthreads();
        if (tid == 0) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < NN_THREADS / 32; ++w) t += wsum[w];
            block_total += t;
        }
    }
    if (tid == 0) partial[(int64_t)e * max_blocks + blockIdx.x] = block_total;
}

// out[e] (+)= mean over the cloud: sum of block partials (double, fixed order) / count
__global__ void nn_finalize_kernel(const float* __restrict__ partial, int max_blocks, const int64_t* __restrict__ q_off,
                                   float* __restrict__ out, int out_stride, int accumulate, int E) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    double s = 0.0;
    for (int b = 0; b < max_blocks; ++b) s += (double)partial[(int64_t)e * max_blocks + b];
    const int64_t n = q_off[e + 1] - q_off[e];
    const float v = n > 0 ? (float)(s / (double)n) : 0.f;
    out[e * out_stride] = accumulate ? out[e * out_stride] + v : v;
}


// ---------------------------------------------------------------------------------------------------------------------
// uniform-grid build (all clouds per launch; cloud = blockIdx.y).  cells [E, C^3] i32, tile_sums [E, 1024] i32.
constexpr int GRID_BUILD_THREADS = 256;
constexpr int GRID_BUILD_BLOCKS = 128;         // point blocks per cloud (grid-stride)
constexpr int SCAN_THREADS = 1024;
constexpr int SCAN_TILE = SCAN_THREADS * 4;     // cells per scan block

__device__ __forceinline__ float block_minmax(float v, bool is_min, float* scratch) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float t = __shfl_xor_sync(0xffffffffu, v, o);
        v = is_min ? fminf(v, t) : fmaxf(v, t);
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = scratch[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r = is_min ? fminf(r, scratch[w]) : fmaxf(r, scratch[w]);
    return r;
}

// one block per cloud: bounding box -> NNGridMeta
__global__ void __launch_bounds__(SCAN_THREADS)
nn_bbox_kernel(const float* __restrict__ r, const int64_t* __restrict__ r_off, NNGridMeta* __restrict__ metas, int cells_per_axis) {
    __shared__ float scratch[SCAN_THREADS / 32];
    const int e = blockIdx.x;
    const int64_t r0 = r_off[e], n = r_off[e + 1] - r0;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int64_t i = threadIdx.x; i < n; i += SCAN_THREADS) {
        const float* p = r + (r0 + i) * 3;
#pragma unroll
        for (int a = 0; a < 3; ++a) { lo[a] = fminf(lo[a], p[a]); hi[a] = fmaxf(hi[a], p[a]); }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) { lo[a] = block_minmax(lo[a], true, scratch); hi[a] = block_minmax(hi[a], false, scratch); }
    if (threadIdx.x == 0) {
        NNGridMeta m;
#pragma unroll
        for (int k = 0; k < 7; ++k) m.pad[k] = 0;
        nn_make_meta(m, lo, hi, (int)n, cells_per_axis);
        metas[e] = m;
    }
}

// FILL = false: cells[c] += 1 per point;  FILL = true: slot = cells[c]++ (after the exclusive scan) and the point is stored
template <bool FILL>
__global__ void __launch_bounds__(GRID_BUILD_THREADS)
nn_bin_kernel(const float* __restrict__ r, const int64_t* __restrict__ r_off, const NNGridMeta* __restrict__ metas,
              int* __restrict__ cells, int64_t cells_stride, Float4* __restrict__ sorted) {
    const int e = blockIdx.y;
    const NNGridMeta m = metas[e];
    const int64_t r0 = r_off[e];
    int* cl = cells + (int64_t)e * cells_stride;
    for (int i = blockIdx.x * GRID_BUILD_THREADS + threadIdx.x; i < m.n; i += gridDim.x * GRID_BUILD_THREADS) {
        const float* p = r + (r0 + i) * 3;
        const float x = p[0], y = p[1], z = p[2];
        int c[3];
        nn_cell_of(m, x, y, z, c);
        const int lin = nn_cell_linear(m, c);
        if (FILL) {
            const int slot = atomicAdd(&cl[lin], 1);
            Float4 v; v.x = x; v.y = y; v.z = z; v.w = 0.f;
            sorted[r0 + slot] = v;
        } else {
            atomicAdd(&cl[lin], 1);
        }
    }
}

// exclusive scan of blockDim.x values (one per thread); returns the exclusive prefix, *total = sum over the block
__device__ __forceinline__ int block_exclusive_scan(int v, int* warp_tot, int* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    __syncthreads();
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int t = warp_tot[lane], s = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int q = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += q;
        }
        warp_tot[lane] = s - t;
        if (lane == 31) *total = s;
    }
    __syncthreads();
    return warp_tot[warp] + incl - v;
}

// phase 1: each block turns its tile of 4096 counts into tile-local exclusive offsets and records the tile total
__global__ void __launch_bounds__(SCAN_THREADS)
nn_scan_tiles_kernel(const NNGridMeta* __restrict__ metas, int* __restrict__ cells, int64_t cells_stride, int* __restrict__ tile_sums) {
    __shared__ int warp_tot[32];
    __shared__ int total;
    const int e = blockIdx.y, tile = blockIdx.x;
    const NNGridMeta m = metas[e];
    const int ncell = m.dims[0] * m.dims[1] * m.dims[2];
    if (tile * SCAN_TILE >= ncell) return;                         // whole block leaves together
    int* cl = cells + (int64_t)e * cells_stride + (int64_t)tile * SCAN_TILE;
    int4 v = reinterpret_cast<int4*>(cl)[threadIdx.x];             // cells beyond ncell are zero (memset)
    const int sum = v.x + v.y + v.z + v.w;
    const int ex = block_exclusive_scan(sum, warp_tot, &total);
    int4 o;
    o.x = ex; o.y = ex + v.x; o.z = o.y + v.y; o.w = o.z + v.z;
    reinterpret_cast<int4*>(cl)[threadIdx.x] = o;
    if (threadIdx.x == 0) tile_sums[e * SCAN_THREADS + tile] = total;
}

// phase 2: one block per cloud scans the (<= 1024) tile totals
__global__ void __launch_bounds__(SCAN_THREADS)
nn_scan_sums_kernel(const NNGridMeta* __restrict__ metas, int* __restrict__ tile_sums) {
    __shared__ int warp_tot[32];
    __shared__ int total;
    const int e = blockIdx.x;
    const NNGridMeta m = metas[e];
    const int ntile = (m.dims[0] * m.dims[1] * m.dims[2] + SCAN_TILE - 1) / SCAN_TILE;
    const int v = (int)threadIdx.x < ntile ? tile_sums[e * SCAN_THREADS + threadIdx.x] : 0;
    const int ex = block_exclusive_scan(v, warp_tot, &total);
    if ((int)threadIdx.x < ntile) tile_sums[e * SCAN_THREADS + threadIdx.x] = ex;
}

// phase 3: add the tile offset
__global__ void __launch_bounds__(SCAN_THREADS)
nn_scan_add_kernel(const NNGridMeta* __restrict__ metas, int* __restrict__ cells, int64_t cells_stride, const int* __restrict__ tile_sums) {
    const int e = blockIdx.y, tile = blockIdx.x;
    const NNGridMeta m = metas[e];
    if (tile == 0 || tile * SCAN_TILE >= m.dims[0] * m.dims[1] * m.dims[2]) return;
    const int add = tile_sums[e * SCAN_THREADS + tile];
    int4* cl = reinterpret_cast<int4*>(cells + (int64_t)e * cells_stride + (int64_t)tile * SCAN_TILE);
    int4 v = cl[threadIdx.x];
    v.x += add; v.y += add; v.z += add; v.w += add;
    cl[threadIdx.x] = v;
}

// same outer structure (and therefore the same summation order) as nn_min_sum_kernel, inner scan = grid search
__global__ void __launch_bounds__(NN_THREADS)
nn_grid_query_kernel(const float* __restrict__ q, const int64_t* __restrict__ q_off, const int64_t* __restrict__ r_off,
                     const NNGridMeta* __restrict__ metas, const int* __restrict__ cells, int64_t cells_stride,
                     const Float4* __restrict__ sorted, float* __restrict__ partial, float* __restrict__ min_out, int max_blocks) {
    __shared__ float wsum[NN_THREADS / 32];
    const int e = blockIdx.y, tid = threadIdx.x;
    const NNGridMeta m = metas[e];
    const int64_t q0 = q_off[e], nq = q_off[e + 1] - q0;
    const int* cell_end = cells + (int64_t)e * cells_stride;
    const Float4* pts = sorted + r_off[e];
    float block_total = 0.f;
    for (int64_t base = (int64_t)blockIdx.x * NN_THREADS; base < nq; base += (int64_t)gridDim.x * NN_THREADS) {
        const int64_t qi = base + tid;
        const bool active = qi < nq;
        float best = FLT_MAX;
        if (active) {
            const float* p = q + (q0 + qi) * 3;
            best = nn_query(m, cell_end, pts, p[0], p[1], p[2]);
            if (min_out) min_out[q0 + qi] = best;
        }
        float s = (active && m.n > 0) ? best : 0.f;
        s = warp_sum(s);
        __syncthreads();
        if ((tid & 31) == 0) wsum[tid >> 5] = s;
        __syncthreads();
        if (tid == 0) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < NN_THREADS / 32; ++w) t += wsum[w];
            block_total += t;
        }
    }
    if (tid == 0) partial[(int64_t)e * max_blocks + blockIdx.x] = block_total;
}

}  // namespace gnbv

using namespace gnbv;

constexpr int CHAMFER_BLOCKS = 296;          // 2 x 148 query blocks per cloud

extern "C" size_t gnbv_chamfer_workspace_bytes(int num_clouds) {
    return num_clouds > 0 ? (size_t)num_clouds * CHAMFER_BLOCKS * sizeof(float) : 0;
}

extern "C" int gnbv_chamfer(const float* x, const int64_t* x_offsets, const float* y, const int64_t* y_offsets, int num_clouds,
                            float* cham_x, float* cham_y, float* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GNBV_REQUIRE(x && x_offsets && y && y_offsets && cham_x && cham_y && workspace && num_clouds > 0, "gnbv_chamfer: bad arguments");
    GNBV_REQUIRE(workspace_bytes >= gnbv_chamfer_workspace_bytes(num_clouds), "gnbv_chamfer: workspace too small");
    dim3 grid(CHAMFER_BLOCKS, num_clouds);
    nn_min_sum_kernel<<<grid, NN_THREADS, 0, stream>>>(x, x_offsets, y, y_offsets, workspace, nullptr, CHAMFER_BLOCKS);
    GNBV_LAUNCH_CHECK("nn_min_sum_kernel");
    nn_finalize_kernel<<<(unsigned)ceil_div(num_clouds, 128), 128, 0, stream>>>(workspace, CHAMFER_BLOCKS, x_offsets, cham_x, 1, 0, num_clouds);
    nn_min_sum_kernel<<<grid, NN_THREADS, 0, stream>>>(y, y_offsets, x, x_offsets, workspace, nullptr, CHAMFER_BLOCKS);
    GNBV_LAUNCH_CHECK("nn_min_sum_kernel");
    nn_finalize_kernel<<<(unsigned)ceil_div(num_clouds, 128), 128, 0, stream>>>(workspace, CHAMFER_BLOCKS, y_offsets, cham_y, 1, 0, num_clouds);
    GNBV_LAUNCH_CHECK("nn_finalize_kernel");
    return GNBV_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
namespace {
struct GridLayout {
    int64_t cells_stride;                     // ints per cloud (C^3 rounded up to whole scan tiles)
    size_t off_partial, off_meta, off_tiles, off_cells, off_sorted, total;
};
GridLayout grid_layout(int E, int64_t total_x, int64_t total_y, int C) {
    GridLayout L;
    const int64_t c3 = (int64_t)C * C * C;
    L.cells_stride = ceil_div(c3, SCAN_TILE) * SCAN_TILE;
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    L.off_partial = 0;
    L.off_meta = up(L.off_partial + (size_t)E * CHAMFER_BLOCKS * sizeof(float));
    L.off_tiles = up(L.off_meta + (size_t)E * sizeof(NNGridMeta));
    L.off_cells = up(L.off_tiles + (size_t)E * SCAN_THREADS * sizeof(int));
    L.off_sorted = up(L.off_cells + (size_t)E * L.cells_stride * sizeof(int));
    L.total = up(L.off_sorted + (size_t)std::max(total_x, total_y) * sizeof(Float4));
    return L;
}

// mean_i min_j |q_i - r_j|^2 for every cloud, grid built over r
int nn_grid_direction(const float* q, const int64_t* q_off, const float* r, const int64_t* r_off, int E, int C,
                      const GridLayout& L, char* ws, float* mean_out, float* min_out, cudaStream_t stream) {
    float* partial = reinterpret_cast<float*>(ws + L.off_partial);
    NNGridMeta* metas = reinterpret_cast<NNGridMeta*>(ws + L.off_meta);
    int* tiles = reinterpret_cast<int*>(ws + L.off_tiles);
    int* cells = reinterpret_cast<int*>(ws + L.off_cells);
    Float4* sorted = reinterpret_cast<Float4*>(ws + L.off_sorted);
    const unsigned ntile = (unsigned)(L.cells_stride / SCAN_TILE);
    GNBV_CUDA_CHECK(cudaMemsetAsync(cells, 0, (size_t)E * L.cells_stride * sizeof(int), stream));
    nn_bbox_kernel<<<E, SCAN_THREADS, 0, stream>>>(r, r_off, metas, C);
    GNBV_LAUNCH_CHECK("nn_bbox_kernel");
    nn_bin_kernel<false><<<dim3(GRID_BUILD_BLOCKS, E), GRID_BUILD_THREADS, 0, stream>>>(r, r_off, metas, cells, L.cells_stride, sorted);
    GNBV_LAUNCH_CHECK("nn_bin_kernel<count>");
    nn_scan_tiles_kernel<<<dim3(ntile, E), SCAN_THREADS, 0, stream>>>(metas, cells, L.cells_stride, tiles);
    GNBV_LAUNCH_CHECK("nn_scan_tiles_kernel");
    nn_scan_sums_kernel<<<E, SCAN_THREADS, 0, stream>>>(metas, tiles);
    GNBV_LAUNCH_CHECK("nn_scan_sums_kernel");
    nn_scan_add_kernel<<<dim3(ntile, E), SCAN_THREADS, 0, stream>>>(metas, cells, L.cells_stride, tiles);
    GNBV_LAUNCH_CHECK("nn_scan_add_kernel");
    nn_bin_kernel<true><<<dim3(GRID_BUILD_BLOCKS, E), GRID_BUILD_THREADS, 0, stream>>>(r, r_off, metas, cells, L.cells_stride, sorted);
    GNBV_LAUNCH_CHECK("nn_bin_kernel<fill>");
    nn_grid_query_kernel<<<dim3(CHAMFER_BLOCKS, E), NN_THREADS, 0, stream>>>(q, q_off, r_off, metas, cells, L.cells_stride, sorted,
                                                                            partial, min_out, CHAMFER_BLOCKS);
    GNBV_LAUNCH_CHECK("nn_grid_query_kernel");
    nn_finalize_kernel<<<(unsigned)ceil_div(E, 128), 128, 0, stream>>>(partial, CHAMFER_BLOCKS, q_off, mean_out, 1, 0, E);
    GNBV_LAUNCH_CHECK("nn_finalize_kernel");
    return GNBV_OK;
}
}  // namespace

extern "C" size_t gnbv_chamfer_grid_workspace_bytes(int num_clouds, int64_t total_x, int64_t total_y, int cells_per_axis) {
    if (num_clouds <= 0 || total_x < 0 || total_y < 0 || cells_per_axis < 1 || cells_per_axis > NN_CELLS_MAX) return 0;
    return grid_layout(num_clouds, total_x, total_y, cells_per_axis).total;
}

extern "C" int gnbv_chamfer_grid(const float* x, const int64_t* x_offsets, const float* y, const int64_t* y_offsets, int num_clouds,
                                 int64_t total_x, int64_t total_y, int cells_per_axis, float* cham_x, float* cham_y,
                                 float* min_x, float* min_y, void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GNBV_REQUIRE(x && x_offsets && y && y_offsets && cham_x && cham_y && workspace && num_clouds > 0, "gnbv_chamfer_grid: bad arguments");
    GNBV_REQUIRE(cells_per_axis >= 1 && cells_per_axis <= NN_CELLS_MAX, "gnbv_chamfer_grid: cells_per_axis must be in [1, %d]", NN_CELLS_MAX);
    GNBV_REQUIRE(total_x >= 0 && total_y >= 0 && total_x < (1LL << 31) && total_y < (1LL << 31),
                 "gnbv_chamfer_grid: point totals must fit 31 bits");
    GNBV_REQUIRE(((uintptr_t)workspace & 255) == 0, "gnbv_chamfer_grid: workspace must be 256-byte aligned");
    const GridLayout L = grid_layout(num_clouds, total_x, total_y, cells_per_axis);
    GNBV_REQUIRE(workspace_bytes >= L.total, "gnbv_chamfer_grid: workspace too small");
    char* ws = static_cast<char*>(workspace);
    int rc = nn_grid_direction(x, x_offsets, y, y_offsets, num_clouds, cells_per_axis, L, ws, cham_x, min_x, stream);
    if (rc != GNBV_OK) return rc;
    return nn_grid_direction(y, y_offsets, x, x_offsets, num_clouds, cells_per_axis, L, ws, cham_y, min_y, stream);
}

/* per-point minima of the brute-force kernel (test / cross-check entry: the grid search must reproduce them bit for bit) */
extern "C" int gnbv_nn_sqdist_brute(const float* q, const int64_t* q_offsets, const float* r, const int64_t* r_offsets, int num_clouds,
                                    float* min_out, float* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GNBV_REQUIRE(q && q_offsets && r && r_offsets && min_out && workspace && num_clouds > 0, "gnbv_nn_sqdist_brute: bad arguments");
    GNBV_REQUIRE(workspace_bytes >= gnbv_chamfer_workspace_bytes(num_clouds), "gnbv_nn_sqdist_brute: workspace too small");
    nn_min_sum_kernel<<<dim3(CHAMFER_BLOCKS, num_clouds), NN_THREADS, 0, stream>>>(q, q_offsets, r, r_offsets, workspace, min_out,
                                                                                  CHAMFER_BLOCKS);
    GNBV_LAUNCH_CHECK("nn_min_sum_kernel");
    return GNBV_OK;
}
