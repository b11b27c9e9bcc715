// nn_grid.cuh -- exact 1-nearest-neighbour search over a uniform cell grid (the per-query logic of
// nn_grid_query_kernel in chamfer.cu).  Plain functions, callable from host and device, so that the search logic (cell
// assignment, shell traversal, termination bound, brute-force escape) is unit-tested on the CPU by
// tests/host/nn_grid_host_test.cu against an O(P1*P2) scan; the kernels in chamfer.cu only add the parallel build.
//
// Layout of one reference cloud after the build (chamfer.cu):
//   * cells: dims[0] x dims[1] x dims[2] cubes of side h covering the cloud's bounding box, z fastest;
//   * cell_end[c] = one past the last sorted point of cell c (cells in linear order, so the points of a run of
//     z-adjacent cells are one contiguous range [cell_end[c0-1], cell_end[c1]));
//   * pts[i] = (x, y, z, 0) of the points, grouped by cell.
//
// Exactness.  After every cell within Chebyshev distance r of the query's (clamped) cell has been visited, any
// unvisited point differs by more than r cells along some axis and is therefore farther than r*h from the query
// (also when the query lies outside the bounding box: its clamped cell is the nearest cell along that axis).  The
// search stops once best <= ((r - 0.01) * h)^2 -- the 1 % slack covers the fp32 rounding of the cell assignment -- or
// when the shells have covered the whole grid.  Distances use the same fp32 expression as the brute-force kernel, so
// both return the identical minimum.  Queries still undecided after NN_RING_MAX shells (far outliers, very sparse
// clouds) scan the whole cloud instead.
#pragma once
#include <float.h>
#include <math.h>

#if defined(__CUDACC__)
#define GNBV_HD __host__ __device__ __forceinline__
#else
#define GNBV_HD static inline
#endif

namespace gnbv {

constexpr int NN_RING_MAX = 6;
constexpr int NN_CELLS_MAX = 160;        // cells per axis (160^3 = 4.1 M cells = 1000 scan tiles, one block scans the tile sums)

struct NNGridMeta {
    float lo[3];        // bounding-box minimum
    float h, inv_h;     // cell side and its reciprocal
    int dims[3];        // cells per axis actually used (<= cells_per_axis)
    int n;              // points in the cloud
    int pad[7];
};
static_assert(sizeof(NNGridMeta) == 64, "NNGridMeta is one 64-byte record per cloud");

struct alignas(16) Float4 { float x, y, z, w; };     // float4 layout (one 128-bit load) without the CUDA vector types on the host

GNBV_HD int nn_clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// meta from a bounding box (lo / hi per axis) and the requested resolution
GNBV_HD void nn_make_meta(NNGridMeta& m, const float lo[3], const float hi[3], int n, int cells_per_axis) {
    float ext = 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a) { m.lo[a] = lo[a]; ext = fmaxf(ext, hi[a] - lo[a]); }
    m.n = n;
    if (!(ext > 0.f) || n <= 0) {        // empty cloud, a single point or coincident points: one cell
        m.h = 1.f; m.inv_h = 1.f; m.dims[0] = m.dims[1] = m.dims[2] = n > 0 ? 1 : 0;
        return;
    }
    m.h = ext / (float)cells_per_axis;
    m.inv_h = 1.f / m.h;
#pragma unroll
    for (int a = 0; a < 3; ++a) m.dims[a] = nn_clampi((int)((hi[a] - lo[a]) * m.inv_h) + 1, 1, cells_per_axis);
}

GNBV_HD void nn_cell_of(const NNGridMeta& m, float x, float y, float z, int c[3]) {
    const float p[3] = {x, y, z};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float t = floorf((p[a] - m.lo[a]) * m.inv_h);
        t = fminf(fmaxf(t, 0.f), (float)(m.dims[a] - 1));     // clamp in float first: far queries must not overflow int
        c[a] = (int)t;
    }
}

GNBV_HD int nn_cell_linear(const NNGridMeta& m, const int c[3]) { return (c[0] * m.dims[1] + c[1]) * m.dims[2] + c[2]; }

GNBV_HD float nn_sqdist(float qx, float qy, float qz, const Float4& p) {
    const float dx = qx - p.x, dy = qy - p.y, dz = qz - p.z;
    return fmaf(dz, dz, fmaf(dy, dy, dx * dx));
}

// points of the cells (ix, iy, z0..z1), a contiguous range of the sorted array
GNBV_HD float nn_visit_run(const NNGridMeta& m, const int* cell_end, const Float4* pts, int ix, int iy, int z0, int z1,
                           float qx, float qy, float qz, float best) {
    const int lin0 = (ix * m.dims[1] + iy) * m.dims[2] + z0, lin1 = lin0 + (z1 - z0);
    const int s = lin0 > 0 ? cell_end[lin0 - 1] : 0, e = cell_end[lin1];
    for (int i = s; i < e; ++i) best = fminf(best, nn_sqdist(qx, qy, qz, pts[i]));
    return best;
}

// min_j |q - pts_j|^2 over one cloud (FLT_MAX for an empty cloud)
GNBV_HD float nn_query(const NNGridMeta& m, const int* cell_end, const Float4* pts, float qx, float qy, float qz) {
    if (m.n <= 0) return FLT_MAX;
    int c[3];
    nn_cell_of(m, qx, qy, qz, c);
    int rmax = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const int far = c[a] > m.dims[a] - 1 - c[a] ? c[a] : m.dims[a] - 1 - c[a];
        rmax = far > rmax ? far : rmax;
    }
    float best = FLT_MAX;
    for (int r = 0; r <= rmax; ++r) {
        if (r > NN_RING_MAX) {           // escape: scan the whole cloud (exact; the visited part is simply seen again)
            for (int i = 0; i < m.n; ++i) best = fminf(best, nn_sqdist(qx, qy, qz, pts[i]));
            return best;
        }
        const int x0 = c[0] - r > 0 ? c[0] - r : 0, x1 = c[0] + r < m.dims[0] - 1 ? c[0] + r : m.dims[0] - 1;
        const int y0 = c[1] - r > 0 ? c[1] - r : 0, y1 = c[1] + r < m.dims[1] - 1 ? c[1] + r : m.dims[1] - 1;
        const int z0 = c[2] - r > 0 ? c[2] - r : 0, z1 = c[2] + r < m.dims[2] - 1 ? c[2] + r : m.dims[2] - 1;
        for (int ix = x0; ix <= x1; ++ix) {
            const bool x_face = (ix == c[0] - r) || (ix == c[0] + r);
            for (int iy = y0; iy <= y1; ++iy) {
                if (x_face || iy == c[1] - r || iy == c[1] + r) {
                    best = nn_visit_run(m, cell_end, pts, ix, iy, z0, z1, qx, qy, qz, best);     // whole z column of the shell
                } else {                                                                          // only its two z caps
                    if (c[2] - r >= 0) best = nn_visit_run(m, cell_end, pts, ix, iy, c[2] - r, c[2] - r, qx, qy, qz, best);
                    if (c[2] + r <= m.dims[2] - 1 && r > 0)
                        best = nn_visit_run(m, cell_end, pts, ix, iy, c[2] + r, c[2] + r, qx, qy, qz, best);
                }
            }
        }
        if (r >= 1) {
            const float lb = ((float)r - 0.01f) * m.h;
            if (best <= lb * lb) break;
        } else if (best == 0.f) {
            break;
        }
    }
    return best;
}

}  // namespace gnbv
