// sort_unique.cu -- in-tree LSD radix sort + unique of 64-bit keys (replaces torch.unique / the CUB sort behind it on the eval
// path: the 1 cm de-duplication of the scanned point history, env_eval_gennbv.py:254-257, operates on packed lattice keys).
//
// One pass sorts by an 8-bit digit, stably, in three launches:
//   rs_hist_kernel    per tile (4096 keys) digit counts -> hist[digit][tile]
//   rs_scan_kernel    one block per digit: exclusive scan of its row over the tiles, row total -> total[digit]
//   rs_scatter_kernel digit bases from total[] (scanned in shared memory), then every key goes to
//                     base[digit] + hist[digit][tile] + (rank of the key among the tile's keys with that digit), the rank being
//                     computed warp by warp with __match_any_sync (keys of a warp in index order) and a per-warp running count.
// Unique: heads (key != predecessor) counted per tile, tile counts scanned by one block, heads compacted in order.
// All counters are 32-bit: n < 2^31 keys.
#include "common.cuh"

#include <algorithm>

namespace gnbv {
namespace {

constexpr int RS_THREADS = 256, RS_ITEMS = 16, RS_TILE = RS_THREADS * RS_ITEMS, RS_WARPS = RS_THREADS / 32;

__global__ void __launch_bounds__(RS_THREADS)
rs_hist_kernel(const uint64_t* __restrict__ keys, int64_t n, int shift, int* __restrict__ hist, int ntiles) {
    __shared__ int h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RS_TILE;
    for (int i = 0; i < RS_ITEMS; ++i) {
        const int64_t idx = base + i * RS_THREADS + threadIdx.x;
        if (idx < n) atomicAdd(&h[(int)((keys[idx] >> shift) & 0xff)], 1);
    }
    __syncthreads();
    hist[(int64_t)threadIdx.x * ntiles + blockIdx.x] = h[threadIdx.x];
}

// block-wide exclusive scan of one int per thread (256 threads); returns the exclusive prefix, *total = block sum
__device__ __forceinline__ int block_exclusive_scan(int v, int* sh /*[RS_WARPS]*/, int* total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) sh[w] = x;
    __syncthreads();
    int wbase = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < RS_WARPS; ++k) { const int s = sh[k]; if (k < w) wbase += s; tot += s; }
    __syncthreads();
    *total = tot;
    return wbase + x - v;
}

// grid = 256 (one block per digit): hist[d][0..ntiles) -> exclusive prefix in place, total[d] = row sum
__global__ void __launch_bounds__(RS_THREADS)
rs_scan_kernel(int* __restrict__ hist, int ntiles, int* __restrict__ total) {
    __shared__ int sh[RS_WARPS];
    int* row = hist + (int64_t)blockIdx.x * ntiles;
    int carry = 0;
    for (int t0 = 0; t0 < ntiles; t0 += RS_THREADS) {
        const int t = t0 + threadIdx.x;
        const int v = t < ntiles ? row[t] : 0;
        int tot;
        const int ex = block_exclusive_scan(v, sh, &tot);
        if (t < ntiles) row[t] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) total[blockIdx.x] = carry;
}

__global__ void __launch_bounds__(RS_THREADS)
rs_scatter_kernel(const uint64_t* __restrict__ in, uint64_t* __restrict__ out, int64_t n, int shift, const int* __restrict__ hist,
                  int ntiles, const int* __restrict__ total) {
    __shared__ int whist[RS_WARPS][256];
    __shared__ int dbase[256];
    __shared__ int sh[RS_WARPS];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    for (int k = 0; k < RS_WARPS; ++k) whist[k][tid] = 0;
    int tot;
    const int ex = block_exclusive_scan(total[tid], sh, &tot);           // digit bases (all tiles of smaller digits come first)
    dbase[tid] = ex + hist[(int64_t)tid * ntiles + blockIdx.x];
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RS_TILE + (int64_t)w * (RS_ITEMS * 32);
    uint64_t key[RS_ITEMS];
    int local[RS_ITEMS];
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const int64_t idx = base + i * 32 + lane;
        const bool valid = idx < n;
        key[i] = valid ? in[idx] : 0;
        const int d = (int)((key[i] >> shift) & 0xff);
        const unsigned mask = __match_any_sync(0xffffffffu, valid ? d : 256 + lane);
        const int rank = __popc(mask & ((1u << lane) - 1));
        const int before = valid ? whist[w][d] : 0;                      // keys of this digit in the warp's earlier items
        __syncwarp();
        if (valid && rank == 0) whist[w][d] = before + __popc(mask);
        __syncwarp();
        local[i] = valid ? before + rank : -1;
    }
    __syncthreads();
    {   // per digit: exclusive prefix over the warps, made absolute
        int acc = dbase[tid];
#pragma unroll
        for (int k = 0; k < RS_WARPS; ++k) { const int c = whist[k][tid]; whist[k][tid] = acc; acc += c; }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i)
        if (local[i] >= 0) out[whist[w][(int)((key[i] >> shift) & 0xff)] + local[i]] = key[i];
}

// heads of runs of equal keys in a sorted array: per tile count, then compaction
__global__ void __launch_bounds__(RS_THREADS)
uq_count_kernel(const uint64_t* __restrict__ keys, int64_t n, int* __restrict__ tile_count) {
    __shared__ int sh[RS_WARPS];
    const int64_t base = (int64_t)blockIdx.x * RS_TILE + (int64_t)threadIdx.x * RS_ITEMS;
    int c = 0;
    for (int i = 0; i < RS_ITEMS; ++i) {
        const int64_t idx = base + i;
        if (idx < n && (idx == 0 || keys[idx] != keys[idx - 1])) ++c;
    }
    int tot;
    block_exclusive_scan(c, sh, &tot);
    if (threadIdx.x == 0) tile_count[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(RS_THREADS)
uq_scan_kernel(int* __restrict__ tile_count, int ntiles, int64_t* __restrict__ count_out) {
    __shared__ int sh[RS_WARPS];
    int carry = 0;
    for (int t0 = 0; t0 < ntiles; t0 += RS_THREADS) {
        const int t = t0 + threadIdx.x;
        const int v = t < ntiles ? tile_count[t] : 0;
        int tot;
        const int ex = block_exclusive_scan(v, sh, &tot);
        if (t < ntiles) tile_count[t] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) *count_out = carry;
}

__global__ void __launch_bounds__(RS_THREADS)
uq_compact_kernel(const uint64_t* __restrict__ keys, int64_t n, const int* __restrict__ tile_base, uint64_t* __restrict__ out) {
    __shared__ int sh[RS_WARPS];
    const int64_t base = (int64_t)blockIdx.x * RS_TILE + (int64_t)threadIdx.x * RS_ITEMS;
    uint64_t k[RS_ITEMS];
    bool head[RS_ITEMS];
    int c = 0;
    for (int i = 0; i < RS_ITEMS; ++i) {
        const int64_t idx = base + i;
        k[i] = idx < n ? keys[idx] : 0;
        head[i] = idx < n && (idx == 0 || k[i] != keys[idx - 1]);
        c += head[i] ? 1 : 0;
    }
    int tot;
    int pos = tile_base[blockIdx.x] + block_exclusive_scan(c, sh, &tot);
    for (int i = 0; i < RS_ITEMS; ++i)
        if (head[i]) out[pos++] = k[i];
}

struct SuWs { size_t buf, hist, total, tiles, bytes; int ntiles; };
SuWs make_su_ws(int64_t n) {
    SuWs w;
    w.ntiles = (int)ceil_div(n, RS_TILE);
    size_t o = 0;
    auto take = [&](size_t b) { size_t r = o; o += (b + 255) & ~(size_t)255; return r; };
    w.buf = take((size_t)n * 8);
    w.hist = take((size_t)256 * w.ntiles * 4);
    w.total = take(256 * 4);
    w.tiles = take((size_t)w.ntiles * 4);
    w.bytes = o;
    return w;
}

}  // namespace
}  // namespace gnbv

using namespace gnbv;

extern "C" size_t gnbv_sort_unique_workspace_bytes(int64_t n) { return n > 0 ? make_su_ws(n).bytes : 0; }

extern "C" int gnbv_sort_unique_u64(uint64_t* keys, int64_t n, int key_bits, uint64_t* unique_out, int64_t* count_out,
                                    void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GNBV_REQUIRE(keys && unique_out && count_out && workspace && n > 0 && n < (int64_t)1 << 31 && key_bits > 0 && key_bits <= 64,
                 "gnbv_sort_unique_u64: bad arguments");
    const SuWs w = make_su_ws(n);
    GNBV_REQUIRE(workspace_bytes >= w.bytes && ((uintptr_t)workspace & 255) == 0, "gnbv_sort_unique_u64: workspace %zu B < %zu B",
                 workspace_bytes, w.bytes);
    uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
    uint64_t* bufs[2] = {keys, reinterpret_cast<uint64_t*>(ws + w.buf)};
    int* hist = reinterpret_cast<int*>(ws + w.hist);
    int* total = reinterpret_cast<int*>(ws + w.total);
    int* tiles = reinterpret_cast<int*>(ws + w.tiles);
    const int passes = (key_bits + 7) / 8;
    int cur = 0;
    for (int p = 0; p < passes; ++p, cur ^= 1) {
        rs_hist_kernel<<<w.ntiles, RS_THREADS, 0, stream>>>(bufs[cur], n, 8 * p, hist, w.ntiles);
        rs_scan_kernel<<<256, RS_THREADS, 0, stream>>>(hist, w.ntiles, total);
        rs_scatter_kernel<<<w.ntiles, RS_THREADS, 0, stream>>>(bufs[cur], bufs[cur ^ 1], n, 8 * p, hist, w.ntiles, total);
    }
    GNBV_LAUNCH_CHECK("radix sort pass");
    const uint64_t* sorted = bufs[cur];                                   // (an odd pass count leaves the result in the workspace)
    uq_count_kernel<<<w.ntiles, RS_THREADS, 0, stream>>>(sorted, n, tiles);
    uq_scan_kernel<<<1, RS_THREADS, 0, stream>>>(tiles, w.ntiles, count_out);
    uq_compact_kernel<<<w.ntiles, RS_THREADS, 0, stream>>>(sorted, n, tiles, unique_out);
    GNBV_LAUNCH_CHECK("unique compaction");
    return GNBV_OK;
}
