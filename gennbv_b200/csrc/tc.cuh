// tc.cuh -- minimal hand-written tcgen05 / TMEM / mbarrier layer for sm_100a (inline PTX, no CUTLASS).
//
// Used by the split-precision ("3xTF32") tensor-core contractions of the encoder.  An fp32 operand x is split into
//   hi = x with the 13 low mantissa bits cleared (exactly representable in TF32),   lo = x - hi (exact in fp32),
// and a*b is accumulated as hi*hi + hi*lo + lo*hi in the fp32 TMEM accumulator: the dropped lo*lo term and the TF32
// truncation of lo are ~2^-21 relative, inside the 1e-4 parity budget that plain TF32 / BF16 operands would break.
//
// Shared-memory operand tiles use the canonical K-major, no-swizzle UMMA layout (cute: `((8,n),2):((1,SBO),LBO)` in
// 16-byte units): 8-row x 16-byte core matrices stored contiguously (128 B); core matrices adjacent along K are
// LBO = 128 B apart, 8-row groups are SBO = (BK/4)*128 B apart.  Element (r, k) of a tile with BK fp32 columns lives at
//   (r/8)*SBO + (k/4)*128 + (r%8)*16 + (k%4)*4   bytes.
// One tcgen05.mma.kind::tf32 consumes K = 8 elements = two core matrices; the descriptor start address advances by
// 256 B per K step.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gnbv {
namespace tc {

constexpr int BK = 32;                        // fp32 elements of K per shared-memory stage
constexpr uint32_t LBO_BYTES = 128;
constexpr uint32_t SBO_BYTES = (BK / 4) * 128;   // 1024
constexpr int KSTEPS = BK / 8;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint32_t tile_offset(int r, int k) {       // byte offset of element (r,k) in a K-major tile
    return (uint32_t)((r >> 3) * SBO_BYTES + (k >> 2) * 128 + (r & 7) * 16 + (k & 3) * 4);
}

__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    lo = x - hi;
}

// One lane of a converged warp (elect.sync).  Together with a PROVABLY warp-uniform role branch (warp index taken through
// __shfl_sync) this lets ptxas keep the tcgen05.mma operands in uniform registers; under a plain `if (lane == 0)` every MMA is
// wrapped in a divergence "waterfall" (ELECT / R2UR.BROADCAST / BRA.U.ANY, ~12 instructions and several R2UR latencies each).
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred;
}
__device__ __forceinline__ int uniform_warp_index() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }

// ---- mbarrier ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// Bounded wait (a stuck barrier must never hang the GPU): returns false after ~2^24 polls.
// A failed poll backs off with nanosleep: a spinning warp otherwise takes its full share of the scheduler's issue slots away
// from the warps doing the work it waits for (measured: 17 warps with ~12 of them spinning ran the working warps 4x slower).
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, uint32_t sleep_ns = 32) {
#pragma unroll 1
    for (uint32_t it = 0; it < (1u << 24); ++it) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok)
                     : "r"(bar), "r"(parity)
                     : "memory");
        if (ok) return true;
        if (sleep_ns) asm volatile("nanosleep.u32 %0;" ::"r"(sleep_ns));
    }
    return false;
}

// ---- proxies / fences ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMEM ------------------------------------------------------------------------------------------------------------
// ncols: power of two in [32, 512].  Executed by ONE full warp; the base address lands in *result_smem.
__device__ __forceinline__ void tmem_alloc(uint32_t result_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(result_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// 16 consecutive fp32 columns of this thread's TMEM lane (lane = 32*(warp%4) + laneid)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- descriptors + MMA ---------------------------------------------------------------------------------------------
// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version = 1 [46,48), base_offset 0, layout_type SWIZZLE_NONE = 0 [61,64)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) { return make_smem_desc(smem_addr, LBO_BYTES, SBO_BYTES); }

// instruction descriptor (cute::UMMA::InstrDescriptor) for kind::tf32, fp32 accumulate, K-major A and B
__device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N) {
    uint32_t d = 0;
    d |= 1u << 4;                          // c_format  = F32
    d |= 2u << 7;                          // a_format  = TF32
    d |= 2u << 10;                         // b_format  = TF32
    d |= (uint32_t)(N >> 3) << 17;         // n_dim
    d |= (uint32_t)(M >> 4) << 24;         // m_dim
    return d;
}

// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// arrive on `bar` when all previously issued MMAs of this thread have completed (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// The three products of the split-precision scheme for one BK stage: tiles are (hi, lo) pairs of A and B.
__device__ __forceinline__ void mma_stage_3xtf32(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo,
                                                 uint32_t idesc, bool first_stage) {
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks) {
        const uint32_t o = ks * 256;
        const uint64_t dah = make_smem_desc(a_hi + o), dal = make_smem_desc(a_lo + o);
        const uint64_t dbh = make_smem_desc(b_hi + o), dbl = make_smem_desc(b_lo + o);
        mma_tf32(tmem_d, dal, dbh, idesc, (first_stage && ks == 0) ? 0u : 1u);     // small terms first
        mma_tf32(tmem_d, dah, dbl, idesc, 1u);
        mma_tf32(tmem_d, dah, dbh, idesc, 1u);
    }
}

}  // namespace tc
}  // namespace gnbv
