// mma.cuh -- warp-level tensor-core helpers shared by the conv kernels: mma.sync m16n8k8 (TF32 operands, fp32 accumulate)
// and the split-precision ("3xTF32") operand split.
//
// Fragment layouts (PTX ISA, mma.m16n8k8 .tf32), g = lane >> 2, t = lane & 3:
//   A (16x8, row):  a0 = (row g, k t)  a1 = (row g+8, k t)  a2 = (row g, k t+4)  a3 = (row g+8, k t+4)
//   B (8x8, col):   b0 = (k t, n g)    b1 = (k t+4, n g)
//   C/D (16x8):     c0 = (row g, n 2t) c1 = (row g, n 2t+1) c2 = (row g+8, n 2t) c3 = (row g+8, n 2t+1)
// The k order inside a k-step is free as long as A and B agree, which the kernels use to make their loads contiguous.
#pragma once
#include <stdint.h>

namespace gnbv {

__device__ __forceinline__ uint32_t to_tf32(float x) {           // round-to-nearest TF32 (expands to ~4 instructions on sm_100a)
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ void mma_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// Cheap split for values produced in registers: hi = v rounded to TF32 (add half an ulp of the 13 dropped bits, mask),
// lo = v - hi (exact in fp32), handed over as is -- the tensor core reads only its upper 19 bits, an error below 2^-21 |v|.
// Valid for finite |v| far from FLT_MAX (the unchecked add could overflow the exponent there).
__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
    hi = (__float_as_uint(v) + 0x1000u) & 0xffffe000u;
    lo = __float_as_uint(v - __uint_as_float(hi));
}

}  // namespace gnbv
