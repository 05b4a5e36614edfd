// lof_gemm_common.cuh -- device helpers shared by the exact (SIMT) and chunked (tcgen05) GEMM kernels.
#pragma once

#include "lof_core.cuh"

namespace lof {

// internal flag of launch_gemm_tc (not part of the public flag set): A / B are 8-bit representations, B is K-major
constexpr uint32_t LOF_GEMM_INTERNAL_PACKED_OPERANDS = 0x80000000u;

// round-to-odd fused multiply-add, bit-identical to the host routine
// lof_round_to_odd_fma (blis_lof_gemm.hpp:63-99): RN result when it is exact or
// non-finite, otherwise the neighbour (of the exact value) with an odd mantissa LSB.
__device__ __forceinline__ float rto_fma(float a, float b, float c)
{
    const float rd = __fmaf_rd(a, b, c);
    const float ru = __fmaf_ru(a, b, c);
    const uint32_t urd = __float_as_uint(rd), uru = __float_as_uint(ru);
    // any directed result with an all-ones exponent: overflow / inf / NaN operands -> take RN
    if (((urd & 0x7F800000u) == 0x7F800000u) || ((uru & 0x7F800000u) == 0x7F800000u)) {
        const float rn = __fmaf_rn(a, b, c);
        if ((__float_as_uint(rn) & 0x7F800000u) == 0x7F800000u) return rn;   // !isfinite(rn)
        // RN stayed finite (exact value between FLT_MAX and the overflow threshold): inexact,
        // the toward-zero neighbour is +-FLT_MAX whose LSB is already odd.
        return rn;
    }
    if (rd == ru) return ru;                         // exact; an exact zero (rd = -0, ru = +0) takes RN's sign = ru
    const float rz = (rd >= 0.0f) ? rd : ru;         // sign(rd) == sign(exact) when inexact
    return __uint_as_float(__float_as_uint(rz) | 1u);
}

// Round-to-odd ADD for the chunked GEMM:  RZ(a + b) with the LSB OR-ed with "inexact".  Branch-free: the
// toward-zero result plus a comparison of the two directed sums (an exact zero gives rd = -0, ru = +0, which
// compare equal, and RZ returns +0 like RN).  A finite sum beyond FLT_MAX gives +-FLT_MAX (already odd);
// infinities and NaNs propagate.  CPU statement: oracle/lof_oracle.c lofo_rto_add.
__device__ __forceinline__ float rto_add(float a, float b)
{
    const float rz = __fadd_rz(a, b);
    const float rd = __fadd_rd(a, b);
    const float ru = __fadd_ru(a, b);
    return __uint_as_float(__float_as_uint(rz) | (rd != ru ? 1u : 0u));
}

}  // namespace lof
