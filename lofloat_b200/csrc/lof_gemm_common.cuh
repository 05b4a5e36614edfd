// lof_gemm_common.cuh -- device helpers shared by the exact (SIMT) and chunked (tcgen05) GEMM kernels.
#pragma once

#include "lof_core.cuh"

namespace lof {

// internal flag of launch_gemm_tc (not part of the public flag set): A / B are 8-bit representations, B is K-major
constexpr uint32_t LOF_GEMM_INTERNAL_PACKED_OPERANDS = 0x80000000u;
// internal: the packed fold kernel issues two N = 128 MMAs per k-step into four TMEM stages (option gemm.fold_split = 2)
constexpr uint32_t LOF_GEMM_INTERNAL_FOLD_SPLIT2 = 0x40000000u;
// internal: the MX GEMM records CTA 0's cycle breakdown (option gemm.mx_prof, read back with lof_debug_mx_prof)
constexpr uint32_t LOF_GEMM_INTERNAL_PROF = 0x20000000u;

// Operand addressing of the exact GEMM beyond plain row-major: the callers fused around lof_gemm.
//   b_kind 0  B is K x N row-major (lof_gemm_f32 / lof_gemm_batched_f32)
//   b_kind 1  B is N x K row-major -- an nn.Linear weight used as stored, replacing `weight.t().contiguous()` of
//             _lof_linear (LoFloat/layers.py:152-168)
//   b_kind 2  implicit im2col: B(k, n) is gathered from an NCHW image, A is the OIHW filter bank; grid.z = image x group.
//             k_order 0 = (c, r, s) ascending, the F.unfold order of LoF_Conv2d's im2col backend (layers.py:710-745);
//             k_order 1 = (r, s, c), the NHWC / KRSC order of the implicit-GEMM backend (Lof_kernel.cu:156-228).
// `bias` (already quantised) is added to alpha * acc in plain fp32 like `out + b_q` (layers.py:392-398, 747-751):
// bias_axis 0 = one value per row of C (conv: output channel), 1 = per column (linear: output feature).
struct ExactView {
    int b_kind = 0, k_order = 0, bias_axis = 0;
    const float* bias = nullptr;
    int C = 0, H = 0, W = 0, Cg = 0, R = 0, S = 0, Q = 0, ph = 0, pw = 0, sh = 1, sw = 1, dh = 1, dw = 1, groups = 1;
    int swap_xy = 0;        // M tiles on grid.x (more than 65535 of them: flattened token dimensions), N tiles on grid.y
};

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

// Round-to-odd ADD for the chunked GEMM:  RZ(a + b) with the LSB OR-ed with "inexact".  Branch-free: of the two directed
// sums rd <= exact <= ru take the one whose LSB is odd -- inexact: they are adjacent floats, exactly one is odd, and it is
// RZ | 1; exact: they are equal (an exact zero gives rd = -0, even, and takes ru = +0 like RZ and RN do).  A finite sum beyond
// FLT_MAX gives +-FLT_MAX (odd); infinities and NaNs propagate.  CPU statement: oracle/lof_oracle.c lofo_rto_add.
__device__ __forceinline__ float rto_add(float a, float b)
{
    const float rd = __fadd_rd(a, b);
    const float ru = __fadd_ru(a, b);
    return (__float_as_uint(rd) & 1u) ? rd : ru;
}


// Mantissa rounding in the FLOAT domain (the chunked GEMM's per-chunk fold and the exact GEMM's optimistic chain): with
// m = sign and exponent of t (+-2^e, one LOP3) and K = m * 2^r (an exact product),
//     s = fma(m, 2^r, t) = round(t + K)      -- |t + K| stays inside the binade [2^(e+r), 2^(e+r+1)), whose ulp is 2^(e+r-23):
//     y = fma(m, -2^r, s) = s - K   (exact)     the hardware rounds t on the grid of a (23 - r)-bit mantissa
// The rounding MODE is the first FMA's: .rn = RoundToNearestEven (K / ulp = 2^23 is even, so ties go to even y), .rz =
// RoundTowardsZero (K has t's sign), .rm / .rp = RoundDown / RoundUp.  Equal to the integer routine (round_bits + mask,
// lo_float.h:2715-2756) for every NORMAL t with e + r < 127 and 1 <= r <= 22, checked on all 2^32 patterns x 22 widths x 4 modes
// on the device (lof_selftest_gemm_rne_round); t = +-0 gives +-0; subnormal t is NOT rounded (K = 0) and larger t / inf /
// NaN give a non-finite y -- callers guarantee the range (operands in an 8-bit format: sums are multiples of 2^-32 below 2^64).
// One alu-pipe and two fma-pipe instructions instead of ~7 alu-pipe ones (the alu pipe is half rate).
template <int MODE>
__device__ __forceinline__ float round_mant_float(float t, float two_r)
{
    const float m = __uint_as_float(__float_as_uint(t) & 0xFF800000u);
    float s;
    if constexpr (MODE == LOF_RoundTowardsZero) s = __fmaf_rz(m, two_r, t);
    else if constexpr (MODE == LOF_RoundDown) s = __fmaf_rd(m, two_r, t);
    else if constexpr (MODE == LOF_RoundUp) s = __fmaf_ru(m, two_r, t);
    else s = __fmaf_rn(m, two_r, t);
    return __fmaf_rn(m, -two_r, s);
}
__host__ __device__ constexpr bool mode_has_float_round(int mode)
{
    return mode == LOF_RoundToNearestEven || mode == LOF_RoundTowardsZero || mode == LOF_RoundDown || mode == LOF_RoundUp;
}

}  // namespace lof
