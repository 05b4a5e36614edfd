// lof_pack.cuh -- per-launch constants of the storage codec (lof_pack.cu) and its launchers.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "lof_core.cuh"
#include "lof_mx_kernels.cuh"

namespace lof {

// FloatingPointParams + ProjSpec folded once per call (lo_pack.hpp:121-145 evaluates these per element)
struct CodecPlan {
    int32_t n, mant, bias;          // bitwidth, mantissa bits, exponent bias
    int32_t to_signed;
    int32_t exp_offset;             // kExponentOffset = bias - 127
    int32_t digit_drop;             // -kDigitShift = 23 - mant  (> 0: formats of <= 16 bits)
    int32_t to_min_exp;             // 1 - bias
    int32_t narrower;               // pair(toMaxExp, toDigits) < pair(128, 24)
    int32_t has_inf;                // OV_behavior != Saturating
    int32_t sign_is_nan;            // IsNaN(sign bit only): the format has no -0
    int32_t len;                    // stoch_length
    int32_t gc, reverse;            // group count, PackEndian::Reverse
    uint32_t signmask, rep_mask;
    uint32_t max_finite_rep;
    uint32_t inf_in_rep;            // what +-inf encodes to   (lo_pack.hpp:165-169)
    uint32_t ovf_rep;               // what an overflow encodes to (:261-264)
    uint32_t neg_unsigned_rep;      // negative input, unsigned format (:156-161)
    uint32_t qnan;
    uint32_t inf_mask, inf_value, nan_mask, nan_value, nan_mant_mask;
    uint32_t rev_xor;               // slot <-> element map inside a group: index ^ rev_xor (group counts are powers of two)
    uint32_t recip_n;               // ceil(2^32 / n) for n > 1: floor(x / n) == __umulhi(x, recip_n) for x < 2^16
    // magic-add fast path (RNE, toward zero, away, round-to-odd on formats whose whole range sits inside fp32's normal
    // range): rep = ((max(e, emin) - emin) << m) + ((|x| (+) c) - c), c = 2^(max(e, emin) + 23 - m)   (lof_pack.cu)
    uint32_t fast_ok;
    uint32_t f_cadd, f_cfloor;      // (23 - m) << 23, (emin_biased + 23 - m) << 23
    uint32_t f_big;                 // |x| bits >= f_big: certain overflow, infinity or NaN
    uint32_t dec_ok, d_special;     // arithmetic decode eligible; every infinity / NaN code has magnitude >= d_special (or is the sign-bit-only pattern)
    uint32_t lanes;                 // pack.fast = 2: register-lane kernels (bit width / sign policy as template parameters) where eligible
};

cudaError_t launch_encode(int mode, const float* src, uint8_t* dst, int64_t count, uint64_t packed_bytes, const CodecPlan& p,
                          const RngArgs& rng, uint64_t first_index, int num_sms, cudaStream_t stream);
cudaError_t launch_round_strict(int mode, const float* src, float* dst, int64_t count, const CodecPlan& p, const RngArgs& rng,
                                uint64_t first_index, int num_sms, cudaStream_t stream);
cudaError_t launch_decode(const uint8_t* src, float* dst, int64_t count, uint64_t packed_bytes, const CodecPlan& p, int num_sms,
                          cudaStream_t stream);

cudaError_t launch_mx_encode(int mode, const float* in, int64_t n, const MxArgs& a, const CodecPlan& scale_codec, const CodecPlan& elem_codec,
                             const RngArgs& rng, const RngArgs& rng_scale, uint64_t first_index, uint8_t* scales, uint8_t* elems, int num_sms,
                             bool fast, const MagicPlan& scale_magic, const FastPlan& scale_fast, cudaStream_t stream);
cudaError_t launch_mx_decode(const uint8_t* scales, const uint8_t* elems, float* out, int64_t n, int block_size, const CodecPlan& scale_codec,
                             const CodecPlan& elem_codec, int num_sms, cudaStream_t stream);

}  // namespace lof
