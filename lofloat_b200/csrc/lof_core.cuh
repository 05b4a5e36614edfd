// lof_core.cuh -- device rounding core of the B200-native LoFloat hot path.
//
// Integer formulation of the reference's scalar rounding (behavioural spec: SURVEY.md
// Appendix A; reference: /root/reference/src/lo_float.h:570-1272 RoundMantissa + helpers,
// :2715-2756 mantissa-only virtual_round, :2843-2939 format-params virtual_round).
// Everything the reference recomputes per element (two std::pow calls, ToMax/ToMin
// exponents, the ~100-byte parameter struct) is folded once on the host into a 48-byte
// RoundPlan that lives in the kernel parameter bank.
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "../../include/lofloat_b200.h"

namespace lof {

// ---------------------------------------------------------------------------------------
// Plan: per-(format, source type) constants, computed on the host (lof_plan.cpp logic in
// lof_api.cu::make_plan).
// ---------------------------------------------------------------------------------------
struct RoundPlan {
    float    min_val;       // (float)(2^ToMin_exp * 2^-m)            lo_float.h:2861
    float    half_val;      // min_val * 0.5f                         lo_float.h:2894
    uint64_t min_mag_bits;  // container bits of from_double(min_val) lo_float.h:2905
    int32_t  min_exp_b;     // ToMin_exp + container bias  (compare with the raw exponent field)
    int32_t  max_exp_b;     // ToMax_exp + container bias
    int32_t  shiftN;        // container mantissa bits - target mantissa bits (= -kDigitShift)
    int32_t  stoch_len;
    int32_t  saturating;    // OV_behavior == Saturating
    int32_t  pad_;
};

struct RngArgs {
    uint32_t key0, key1;    // seed
    uint32_t off0, off1;    // counter offset (already xor-ed with the stream constant)
    uint32_t cls;           // event class: how many random bits one rounding event consumes (rand_class)
    uint32_t pad_;
};

// Random bits per rounding event -> packing class of the event stream (include/lofloat_b200.h, lof_rng_t):
//   class 0 (<= 8 bits):  element i uses byte (i & 3) of the 32-bit word that belongs to its 4-element group
//                         v = i >> 2:  word (v >> 8) & 3 of Philox block ((v >> 10) << 8) | (v & 255)
//                         -- inside every aligned run of 4096 elements, the 4 groups v, v + 256, v + 512, v + 768
//                         share one Philox block, which is exactly what one thread of the vector kernel processes;
//   class 1 (<= 16 bits): element i uses half (i & 1) of dense word i >> 1;
//   class 2 (> 16 bits):  element i uses dense word i.
// Dense word q = word (q & 3) of Philox block (q >> 2).
__host__ __device__ inline int rand_bits_needed(int mode, int len)
{
    if (mode == LOF_True_StochasticRounding) return 32;
    if (mode == LOF_ProbabilisticRounding) return 1;
    if (mode == LOF_StochasticRoundingD) return len + 1;
    return len;            // SR-A / SR-B / SR-C (deterministic modes never draw)
}
__host__ __device__ inline uint32_t rand_class(int mode, int len)
{
    const int b = rand_bits_needed(mode, len);
    return b <= 8 ? 0u : b <= 16 ? 1u : 2u;
}

constexpr uint32_t kStreamElem  = 0u;
constexpr uint32_t kStreamScale = 0x5ca1e000u;
constexpr uint32_t kStreamGemm  = 0x6e3a0000u;

// ---------------------------------------------------------------------------------------
// Philox4x32-10, counter = (block index lo/hi, offset lo/hi), key = seed.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int round = 0; round < 10; ++round) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;      // one IMAD.WIDE.U32 each
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c0 = n0; c1 = (uint32_t)p1; c2 = n2; c3 = (uint32_t)p0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ uint4 philox_block(const RngArgs& r, uint64_t block)
{
    return philox4x32_10((uint32_t)block, (uint32_t)(block >> 32), r.off0, r.off1, r.key0, r.key1);
}

// one word for one global index (used by scalar / tail paths)
__device__ __forceinline__ uint32_t philox_word(const RngArgs& r, uint64_t index)
{
    const uint4 b = philox_block(r, index >> 2);
    const uint32_t lane = (uint32_t)index & 3u;
    return lane == 0 ? b.x : lane == 1 ? b.y : lane == 2 ? b.z : b.w;
}

// the four per-element event words of the 4-element group v (elements 4v .. 4v + 3)
__device__ __forceinline__ uint4 philox_group4(const RngArgs& r, uint64_t v)
{
    if (r.cls == 2u) return philox_block(r, v);
    if (r.cls == 1u) {
        const uint4 b = philox_block(r, v >> 1);
        const uint32_t w0 = (v & 1u) ? b.z : b.x, w1 = (v & 1u) ? b.w : b.y;
        return make_uint4(w0 & 0xFFFFu, w0 >> 16, w1 & 0xFFFFu, w1 >> 16);
    }
    const uint4 b = philox_block(r, ((v >> 10) << 8) | (v & 255u));
    const uint32_t u = (uint32_t)(v >> 8) & 3u;
    const uint32_t w = u == 0 ? b.x : u == 1 ? b.y : u == 2 ? b.z : b.w;
    return make_uint4(w & 0xFFu, (w >> 8) & 0xFFu, (w >> 16) & 0xFFu, w >> 24);
}
__device__ __forceinline__ uint4 bytes_of(uint32_t w) { return make_uint4(w & 0xFFu, (w >> 8) & 0xFFu, (w >> 16) & 0xFFu, w >> 24); }

// event word of one global element index (scalar / tail paths)
__device__ __forceinline__ uint32_t philox_event(const RngArgs& r, uint64_t index)
{
    const uint4 g = philox_group4(r, index >> 2);
    const uint32_t e = (uint32_t)index & 3u;
    return e == 0 ? g.x : e == 1 ? g.y : e == 2 ? g.z : g.w;
}

template <int MODE>
struct ModeTraits {
    static constexpr bool kUsesRng =
        MODE == LOF_StochasticRoundingA || MODE == LOF_StochasticRoundingB || MODE == LOF_StochasticRoundingC ||
        MODE == LOF_StochasticRoundingD || MODE == LOF_True_StochasticRounding || MODE == LOF_ProbabilisticRounding;
};

__host__ __device__ inline bool mode_uses_rng(int mode)
{
    return mode == LOF_StochasticRoundingA || mode == LOF_StochasticRoundingB || mode == LOF_StochasticRoundingC ||
           mode == LOF_StochasticRoundingD || mode == LOF_True_StochasticRounding || mode == LOF_ProbabilisticRounding;
}

// ---------------------------------------------------------------------------------------
// RoundMantissa: add the mode-dependent increment so that masking off the low r bits
// yields the rounded magnitude (carry may ripple into the exponent -- intended).
// B is the container integer type (uint16_t / uint32_t / uint64_t); arithmetic is done in
// a type at least as wide and truncated back, which matches the reference's `Bits`.
// ---------------------------------------------------------------------------------------
template <typename B> struct Wide { using type = uint32_t; };
template <> struct Wide<uint64_t> { using type = uint64_t; };

template <int MODE, typename B>
__device__ __forceinline__ B round_bits(B bits_in, int r, int len, bool pos, uint32_t w)
{
    using W = typename Wide<B>::type;
    const W bits = bits_in;
    const W one = 1;
    W out;
    if constexpr (MODE == LOF_RoundToNearestEven) {                  // lo_float.h:570-595
        out = (r == 0) ? bits : bits + ((bits >> r) & one) + ((one << (r - 1)) - one);
    } else if constexpr (MODE == LOF_RoundToNearestOdd) {            // :1030-1055
        out = (r == 0) ? bits : bits + ((~bits >> r) & one) + ((one << (r - 1)) - one);
    } else if constexpr (MODE == LOF_RoundTiesTowardsZero) {         // :974-996
        out = (r == 0) ? bits : bits + ((one << (r - 1)) - one);
    } else if constexpr (MODE == LOF_RoundTowardsZero) {             // :946-963
        out = bits;                                                   // caller masks
    } else if constexpr (MODE == LOF_RoundAwayFromZero) {            // :1000-1026
        const W low = (one << r) - one;
        out = (bits & ~low) + ((bits & low) != 0 ? (one << r) : W(0));
    } else if constexpr (MODE == LOF_RoundUp) {                      // :1061-1098
        const W low = (one << r) - one;
        out = (bits & ~low) + ((pos && (bits & low) != 0) ? (one << r) : W(0));
    } else if constexpr (MODE == LOF_RoundDown) {                    // :1121-1154
        const W low = (one << r) - one;
        out = (bits & ~low) + ((!pos && (bits & low) != 0) ? (one << r) : W(0));
    } else if constexpr (MODE == LOF_RoundTiesToAway) {              // :1176-1199
        const W low = (one << r) - one;
        out = (bits & ~low) + (((bits >> (r - 1)) & one) << r);
    } else if constexpr (MODE == LOF_RoundToOdd) {                   // :1203-1229
        const W low = (one << r) - one;
        out = (bits & ~low) | ((bits & low) != 0 ? (one << r) : W(0));
    } else if constexpr (MODE == LOF_StochasticRoundingA) {          // :653-694
        out = (len <= 0) ? bits : bits + (W(w & (uint32_t)((1ull << len) - 1ull)) << (r - len));
    } else if constexpr (MODE == LOF_StochasticRoundingB) {          // :738-753
        const W samp = w & (uint32_t)((1ull << len) - 1ull);
        const W to_add = (samp & ((one << len) - one)) + one;
        const W lower = bits & ((one << r) - one);
        out = bits + (lower > 0 ? (to_add << (r - len)) : W(0));
    } else if constexpr (MODE == LOF_StochasticRoundingC) {          // :798-812
        const uint32_t samp = w & (uint32_t)((1ull << len) - 1ull);
        const uint32_t top = ((samp >> 1) + (samp & 1u)) << (r - len + 1);   // shifted as unsigned int
        const W bottom = bits & ((one << r) - one);
        out = (len == 0 || bottom == 0) ? bits : bits + W(top);
    } else if constexpr (MODE == LOF_StochasticRoundingD) {          // :820-860
        const W to_add = W(w & (uint32_t)((1ull << len) - 1ull)) + W((w >> len) & 1u);
        out = (len <= 0) ? bits : bits + (to_add << (r - len));
    } else if constexpr (MODE == LOF_True_StochasticRounding) {      // :916-942, U = w / 2^32
        const W low = (one << r) - one;
        const uint64_t tail = bits & low;
        // w/2^32 < tail/2^r  <=>  w * 2^r < tail * 2^32
        const bool up = (r <= 32) ? ((uint64_t)w < (tail << (32 - r))) : (((uint64_t)w << (r - 32)) < tail);
        out = (bits & ~low) + (up ? (one << r) : W(0));
    } else if constexpr (MODE == LOF_ProbabilisticRounding) {        // :635-648
        const W low = (one << r) - one;
        out = (bits & ~low) + (((bits & low) != 0 ? W(w & 1u) : W(0)) << r);
    } else {
        out = bits;
    }
    return (B)out;
}

// ---------------------------------------------------------------------------------------
// Source container traits (reference Traits<From>, lo_float.h:2623-2668; bf16 is new).
// ---------------------------------------------------------------------------------------
template <int DTYPE> struct Src;

template <> struct Src<LOF_F32> {
    using bits_t = uint32_t;
    using scale_t = float;
    static constexpr int kBits = 32, kMant = 23, kBias = 127;
    static constexpr bits_t kQNaN = 0x7FC00000u, kInf = 0x7F800000u, kAbsMask = 0x7FFFFFFFu;
    __device__ static __forceinline__ float to_float(bits_t a) { return __uint_as_float(a); }
    __device__ static __forceinline__ bits_t k_times_min(bits_t k, const RoundPlan& p)
    { return __float_as_uint(__fmul_rn(__uint2float_rn(k), p.min_val)); }
    __host__ static scale_t make_scale(double s) { return (float)s; }
    __device__ static __forceinline__ bool scale_is_one(scale_t s) { return s == 1.0f; }
    // NaN operands: the CPU reference (x86 mulss) returns the operand NaN quieted, payload and sign
    // kept, while FMUL returns the canonical NaN; the mantissa-only overload rounds those bits.
    __device__ static __forceinline__ bits_t apply_scale(bits_t v, scale_t s)
    { return ((v & kAbsMask) > kInf) ? (v | 0x00400000u) : __float_as_uint(__fmul_rn(s, __uint_as_float(v))); }
};

template <> struct Src<LOF_F64> {
    using bits_t = uint64_t;
    using scale_t = double;
    static constexpr int kBits = 64, kMant = 52, kBias = 1023;
    static constexpr bits_t kQNaN = 0x7FF8000000000000ull, kInf = 0x7FF0000000000000ull,
                            kAbsMask = 0x7FFFFFFFFFFFFFFFull;
    __device__ static __forceinline__ float to_float(bits_t a) { return __double2float_rn(__longlong_as_double((long long)a)); }
    __device__ static __forceinline__ bits_t k_times_min(bits_t k, const RoundPlan& p)
    { return (bits_t)__double_as_longlong(__dmul_rn(__ull2double_rn(k), (double)p.min_val)); }
    __host__ static scale_t make_scale(double s) { return s; }
    __device__ static __forceinline__ bool scale_is_one(scale_t s) { return s == 1.0; }
    __device__ static __forceinline__ bits_t apply_scale(bits_t v, scale_t s)
    {
        if ((v & kAbsMask) > kInf) return v | 0x0008000000000000ull;     // quieted operand NaN (x86 mulsd)
        return (bits_t)__double_as_longlong(__dmul_rn(s, __longlong_as_double((long long)v)));
    }
};

template <> struct Src<LOF_F16> {
    using bits_t = uint16_t;
    using scale_t = float;   // value of static_cast<c10::Half>(scale), widened
    static constexpr int kBits = 16, kMant = 10, kBias = 15;
    static constexpr bits_t kQNaN = 0x7E00u, kInf = 0x7C00u, kAbsMask = 0x7FFFu;
    __device__ static __forceinline__ float to_float(bits_t a) { return __half2float(__ushort_as_half(a)); }
    __device__ static __forceinline__ bits_t k_times_min(bits_t k, const RoundPlan& p)
    { return __half_as_ushort(__float2half_rn(__fmul_rn(__uint2float_rn((uint32_t)k), p.min_val))); }
    __host__ static scale_t make_scale(double s) { return __half2float(__float2half_rn((float)s)); }
    __device__ static __forceinline__ bool scale_is_one(scale_t s) { return s == 1.0f; }
    __device__ static __forceinline__ bits_t apply_scale(bits_t v, scale_t s)   // Half*Half: float multiply, RN to half
    {
        if ((v & kAbsMask) > kInf) return (bits_t)((v & 0x8000u) | 0x7E00u);  // c10 float->half: NaN -> sign | 0x7E00
        return __half_as_ushort(__float2half_rn(__fmul_rn(s, __half2float(__ushort_as_half(v)))));
    }
};

template <> struct Src<LOF_BF16> {
    using bits_t = uint16_t;
    using scale_t = float;
    static constexpr int kBits = 16, kMant = 7, kBias = 127;
    static constexpr bits_t kQNaN = 0x7FC0u, kInf = 0x7F80u, kAbsMask = 0x7FFFu;
    __device__ static __forceinline__ float to_float(bits_t a) { return __uint_as_float((uint32_t)a << 16); }
    __device__ static __forceinline__ bits_t k_times_min(bits_t k, const RoundPlan& p)
    { return __bfloat16_as_ushort(__float2bfloat16_rn(__fmul_rn(__uint2float_rn((uint32_t)k), p.min_val))); }
    __host__ static scale_t make_scale(double s) { return __bfloat162float(__float2bfloat16_rn((float)s)); }
    __device__ static __forceinline__ bool scale_is_one(scale_t s) { return s == 1.0f; }
    __device__ static __forceinline__ bits_t apply_scale(bits_t v, scale_t s)
    {
        if ((v & kAbsMask) > kInf) return (bits_t)(v | 0x0040u);             // quieted operand NaN
        return __bfloat16_as_ushort(__float2bfloat16_rn(__fmul_rn(s, __uint_as_float((uint32_t)v << 16))));
    }
};

// ---------------------------------------------------------------------------------------
// virtual_round(value, mantissa_bits, ps)                       lo_float.h:2715-2756
// r = container mantissa bits - requested mantissa bits; r <= 0 is the identity.
// ---------------------------------------------------------------------------------------
template <int DTYPE, int MODE>
__device__ __forceinline__ typename Src<DTYPE>::bits_t
round_mantissa_elem(typename Src<DTYPE>::bits_t v, int r, int len, uint32_t w)
{
    using S = Src<DTYPE>;
    using B = typename S::bits_t;
    if (r <= 0) return v;
    const B sign = v & (B)~S::kAbsMask;
    B a = v & S::kAbsMask;
    a = round_bits<MODE, B>(a, r, len, sign == 0, w);
    a &= (B) ~(((B)1 << r) - (B)1);
    return a | sign;
}

// ---------------------------------------------------------------------------------------
// virtual_round(value, FloatingPointParams, ps)                 lo_float.h:2843-2939
// ---------------------------------------------------------------------------------------
template <int DTYPE, int MODE>
__device__ __forceinline__ typename Src<DTYPE>::bits_t
round_params_elem(typename Src<DTYPE>::bits_t v, const RoundPlan& p, uint32_t w)
{
    using S = Src<DTYPE>;
    using B = typename S::bits_t;
    const B sign = v & (B)~S::kAbsMask;
    B a = v & S::kAbsMask;

    if (a > S::kInf) return S::kQNaN;                                 // :2869  NaN -> +qNaN

    const float av = S::to_float(a);
    if (av < p.min_val) {                                             // :2872-2908  below denorm_min
        const bool pos = (sign == 0);
        bool to_min;
        if constexpr (MODE == LOF_RoundTowardsZero) to_min = false;
        else if constexpr (MODE == LOF_RoundToOdd || MODE == LOF_RoundAwayFromZero) to_min = av > 0.0f;
        else if constexpr (MODE == LOF_RoundUp) to_min = pos && av > 0.0f;
        else if constexpr (MODE == LOF_RoundDown) to_min = !pos && av > 0.0f;
        else if constexpr (MODE == LOF_RoundToNearestEven || MODE == LOF_RoundTiesTowardsZero)
            to_min = av > p.half_val;
        else if constexpr (MODE == LOF_RoundToNearestOdd || MODE == LOF_RoundTiesToAway)
            to_min = av >= p.half_val;
        else to_min = false;                                          // stochastic: flush to 0
        const B mag = to_min ? (B)p.min_mag_bits : (B)0;
        return mag | sign;
    }

    int field = (int)(a >> S::kMant);
    if (field < p.min_exp_b) {                                        // :2915-2923  target-subnormal
        const int shift = p.shiftN + (p.min_exp_b - field);
        if (shift <= 0) return v;   // container has fewer mantissa bits than the target: exact
        B sig = (a & (((B)1 << S::kMant) - (B)1)) | ((B)1 << S::kMant);
        sig = round_bits<MODE, B>(sig, shift, p.stoch_len, sign == 0, w);
        const B k = (B)(sig >> shift);
        return S::k_times_min(k, p) | sign;
    }

    if (p.shiftN > 0) {                                               // :2925-2933
        a = round_bits<MODE, B>(a, p.shiftN, p.stoch_len, sign == 0, w);
        a &= (B) ~(((B)1 << p.shiftN) - (B)1);
    }
    field = (int)(a >> S::kMant);                                     // :2935-2936
    if (field > p.max_exp_b) return p.saturating ? v : S::kInf;
    return a | sign;
}

// runtime-mode dispatch (used where the call is rare: MX scale rounding)
template <int DTYPE>
__device__ __forceinline__ typename Src<DTYPE>::bits_t
round_params_elem_rt(int mode, typename Src<DTYPE>::bits_t v, const RoundPlan& p, uint32_t w)
{
    switch (mode) {
#define LOF_CASE(M) case M: return round_params_elem<DTYPE, M>(v, p, w);
        LOF_CASE(LOF_RoundToNearestEven) LOF_CASE(LOF_RoundTowardsZero) LOF_CASE(LOF_RoundAwayFromZero)
        LOF_CASE(LOF_StochasticRoundingA) LOF_CASE(LOF_RoundToNearestOdd) LOF_CASE(LOF_RoundDown)
        LOF_CASE(LOF_RoundUp) LOF_CASE(LOF_RoundTiesToAway) LOF_CASE(LOF_StochasticRoundingB)
        LOF_CASE(LOF_StochasticRoundingC) LOF_CASE(LOF_True_StochasticRounding)
        LOF_CASE(LOF_ProbabilisticRounding) LOF_CASE(LOF_RoundToOdd) LOF_CASE(LOF_RoundTiesTowardsZero)
        LOF_CASE(LOF_StochasticRoundingD)
#undef LOF_CASE
        default: return v;
    }
}

}  // namespace lof
