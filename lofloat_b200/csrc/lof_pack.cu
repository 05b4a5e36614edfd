// lof_pack.cu -- low-bit storage codec for sm_100a: fp32 -> packed n-bit representations and back.
//
// Replaces the reference's one-thread-per-group kernels (/root/reference/src/lo_pack.hpp:474-529,
// encode_float_kernel / decode_float_kernel: 64-float local arrays, bit-at-a-time put_bits / get_bits, scalar
// uncoalesced loads).  Semantics are those of lo_pack.hpp:
//   convert_float_to_rep   (:109-267)  runtime twin of ConvertImpl<float, Templated_Float<Fp>>::run(): strict typed
//                                       conversion honouring Saturation_Mode, the Inf / NaN checkers, signedness;
//   convert_rep_to_float   (:277-319)
//   group packing          (:58-102)   n in [1,16]; a group of lcm(n, 8 or 16) bits, LSB-first, Forward / Reverse slots.
// HBM-bound streaming op: 4 B read + n/8 B written per element (encode), the reverse for decode.  Data path per CTA
// tile of 4096 elements: coalesced 128-bit loads -> rep per element in registers -> 16-bit reps staged in shared
// memory -> every thread assembles whole 32-bit words of the packed stream (any n, either slot order) -> coalesced
// 32-bit stores.  The bit position of slot s is s * n in BOTH slot orders (groups tile the stream exactly), so the
// packed stream is a plain LSB-first bit stream and only the slot -> element map depends on the order.
#include <cuda_runtime.h>
#include <stdint.h>

#include "lof_core.cuh"
#include "lof_mx_kernels.cuh"
#include "lof_pack.cuh"

namespace lof {

namespace {

constexpr int kPackThreads = 256;
constexpr int kPackPerThread = 16;
constexpr int kLutBits = 12;
constexpr int kPackTile = kPackThreads * kPackPerThread;     // 4096 elements: a multiple of every group count (1..16)

__device__ __forceinline__ bool codec_is_nan(const CodecPlan& p, uint32_t rep)
{
    return (rep & p.nan_mask) == p.nan_value && (p.nan_mant_mask == 0u || (rep & p.nan_mant_mask) != 0u);
}
__device__ __forceinline__ uint32_t with_sign(const CodecPlan& p, uint32_t mag, bool neg)     // lo_pack.hpp:147-152
{
    if (!neg || !p.to_signed) return mag;
    if (mag == 0u && p.sign_is_nan) return 0u;
    return mag | p.signmask;
}

// lo_pack.hpp:109-267.  Every RoundMantissa call there leaves `pos` at its default (false), lo_float.h:1232.
template <int MODE>
__device__ __forceinline__ uint32_t float_to_rep(uint32_t fb, const CodecPlan& p, uint32_t w)
{
    const bool neg = (fb >> 31) != 0;
    if (!p.to_signed && neg) return p.neg_unsigned_rep;                                        // :156-161
    const uint32_t from_bits = fb & 0x7FFFFFFFu;
    if (from_bits >= 0x7F800000u) {
        if (from_bits > 0x7F800000u) return p.qnan;                                            // :170
        return with_sign(p, p.inf_in_rep, neg);                                                // :165-169
    }
    if (from_bits == 0u) return with_sign(p, 0u, neg);
    const int biased_from = (int)(from_bits >> 23);
    const int ds = p.digit_drop;                                                               // -kDigitShift = 23 - m > 0

    if (p.to_min_exp < -125 && biased_from == 0) {                                             // branch 1, :176-196
        uint64_t bits = from_bits;
        const int nf = __clz((int)from_bits) - (32 - 23) + 1;
        const int be = p.exp_offset - nf + 1;
        if (be <= 0) {
            if (p.exp_offset < 64) bits <<= p.exp_offset;
        } else {
            bits <<= nf;
            bits &= ~(1ull << 23);
            bits |= (uint64_t)be << 23;
        }
        bits = round_bits<MODE, uint64_t>(bits, ds, p.len, false, w) >> ds;
        return with_sign(p, (uint32_t)bits, neg);
    }
    if (p.to_min_exp > -125) {                                                                 // branch 2, :199-243
        const int biased_to = biased_from - 127 + p.bias;
        if (biased_to <= 0) {
            const uint32_t lead = biased_from > 0 ? 1u : 0u;
            const int sh = ds - biased_to + (int)lead;
            const uint32_t sig = (from_bits & 0x007FFFFFu) | (lead << 23);
            uint32_t bits;
            if (sh <= 24) {
                bits = round_bits<MODE, uint32_t>(sig, sh, p.len, false, w) >> sh;
            } else if constexpr (MODE == LOF_RoundToOdd || MODE == LOF_RoundAwayFromZero) {
                bits = sig > 0u ? 1u : 0u;
            } else if constexpr (MODE == LOF_RoundUp) {
                bits = (sig != 0u && !neg) ? 1u : 0u;
            } else if constexpr (MODE == LOF_RoundDown) {
                bits = (sig != 0u && neg) ? 1u : 0u;
            } else if constexpr (ModeTraits<MODE>::kUsesRng) {
                bits = sh < 64 ? (uint32_t)(round_bits<MODE, uint64_t>((uint64_t)sig, sh, p.len, false, w) >> sh) : 0u;
            } else {
                bits = 0u;
            }
            return with_sign(p, bits, neg);
        }
    }
    // normal path, :246-266
    uint64_t rounded = round_bits<MODE, uint64_t>((uint64_t)from_bits, ds, p.len, false, w);
    rounded &= ~((1ull << ds) - 1ull);
    rounded += (uint64_t)(int64_t)p.exp_offset << 23;
    if (p.narrower && rounded > ((uint64_t)p.max_finite_rep << ds)) return with_sign(p, p.ovf_rep, neg);
    return with_sign(p, (uint32_t)(rounded >> ds), neg);
}

// Magic-add formulation of the same conversion for the IEEE-like modes (see lof_fast_f32.cuh for the idea): with
// c = 2^(max(e, emin) + 23 - m) the sum |x| (+) c lands in c's binade, whose ulp is exactly the target quantum of x
// (2^(e - m), or denorm_min below the smallest normal), so (bits(|x| (+) c) - bits(c)) is the rounded significand in
// quanta -- including a carry into the next binade -- and the exponent field is max(e, emin) - emin.  Eligibility is
// decided on the host (make_codec_plan); equality with the integer routine above is established exhaustively over
// all 2^32 inputs on the device (tests/test_gpu_parity.py::test_pack_fast_path_equals_exact_every_fp32_pattern).
// launch-uniform constants of the fast path, copied out of the parameter bank once per thread
struct FastConsts {
    uint32_t cadd, cfloor, big, max_finite, ovf, inf_in, qnan, signmask, neg_unsigned;
    int ds, sign_shift;
    bool to_signed, sign_is_nan;
    __device__ __forceinline__ explicit FastConsts(const CodecPlan& p)
        : cadd(p.f_cadd), cfloor(p.f_cfloor), big(p.f_big), max_finite(p.max_finite_rep), ovf(p.ovf_rep), inf_in(p.inf_in_rep),
          qnan(p.qnan), signmask(p.signmask), neg_unsigned(p.neg_unsigned_rep), ds(p.digit_drop), sign_shift(32 - p.n),
          to_signed(p.to_signed != 0), sign_is_nan(p.sign_is_nan != 0)
    {
        // keep the per-element constants in registers: ptxas otherwise re-loads them from the parameter bank for every element
        // (7 LDC / LDCU per conversion in the first build, ncu r01i)
        asm volatile("" : "+r"(cadd), "+r"(cfloor), "+r"(big), "+r"(max_finite), "+r"(ovf), "+r"(signmask), "+r"(ds), "+r"(sign_shift));
    }
};

// ~15 instructions per element: the stream has to stay under ~20 to be HBM- rather than ALU-bound (ncu r01g).
template <int MODE>
__device__ __forceinline__ uint32_t float_to_rep_fast(uint32_t x, const FastConsts& k)
{
    const uint32_t a = x & 0x7FFFFFFFu;
    const uint32_t c = max((a & 0x7F800000u) + k.cadd, k.cfloor);
    const float cf = __uint_as_float(c), ax = __uint_as_float(a);
    float t;
    if constexpr (MODE == LOF_RoundToNearestEven) t = __fadd_rn(ax, cf);
    else if constexpr (MODE == LOF_RoundTowardsZero) t = __fadd_rz(ax, cf);
    else if constexpr (MODE == LOF_RoundAwayFromZero) t = __fadd_ru(ax, cf);
    else {                                                    // RoundToOdd: truncate, LSB |= inexact
        const float tz = __fadd_rz(ax, cf), tu = __fadd_ru(ax, cf);
        t = __uint_as_float(__float_as_uint(tz) | (tz != tu ? 1u : 0u));
    }
    // (c - cfloor) has 23 zero low bits, so >> ds places max(e, emin) - emin directly above the m mantissa bits
    uint32_t mag = (__float_as_uint(t) + ((c - k.cfloor) >> k.ds)) - c;
    mag = mag > k.max_finite ? k.ovf : mag;
    uint32_t xs = x;                                          // carries the sign that is attached to the result
    if (a >= k.big) {                                         // certain overflow, infinity or NaN: rare
        mag = a > 0x7F800000u ? k.qnan : (a == 0x7F800000u ? k.inf_in : k.ovf);
        if (a > 0x7F800000u) xs = a;                          // a NaN result carries no sign (lo_pack.hpp:170)
    }
    if (k.to_signed) {                                        // launch-uniform
        uint32_t sbit = (xs >> k.sign_shift) & k.signmask;
        if (k.sign_is_nan) sbit = mag == 0u ? 0u : sbit;      // no -0 when the sign-bit-only pattern is the NaN (:147-152)
        return mag | sbit;
    }
    return (int32_t)x < 0 ? k.neg_unsigned : mag;             // any negative input incl. -0, -inf, -NaN (:156-161)
}

// ------------------------------------------------------------------------------------------------
// Third-generation fast encode ("register lanes"): the SASS of the kernels above spends ~30 instructions per element, two
// thirds of them on launch-uniform decisions (bit width, signedness, slot order, bounds) and on the per-element test for
// certain overflow / infinity / NaN.  Here the bit width N and the sign policy SGN are template parameters, the rare
// inputs (infinities, NaNs) are detected ONCE per thread tile on a running maximum of the exponent fields, and the sign bits of the four
// elements of a vector are attached with byte permutes on the packed word: ~8 alu-pipe + 3 fma-pipe instructions per
// element, which keeps a 4.5 .. 6 B / element stream HBM-bound.  Same arithmetic as float_to_rep_fast (equality on every
// fp32 pattern: tests/test_gpu_parity.py::test_pack_fast_path_equals_exact_every_fp32_pattern, which now runs these).
// SGN: 0 = unsigned target, 1 = signed, 2 = signed and the sign-bit-only pattern is the NaN (no -0).
struct LaneConsts {
    uint32_t e0, cfloor, clampv, dcap, neg_unsigned;
    int ds;
    __device__ __forceinline__ explicit LaneConsts(const CodecPlan& p)
        : e0(p.f_cfloor - p.f_cadd), cfloor(p.f_cfloor), clampv(p.ovf_rep), dcap(p.f_big - (p.f_cfloor - p.f_cadd) - 0x00800000u),
          neg_unsigned(p.neg_unsigned_rep), ds(p.digit_drop)
    {
        asm volatile("" : "+r"(e0), "+r"(cfloor), "+r"(clampv), "+r"(dcap), "+r"(ds));
    }
};
constexpr uint32_t kLaneRare = 0x7F800000u;       // running maximum of the exponent fields >= kLaneRare: an infinity or a NaN among the inputs

// magnitude code of x (no sign attached) and the word whose bit 31 is the sign to attach.  The exponent distance d above the
// smallest normal is clamped to the format's top binade 2^E: a finite |x| >= 2^(E + 1) (certain overflow -- common for the narrow
// formats on unscaled data) then adds at least 2^(m + 1) quanta to c = 2^(E + 23 - m), so the code comes out >= 2^(ebits + m), above
// every magnitude, and the min below turns it into ovf_rep like any other overflow (ovf_rep is max_finite_rep or max_finite_rep + 1,
// make_codec_plan).  Only infinities and NaNs (emax accumulates the exponent fields) need the second-generation routine.
template <int MODE, int SGN>
__device__ __forceinline__ uint32_t lane_mag(uint32_t x, const LaneConsts& k, uint32_t& emax, uint32_t& sx)
{
    const uint32_t ef = x & 0x7F800000u;
    emax = max(emax, ef);
    const int32_t d = min(max((int32_t)(ef - k.e0), 0), (int32_t)k.dcap);
    const uint32_t c = (uint32_t)d + k.cfloor;
    const float cf = __uint_as_float(c), ax = fabsf(__uint_as_float(x));
    float t;
    if constexpr (MODE == LOF_RoundToNearestEven) t = __fadd_rn(ax, cf);
    else if constexpr (MODE == LOF_RoundTowardsZero) t = __fadd_rz(ax, cf);
    else if constexpr (MODE == LOF_RoundAwayFromZero) t = __fadd_ru(ax, cf);
    else {                                                    // RoundToOdd: truncate, LSB |= inexact
        const float tz = __fadd_rz(ax, cf), tu = __fadd_ru(ax, cf);
        t = __uint_as_float(__float_as_uint(tz) | (tz != tu ? 1u : 0u));
    }
    uint32_t mag = min(__float_as_uint(t) - c + ((uint32_t)d >> k.ds), k.clampv);
    if constexpr (SGN == 0) { mag = (int32_t)x < 0 ? k.neg_unsigned : mag; sx = 0u; }
    else if constexpr (SGN == 2) sx = mag == 0u ? 0u : x;
    else sx = x;
    return mag;
}

// the rare inputs of a lane (infinity, NaN): the second-generation routine, split into magnitude + sign carrier
template <int MODE>
__device__ __noinline__ uint2 lane_mag_rare(uint32_t x, const CodecPlan& p)      // (magnitude, sign carrier)
{
    const FastConsts fk(p);
    const uint32_t rep = float_to_rep_fast<MODE>(x, fk) & p.rep_mask;
    return make_uint2(rep & ~p.signmask, (rep & p.signmask) ? 0x80000000u : 0u);
}

// four codes of one 16-byte vector -> packed bits.  pack8: one word; pack16: two words; pack4: 16 bits; pack6: 24 bits.
// SGN = 2 (the sign-bit-only pattern is the NaN: a zero magnitude takes no sign): instead of a compare + select per element the sign
// mask is cleared on the packed word for every field whose magnitude bits are all zero -- z + ones carries into the field's top bit
// exactly when z != 0 and never beyond it (z <= ones).  `mags` = the magnitude bits of every field, `signs` = the sign bit of every field.
template <int SGN>
__device__ __forceinline__ uint32_t no_minus_zero(uint32_t w, uint32_t mags, uint32_t signs)
{
    if constexpr (SGN != 2) return w;
    const uint32_t z = w & mags;
    return w & (((z + mags) | z) | ~signs);
}

template <int SGN>
__device__ __forceinline__ uint32_t pack8(const uint32_t (&m)[4], const uint32_t (&s)[4])
{
    uint32_t w = __byte_perm(__byte_perm(m[0], m[1], 0x0040), __byte_perm(m[2], m[3], 0x0040), 0x5410);
    if constexpr (SGN != 0) {
        const uint32_t sg = __byte_perm(__byte_perm(s[0], s[1], 0x0073), __byte_perm(s[2], s[3], 0x0073), 0x5410);
        w |= sg & 0x80808080u;
    }
    return no_minus_zero<SGN>(w, 0x7F7F7F7Fu, 0x80808080u);
}
template <int SGN>
__device__ __forceinline__ uint2 pack16(const uint32_t (&m)[4], const uint32_t (&s)[4])
{
    uint2 w = make_uint2(__byte_perm(m[0], m[1], 0x5410), __byte_perm(m[2], m[3], 0x5410));
    if constexpr (SGN != 0) {
        w.x |= __byte_perm(s[0], s[1], 0x7632) & 0x80008000u;
        w.y |= __byte_perm(s[2], s[3], 0x7632) & 0x80008000u;
    }
    w.x = no_minus_zero<SGN>(w.x, 0x7FFF7FFFu, 0x80008000u);
    w.y = no_minus_zero<SGN>(w.y, 0x7FFF7FFFu, 0x80008000u);
    return w;
}
template <int SGN>
__device__ __forceinline__ uint32_t pack4(const uint32_t (&m)[4], const uint32_t (&s)[4], bool reverse)
{
    uint32_t q[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) q[e] = SGN != 0 ? (m[e] | ((s[e] >> 28) & 8u)) : m[e];
    uint32_t half = no_minus_zero<SGN>((q[1] * 16u + q[0]) + (q[3] * 16u + q[2]) * 256u, 0x7777u, 0x8888u);
    if (reverse) half = ((half & 0x0F0Fu) << 4) | ((half >> 4) & 0x0F0Fu);      // slot order inside a byte (launch-uniform)
    return half;
}
template <int SGN>
__device__ __forceinline__ uint32_t pack6(const uint32_t (&m)[4], const uint32_t (&s)[4], bool reverse)
{
    uint32_t q[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) q[e] = SGN != 0 ? (m[e] | ((s[e] >> 26) & 32u)) : m[e];
    const uint32_t chunk = reverse ? (q[2] * 64u + q[3]) + (q[0] * 64u + q[1]) * 4096u : (q[1] * 64u + q[0]) + (q[3] * 64u + q[2]) * 4096u;
    return no_minus_zero<SGN>(chunk, 0x7DF7DFu, 0x820820u);
}
// 6-bit lanes: four consecutive lanes hold the 24-bit chunks c0..c3 of 16 consecutive elements = three words
// (c0 | c1 << 24, c1 >> 8 | c2 << 16, c2 >> 16 | c3 << 8); lane r = lane & 3 builds word r from its own and the next lane's chunk
__device__ __forceinline__ uint32_t word6(uint32_t chunk, int r)
{
    const uint32_t next = __shfl_down_sync(0xFFFFFFFFu, chunk, 1);
    return __funnelshift_r(chunk | (next << 24), next >> 8, 8 * r);
}

__device__ __noinline__ float rep_to_float_wide(uint32_t ef, uint32_t mf, bool sign, int mant, int bias)
{
    const double v = ef == 0u ? ldexp((double)mf, 1 - bias - mant)
                              : ldexp(1.0 + (double)mf / (double)(1u << mant), (int)ef - bias);
    return (float)(sign ? -v : v);
}

// lo_pack.hpp:277-319
__device__ __forceinline__ float rep_to_float(uint32_t rep, const CodecPlan& p)
{
    rep &= p.rep_mask;
    const bool sign = p.to_signed && ((rep >> (p.n - 1)) & 1u);
    const uint32_t mag = p.to_signed ? (rep & (p.signmask - 1u)) : rep;
    if (p.has_inf && (rep & p.inf_mask) == p.inf_value) return __uint_as_float(sign ? 0xFF800000u : 0x7F800000u);
    if (codec_is_nan(p, rep)) return __uint_as_float(0x7FC00000u);
    const uint32_t ef = mag >> p.mant, mf = mag & ((1u << p.mant) - 1u);
    const uint32_t s = sign ? 0x80000000u : 0u;
    if (ef == 0u) {
        const int e = 1 - p.bias - p.mant;                    // value = mf * 2^e
        if (e >= -126 && e <= 100) return __uint_as_float(__float_as_uint(__fmul_rn(__uint2float_rn(mf), __uint_as_float((uint32_t)(e + 127) << 23))) | s);
        return rep_to_float_wide(ef, mf, sign, p.mant, p.bias);
    }
    const int E = (int)ef - p.bias + 127;
    if (E >= 1 && E <= 254) return __uint_as_float(s | ((uint32_t)E << 23) | (mf << (23 - p.mant)));
    return rep_to_float_wide(ef, mf, sign, p.mant, p.bias);
}


template <int MODE, bool FAST>
__global__ void __launch_bounds__(kPackThreads)
encode_kernel(const float* __restrict__ src, uint8_t* __restrict__ dst, int64_t count, uint64_t packed_bytes,
              const CodecPlan p, const RngArgs rng, uint64_t first_index, int vec_ok, int word_ok)
{
    __shared__ uint16_t reps[kPackTile];
    const FastConsts fk(p);
    const int64_t ntiles = (count + kPackTile - 1) / kPackTile;
    const int words_per_tile = kPackTile / 32 * p.n;          // 128 n
    const bool direct = word_ok && (p.n == 8 || p.n == 16 || p.n == 4);
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t base = tile * kPackTile;
        // ---- phase 1: one rep per element (all four 16-byte loads of the thread are issued before the first use) ----
        uint4 v[kPackPerThread / 4];
#pragma unroll
        for (int u = 0; u < kPackPerThread / 4; ++u) {
            const int64_t i = base + (u * kPackThreads + (int)threadIdx.x) * 4;
            v[u] = make_uint4(0u, 0u, 0u, 0u);
            if (vec_ok && i + 3 < count) {
                v[u] = __ldcs(reinterpret_cast<const uint4*>(src + i));
            } else {
                if (i < count) v[u].x = __float_as_uint(src[i]);
                if (i + 1 < count) v[u].y = __float_as_uint(src[i + 1]);
                if (i + 2 < count) v[u].z = __float_as_uint(src[i + 2]);
                if (i + 3 < count) v[u].w = __float_as_uint(src[i + 3]);
            }
        }
#pragma unroll
        for (int u = 0; u < kPackPerThread / 4; ++u) {
            const int local = (u * kPackThreads + (int)threadIdx.x) * 4;
            const int64_t i = base + local;
            const uint32_t x[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
            uint32_t r[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if constexpr (FAST) {
                    r[e] = float_to_rep_fast<MODE>(x[e], fk);      // elements past `count` were loaded as +0.0 -> rep 0
                } else {
                    uint32_t w = 0u;
                    if constexpr (ModeTraits<MODE>::kUsesRng) w = philox_event(rng, first_index + (uint64_t)(i + e));
                    r[e] = (i + e < count) ? float_to_rep<MODE>(x[e], p, w) : 0u;    // partial group: encode(0.0f) == 0 (:425-429)
                }
            }
            if (direct) {
                // n = 8 / 16 / 4: the thread's four consecutive elements are whole groups and land on byte boundaries of
                // their own -> packed straight from registers (no staging, no barrier); slot order only swaps nibbles
                if (p.n == 8) {
                    const uint32_t word = (r[0] & 0xFFu) | ((r[1] & 0xFFu) << 8) | ((r[2] & 0xFFu) << 16) | (r[3] << 24);
                    if (i + 3 < count) __stcs(reinterpret_cast<uint32_t*>(dst + i), word);
                    else for (int b = 0; b < 4 && i + b < count; ++b) dst[i + b] = (uint8_t)(word >> (8 * b));
                } else if (p.n == 16) {
                    const uint32_t w0 = (r[0] & 0xFFFFu) | (r[1] << 16), w1 = (r[2] & 0xFFFFu) | (r[3] << 16);
                    if (i + 3 < count) {
                        __stcs(reinterpret_cast<uint32_t*>(dst + 2 * i), w0);
                        __stcs(reinterpret_cast<uint32_t*>(dst + 2 * i + 4), w1);
                    } else {
                        const uint32_t ww[2] = {w0, w1};
                        for (int e = 0; e < 4 && i + e < count; ++e) reinterpret_cast<uint16_t*>(dst)[i + e] = (uint16_t)(ww[e >> 1] >> (16 * (e & 1)));
                    }
                } else {    // n == 4, two elements per byte
                    const uint32_t lo0 = p.reverse ? r[1] : r[0], hi0 = p.reverse ? r[0] : r[1];
                    const uint32_t lo1 = p.reverse ? r[3] : r[2], hi1 = p.reverse ? r[2] : r[3];
                    const uint32_t half = (lo0 & 0xFu) | ((hi0 & 0xFu) << 4) | ((lo1 & 0xFu) << 8) | ((hi1 & 0xFu) << 12);
                    const uint64_t byte = (uint64_t)i >> 1;
                    const uint32_t pair = half | (__shfl_down_sync(0xFFFFFFFFu, half, 1) << 16);       // even lane: 8 elements = one word
                    const uint64_t wbyte = byte & ~(uint64_t)3;
                    if (wbyte + 4u <= packed_bytes) {
                        if ((threadIdx.x & 1) == 0) __stcs(reinterpret_cast<uint32_t*>(dst + byte), pair);
                    } else if (byte + 2u <= packed_bytes) {
                        __stcs(reinterpret_cast<uint16_t*>(dst + byte), (uint16_t)half);
                    } else if (byte < packed_bytes) {
                        dst[byte] = (uint8_t)half;
                    }
                }
            } else {
                *reinterpret_cast<uint2*>(&reps[local]) = make_uint2((r[0] & 0xFFFFu) | (r[1] << 16), (r[2] & 0xFFFFu) | (r[3] << 16));
            }
        }
        if (direct) continue;
        __syncthreads();
        // ---- phase 2: whole 32-bit words of the packed stream ----
        const uint64_t tile_byte = (uint64_t)tile * (uint64_t)words_per_tile * 4u;
        for (int k = threadIdx.x; k < words_per_tile; k += kPackThreads) {
            const uint64_t byte = tile_byte + (uint64_t)k * 4u;
            if (byte >= packed_bytes) break;
            const int bit0 = k * 32;
            uint32_t word = 0u;
            int slot = p.n == 1 ? bit0 : (int)__umulhi((uint32_t)bit0, p.recip_n);
            for (int sb = slot * p.n; sb < bit0 + 32; ++slot, sb += p.n) {
                const uint32_t rep = reps[slot ^ p.rev_xor] & p.rep_mask;
                const int sh = sb - bit0;
                word |= sh >= 0 ? (rep << sh) : (rep >> (-sh));
            }
            if (word_ok && byte + 4u <= packed_bytes) {
                __stcs(reinterpret_cast<uint32_t*>(dst + byte), word);
            } else {
                for (int b = 0; b < 4 && byte + (uint64_t)b < packed_bytes; ++b) dst[byte + b] = (uint8_t)(word >> (8 * b));
            }
        }
        __syncthreads();
    }
}

// whole tiles of 4096 elements, 16-byte aligned source, 8-byte aligned destination (launch_encode sends the remainder to
// encode_kernel); N in {8, 16, 4, 6}
template <int MODE, int N, int SGN>
__global__ void __launch_bounds__(kPackThreads)
encode_lane_kernel(const uint4* __restrict__ src, uint8_t* __restrict__ dst, int64_t ntiles, const CodecPlan p)
{
    const LaneConsts k(p);
    const bool reverse = p.reverse != 0;
    constexpr int VPT = kPackPerThread / 4;                   // vectors per thread and tile
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t vbase = tile * (kPackTile / 4) + threadIdx.x;
        uint4 v[VPT];
#pragma unroll
        for (int u = 0; u < VPT; ++u) v[u] = __ldcs(src + vbase + u * kPackThreads);
        uint32_t m[VPT][4], s[VPT][4], emax = 0u;
#pragma unroll
        for (int u = 0; u < VPT; ++u) {
            const uint32_t x[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) m[u][e] = lane_mag<MODE, SGN == 2 ? 1 : SGN>(x[e], k, emax, s[u][e]);      // SGN = 2: the no -0 rule is applied on the packed word
        }
        const bool rare = emax >= kLaneRare;
        if (rare) {                                           // an infinity / NaN among the thread's 16 elements: redone exactly (incl. the no -0 rule)
#pragma unroll
            for (int u = 0; u < VPT; ++u) {
                const uint32_t x[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) { const uint2 r = lane_mag_rare<MODE>(x[e], p); m[u][e] = r.x; s[u][e] = r.y; }
            }
        }
#pragma unroll
        for (int u = 0; u < VPT; ++u) {
            const int64_t vi = vbase + u * kPackThreads;
            constexpr int S1 = SGN == 2 ? 1 : SGN;            // packer of the exactly redone elements (a sign-bit-only NaN code must survive)
            if constexpr (N == 8) {
                __stcs(reinterpret_cast<uint32_t*>(dst) + vi, rare ? pack8<S1>(m[u], s[u]) : pack8<SGN>(m[u], s[u]));
            } else if constexpr (N == 16) {
                __stcs(reinterpret_cast<uint2*>(dst) + vi, rare ? pack16<S1>(m[u], s[u]) : pack16<SGN>(m[u], s[u]));
            } else if constexpr (N == 4) {
                __stcs(reinterpret_cast<uint16_t*>(dst) + vi, (uint16_t)(rare ? pack4<S1>(m[u], s[u], reverse) : pack4<SGN>(m[u], s[u], reverse)));
            } else {
                const int r = (int)threadIdx.x & 3;
                const uint32_t w = word6(rare ? pack6<S1>(m[u], s[u], reverse) : pack6<SGN>(m[u], s[u], reverse), r);
                if (r != 3) __stcs(reinterpret_cast<uint32_t*>(dst) + (vi >> 2) * 3 + r, w);
            }
        }
    }
}

__global__ void __launch_bounds__(kPackThreads)
decode_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, int64_t count, uint64_t packed_bytes,
              const CodecPlan p, int vec_ok, int word_ok)
{
    __shared__ uint32_t words[kPackTile / 32 * 16 + 1];
    __shared__ float lut[1 << kLutBits];                      // n <= 12: every representation decoded once per CTA
    const bool use_lut = p.n <= kLutBits;
    if (use_lut) {
        for (int r = threadIdx.x; r < (1 << p.n); r += kPackThreads) lut[r] = rep_to_float((uint32_t)r, p);
        __syncthreads();
    }
    const int64_t ntiles = (count + kPackTile - 1) / kPackTile;
    const int words_per_tile = kPackTile / 32 * p.n;
    const bool direct = word_ok && (p.n == 8 || p.n == 16 || p.n == 4);
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t base = tile * kPackTile;
        if (direct) {
            // n = 8 / 16 / 4: each thread reads the bytes of its own four elements (mirror of the encode fast lane)
            uint32_t raw[kPackPerThread / 4][2];
#pragma unroll
            for (int u = 0; u < kPackPerThread / 4; ++u) {
                const int64_t i = base + (u * kPackThreads + (int)threadIdx.x) * 4;
                raw[u][0] = raw[u][1] = 0u;
                if (i >= count) continue;
                if (p.n == 8) {
                    if (i + 3 < count) raw[u][0] = __ldcs(reinterpret_cast<const uint32_t*>(src + i));
                    else for (int b = 0; b < 4 && i + b < count; ++b) raw[u][0] |= (uint32_t)src[i + b] << (8 * b);
                } else if (p.n == 16) {
                    if (i + 3 < count) {
                        raw[u][0] = __ldcs(reinterpret_cast<const uint32_t*>(src + 2 * i));
                        raw[u][1] = __ldcs(reinterpret_cast<const uint32_t*>(src + 2 * i + 4));
                    } else {
                        for (int e = 0; e < 4 && i + e < count; ++e)
                            raw[u][e >> 1] |= (uint32_t)reinterpret_cast<const uint16_t*>(src)[i + e] << (16 * (e & 1));
                    }
                } else {
                    const uint64_t byte = (uint64_t)i >> 1;
                    if (byte + 2u <= packed_bytes) raw[u][0] = __ldcs(reinterpret_cast<const uint16_t*>(src + byte));
                    else if (byte < packed_bytes) raw[u][0] = src[byte];
                }
            }
#pragma unroll
            for (int u = 0; u < kPackPerThread / 4; ++u) {
                const int64_t i = base + (u * kPackThreads + (int)threadIdx.x) * 4;
                float o[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    uint32_t rep;
                    if (p.n == 8) rep = (raw[u][0] >> (8 * e)) & 0xFFu;
                    else if (p.n == 16) rep = (raw[u][e >> 1] >> (16 * (e & 1))) & 0xFFFFu;
                    else rep = (raw[u][0] >> (4 * (e ^ (int)p.rev_xor))) & 0xFu;
                    o[e] = use_lut ? lut[rep] : rep_to_float(rep, p);
                }
                if (vec_ok && i + 3 < count) {
                    __stcs(reinterpret_cast<float4*>(dst + i), make_float4(o[0], o[1], o[2], o[3]));
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e) if (i + e < count) dst[i + e] = o[e];
                }
            }
            continue;
        }
        const uint64_t tile_byte = (uint64_t)tile * (uint64_t)words_per_tile * 4u;
        for (int k = threadIdx.x; k <= words_per_tile; k += kPackThreads) {
            const uint64_t byte = tile_byte + (uint64_t)k * 4u;
            uint32_t wv = 0u;
            if (k < words_per_tile && byte < packed_bytes) {
                if (word_ok && byte + 4u <= packed_bytes) wv = __ldcs(reinterpret_cast<const uint32_t*>(src + byte));
                else for (int b = 0; b < 4 && byte + (uint64_t)b < packed_bytes; ++b) wv |= (uint32_t)src[byte + b] << (8 * b);
            }
            words[k] = wv;
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < kPackPerThread / 4; ++u) {
            const int local = (u * kPackThreads + (int)threadIdx.x) * 4;
            const int64_t i = base + local;
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int el = local + e;
                const int slot = el ^ (int)p.rev_xor;                          // the slot <-> element map is an involution
                const int bit = slot * p.n;
                const uint64_t two = (uint64_t)words[bit >> 5] | ((uint64_t)words[(bit >> 5) + 1] << 32);
                const uint32_t rep = (uint32_t)(two >> (bit & 31)) & p.rep_mask;
                o[e] = use_lut ? lut[rep] : rep_to_float(rep, p);
            }
            if (vec_ok && i + 3 < count) {
                __stcs(reinterpret_cast<float4*>(dst + i), make_float4(o[0], o[1], o[2], o[3]));
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) if (i + e < count) dst[i + e] = o[e];
            }
        }
        __syncthreads();
    }
}

// Register-lane decode (mirror of encode_lane_kernel): a finite code is  (mag << (23 - m)) reinterpreted as fp32 times
// 2^(127 - bias) -- exact for normal codes (exponent re-bias) and for subnormal codes (the fp32 subnormal mf * 2^-(126 + m)
// scaled to mf * 2^(1 - bias - m)) -- with the sign bit moved to bit 31.  Infinity / NaN codes are found on the host
// (make_codec_plan: every special code has magnitude >= d_special, or is the sign-bit-only pattern when SGN = 2) and
// detected per thread tile on a running maximum; such tiles are redone with rep_to_float.  ~6 alu-pipe instructions per
// element instead of a shared-memory table look-up behind launch-uniform branches.
__device__ __noinline__ float rep_to_float_rare(uint32_t rep, const CodecPlan& p) { return rep_to_float(rep, p); }

struct DecConsts {
    uint32_t magmask, special, signmask;
    float k;
    int ds;
    __device__ __forceinline__ explicit DecConsts(const CodecPlan& p)
        : magmask(p.to_signed ? p.signmask - 1u : p.rep_mask), special(p.d_special), signmask(p.signmask), k(0.0f), ds(p.digit_drop)
    {
        k = __uint_as_float((uint32_t)(254 - p.bias) << 23);
        asm volatile("" : "+r"(magmask), "+r"(special), "+r"(ds), "+f"(k));
    }
};

// the N-bit code at bit `pos` of `chunk` (compile-time position): value, running maximum of the magnitudes and (SGN = 2) running
// minimum of rep ^ signmask (0 <=> the sign-bit-only pattern occurred)
template <int N, int SGN>
__device__ __forceinline__ float lane_decode(uint32_t chunk, int pos, const DecConsts& k, uint32_t& mmax, uint32_t& nmin)
{
    const uint32_t rep = chunk >> pos;
    const uint32_t mag = rep & k.magmask;
    mmax = max(mmax, mag);
    if constexpr (SGN == 2) nmin = min(nmin, (rep & (k.magmask | k.signmask)) ^ k.signmask);
    const float f = __fmul_rn(__uint_as_float(mag << k.ds), k.k);
    if constexpr (SGN == 0) return f;
    else return __uint_as_float(__float_as_uint(f) | ((chunk << (32 - N - pos)) & 0x80000000u));
}

template <int N, int SGN>
__global__ void __launch_bounds__(kPackThreads)
decode_lane_kernel(const uint8_t* __restrict__ src, float4* __restrict__ dst, int64_t ntiles, const CodecPlan p)
{
    const DecConsts k(p);
    const bool reverse = p.reverse != 0;
    constexpr int VPT = kPackPerThread / 4;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t vbase = tile * (kPackTile / 4) + threadIdx.x;
        uint32_t c0[VPT], c1[VPT];
#pragma unroll
        for (int u = 0; u < VPT; ++u) {
            const int64_t vi = vbase + u * kPackThreads;
            c1[u] = 0u;
            if constexpr (N == 8) c0[u] = __ldcs(reinterpret_cast<const uint32_t*>(src) + vi);
            else if constexpr (N == 16) { const uint2 w = __ldcs(reinterpret_cast<const uint2*>(src) + vi); c0[u] = w.x; c1[u] = w.y; }
            else if constexpr (N == 4) c0[u] = __ldcs(reinterpret_cast<const uint16_t*>(src) + vi);
            else {
                const int r = (int)threadIdx.x & 3;
                c0[u] = r != 3 ? __ldcs(reinterpret_cast<const uint32_t*>(src) + (vi >> 2) * 3 + r) : 0u;
            }
        }
        float o[VPT][4];
        uint32_t mmax = 0u, nmin = 0xFFFFFFFFu;
#pragma unroll
        for (int u = 0; u < VPT; ++u) {
            if constexpr (N == 6) {        // lanes 4g .. 4g + 3 hold the three words of 16 codes: lane r takes bits [24 r, 24 r + 24)
                const int lane = (int)threadIdx.x & 31, r = lane & 3, g = lane & ~3;
                const uint32_t lo = __shfl_sync(0xFFFFFFFFu, c0[u], g + max(r - 1, 0)), hi = __shfl_sync(0xFFFFFFFFu, c0[u], g + min(r, 2));
                c0[u] = __funnelshift_r(lo, hi, (32 - 8 * r) & 31) & 0x00FFFFFFu;
            }
            if constexpr (N == 4) { if (reverse) c0[u] = ((c0[u] & 0x0F0Fu) << 4) | ((c0[u] >> 4) & 0x0F0Fu); }
            if constexpr (N == 16) {
                o[u][0] = lane_decode<16, SGN>(c0[u], 0, k, mmax, nmin); o[u][1] = lane_decode<16, SGN>(c0[u], 16, k, mmax, nmin);
                o[u][2] = lane_decode<16, SGN>(c1[u], 0, k, mmax, nmin); o[u][3] = lane_decode<16, SGN>(c1[u], 16, k, mmax, nmin);
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) o[u][e] = lane_decode<N, SGN>(c0[u], N * e, k, mmax, nmin);
            }
        }
        if (mmax >= k.special || (SGN == 2 && nmin == 0u)) {         // an infinity / NaN code among the thread's 16: rare
#pragma unroll
            for (int u = 0; u < VPT; ++u) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const uint32_t rep = N == 16 ? ((e < 2 ? c0[u] : c1[u]) >> (16 * (e & 1))) : (c0[u] >> (N * e));
                    o[u][e] = rep_to_float_rare(rep, p);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < VPT; ++u) {
            const bool flip = N == 6 && reverse;
            __stcs(dst + vbase + u * kPackThreads, flip ? make_float4(o[u][3], o[u][2], o[u][1], o[u][0]) : make_float4(o[u][0], o[u][1], o[u][2], o[u][3]));
        }
    }
}

// Strict typed-semantics rounding as an elementwise op: out = decode(encode(x)) in one
// pass, i.e. (float)Templated_Float<Fp>(x) -- ConvertImpl::run (lo_float.h:3057-3364) through its runtime twin in
// lo_pack.hpp -- with Saturation_Mode (OvfInf / SatFinite / SatPropagate), signed infinities, clamp-to-max and the
// NegtoZero / NegtoNaN policies, unlike virtual_round's container semantics.  8 B / element, HBM-bound.
template <int MODE, bool FAST>
__global__ void __launch_bounds__(kPackThreads)
round_strict_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t count, const CodecPlan p, const RngArgs rng,
                    uint64_t first_index, int vec_ok)
{
    __shared__ float lut[1 << kLutBits];
    const bool use_lut = p.n <= kLutBits;
    if (use_lut) {
        for (int r = threadIdx.x; r < (1 << p.n); r += kPackThreads) lut[r] = rep_to_float((uint32_t)r, p);
        __syncthreads();
    }
    const FastConsts fk(p);
    const int64_t ntiles = (count + kPackTile - 1) / kPackTile;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t base = tile * kPackTile;
        uint4 v[kPackPerThread / 4];
#pragma unroll
        for (int u = 0; u < kPackPerThread / 4; ++u) {
            const int64_t i = base + (u * kPackThreads + (int)threadIdx.x) * 4;
            v[u] = make_uint4(0u, 0u, 0u, 0u);
            if (vec_ok && i + 3 < count) {
                v[u] = __ldcs(reinterpret_cast<const uint4*>(src + i));
            } else {
                if (i < count) v[u].x = __float_as_uint(src[i]);
                if (i + 1 < count) v[u].y = __float_as_uint(src[i + 1]);
                if (i + 2 < count) v[u].z = __float_as_uint(src[i + 2]);
                if (i + 3 < count) v[u].w = __float_as_uint(src[i + 3]);
            }
        }
#pragma unroll
        for (int u = 0; u < kPackPerThread / 4; ++u) {
            const int64_t i = base + (u * kPackThreads + (int)threadIdx.x) * 4;
            const uint32_t x[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                uint32_t rep;
                if constexpr (FAST) {
                    rep = float_to_rep_fast<MODE>(x[e], fk);
                } else {
                    uint32_t w = 0u;
                    if constexpr (ModeTraits<MODE>::kUsesRng) w = philox_event(rng, first_index + (uint64_t)(i + e));
                    rep = float_to_rep<MODE>(x[e], p, w);
                }
                rep &= p.rep_mask;
                o[e] = use_lut ? lut[rep] : rep_to_float(rep, p);
            }
            if (vec_ok && i + 3 < count) {
                __stcs(reinterpret_cast<float4*>(dst + i), make_float4(o[0], o[1], o[2], o[3]));
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) if (i + e < count) dst[i + e] = o[e];
            }
        }
    }
}

// Register-lane strict rounding: lane_mag (encode) composed with the arithmetic decode of lane_decode, no table, no bit
// width dependence (nothing is packed).  An overflow code above max_finite_rep (the infinity code under OvfInf) and the
// negative-input code of unsigned formats decode to launch constants; infinity / NaN inputs redo the thread tile with the
// second-generation routine.  Whole tiles, 16-byte aligned tensors.
template <int MODE, int SGN>
__global__ void __launch_bounds__(kPackThreads)
round_strict_lane_kernel(const uint4* __restrict__ src, float4* __restrict__ dst, int64_t ntiles, const CodecPlan p)
{
    const LaneConsts k(p);
    const DecConsts dk(p);
    const uint32_t max_finite = p.max_finite_rep;
    const float ovf_value = fabsf(rep_to_float_rare(p.ovf_rep, p));                   // max finite, or +inf
    const float neg_value = SGN == 0 ? rep_to_float_rare(p.neg_unsigned_rep, p) : 0.0f;
    // codes are <= ovf_rep.  An infinity code above max_finite_rep as ovf_rep is the select below; infinity / NaN codes that some
    // formats keep inside [0, max_finite_rep] (e4m3fn's 0x7F, the top codes of a saturating e5m2, the second infinity pattern 0x7E of the
    // unsigned extended P3109 formats) are reachable by rounding: running maximum of the codes, tile redone through the exact decode
    const uint32_t special = p.d_special > p.max_finite_rep ? 0xFFFFFFFFu : p.d_special;
    constexpr int VPT = kPackPerThread / 4;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t vbase = tile * (kPackTile / 4) + threadIdx.x;
        uint4 v[VPT];
#pragma unroll
        for (int u = 0; u < VPT; ++u) v[u] = __ldcs(src + vbase + u * kPackThreads);
        float o[VPT][4];
        uint32_t emax = 0u, mmax = 0u;
#pragma unroll
        for (int u = 0; u < VPT; ++u) {
            const uint32_t x[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                uint32_t sx;
                // unsigned targets: the magnitude path on |x|, negative inputs replaced by their constant below
                const uint32_t mag = lane_mag<MODE, SGN == 0 ? 1 : SGN>(x[e], k, emax, sx);
                mmax = max(mmax, mag);
                float f = __fmul_rn(__uint_as_float(mag << dk.ds), dk.k);
                f = mag > max_finite ? ovf_value : f;
                if constexpr (SGN == 0) f = (int32_t)x[e] < 0 ? neg_value : f;
                else f = __uint_as_float(__float_as_uint(f) | (sx & 0x80000000u));
                o[u][e] = f;
            }
        }
        if (emax >= kLaneRare || mmax >= special) {           // an infinity / NaN input, or a NaN code, among the thread's 16 elements: rare
#pragma unroll
            for (int u = 0; u < VPT; ++u) {
                const uint32_t x[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const uint2 r = lane_mag_rare<MODE>(x[e], p);
                    o[u][e] = rep_to_float_rare(r.x | (r.y ? p.signmask : 0u), p);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < VPT; ++u) __stcs(dst + vbase + u * kPackThreads, make_float4(o[u][0], o[u][1], o[u][2], o[u][3]));
    }
}

// ------------------------------------------------------------------------------------------------
// MX quantisation with MATERIALISED scales: virtual_mx_round's block / amax / scale / element rounding
// (lo_float.h:4311-4383, same device routines as lof_mx_kernels.cuh) composed with the codec above:
// one scale byte per block + packed element representations (4-, 6- or 8-bit) instead of fp32 fake-quant values.
// A lane owns 4 consecutive elements (one 128-bit load), a block is BS / 4 lanes of one warp (amax by shfl.xor).
// 4 B read + (n/8 + 1/BS) B written per element.
// ------------------------------------------------------------------------------------------------
template <int MODE, bool FAST>
__global__ void __launch_bounds__(kPackThreads)
mx_encode_kernel(const float4* __restrict__ in, int64_t nvec, MxArgs a, MxScaleQuick sq, CodecPlan scale_codec, CodecPlan elem_codec,
                 RngArgs rng, RngArgs rng_scale, uint64_t first_index, uint8_t* __restrict__ scales, uint8_t* __restrict__ elems)
{
    constexpr int U = 2;                                         // vectors in flight per thread
    const int lanes = a.block_size >> 2;                         // lanes per block: 4 .. 32
    const int log2_bs = 31 - __clz(a.block_size);                // block sizes are powers of two: no 64-bit divisions in the loop
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * kPackThreads;
    const int64_t nvec_pad = (nvec + 31) & ~(int64_t)31;         // whole warps stay converged for the shuffles
    const FastConsts fk(elem_codec);
    for (int64_t v0 = (int64_t)blockIdx.x * kPackThreads + threadIdx.x; v0 < nvec_pad; v0 += U * stride) {
        float4 xv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t v = v0 + u * stride;
            xv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (v < nvec) xv[u] = __ldcs(in + v);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t v = v0 + u * stride;
            if (v >= nvec_pad) break;                            // warp-uniform
            const bool live_lane = v < nvec;
            const float4 x = xv[u];
            float amax = fmaxf(fmaxf(fabsf(x.x), fabsf(x.y)), fmaxf(fabsf(x.z), fabsf(x.w)));       // NaN-dropping like std::max(amax, |x|)
            for (int o = lanes >> 1; o >= 1; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xFFFFFFFFu, amax, o));
            const uint64_t g0 = first_index + (uint64_t)v * 4u;
            // the block's scale and its stored byte: computed by the block's first lane only, then broadcast
            const bool head = (lane & (lanes - 1)) == 0;
            float scale = 0.0f;
            if (head && live_lane) {
                scale = mx_scale_quick(amax, a, sq, rng_scale, g0 >> log2_bs);
                // stored scale byte: for exponent-only unsigned formats (E8M0) a normal power-of-two scale encodes as its
                // re-biased exponent field (lo_pack.hpp's normal path with nothing to round); anything else: exact routine
                const uint32_t sb = __float_as_uint(scale);
                const int code = (int)(sb >> 23) + scale_codec.exp_offset;
                uint32_t byte;
                if (scale_codec.mant == 0 && !scale_codec.to_signed && (sb & 0x807FFFFFu) == 0u && (sb >> 23) >= 1u && (sb >> 23) <= 254u &&
                    code >= 1 && code <= (int)scale_codec.max_finite_rep)
                    byte = (uint32_t)code;
                else
                    byte = float_to_rep<LOF_RoundTowardsZero>(sb, scale_codec, 0u);
                scales[(v * 4) >> log2_bs] = (uint8_t)byte;
            }
            scale = __shfl_sync(0xFFFFFFFFu, scale, lane & ~(lanes - 1));
            if (!live_lane) continue;
            const MxBlockCtx ctx = mx_make_ctx(scale);
            uint32_t r[4] = {0u, 0u, 0u, 0u};
            if (ctx.live) {
                const float xs[4] = {x.x, x.y, x.z, x.w};
                uint4 w4 = make_uint4(0u, 0u, 0u, 0u);
                if constexpr (ModeTraits<MODE>::kUsesRng) w4 = philox_group4(rng, g0 >> 2);         // g0 % 4 == 0 (host-checked)
                const uint32_t w[4] = {w4.x, w4.y, w4.z, w4.w};
                float y[4];
                if (ctx.pow2) {                                  // E8M0-like scale: x / scale == x * 2^-k exactly
#pragma unroll
                    for (int e = 0; e < 4; ++e) y[e] = __fmul_rn(xs[e], ctx.recip);
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e) y[e] = __fdiv_rn(xs[e], ctx.scale);
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if constexpr (FAST) {
                        // rep(virtual_round(y)) == the codec's own rounding of y for the IEEE-like modes on eligible formats
                        // (both are the correctly rounded grid value; overflow clamps to the same code): one magic add
                        r[e] = float_to_rep_fast<MODE>(__float_as_uint(y[e]), fk) & elem_codec.rep_mask;
                    } else {
                        const uint32_t q = round_params_elem<LOF_F32, MODE>(__float_as_uint(y[e]), a.elem, w[e]);
                        r[e] = float_to_rep<LOF_RoundTowardsZero>(q, elem_codec, 0u) & elem_codec.rep_mask;   // q is on the grid: exact
                    }
                }
            }
            if (elem_codec.reverse && elem_codec.gc > 1) {           // slot order inside a group (gc = 2 for 4-bit, 4 for 6-bit)
                if (elem_codec.gc == 2) { uint32_t t = r[0]; r[0] = r[1]; r[1] = t; t = r[2]; r[2] = r[3]; r[3] = t; }
                else { uint32_t t = r[0]; r[0] = r[3]; r[3] = t; t = r[1]; r[1] = r[2]; r[2] = t; }
            }
            if (elem_codec.n == 8) {
                __stcs(reinterpret_cast<uint32_t*>(elems) + v, r[0] | (r[1] << 8) | (r[2] << 16) | (r[3] << 24));
            } else if (elem_codec.n == 4) {
                __stcs(reinterpret_cast<uint16_t*>(elems) + v, (uint16_t)(r[0] | (r[1] << 4) | (r[2] << 8) | (r[3] << 12)));
            } else {                                                 // 6-bit: four elements = one 3-byte group
                const uint32_t g = r[0] | (r[1] << 6) | (r[2] << 12) | (r[3] << 18);
                uint8_t* d = elems + v * 3;
                d[0] = (uint8_t)g; d[1] = (uint8_t)(g >> 8); d[2] = (uint8_t)(g >> 16);
            }
        }
    }
}

// Second-generation MX encode for the eligible bulk (fast element path, block sizes 16 / 32 / 64): ncu of the kernel above
// (profiles/r01j) shows ~300 warp instructions per 128 elements, a third of them the block-scale path, which only the
// first lane of each block executes but the whole warp pays for.  Here a warp step covers exactly 32 MX blocks
// (LANES = BS / 4 vectors per lane); the 32 block maxima are transposed onto the 32 lanes by shuffles, EVERY lane rounds
// one scale (no divergence, 32 scale bytes stored as one coalesced row), and scale / reciprocal travel back by shuffle.
template <int MODE, int LANES>
__global__ void __launch_bounds__(kPackThreads)
mx_encode_fast_kernel(const float4* __restrict__ in, int64_t nvec, MxArgs a, MxScaleQuick sq, CodecPlan scale_codec, CodecPlan elem_codec,
                      RngArgs rng_scale, uint64_t first_index, uint8_t* __restrict__ scales, uint8_t* __restrict__ elems)
{
    constexpr int V = LANES;                         // vectors (rows of 128 floats) per lane and warp step
    constexpr int BPR = 32 / LANES;                  // blocks per row
    constexpr int kWarps = kPackThreads / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t nsteps = (nvec + 32 * V - 1) / (32 * V);
    const int64_t nblocks = (nvec * 4) / a.block_size;
    const int log2_bs = 31 - __clz(a.block_size);
    const FastConsts fk(elem_codec);
    for (int64_t step = (int64_t)blockIdx.x * kWarps + warp; step < nsteps; step += (int64_t)gridDim.x * kWarps) {
        const int64_t v0 = step * (32 * V) + lane;
        float4 x[V];
#pragma unroll
        for (int u = 0; u < V; ++u) {
            const int64_t v = v0 + u * 32;
            x[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (v < nvec) x[u] = __ldcs(in + v);
        }
        // block maxima (all LANES lanes of a block end up with it), then lane j takes block j = (row j / BPR, group j % BPR)
        float mine = 0.0f;
#pragma unroll
        for (int u = 0; u < V; ++u) {
            float m = fmaxf(fmaxf(fabsf(x[u].x), fabsf(x[u].y)), fmaxf(fabsf(x[u].z), fabsf(x[u].w)));
#pragma unroll
            for (int o = LANES / 2; o >= 1; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
            const float t = __shfl_sync(0xFFFFFFFFu, m, (lane % BPR) * LANES);
            if (lane / BPR == u) mine = t;
        }
        const int64_t gblock_local = step * 32 + lane;
        const uint64_t gblock = (first_index >> log2_bs) + (uint64_t)gblock_local;
        const float scale = mx_scale_quick(mine, a, sq, rng_scale, gblock);
        const uint32_t sb = __float_as_uint(scale);
        // reciprocal for the consumers: exact 2^-k when the scale is a normal power of two, 0 = "use the division"
        const bool pow2 = (sb & 0x807FFFFFu) == 0u && (sb - 0x00800000u) < 0x7E800000u;
        const float recip = pow2 ? __uint_as_float(0x7F000000u - sb) : 0.0f;
        if (gblock_local < nblocks) {
            const int code = (int)(sb >> 23) + scale_codec.exp_offset;
            uint32_t byte;
            if (scale_codec.mant == 0 && !scale_codec.to_signed && pow2 && (sb >> 23) <= 254u && code >= 1 && code <= (int)scale_codec.max_finite_rep)
                byte = (uint32_t)code;
            else
                byte = float_to_rep<LOF_RoundTowardsZero>(sb, scale_codec, 0u);
            scales[gblock_local] = (uint8_t)byte;
        }
#pragma unroll
        for (int u = 0; u < V; ++u) {
            const int owner = u * BPR + lane / LANES;
            const float sc = __shfl_sync(0xFFFFFFFFu, scale, owner), rc = __shfl_sync(0xFFFFFFFFu, recip, owner);
            const int64_t v = v0 + u * 32;
            if (v >= nvec) continue;
            uint32_t r[4] = {0u, 0u, 0u, 0u};
            if (sc > 0.0f) {
                const float xs[4] = {x[u].x, x[u].y, x[u].z, x[u].w};
                float y[4];
                if (rc != 0.0f) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) y[e] = __fmul_rn(xs[e], rc);
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e) y[e] = __fdiv_rn(xs[e], sc);
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) r[e] = float_to_rep_fast<MODE>(__float_as_uint(y[e]), fk) & elem_codec.rep_mask;
            }
            if (elem_codec.reverse && elem_codec.gc > 1) {
                if (elem_codec.gc == 2) { uint32_t t = r[0]; r[0] = r[1]; r[1] = t; t = r[2]; r[2] = r[3]; r[3] = t; }
                else { uint32_t t = r[0]; r[0] = r[3]; r[3] = t; t = r[1]; r[1] = r[2]; r[2] = t; }
            }
            if (elem_codec.n == 8) {
                __stcs(reinterpret_cast<uint32_t*>(elems) + v, r[0] | (r[1] << 8) | (r[2] << 16) | (r[3] << 24));
            } else if (elem_codec.n == 4) {
                __stcs(reinterpret_cast<uint16_t*>(elems) + v, (uint16_t)(r[0] | (r[1] << 4) | (r[2] << 8) | (r[3] << 12)));
            } else {
                const uint32_t g = r[0] | (r[1] << 6) | (r[2] << 12) | (r[3] << 18);
                uint8_t* d = elems + v * 3;
                d[0] = (uint8_t)g; d[1] = (uint8_t)(g >> 8); d[2] = (uint8_t)(g >> 16);
            }
        }
    }
}

// Third-generation MX encode: the warp-step structure of mx_encode_fast_kernel with the register-lane element path
// (lane_mag + pack8 / pack4 / pack6, bit width as a template parameter, rare inputs detected once per vector).  Signed
// element formats (SGN = 1: with a -0, SGN = 2: P3109-style, the sign-bit-only pattern is the NaN); anything else stays on the kernels above.
template <int MODE, int LANES, int N, int SGN>
__global__ void __launch_bounds__(kPackThreads)
mx_encode_lane_kernel(const float4* __restrict__ in, int64_t nvec, MxArgs a, MxScaleQuick sq, CodecPlan scale_codec, CodecPlan elem_codec,
                      RngArgs rng_scale, uint64_t first_index, uint8_t* __restrict__ scales, uint8_t* __restrict__ elems)
{
    constexpr int V = LANES;                         // vectors (rows of 128 floats) per lane and warp step
    constexpr int BPR = 32 / LANES;                  // blocks per row
    constexpr int kWarps = kPackThreads / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t nsteps = (nvec + 32 * V - 1) / (32 * V);
    const int64_t nblocks = (nvec * 4) / a.block_size;
    const int log2_bs = 31 - __clz(a.block_size);
    const LaneConsts k(elem_codec);
    const bool reverse = elem_codec.reverse != 0;
    const bool e8m0_like = scale_codec.mant == 0 && !scale_codec.to_signed;
    for (int64_t step = (int64_t)blockIdx.x * kWarps + warp; step < nsteps; step += (int64_t)gridDim.x * kWarps) {
        const int64_t v0 = step * (32 * V) + lane;
        float4 x[V];
#pragma unroll
        for (int u = 0; u < V; ++u) {
            const int64_t v = v0 + u * 32;
            x[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (v < nvec) x[u] = __ldcs(in + v);
        }
        float mine = 0.0f;
#pragma unroll
        for (int u = 0; u < V; ++u) {
            float m = fmaxf(fmaxf(fabsf(x[u].x), fabsf(x[u].y)), fmaxf(fabsf(x[u].z), fabsf(x[u].w)));
#pragma unroll
            for (int o = LANES / 2; o >= 1; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
            const float t = __shfl_sync(0xFFFFFFFFu, m, (lane % BPR) * LANES);
            if (lane / BPR == u) mine = t;
        }
        const int64_t gblock_local = step * 32 + lane;
        const uint64_t gblock = (first_index >> log2_bs) + (uint64_t)gblock_local;
        const float scale = mx_scale_quick(mine, a, sq, rng_scale, gblock);
        const uint32_t sb = __float_as_uint(scale);
        const bool pow2 = (sb & 0x807FFFFFu) == 0u && (sb - 0x00800000u) < 0x7E800000u;
        // what the consumers multiply by: exact 2^-k for a normal power-of-two scale, 0 for a dead block (scale <= 0 or NaN: codes 0),
        // NaN = "divide by the scale" (never for E8M0-like scale formats)
        const float recip = pow2 ? __uint_as_float(0x7F000000u - sb) : (scale > 0.0f ? __uint_as_float(0x7FC00000u) : 0.0f);
        if (gblock_local < nblocks) {
            const int code = (int)(sb >> 23) + scale_codec.exp_offset;
            uint32_t byte;
            if (e8m0_like && pow2 && (sb >> 23) <= 254u && code >= 1 && code <= (int)scale_codec.max_finite_rep)
                byte = (uint32_t)code;
            else
                byte = float_to_rep<LOF_RoundTowardsZero>(sb, scale_codec, 0u);
            scales[gblock_local] = (uint8_t)byte;
        }
#pragma unroll
        for (int u = 0; u < V; ++u) {
            const int owner = u * BPR + lane / LANES;
            const float rc = __shfl_sync(0xFFFFFFFFu, recip, owner), sc = __shfl_sync(0xFFFFFFFFu, scale, owner);
            const int64_t v = v0 + u * 32;
            float y[4] = {x[u].x, x[u].y, x[u].z, x[u].w};
            if (rc != rc) {                                                  // non-power-of-two scale: the reference's division
#pragma unroll
                for (int e = 0; e < 4; ++e) y[e] = __fdiv_rn(y[e], sc);
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) y[e] = __fmul_rn(y[e], rc);
            }
            uint32_t m[4], s[4], emax = 0u;
#pragma unroll
            for (int e = 0; e < 4; ++e) m[e] = lane_mag<MODE, 1>(__float_as_uint(y[e]), k, emax, s[e]);      // SGN = 2: no -0 on the packed word
            const bool rare = emax >= kLaneRare;
            constexpr int S1 = 1;
            if (rare) {
#pragma unroll
                for (int e = 0; e < 4; ++e) { const uint2 r = lane_mag_rare<MODE>(__float_as_uint(y[e]), elem_codec); m[e] = r.x; s[e] = r.y; }
            }
            if (rc == 0.0f) {                                                // dead block: all-zero codes (also drops the signs)
#pragma unroll
                for (int e = 0; e < 4; ++e) { m[e] = 0u; s[e] = 0u; }
            }
            if constexpr (N == 8) {
                if (v < nvec) __stcs(reinterpret_cast<uint32_t*>(elems) + v, rare ? pack8<S1>(m, s) : pack8<SGN>(m, s));
            } else if constexpr (N == 4) {
                if (v < nvec) __stcs(reinterpret_cast<uint16_t*>(elems) + v, (uint16_t)(rare ? pack4<S1>(m, s, reverse) : pack4<SGN>(m, s, reverse)));
            } else {
                const int r = lane & 3;
                const uint32_t w = word6(rare ? pack6<S1>(m, s, reverse) : pack6<SGN>(m, s, reverse), r);     // nvec % 4 == 0 (block sizes >= 16): groups of 4 lanes are whole
                if (r != 3 && v < nvec) __stcs(reinterpret_cast<uint32_t*>(elems) + (v >> 2) * 3 + r, w);
            }
        }
    }
}

__device__ __forceinline__ void ld_nc_256(const float4* p, float4& a, float4& b)        // p 32-byte aligned
{
    asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}

// Fourth-generation MX encode: ONE MX block per thread (VPB = block_size / 4 vectors of four elements, 64 .. 256 contiguous bytes).
// No shuffles at all -- the block maximum, the scale, its reciprocal and the scale byte are thread-local -- and a thread stores its
// block's codes as one contiguous run (16-byte stores), so a warp writes 32 runs back to back.  The loads are 16 bytes per thread at
// a stride of one block: every 128-byte line is fetched by 8 / VPB-dependent neighbouring instructions of the same thread and stays in
// L1 in between (ld.global.nc).  Element path: lane_mag + the register packers, as in the lane kernel above.
template <int MODE, int VPB, int N, int SGN>
__global__ void __launch_bounds__(kPackThreads)       // 68 registers = 3 CTAs per SM; forcing 4 (<= 64 registers) measured slower for 8-bit codes (4.77 vs 4.99 TB/s)
mx_encode_block_kernel(const float4* __restrict__ in, int64_t nblocks, MxArgs a, MxScaleQuick sq, CodecPlan scale_codec, CodecPlan elem_codec,
                       RngArgs rng_scale, uint64_t first_block, uint8_t* __restrict__ scales, uint8_t* __restrict__ elems)
{
    const LaneConsts k(elem_codec);
    const bool reverse = elem_codec.reverse != 0;
    const bool e8m0_like = scale_codec.mant == 0 && !scale_codec.to_signed;
    const int64_t stride = (int64_t)gridDim.x * kPackThreads;
    for (int64_t b = (int64_t)blockIdx.x * kPackThreads + threadIdx.x; b < nblocks; b += stride) {
        float4 x[VPB];
#pragma unroll
        for (int u = 0; u < VPB; u += 2) ld_nc_256(in + b * VPB + u, x[u], x[u + 1]);     // 32-byte loads: half the L1 tag look-ups of the strided pattern
        float amax = 0.0f, nansum = 0.0f;
#pragma unroll
        for (int u = 0; u < VPB; ++u) {
            amax = fmaxf(amax, fmaxf(fmaxf(fabsf(x[u].x), fabsf(x[u].y)), fmaxf(fabsf(x[u].z), fabsf(x[u].w))));
            nansum += (x[u].x + x[u].y) + (x[u].z + x[u].w);             // fma pipe (idle here): NaN iff the block holds a NaN (or +inf and -inf)
        }
        // with a power-of-two scale, NaN elements are the only inputs the lane path cannot encode (x * 2^-k is NaN only for a NaN x; an
        // overflowing product clamps to max_finite_rep like the exact routine does for these formats without infinities): one test per
        // block instead of a running maximum per element
        const bool nan_in = nansum != nansum;
        const float scale = mx_scale_quick(amax, a, sq, rng_scale, first_block + (uint64_t)b);
        const uint32_t sb = __float_as_uint(scale);
        const bool pow2 = (sb & 0x807FFFFFu) == 0u && (sb - 0x00800000u) < 0x7E800000u;
        const bool live = scale > 0.0f;                                  // dead block (scale <= 0 or NaN): all-zero codes
        const bool rare = nan_in || !pow2;                               // a non-power-of-two scale (e.g. the infinite scale of a block with an infinite element:
                                                                         // inf / inf = NaN) takes the exact element routine as well
        const float recip = __uint_as_float(0x7F000000u - sb);          // exact 2^-k when pow2
        {
            const int code = (int)(sb >> 23) + scale_codec.exp_offset;
            uint32_t byte;
            if (e8m0_like && pow2 && (sb >> 23) <= 254u && code >= 1 && code <= (int)scale_codec.max_finite_rep) byte = (uint32_t)code;
            else byte = float_to_rep<LOF_RoundTowardsZero>(sb, scale_codec, 0u);
            scales[b] = (uint8_t)byte;
        }
        uint32_t w[VPB];
#pragma unroll
        for (int u = 0; u < VPB; ++u) {
            float y[4] = {x[u].x, x[u].y, x[u].z, x[u].w};
            if (pow2) {
#pragma unroll
                for (int e = 0; e < 4; ++e) y[e] = __fmul_rn(y[e], recip);
            } else {                                                     // non-power-of-two scale: the reference's division (never for E8M0)
#pragma unroll
                for (int e = 0; e < 4; ++e) y[e] = __fdiv_rn(y[e], scale);
            }
            uint32_t m[4], s[4], unused = 0u;
#pragma unroll
            for (int e = 0; e < 4; ++e) m[e] = lane_mag<MODE, 1>(__float_as_uint(y[e]), k, unused, s[e]);
            if (rare) {
#pragma unroll
                for (int e = 0; e < 4; ++e) { const uint2 r = lane_mag_rare<MODE>(__float_as_uint(y[e]), elem_codec); m[e] = r.x; s[e] = r.y; }
            }
            uint32_t c;
            if constexpr (N == 8) c = rare ? pack8<1>(m, s) : pack8<SGN>(m, s);
            else if constexpr (N == 4) c = rare ? pack4<1>(m, s, reverse) : pack4<SGN>(m, s, reverse);
            else c = rare ? pack6<1>(m, s, reverse) : pack6<SGN>(m, s, reverse);
            w[u] = live ? c : 0u;
        }
        if constexpr (N == 8) {
            uint4* d = reinterpret_cast<uint4*>(elems) + b * (VPB / 4);
#pragma unroll
            for (int q = 0; q < VPB / 4; ++q) __stcs(d + q, make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]));
        } else if constexpr (N == 4) {                                    // 16 bits per vector
            uint32_t h[VPB / 2];
#pragma unroll
            for (int q = 0; q < VPB / 2; ++q) h[q] = w[2 * q] | (w[2 * q + 1] << 16);
            if constexpr (VPB == 4) __stcs(reinterpret_cast<uint2*>(elems) + b, make_uint2(h[0], h[1]));
            else {
                uint4* d = reinterpret_cast<uint4*>(elems) + b * (VPB / 8);
#pragma unroll
                for (int q = 0; q < VPB / 8; ++q) __stcs(d + q, make_uint4(h[4 * q], h[4 * q + 1], h[4 * q + 2], h[4 * q + 3]));
            }
        } else {                                                          // 24 bits per vector: four vectors = three words
            uint32_t t[VPB / 4 * 3];
#pragma unroll
            for (int q = 0; q < VPB / 4; ++q) {
                t[3 * q] = w[4 * q] | (w[4 * q + 1] << 24);
                t[3 * q + 1] = (w[4 * q + 1] >> 8) | (w[4 * q + 2] << 16);
                t[3 * q + 2] = (w[4 * q + 2] >> 16) | (w[4 * q + 3] << 8);
            }
            uint32_t* d = reinterpret_cast<uint32_t*>(elems) + b * (VPB / 4 * 3);
#pragma unroll
            for (int q = 0; q < VPB / 4 * 3; ++q) __stcs(d + q, t[q]);
        }
    }
}

// element * scale of the MX decode with the NaN bit patterns of the oracle's CPU arithmetic (x86 SSE, like mx_quant in lof_mx_kernels.cuh):
// a NaN operand (both decode to +qNaN) propagates as 0x7FC00000, an invalid product (inf code * zero scale) is the negative
// "indefinite" 0xFFC00000; the GPU multiply would return 0x7FFFFFFF for all of them.  One test per vector on the sum of the products.
__device__ __forceinline__ float mx_nan_fix(float r, float a, float s)
{
    return r != r ? __uint_as_float((a != a || s != s) ? 0x7FC00000u : 0xFFC00000u) : r;
}
__device__ __forceinline__ float4 mx_scale_mul(float a0, float a1, float a2, float a3, float s)
{
    float4 r = make_float4(__fmul_rn(a0, s), __fmul_rn(a1, s), __fmul_rn(a2, s), __fmul_rn(a3, s));
    const float chk = (r.x + r.y) + (r.z + r.w);
    if (chk != chk) { r.x = mx_nan_fix(r.x, a0, s); r.y = mx_nan_fix(r.y, a1, s); r.z = mx_nan_fix(r.z, a2, s); r.w = mx_nan_fix(r.w, a3, s); }
    return r;
}

__global__ void __launch_bounds__(kPackThreads)
mx_decode_kernel(const uint8_t* __restrict__ scales, const uint8_t* __restrict__ elems, float4* __restrict__ out, int64_t nvec,
                 int block_size, CodecPlan scale_codec, CodecPlan elem_codec)
{
    __shared__ float lut_e[256], lut_s[256];
    for (int r = threadIdx.x; r < 256; r += kPackThreads) {
        lut_e[r] = rep_to_float((uint32_t)r & elem_codec.rep_mask, elem_codec);
        lut_s[r] = rep_to_float((uint32_t)r & scale_codec.rep_mask, scale_codec);
    }
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * kPackThreads;
    for (int64_t v = (int64_t)blockIdx.x * kPackThreads + threadIdx.x; v < nvec; v += stride) {
        uint32_t r[4];
        if (elem_codec.n == 8) {
            const uint32_t w = __ldcs(reinterpret_cast<const uint32_t*>(elems) + v);
            r[0] = w & 0xFFu; r[1] = (w >> 8) & 0xFFu; r[2] = (w >> 16) & 0xFFu; r[3] = w >> 24;
        } else if (elem_codec.n == 4) {
            const uint32_t w = __ldcs(reinterpret_cast<const uint16_t*>(elems) + v);
            r[0] = w & 0xFu; r[1] = (w >> 4) & 0xFu; r[2] = (w >> 8) & 0xFu; r[3] = (w >> 12) & 0xFu;
        } else {
            const uint8_t* d = elems + v * 3;
            const uint32_t g = (uint32_t)d[0] | ((uint32_t)d[1] << 8) | ((uint32_t)d[2] << 16);
            r[0] = g & 0x3Fu; r[1] = (g >> 6) & 0x3Fu; r[2] = (g >> 12) & 0x3Fu; r[3] = (g >> 18) & 0x3Fu;
        }
        if (elem_codec.reverse && elem_codec.gc > 1) {
            if (elem_codec.gc == 2) { uint32_t t = r[0]; r[0] = r[1]; r[1] = t; t = r[2]; r[2] = r[3]; r[3] = t; }
            else { uint32_t t = r[0]; r[0] = r[3]; r[3] = t; t = r[1]; r[1] = r[2]; r[2] = t; }
        }
        const float s = lut_s[scales[(v * 4) / block_size]];
        __stcs(out + v, mx_scale_mul(lut_e[r[0]], lut_e[r[1]], lut_e[r[2]], lut_e[r[3]], s));
    }
}

// Register-lane MX decode: the tile structure of decode_lane_kernel (four vectors and their scale bytes in flight per thread) with the
// arithmetic element decode of lane_decode; the scale byte goes through a 256-entry table (one look-up per four elements).  Whole
// tiles of 4096 elements (launch_mx_decode sends the remainder to mx_decode_kernel), power-of-two block sizes, signed element formats.
template <int N, int SGN>
__global__ void __launch_bounds__(kPackThreads)
mx_decode_lane_kernel(const uint8_t* __restrict__ scales, const uint8_t* __restrict__ elems, float4* __restrict__ out, int64_t ntiles,
                      int log2_bs, const CodecPlan scale_codec, const CodecPlan elem_codec)
{
    __shared__ float lut_s[256];
    for (int r = threadIdx.x; r < 256; r += kPackThreads) lut_s[r] = rep_to_float_rare((uint32_t)r & scale_codec.rep_mask, scale_codec);
    __syncthreads();
    const DecConsts k(elem_codec);
    const bool reverse = elem_codec.reverse != 0;
    constexpr int VPT = kPackPerThread / 4;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t vbase = tile * (kPackTile / 4) + threadIdx.x;
        uint32_t c[VPT], sb[VPT];
#pragma unroll
        for (int u = 0; u < VPT; ++u) {
            const int64_t vi = vbase + u * kPackThreads;
            if constexpr (N == 8) c[u] = __ldcs(reinterpret_cast<const uint32_t*>(elems) + vi);
            else if constexpr (N == 4) c[u] = __ldcs(reinterpret_cast<const uint16_t*>(elems) + vi);
            else {
                const int r = (int)threadIdx.x & 3;
                c[u] = r != 3 ? __ldcs(reinterpret_cast<const uint32_t*>(elems) + (vi >> 2) * 3 + r) : 0u;
            }
            sb[u] = __ldg(scales + ((vi * 4) >> log2_bs));
        }
#pragma unroll
        for (int u = 0; u < VPT; ++u) {
            if constexpr (N == 6) {
                const int lane = (int)threadIdx.x & 31, r = lane & 3, g = lane & ~3;
                const uint32_t lo = __shfl_sync(0xFFFFFFFFu, c[u], g + max(r - 1, 0)), hi = __shfl_sync(0xFFFFFFFFu, c[u], g + min(r, 2));
                c[u] = __funnelshift_r(lo, hi, (32 - 8 * r) & 31) & 0x00FFFFFFu;
            }
            if constexpr (N == 4) { if (reverse) c[u] = ((c[u] & 0x0F0Fu) << 4) | ((c[u] >> 4) & 0x0F0Fu); }
            const float sc = lut_s[sb[u]];
            float o[4];
            uint32_t mmax = 0u, nmin = 0xFFFFFFFFu;
#pragma unroll
            for (int e = 0; e < 4; ++e) o[e] = lane_decode<N, SGN>(c[u], N * e, k, mmax, nmin);
            if (mmax >= k.special || (SGN == 2 && nmin == 0u)) {
#pragma unroll
                for (int e = 0; e < 4; ++e) o[e] = rep_to_float_rare(c[u] >> (N * e), elem_codec);
            }
            const bool flip = N == 6 && reverse;
            __stcs(out + vbase + u * kPackThreads, flip ? mx_scale_mul(o[3], o[2], o[1], o[0], sc) : mx_scale_mul(o[0], o[1], o[2], o[3], sc));
        }
    }
}

unsigned pack_grid(int64_t count, int num_sms)
{
    const int64_t ntiles = (count + kPackTile - 1) / kPackTile;
    const int64_t cap = (int64_t)num_sms /* = SMs x resident CTAs per SM, chosen by the caller */;
    return (unsigned)(ntiles < 1 ? 1 : ntiles < cap ? ntiles : cap);
}

}  // namespace

cudaError_t launch_encode(int mode, const float* src, uint8_t* dst, int64_t count, uint64_t packed_bytes, const CodecPlan& p,
                          const RngArgs& rng, uint64_t first_index, int num_sms, cudaStream_t stream)
{
    const int vec_ok = (reinterpret_cast<uintptr_t>(src) & 15u) == 0;
    const int word_ok = (reinterpret_cast<uintptr_t>(dst) & 3u) == 0;
    unsigned grid = pack_grid(count, num_sms);
    const int fast_mode = mode == LOF_RoundToNearestEven || mode == LOF_RoundTowardsZero || mode == LOF_RoundAwayFromZero || mode == LOF_RoundToOdd;
    if (p.fast_ok && p.lanes && fast_mode && (p.n == 8 || p.n == 16 || p.n == 4 || p.n == 6) && vec_ok &&
        (reinterpret_cast<uintptr_t>(dst) & 7u) == 0 && count >= kPackTile) {
        // whole tiles on the register-lane kernel, the remainder (< 4096 elements) on encode_kernel below
        const int64_t ntiles = count / kPackTile;
        const unsigned lgrid = pack_grid(ntiles * kPackTile, num_sms);
        const uint4* vsrc = reinterpret_cast<const uint4*>(src);
        const int sgn = !p.to_signed ? 0 : (p.sign_is_nan ? 2 : 1);
#define LOF_LANE3(M, N, S) encode_lane_kernel<M, N, S><<<lgrid, kPackThreads, 0, stream>>>(vsrc, dst, ntiles, p)
#define LOF_LANE2(M, N) do { if (sgn == 0) LOF_LANE3(M, N, 0); else if (sgn == 1) LOF_LANE3(M, N, 1); else LOF_LANE3(M, N, 2); } while (0)
#define LOF_LANE(M) case M: if (p.n == 8) LOF_LANE2(M, 8); else if (p.n == 16) LOF_LANE2(M, 16); else if (p.n == 4) LOF_LANE2(M, 4); else LOF_LANE2(M, 6); break;
        switch (mode) {
            LOF_LANE(LOF_RoundToNearestEven) LOF_LANE(LOF_RoundTowardsZero) LOF_LANE(LOF_RoundAwayFromZero) LOF_LANE(LOF_RoundToOdd)
        default: break;
        }
#undef LOF_LANE
#undef LOF_LANE2
#undef LOF_LANE3
        cudaError_t e = cudaGetLastError();
        const int64_t done = ntiles * kPackTile;
        if (e != cudaSuccess || done == count) return e;
        const uint64_t done_bytes = (uint64_t)done / 8u * (uint64_t)p.n;
        src += done; dst += done_bytes; count -= done; packed_bytes -= done_bytes; first_index += (uint64_t)done;
        grid = pack_grid(count, num_sms);
    }
    if (p.fast_ok) {
        switch (mode) {
#define LOF_CASE(M) case M: encode_kernel<M, true><<<grid, kPackThreads, 0, stream>>>(src, dst, count, packed_bytes, p, rng, first_index, vec_ok, word_ok); return cudaGetLastError();
            LOF_CASE(LOF_RoundToNearestEven) LOF_CASE(LOF_RoundTowardsZero) LOF_CASE(LOF_RoundAwayFromZero) LOF_CASE(LOF_RoundToOdd)
#undef LOF_CASE
        default: break;
        }
    }
    switch (mode) {
#define LOF_CASE(M) case M: encode_kernel<M, false><<<grid, kPackThreads, 0, stream>>>(src, dst, count, packed_bytes, p, rng, first_index, vec_ok, word_ok); break;
        LOF_CASE(LOF_RoundToNearestEven) LOF_CASE(LOF_RoundTowardsZero) LOF_CASE(LOF_RoundAwayFromZero)
        LOF_CASE(LOF_StochasticRoundingA) LOF_CASE(LOF_RoundToNearestOdd) LOF_CASE(LOF_RoundDown)
        LOF_CASE(LOF_RoundUp) LOF_CASE(LOF_RoundTiesToAway) LOF_CASE(LOF_StochasticRoundingB)
        LOF_CASE(LOF_StochasticRoundingC) LOF_CASE(LOF_True_StochasticRounding)
        LOF_CASE(LOF_ProbabilisticRounding) LOF_CASE(LOF_RoundToOdd) LOF_CASE(LOF_RoundTiesTowardsZero)
        LOF_CASE(LOF_StochasticRoundingD)
#undef LOF_CASE
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t launch_round_strict(int mode, const float* src, float* dst, int64_t count, const CodecPlan& p, const RngArgs& rng,
                                uint64_t first_index, int num_sms, cudaStream_t stream)
{
    const int vec_ok = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15u) == 0;
    unsigned grid = pack_grid(count, num_sms);
    const int fast_mode = mode == LOF_RoundToNearestEven || mode == LOF_RoundTowardsZero || mode == LOF_RoundAwayFromZero || mode == LOF_RoundToOdd;
    if (p.fast_ok && p.lanes && p.dec_ok && fast_mode && vec_ok && count >= kPackTile) {
        // whole tiles on the register-lane kernel, the remainder (< 4096 elements) on round_strict_kernel below
        const int64_t ntiles = count / kPackTile;
        const unsigned lgrid = pack_grid(ntiles * kPackTile, num_sms);
        const uint4* vsrc = reinterpret_cast<const uint4*>(src);
        float4* vdst = reinterpret_cast<float4*>(dst);
        const int sgn = !p.to_signed ? 0 : (p.sign_is_nan ? 2 : 1);
#define LOF_LANE2(M, S) round_strict_lane_kernel<M, S><<<lgrid, kPackThreads, 0, stream>>>(vsrc, vdst, ntiles, p)
#define LOF_LANE(M) case M: if (sgn == 0) LOF_LANE2(M, 0); else if (sgn == 1) LOF_LANE2(M, 1); else LOF_LANE2(M, 2); break;
        switch (mode) {
            LOF_LANE(LOF_RoundToNearestEven) LOF_LANE(LOF_RoundTowardsZero) LOF_LANE(LOF_RoundAwayFromZero) LOF_LANE(LOF_RoundToOdd)
        default: break;
        }
#undef LOF_LANE
#undef LOF_LANE2
        cudaError_t e = cudaGetLastError();
        const int64_t done = ntiles * kPackTile;
        if (e != cudaSuccess || done == count) return e;
        src += done; dst += done; count -= done; first_index += (uint64_t)done;
        grid = pack_grid(count, num_sms);
    }
    if (p.fast_ok) {
        switch (mode) {
#define LOF_CASE(M) case M: round_strict_kernel<M, true><<<grid, kPackThreads, 0, stream>>>(src, dst, count, p, rng, first_index, vec_ok); return cudaGetLastError();
            LOF_CASE(LOF_RoundToNearestEven) LOF_CASE(LOF_RoundTowardsZero) LOF_CASE(LOF_RoundAwayFromZero) LOF_CASE(LOF_RoundToOdd)
#undef LOF_CASE
        default: break;
        }
    }
    switch (mode) {
#define LOF_CASE(M) case M: round_strict_kernel<M, false><<<grid, kPackThreads, 0, stream>>>(src, dst, count, p, rng, first_index, vec_ok); break;
        LOF_CASE(LOF_RoundToNearestEven) LOF_CASE(LOF_RoundTowardsZero) LOF_CASE(LOF_RoundAwayFromZero)
        LOF_CASE(LOF_StochasticRoundingA) LOF_CASE(LOF_RoundToNearestOdd) LOF_CASE(LOF_RoundDown)
        LOF_CASE(LOF_RoundUp) LOF_CASE(LOF_RoundTiesToAway) LOF_CASE(LOF_StochasticRoundingB)
        LOF_CASE(LOF_StochasticRoundingC) LOF_CASE(LOF_True_StochasticRounding)
        LOF_CASE(LOF_ProbabilisticRounding) LOF_CASE(LOF_RoundToOdd) LOF_CASE(LOF_RoundTiesTowardsZero)
        LOF_CASE(LOF_StochasticRoundingD)
#undef LOF_CASE
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t launch_mx_encode(int mode, const float* in, int64_t n, const MxArgs& a, const CodecPlan& scale_codec, const CodecPlan& elem_codec,
                             const RngArgs& rng, const RngArgs& rng_scale, uint64_t first_index, uint8_t* scales, uint8_t* elems, int num_sms,
                             bool fast, const MagicPlan& scale_magic, const FastPlan& scale_fast, cudaStream_t stream)
{
    MxScaleQuick sq;
    sq.magic = scale_magic;
    sq.fast = scale_fast;
    sq.quick_mode_ok = a.scale_mode == LOF_RoundToNearestEven || a.scale_mode == LOF_RoundTowardsZero || a.scale_mode == LOF_RoundAwayFromZero ||
                       a.scale_mode == LOF_RoundUp || a.scale_mode == LOF_RoundDown || a.scale_mode == LOF_RoundToOdd;
    const int64_t nvec = n / 4;
    const int64_t want = (nvec + kPackThreads - 1) / kPackThreads, cap = (int64_t)num_sms /* = SMs x resident CTAs per SM, chosen by the caller */;
    const unsigned grid = (unsigned)(want < 1 ? 1 : want < cap ? want : cap);
    const float4* vin = reinterpret_cast<const float4*>(in);
    // eligible bulk: fast element path, deterministic scale mode, 32 whole blocks per warp step
    if (fast && !mode_uses_rng(a.scale_mode) && (a.block_size == 16 || a.block_size == 32 || a.block_size == 64)) {
        constexpr int kWarps = kPackThreads / 32;
        const int lanes = a.block_size / 4;
        const int64_t nsteps = (nvec + 32 * lanes - 1) / (32 * lanes);
        const int64_t want2 = (nsteps + kWarps - 1) / kWarps;
        const unsigned grid2 = (unsigned)(want2 < 1 ? 1 : want2 < cap ? want2 : cap);
        if (elem_codec.lanes >= 2 && elem_codec.to_signed && (elem_codec.n == 8 || elem_codec.n == 4 || elem_codec.n == 6) &&
            (reinterpret_cast<uintptr_t>(elems) & 15u) == 0 && (reinterpret_cast<uintptr_t>(in) & 31u) == 0 && first_index % (uint64_t)a.block_size == 0) {
            // one MX block per thread (pack.fast = 2, default)
            const int64_t nblocks = n / a.block_size;
            const int64_t want3 = (nblocks + kPackThreads - 1) / kPackThreads;
            const unsigned grid3 = (unsigned)(want3 < 1 ? 1 : want3 < cap ? want3 : cap);
            const uint64_t first_block = first_index / (uint64_t)a.block_size;
            const bool nz = elem_codec.sign_is_nan != 0;
#define LOF_MXB4(M, V, N, S) mx_encode_block_kernel<M, V, N, S><<<grid3, kPackThreads, 0, stream>>>(vin, nblocks, a, sq, scale_codec, elem_codec, rng_scale, first_block, scales, elems)
#define LOF_MXB3(M, V, N) do { if (nz) LOF_MXB4(M, V, N, 2); else LOF_MXB4(M, V, N, 1); } while (0)
#define LOF_MXB2(M, V) do { if (elem_codec.n == 8) LOF_MXB3(M, V, 8); else if (elem_codec.n == 4) LOF_MXB3(M, V, 4); else LOF_MXB3(M, V, 6); } while (0)
#define LOF_MXB(M) case M: if (lanes == 4) LOF_MXB2(M, 4); else if (lanes == 8) LOF_MXB2(M, 8); else LOF_MXB2(M, 16); return cudaGetLastError();
            switch (mode) {
                LOF_MXB(LOF_RoundToNearestEven) LOF_MXB(LOF_RoundTowardsZero) LOF_MXB(LOF_RoundAwayFromZero) LOF_MXB(LOF_RoundToOdd)
            default: break;
            }
#undef LOF_MXB
#undef LOF_MXB2
#undef LOF_MXB3
#undef LOF_MXB4
        }
        if (elem_codec.lanes && elem_codec.to_signed && (elem_codec.n == 8 || elem_codec.n == 4 || elem_codec.n == 6)) {
            const bool nz = elem_codec.sign_is_nan != 0;         // P3109-style element formats: no -0
#define LOF_MXL4(M, L, N, S) mx_encode_lane_kernel<M, L, N, S><<<grid2, kPackThreads, 0, stream>>>(vin, nvec, a, sq, scale_codec, elem_codec, rng_scale, first_index, scales, elems)
#define LOF_MXL3(M, L, N) do { if (nz) LOF_MXL4(M, L, N, 2); else LOF_MXL4(M, L, N, 1); } while (0)
#define LOF_MXL2(M, L) do { if (elem_codec.n == 8) LOF_MXL3(M, L, 8); else if (elem_codec.n == 4) LOF_MXL3(M, L, 4); else LOF_MXL3(M, L, 6); } while (0)
#define LOF_MXL(M) case M: if (lanes == 4) LOF_MXL2(M, 4); else if (lanes == 8) LOF_MXL2(M, 8); else LOF_MXL2(M, 16); return cudaGetLastError();
            switch (mode) {
                LOF_MXL(LOF_RoundToNearestEven) LOF_MXL(LOF_RoundTowardsZero) LOF_MXL(LOF_RoundAwayFromZero) LOF_MXL(LOF_RoundToOdd)
            default: break;
            }
#undef LOF_MXL
#undef LOF_MXL2
#undef LOF_MXL3
#undef LOF_MXL4
        }
#define LOF_CASE2(M, L) mx_encode_fast_kernel<M, L><<<grid2, kPackThreads, 0, stream>>>(vin, nvec, a, sq, scale_codec, elem_codec, rng_scale, first_index, scales, elems)
#define LOF_CASE(M) case M: if (lanes == 4) LOF_CASE2(M, 4); else if (lanes == 8) LOF_CASE2(M, 8); else LOF_CASE2(M, 16); return cudaGetLastError();
        switch (mode) {
            LOF_CASE(LOF_RoundToNearestEven) LOF_CASE(LOF_RoundTowardsZero) LOF_CASE(LOF_RoundAwayFromZero) LOF_CASE(LOF_RoundToOdd)
        default: break;
        }
#undef LOF_CASE
#undef LOF_CASE2
    }
    if (fast) {
        switch (mode) {
#define LOF_CASE(M) case M: mx_encode_kernel<M, true><<<grid, kPackThreads, 0, stream>>>(vin, nvec, a, sq, scale_codec, elem_codec, rng, rng_scale, first_index, scales, elems); return cudaGetLastError();
            LOF_CASE(LOF_RoundToNearestEven) LOF_CASE(LOF_RoundTowardsZero) LOF_CASE(LOF_RoundAwayFromZero) LOF_CASE(LOF_RoundToOdd)
#undef LOF_CASE
        default: break;
        }
    }
    switch (mode) {
#define LOF_CASE(M) case M: mx_encode_kernel<M, false><<<grid, kPackThreads, 0, stream>>>(vin, nvec, a, sq, scale_codec, elem_codec, rng, rng_scale, first_index, scales, elems); break;
        LOF_CASE(LOF_RoundToNearestEven) LOF_CASE(LOF_RoundTowardsZero) LOF_CASE(LOF_RoundAwayFromZero)
        LOF_CASE(LOF_StochasticRoundingA) LOF_CASE(LOF_RoundToNearestOdd) LOF_CASE(LOF_RoundDown)
        LOF_CASE(LOF_RoundUp) LOF_CASE(LOF_RoundTiesToAway) LOF_CASE(LOF_StochasticRoundingB)
        LOF_CASE(LOF_StochasticRoundingC) LOF_CASE(LOF_True_StochasticRounding)
        LOF_CASE(LOF_ProbabilisticRounding) LOF_CASE(LOF_RoundToOdd) LOF_CASE(LOF_RoundTiesTowardsZero)
        LOF_CASE(LOF_StochasticRoundingD)
#undef LOF_CASE
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t launch_mx_decode(const uint8_t* scales, const uint8_t* elems, float* out, int64_t n, int block_size, const CodecPlan& scale_codec,
                             const CodecPlan& elem_codec, int num_sms, cudaStream_t stream)
{
    if (elem_codec.lanes && elem_codec.dec_ok && elem_codec.to_signed && (block_size & (block_size - 1)) == 0 && block_size >= 4 && block_size <= kPackTile && n >= kPackTile &&
        (reinterpret_cast<uintptr_t>(elems) & 3u) == 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0 &&
        (elem_codec.n == 8 || elem_codec.n == 4 || elem_codec.n == 6)) {
        // whole tiles on the register-lane kernel, the remainder (< 4096 elements: whole blocks, byte-aligned packed offset) on mx_decode_kernel
        const int64_t ntiles = n / kPackTile;
        int log2_bs = 0;
        while ((1 << log2_bs) < block_size) ++log2_bs;
        const unsigned lgrid = pack_grid(ntiles * kPackTile, num_sms);
        float4* vout = reinterpret_cast<float4*>(out);
#define LOF_MXD(N) do { if (elem_codec.sign_is_nan) mx_decode_lane_kernel<N, 2><<<lgrid, kPackThreads, 0, stream>>>(scales, elems, vout, ntiles, log2_bs, scale_codec, elem_codec); \
                        else mx_decode_lane_kernel<N, 1><<<lgrid, kPackThreads, 0, stream>>>(scales, elems, vout, ntiles, log2_bs, scale_codec, elem_codec); } while (0)
        if (elem_codec.n == 8) LOF_MXD(8); else if (elem_codec.n == 4) LOF_MXD(4); else LOF_MXD(6);
#undef LOF_MXD
        cudaError_t e = cudaGetLastError();
        const int64_t done = ntiles * kPackTile;
        if (e != cudaSuccess || done == n) return e;
        scales += done / block_size; elems += (uint64_t)done / 8u * (uint64_t)elem_codec.n; out += done; n -= done;
    }
    const int64_t nvec = n / 4;
    const int64_t want = (nvec + kPackThreads - 1) / kPackThreads, cap = (int64_t)num_sms /* = SMs x resident CTAs per SM, chosen by the caller */;
    const unsigned grid = (unsigned)(want < 1 ? 1 : want < cap ? want : cap);
    mx_decode_kernel<<<grid, kPackThreads, 0, stream>>>(scales, elems, reinterpret_cast<float4*>(out), nvec, block_size, scale_codec, elem_codec);
    return cudaGetLastError();
}

cudaError_t launch_decode(const uint8_t* src, float* dst, int64_t count, uint64_t packed_bytes, const CodecPlan& p, int num_sms,
                          cudaStream_t stream)
{
    const int vec_ok = (reinterpret_cast<uintptr_t>(dst) & 15u) == 0;
    const int word_ok = (reinterpret_cast<uintptr_t>(src) & 3u) == 0;
    if (p.lanes && p.dec_ok && (p.n == 8 || p.n == 16 || p.n == 4 || p.n == 6) && vec_ok && (reinterpret_cast<uintptr_t>(src) & 7u) == 0 &&
        count >= kPackTile) {
        // whole tiles on the register-lane kernel, the remainder (< 4096 elements) on decode_kernel below
        const int64_t ntiles = count / kPackTile;
        const unsigned lgrid = pack_grid(ntiles * kPackTile, num_sms);
        float4* vdst = reinterpret_cast<float4*>(dst);
        const int sgn = !p.to_signed ? 0 : (p.sign_is_nan ? 2 : 1);
#define LOF_LANE3(N, S) decode_lane_kernel<N, S><<<lgrid, kPackThreads, 0, stream>>>(src, vdst, ntiles, p)
#define LOF_LANE2(N) do { if (sgn == 0) LOF_LANE3(N, 0); else if (sgn == 1) LOF_LANE3(N, 1); else LOF_LANE3(N, 2); } while (0)
        if (p.n == 8) LOF_LANE2(8); else if (p.n == 16) LOF_LANE2(16); else if (p.n == 4) LOF_LANE2(4); else LOF_LANE2(6);
#undef LOF_LANE2
#undef LOF_LANE3
        cudaError_t e = cudaGetLastError();
        const int64_t done = ntiles * kPackTile;
        if (e != cudaSuccess || done == count) return e;
        const uint64_t done_bytes = (uint64_t)done / 8u * (uint64_t)p.n;
        src += done_bytes; dst += done; count -= done; packed_bytes -= done_bytes;
    }
    decode_kernel<<<pack_grid(count, num_sms), kPackThreads, 0, stream>>>(src, dst, count, packed_bytes, p, vec_ok, word_ok);
    return cudaGetLastError();
}

}  // namespace lof
