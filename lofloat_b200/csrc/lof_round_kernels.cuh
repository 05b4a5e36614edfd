// lof_round_kernels.cuh -- elementwise virtual_round kernels for sm_100a.
//
// Replaces the reference's one-thread-per-element, scalar-load kernels
// (/root/reference/src/Lof_kernel.cu:37-60).  This is an HBM-bound streaming op
// (8 B/element for fp32: 4 read + 4 written), so the design is about keeping enough bytes
// in flight and issuing as few instructions per byte as possible:
//   * 128-bit ld.global / st.global with streaming (evict-first) hints, fully coalesced;
//   * UNROLL independent 16-byte loads per thread issued before any use;
//   * a persistent, grid-strided CTA set sized as a multiple of the SM count;
//   * rounding mode and source type are template parameters (no per-element switch), the
//     format constants sit in the kernel parameter bank as a RoundPlan;
//   * stochastic modes draw ONE Philox4x32-10 block per 4 consecutive elements, keyed on
//     the global element index (results are independent of the launch geometry).
#pragma once
#include <cuda_fp16.h>

#include "lof_core.cuh"
#include "lof_fast_f32.cuh"

namespace lof {

constexpr int kRoundThreads = 256;

__device__ __forceinline__ uint4 ld_stream(const uint4* p)
{
    uint4 v;
    asm volatile("ld.global.cs.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream(uint4* p, const uint4& v)
{
    asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w));
}

// What the kernel does per element.
enum : int {
    kKindMantissa = 0,      // virtual_round(value, mantissa_bits)
    kKindParams = 1,        // virtual_round(value, params): exact integer routine, any format / container
    kKindFastSat = 2,       // fp32 branch-free path (lof_fast_f32.cuh), OV_behavior == Saturating
    kKindFastExt = 3,       //   "                                      , OV_behavior == Extended
    kKindFastSatChk = 4,    //   + fp32-subnormal inputs routed to the exact routine (denorm_min < 2^-126)
    kKindFastExtChk = 5,
    kKindMagicSat = 6,      // fp32 "magic add" path (lof_fast_f32.cuh, second generation): 6 IEEE-like modes only
    kKindMagicExt = 7,
    kKindMagicSatChk = 8,   //   + exact routine for fp32-subnormal / out-of-range-constant inputs
    kKindMagicExtChk = 9
};
constexpr bool kind_is_fast(int k) { return k >= kKindFastSat && k <= kKindFastExtChk; }
constexpr bool kind_is_magic(int k) { return k >= kKindMagicSat; }

template <int DTYPE, int MODE, int KIND>
struct ElemOp {
    using S = Src<DTYPE>;
    using B = typename S::bits_t;
    RoundPlan plan;                 // params overloads
    FastPlan fast;                  // fast fp32 kinds
    MagicPlan magic;                // magic-add fp32 kinds
    int r;                          // mantissa overload: bits to drop
    int len;
    typename S::scale_t scale;
    bool scale_one;

    // exact routines (any container)
    __device__ __forceinline__ B operator()(B v, uint32_t w) const
    {
        if (!scale_one) v = S::apply_scale(v, scale);   // Lof_kernel.cu:41,53 fused pre-scale
        if constexpr (KIND == kKindMantissa) return round_mantissa_elem<DTYPE, MODE>(v, r, len, w);
        else return round_params_elem<DTYPE, MODE>(v, plan, w);          // also the scalar tail / rare lanes of the fast kinds
    }
};

template <int MODE, int KIND>
__device__ __forceinline__ uint4 round_vec16_fast(uint4 v, const ElemOp<LOF_F32, MODE, KIND>& op, const uint4& w)
{
    constexpr bool SAT = (KIND == kKindFastSat || KIND == kKindFastSatChk);
    constexpr bool CHK = (KIND == kKindFastSatChk || KIND == kKindFastExtChk);
    // scale * x unconditionally: x * 1.0f is the identity on every non-NaN bit pattern and NaN ends as +qNaN anyway
    const uint32_t x0 = __float_as_uint(__fmul_rn(op.scale, __uint_as_float(v.x)));
    const uint32_t x1 = __float_as_uint(__fmul_rn(op.scale, __uint_as_float(v.y)));
    const uint32_t x2 = __float_as_uint(__fmul_rn(op.scale, __uint_as_float(v.z)));
    const uint32_t x3 = __float_as_uint(__fmul_rn(op.scale, __uint_as_float(v.w)));
    bool rare = false;
    uint4 o;
    o.x = round_fast_f32<MODE, SAT, CHK>(x0, op.fast, w.x, rare);
    o.y = round_fast_f32<MODE, SAT, CHK>(x1, op.fast, w.y, rare);
    o.z = round_fast_f32<MODE, SAT, CHK>(x2, op.fast, w.z, rare);
    o.w = round_fast_f32<MODE, SAT, CHK>(x3, op.fast, w.w, rare);
    if (rare) {     // an fp32-subnormal input in this vector: redo it with the exact routine
        o.x = exact_params_f32<MODE>(x0, op.plan, w.x);
        o.y = exact_params_f32<MODE>(x1, op.plan, w.y);
        o.z = exact_params_f32<MODE>(x2, op.plan, w.z);
        o.w = exact_params_f32<MODE>(x3, op.plan, w.w);
    }
    return o;
}

template <int MODE, int KIND>
__device__ __forceinline__ uint4 round_vec16_magic(uint4 v, const ElemOp<LOF_F32, MODE, KIND>& op, const uint4& w)
{
    constexpr bool SAT = (KIND == kKindMagicSat || KIND == kKindMagicSatChk);
    constexpr bool CHK = (KIND == kKindMagicSatChk || KIND == kKindMagicExtChk);
    const uint32_t x0 = __float_as_uint(__fmul_rn(op.scale, __uint_as_float(v.x)));
    const uint32_t x1 = __float_as_uint(__fmul_rn(op.scale, __uint_as_float(v.y)));
    const uint32_t x2 = __float_as_uint(__fmul_rn(op.scale, __uint_as_float(v.z)));
    const uint32_t x3 = __float_as_uint(__fmul_rn(op.scale, __uint_as_float(v.w)));
    bool rare = false;
    uint4 o;
    if constexpr (CHK) {
        // checked kinds: ONE range test per vector on running extrema; NaN inputs trip it too (the host clamps `limit` to 0x7F800001
        // for these kinds) and take the exact routine, so the per-lane NaN select disappears as well
        uint32_t mn = 0xFFFFFFFFu, mx = 0u;
        o.x = round_magic_f32_acc<MODE, SAT>(x0, op.magic, w.x, mn, mx);
        o.y = round_magic_f32_acc<MODE, SAT>(x1, op.magic, w.y, mn, mx);
        o.z = round_magic_f32_acc<MODE, SAT>(x2, op.magic, w.z, mn, mx);
        o.w = round_magic_f32_acc<MODE, SAT>(x3, op.magic, w.w, mn, mx);
        rare = (mn < op.magic.lo_rare) || (mx >= op.magic.limit);
    } else {
        o.x = round_magic_f32<MODE, SAT, CHK>(x0, op.magic, w.x, rare);
        o.y = round_magic_f32<MODE, SAT, CHK>(x1, op.magic, w.y, rare);
        o.z = round_magic_f32<MODE, SAT, CHK>(x2, op.magic, w.z, rare);
        o.w = round_magic_f32<MODE, SAT, CHK>(x3, op.magic, w.w, rare);
    }
    if constexpr (CHK) {
        if (rare) {
            o.x = exact_params_f32<MODE>(x0, op.plan, w.x);
            o.y = exact_params_f32<MODE>(x1, op.plan, w.y);
            o.z = exact_params_f32<MODE>(x2, op.plan, w.z);
            o.w = exact_params_f32<MODE>(x3, op.plan, w.w);
        }
    }
    return o;
}

// 4-byte containers: one 16-byte vector with its four event words already drawn
template <int DTYPE, int MODE, int KIND>
__device__ __forceinline__ uint4 round_vec16_w(uint4 v, const ElemOp<DTYPE, MODE, KIND>& op, const uint4& w)
{
    if constexpr (kind_is_fast(KIND)) return round_vec16_fast<MODE, KIND>(v, op, w);
    else if constexpr (kind_is_magic(KIND)) return round_vec16_magic<MODE, KIND>(v, op, w);
    else {
        v.x = op(v.x, w.x); v.y = op(v.y, w.y); v.z = op(v.z, w.z); v.w = op(v.w, w.w);
        return v;
    }
}

// bf16 container on the fp32 branch-free routines, event words of the 8 elements given (w0: elements 0..3, w1: 4..7): a bf16
// value is the fp32 value (bits << 16); rounding it into a target of <= 7 mantissa bits gives an fp32 result whose low 16 bits
// are zero, i.e. the bf16 result.  Deterministic modes and SR-A / SR-D (host-checked; the random offset R << (r - len) is the
// same number in both containers); fp32-subnormal inputs and whatever the fp32 routine flags go to the exact bf16 routine.
// Equality with the exact container routine: all 65536 bf16 patterns x formats x modes, tests/test_gpu_parity.py.
// exact container routine for one 8-element bf16 vector: the cold path of the fast kinds below, kept out of line so that the
// streaming loop carries ONE rarely taken call instead of sixteen inlined copies of the integer routine
template <int MODE, int KIND>
__device__ __noinline__ uint4 round_vec16_bf16_exact(uint4 v, const ElemOp<LOF_BF16, MODE, KIND>& op, const uint4& w0, const uint4& w1)
{
    auto two = [&](uint32_t packed, uint32_t wa, uint32_t wb) -> uint32_t {
        const uint16_t lo = round_params_elem<LOF_BF16, MODE>((uint16_t)(packed & 0xFFFFu), op.plan, wa);
        const uint16_t hi = round_params_elem<LOF_BF16, MODE>((uint16_t)(packed >> 16), op.plan, wb);
        return (uint32_t)lo | ((uint32_t)hi << 16);
    };
    v.x = two(v.x, w0.x, w0.y); v.y = two(v.y, w0.z, w0.w);
    v.z = two(v.z, w1.x, w1.y); v.w = two(v.w, w1.z, w1.w);
    return v;
}

// bf16 container on the fp32 branch-free routines, event words of the 8 elements given (w0: elements 0..3, w1: 4..7): a bf16
// value is the fp32 value (bits << 16); rounding it into a target of <= 7 mantissa bits gives an fp32 result whose low 16 bits
// are zero, i.e. the bf16 result.  Deterministic modes and SR-A / SR-D (host-checked; the random offset R << (r - len) is the
// same number in both containers); fp32-subnormal inputs and whatever the fp32 routine flags set `rare`: the caller then redoes the
// whole vector with the exact bf16 routine (round_vec16_bf16_exact).
// Equality with the exact container routine: all 65536 bf16 patterns x formats x modes, tests/test_gpu_parity.py.
template <int MODE, int KIND>
__device__ __forceinline__ uint4 round_vec16_bf16_w(uint4 v, const ElemOp<LOF_BF16, MODE, KIND>& op, const uint4& w0, const uint4& w1, bool& rare)
{
    constexpr bool SAT = (KIND == kKindFastSat || KIND == kKindFastSatChk || KIND == kKindMagicSat || KIND == kKindMagicSatChk);
    // unchecked magic kinds (narrow targets such as the 8 / 6 / 4-bit formats, deterministic modes): the magic add is exact
    // for every input incl. the fp32 subnormals, so no lane ever needs the container routine -- ~7 alu-pipe instructions
    // per element less, which is what bounds a 4 B / element stream
    constexpr bool CHK = !(KIND == kKindMagicSat || KIND == kKindMagicExt);
    if constexpr (kind_is_magic(KIND) && CHK) {
        // checked magic kinds: ONE range test per vector on running extrema (the host folds the bf16-subnormal inputs into lo_rare and
        // NaN into limit for this container, lof_api.cu round_common)
        uint32_t mn = 0xFFFFFFFFu, mx = 0u;
        auto two = [&](uint32_t packed, uint32_t wa, uint32_t wb) -> uint32_t {
            const uint32_t olo = round_magic_f32_acc<MODE, SAT>(packed << 16, op.magic, wa, mn, mx);
            const uint32_t ohi = round_magic_f32_acc<MODE, SAT>(packed & 0xFFFF0000u, op.magic, wb, mn, mx);
            return __byte_perm(olo, ohi, 0x7632);
        };
        uint4 o;
        o.x = two(v.x, w0.x, w0.y); o.y = two(v.y, w0.z, w0.w);
        o.z = two(v.z, w1.x, w1.y); o.w = two(v.w, w1.z, w1.w);
        rare |= (mn < op.magic.lo_rare) || (mx >= op.magic.limit);
        return o;
    }
    auto two = [&](uint32_t packed, uint32_t wa, uint32_t wb) -> uint32_t {
        const uint32_t xlo = packed << 16, xhi = packed & 0xFFFF0000u;
        if constexpr (CHK) rare |= ((xlo & 0x7FFFFFFFu) - 1u) < 0x007FFFFFu || ((xhi & 0x7FFFFFFFu) - 1u) < 0x007FFFFFu;
        uint32_t olo, ohi;
        if constexpr (kind_is_magic(KIND)) {
            olo = round_magic_f32<MODE, SAT, CHK>(xlo, op.magic, wa, rare);
            ohi = round_magic_f32<MODE, SAT, CHK>(xhi, op.magic, wb, rare);
        } else {
            olo = round_fast_f32<MODE, SAT, true>(xlo, op.fast, wa, rare);
            ohi = round_fast_f32<MODE, SAT, true>(xhi, op.fast, wb, rare);
        }
        return __byte_perm(olo, ohi, 0x7632);                 // (olo >> 16) | (ohi & 0xFFFF0000)
    };
    uint4 o;
    o.x = two(v.x, w0.x, w0.y); o.y = two(v.y, w0.z, w0.w);
    o.z = two(v.z, w1.x, w1.y); o.w = two(v.w, w1.z, w1.w);
    return o;
}
template <int MODE, int KIND>
__device__ __forceinline__ uint4 round_vec16_bf16_w(uint4 v, const ElemOp<LOF_BF16, MODE, KIND>& op, const uint4& w0, const uint4& w1)
{
    bool rare = false;
    const uint4 o = round_vec16_bf16_w<MODE, KIND>(v, op, w0, w1, rare);
    if (rare) return round_vec16_bf16_exact<MODE, KIND>(v, op, w0, w1);
    return o;
}

// process one 16-byte vector that starts at global element index g0
template <int DTYPE, int MODE, int KIND>
__device__ __forceinline__ uint4 round_vec16(uint4 v, const ElemOp<DTYPE, MODE, KIND>& op,
                                             const RngArgs& rng, uint64_t g0)
{
    using B = typename Src<DTYPE>::bits_t;
    constexpr bool kRng = ModeTraits<MODE>::kUsesRng;
    if constexpr (sizeof(B) == 4) {
        uint4 w = make_uint4(0, 0, 0, 0);
        if constexpr (kRng) w = philox_group4(rng, g0 >> 2);         // g0 % 4 == 0 (host-checked)
        return round_vec16_w<DTYPE, MODE, KIND>(v, op, w);
    } else if constexpr (DTYPE == LOF_BF16 && (kind_is_fast(KIND) || kind_is_magic(KIND))) {
        uint4 w0 = make_uint4(0, 0, 0, 0), w1 = w0;               // event words of the 8 elements (stochastic kinds: SR-A / SR-D, magic path)
        if constexpr (kRng) { w0 = philox_group4(rng, g0 >> 2); w1 = philox_group4(rng, (g0 >> 2) + 1); }
        return round_vec16_bf16_w<MODE, KIND>(v, op, w0, w1);
    } else if constexpr (DTYPE == LOF_F16 && (KIND == kKindMagicSat || KIND == kKindMagicExt)) {
        // fp16 container on the fp32 magic-add routine: half -> float is exact, and the host only selects this kind when every
        // result is representable in fp16 again (target mantissa <= 10 bits, largest exponent <= 15, denorm_min >= 2^-14 -- the
        // last also keeps fp16-subnormal inputs below denorm_min, where the container routine is value-based too).  Deterministic
        // modes, no fused scale.  Equality with the exact container routine: all 65536 patterns, tests/test_gpu_parity.py.
        constexpr bool SAT = (KIND == kKindMagicSat);
        auto two = [&](uint32_t packed) -> uint32_t {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&packed));
            bool rare = false;
            const uint32_t olo = round_magic_f32<MODE, SAT, false>(__float_as_uint(f.x), op.magic, 0u, rare);
            const uint32_t ohi = round_magic_f32<MODE, SAT, false>(__float_as_uint(f.y), op.magic, 0u, rare);
            const __half2 o = __floats2half2_rn(__uint_as_float(olo), __uint_as_float(ohi));
            const uint32_t nan = __hneu2_mask(o, o);              // 0xFFFF per NaN lane: cvt gives 0x7FFF, the container's qNaN is 0x7E00
            return (*reinterpret_cast<const uint32_t*>(&o) & ~nan) | (0x7E007E00u & nan);
        };
        v.x = two(v.x); v.y = two(v.y); v.z = two(v.z); v.w = two(v.w);
    } else if constexpr (sizeof(B) == 2) {
        uint4 w0 = make_uint4(0, 0, 0, 0), w1 = w0;
        if constexpr (kRng) { w0 = philox_group4(rng, g0 >> 2); w1 = philox_group4(rng, (g0 >> 2) + 1); }
        auto two = [&](uint32_t packed, uint32_t wa, uint32_t wb) -> uint32_t {
            const uint16_t lo = op((uint16_t)(packed & 0xFFFFu), wa);
            const uint16_t hi = op((uint16_t)(packed >> 16), wb);
            return (uint32_t)lo | ((uint32_t)hi << 16);
        };
        v.x = two(v.x, w0.x, w0.y); v.y = two(v.y, w0.z, w0.w);
        v.z = two(v.z, w1.x, w1.y); v.w = two(v.w, w1.z, w1.w);
    } else {
        uint32_t wa = 0, wb = 0;
        if constexpr (kRng) {                                         // g0 % 2 == 0
            const uint4 w = philox_group4(rng, g0 >> 2);
            const bool hi = (g0 & 2u) != 0;
            wa = hi ? w.z : w.x; wb = hi ? w.w : w.y;
        }
        const uint64_t a = op(((uint64_t)v.y << 32) | v.x, wa);
        const uint64_t b = op(((uint64_t)v.w << 32) | v.z, wb);
        v.x = (uint32_t)a; v.y = (uint32_t)(a >> 32); v.z = (uint32_t)b; v.w = (uint32_t)(b >> 32);
    }
    return v;
}

// Vector kernel: in/out 16-byte aligned, nvec full vectors, then `ntail` scalar elements.
template <int DTYPE, int MODE, int KIND, int UNROLL>
__global__ void __launch_bounds__(kRoundThreads)
round_vec_kernel(const uint4* in, uint4* out, int64_t nvec, int ntail,
                 ElemOp<DTYPE, MODE, KIND> op, RngArgs rng, uint64_t first_index)
{
    using B = typename Src<DTYPE>::bits_t;
    constexpr int VEC = 16 / (int)sizeof(B);
    constexpr int64_t TILE = (int64_t)kRoundThreads * UNROLL;
    const int64_t ntiles = (nvec + TILE - 1) / TILE;

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        if constexpr (ModeTraits<MODE>::kUsesRng && DTYPE == LOF_BF16 && UNROLL == 4 && (kind_is_fast(KIND) || kind_is_magic(KIND))) {
            // 16-bit containers, class-0 events: a tile is 8192 elements = two aligned runs of 4096; thread t takes, in run
            // t >> 7, the vectors w * 128 + (t & 127), w = 0..3, i.e. the 4-element groups v = w * 256 + 2 (t & 127) + {0, 1}
            // of that run -- all four words of the two Philox blocks (run, 2 (t & 127)) and (run, 2 (t & 127) + 1): one block
            // per 16 elements like the fp32 layout (two blocks per 8-element vector before).  Same event stream as philox_group4.
            if (rng.cls == 0u && (first_index & 4095u) == 0u && (tile + 1) * TILE <= nvec) {
                const uint32_t t = threadIdx.x, kk = t & 127u;
                const uint64_t run = (first_index >> 12) + (uint64_t)tile * 2u + (t >> 7);
                const int64_t vbase = tile * TILE + (int64_t)(t >> 7) * 512 + kk;
                uint4 q[UNROLL];
#pragma unroll
                for (int w = 0; w < UNROLL; ++w) q[w] = ld_stream(in + vbase + w * 128);
                const uint4 ba = philox_block(rng, (run << 8) | (2u * kk)), bb = philox_block(rng, (run << 8) | (2u * kk + 1u));
                const uint32_t wa[4] = {ba.x, ba.y, ba.z, ba.w}, wb[4] = {bb.x, bb.y, bb.z, bb.w};
                bool rare = false;
                uint4 o[UNROLL];
#pragma unroll
                for (int w = 0; w < UNROLL; ++w) o[w] = round_vec16_bf16_w<MODE, KIND>(q[w], op, bytes_of(wa[w]), bytes_of(wb[w]), rare);
                if (rare) {             // an fp32-subnormal / flagged lane somewhere in this thread's 32 elements: redo them exactly
#pragma unroll 1
                    for (int w = 0; w < UNROLL; ++w) o[w] = round_vec16_bf16_exact<MODE, KIND>(q[w], op, bytes_of(wa[w]), bytes_of(wb[w]));
                }
#pragma unroll
                for (int w = 0; w < UNROLL; ++w) st_stream(out + vbase + w * 128, o[w]);
                continue;
            }
        }
        const int64_t base = tile * TILE + threadIdx.x;
        uint4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int64_t idx = base + (int64_t)u * kRoundThreads;
            if (idx < nvec) v[u] = ld_stream(in + idx);
        }
        if constexpr (ModeTraits<MODE>::kUsesRng && sizeof(B) == 4 && UNROLL == 4) {
            // class-0 events (<= 8 random bits): the 4 vectors of this thread in this 4096-element tile share ONE
            // Philox block by construction of the event stream (lof_core.cuh) when the tile grid is aligned with it
            if (rng.cls == 0u && (first_index & 4095u) == 0u) {
                const uint4 b = philox_block(rng, ((first_index >> 12) + (uint64_t)tile) * 256u + threadIdx.x);
                const uint32_t wu[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int u = 0; u < UNROLL; ++u) {
                    const int64_t idx = base + (int64_t)u * kRoundThreads;
                    if (idx < nvec) st_stream(out + idx, round_vec16_w<DTYPE, MODE, KIND>(v[u], op, bytes_of(wu[u])));
                }
                continue;
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int64_t idx = base + (int64_t)u * kRoundThreads;
            if (idx < nvec) {
                const uint4 r = round_vec16<DTYPE, MODE, KIND>(v[u], op, rng, first_index + (uint64_t)idx * VEC);
                st_stream(out + idx, r);
            }
        }
    }
    // scalar tail (< VEC elements), handled by the first threads of CTA 0
    if (blockIdx.x == 0 && (int)threadIdx.x < ntail) {
        const int64_t i = nvec * VEC + threadIdx.x;
        const B* sin = reinterpret_cast<const B*>(in);
        B* sout = reinterpret_cast<B*>(out);
        uint32_t w = 0;
        if constexpr (ModeTraits<MODE>::kUsesRng) w = philox_event(rng, first_index + (uint64_t)i);
        sout[i] = op(sin[i], w);
    }
}

// Scalar kernel: any alignment, any first_index (still coalesced; one element per thread
// per grid-stride step).  Only used for unaligned views / odd shard offsets.
template <int DTYPE, int MODE, int KIND>
__global__ void __launch_bounds__(kRoundThreads)
round_scalar_kernel(const typename Src<DTYPE>::bits_t* in, typename Src<DTYPE>::bits_t* out,
                    int64_t n, ElemOp<DTYPE, MODE, KIND> op, RngArgs rng, uint64_t first_index)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint32_t w = 0;
        if constexpr (ModeTraits<MODE>::kUsesRng) w = philox_event(rng, first_index + (uint64_t)i);
        out[i] = op(in[i], w);
    }
}

struct RoundLaunchCfg {
    int unroll;         // 1, 2, 4 or 8 vectors in flight per thread
    int ctas_per_sm;    // persistent grid = SMs * ctas_per_sm
    int num_sms;
};

template <int DTYPE, int MODE, int KIND>
cudaError_t launch_round_mode(const void* in, void* out, int64_t n, const ElemOp<DTYPE, MODE, KIND>& op,
                              const RngArgs& rng, uint64_t first_index, const RoundLaunchCfg& cfg,
                              cudaStream_t stream)
{
    using B = typename Src<DTYPE>::bits_t;
    constexpr int VEC = 16 / (int)sizeof(B);
    const bool aligned = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0;
    const bool rng_ok = !ModeTraits<MODE>::kUsesRng || (first_index % (VEC < 4 ? VEC : 4)) == 0;
    const int64_t max_ctas = (int64_t)cfg.num_sms * cfg.ctas_per_sm;
    if (aligned && rng_ok && n >= VEC) {
        const int64_t nvec = n / VEC;
        const int ntail = (int)(n - nvec * VEC);
        const int64_t tile = (int64_t)kRoundThreads * cfg.unroll;
        int64_t grid = (nvec + tile - 1) / tile;
        // persistent grid: never more CTAs than are resident at once (a partial second wave would run alone)
        static int resident = 0;
        if (resident == 0) {
            int nb = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, round_vec_kernel<DTYPE, MODE, KIND, 4>, kRoundThreads, 0) != cudaSuccess || nb < 1) nb = 1;
            resident = nb;
        }
        const int64_t cap = (int64_t)cfg.num_sms * (resident < cfg.ctas_per_sm ? resident : cfg.ctas_per_sm);
        if (grid > cap) grid = cap;
        if (grid < 1) grid = 1;
        const uint4* vin = static_cast<const uint4*>(in);
        uint4* vout = static_cast<uint4*>(out);
        // the unroll depth is tunable for fp32 (the benchmarked container); the other containers
        // use 4 vectors in flight per thread
#define LOF_LAUNCH_VEC(U) round_vec_kernel<DTYPE, MODE, KIND, U><<<(unsigned)grid, kRoundThreads, 0, stream>>>( \
            vin, vout, nvec, ntail, op, rng, first_index)
        if constexpr (DTYPE == LOF_F32) {
#ifdef LOF_ROUND_TUNING_VARIANTS          // `make TUNING=1`: the 1 / 2 / 8-vector instantiations the tuning scripts sweep (3x the fp32 rounding kernels)
            switch (cfg.unroll) {
            case 1: LOF_LAUNCH_VEC(1); break;
            case 2: LOF_LAUNCH_VEC(2); break;
            case 8: LOF_LAUNCH_VEC(8); break;
            default: LOF_LAUNCH_VEC(4); break;
            }
#else                                     // product build: 4 vectors in flight per thread (the B200 optimum, scripts/geom_tune2.py); other settings map to it
            LOF_LAUNCH_VEC(4);
#endif
        } else {
            LOF_LAUNCH_VEC(4);
        }
#undef LOF_LAUNCH_VEC
    } else {
        int64_t grid = (n + kRoundThreads - 1) / kRoundThreads;
        if (grid > max_ctas * 2) grid = max_ctas * 2;
        if (grid < 1) grid = 1;
        round_scalar_kernel<DTYPE, MODE, KIND><<<(unsigned)grid, kRoundThreads, 0, stream>>>(
            static_cast<const B*>(in), static_cast<B*>(out), n, op, rng, first_index);
    }
    return cudaGetLastError();
}

// mode dispatch for one (dtype, overload); defined once per dtype in lof_round_<dtype>.cu
template <int DTYPE, int KIND>
cudaError_t launch_round(int mode, const void* in, void* out, int64_t n, const RoundPlan& plan, const FastPlan& fast, const MagicPlan& magic, int r, int len,
                         double scale, const RngArgs& rng, uint64_t first_index, const RoundLaunchCfg& cfg,
                         cudaStream_t stream)
{
    using S = Src<DTYPE>;
    const typename S::scale_t sc = S::make_scale(scale);
    switch (mode) {
#define LOF_CASE(M)                                                                           \
    case M: {                                                                                 \
        if constexpr (kind_is_magic(KIND) && !MagicTraits<M>::kOk) return cudaErrorInvalidValue; else { \
        ElemOp<DTYPE, M, KIND> op;                                                          \
        op.plan = plan; op.fast = fast; op.magic = magic; op.r = r; op.len = len; op.scale = sc; op.scale_one = (sc == (typename S::scale_t)1); \
        return launch_round_mode<DTYPE, M, KIND>(in, out, n, op, rng, first_index, cfg, stream); } \
    }
        LOF_CASE(LOF_RoundToNearestEven) LOF_CASE(LOF_RoundTowardsZero) LOF_CASE(LOF_RoundAwayFromZero)
        LOF_CASE(LOF_StochasticRoundingA) LOF_CASE(LOF_RoundToNearestOdd) LOF_CASE(LOF_RoundDown)
        LOF_CASE(LOF_RoundUp) LOF_CASE(LOF_RoundTiesToAway) LOF_CASE(LOF_StochasticRoundingB)
        LOF_CASE(LOF_StochasticRoundingC) LOF_CASE(LOF_True_StochasticRounding)
        LOF_CASE(LOF_ProbabilisticRounding) LOF_CASE(LOF_RoundToOdd) LOF_CASE(LOF_RoundTiesTowardsZero)
        LOF_CASE(LOF_StochasticRoundingD)
#undef LOF_CASE
    default: return cudaErrorInvalidValue;
    }
}

// explicit-instantiation entry points (one translation unit per dtype)
cudaError_t launch_round_f32(int kind, int mode, const void* in, void* out, int64_t n, const RoundPlan& plan,
                              const FastPlan& fast, const MagicPlan& magic, int r, int len, double scale, const RngArgs& rng, uint64_t first_index,
                             const RoundLaunchCfg& cfg, cudaStream_t stream);
cudaError_t launch_round_f16(int kind, int mode, const void* in, void* out, int64_t n, const RoundPlan& plan,
                              const FastPlan& fast, const MagicPlan& magic, int r, int len, double scale, const RngArgs& rng, uint64_t first_index,
                             const RoundLaunchCfg& cfg, cudaStream_t stream);
cudaError_t launch_round_bf16(int kind, int mode, const void* in, void* out, int64_t n, const RoundPlan& plan,
                              const FastPlan& fast, const MagicPlan& magic, int r, int len, double scale, const RngArgs& rng, uint64_t first_index,
                              const RoundLaunchCfg& cfg, cudaStream_t stream);
cudaError_t launch_round_f64(int kind, int mode, const void* in, void* out, int64_t n, const RoundPlan& plan,
                              const FastPlan& fast, const MagicPlan& magic, int r, int len, double scale, const RngArgs& rng, uint64_t first_index,
                             const RoundLaunchCfg& cfg, cudaStream_t stream);

}  // namespace lof
