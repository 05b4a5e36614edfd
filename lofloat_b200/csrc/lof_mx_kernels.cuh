// lof_mx_kernels.cuh -- MX (microscaling) block rounding kernels for sm_100a, fp32.
//
// Replaces /root/reference/src/mx_round.cuh:27-214 (+ launcher Lof_kernel.cu:70-94).
// Semantics follow the reference's CPU definition (LoPy_bind.cpp:197-215 ==
// lo_float.h:4311-4383): scale = virtual_round(amax / priv_max_normal, scale_fmt),
// out = virtual_round(x / scale, elem_fmt) * scale, a block whose scale is not > 0 becomes 0.
//
// One pass over HBM (4 B read + 4 B written per element, no scale tensor).  Two variants:
//
//  * "shuffle" (this file, mx_shuffle_kernel): a warp owns rows of 128 consecutive floats,
//    each lane one float4; the block amax is a __shfl_xor reduction over the BS/4 lanes of
//    a block.  Works for any pointer alignment >= 16 B and any n (scalar tail kernel for
//    the rest).
//  * "tma" (lof_mx_tma.cuh): TMA-staged [rows x 32 floats] tiles with a 128B swizzle, one
//    thread per 32-element row, no shuffles and no redundant scale computation.
#pragma once

#include "lof_core.cuh"
#include "lof_fast_f32.cuh"

namespace lof {

constexpr int kMxThreads = 256;

struct MxArgs {
    RoundPlan elem;        // private (element) format plan
    RoundPlan scale;       // public (scale) format plan
    FastPlan  fast;        // branch-free element rounding when fast.valid (lof_fast_f32.cuh)
    float     inv_pmn_unused;
    float     pmn;         // priv_max_normal (LoPy_bind.cpp:202)
    int32_t   scale_mode;  // rounding mode of the scale (stoch_len forced to 0, LoPy_bind.cpp:208)
    int32_t   block_size;
};

// scale of one block from its amax (group-uniform)
__device__ __forceinline__ float mx_block_scale(float amax, const MxArgs& a, const RngArgs& rng_scale, uint64_t gblock)
{
    const float ratio = __fdiv_rn(amax, a.pmn);
    uint32_t w = 0;
    if (mode_uses_rng(a.scale_mode)) w = philox_word(rng_scale, gblock);
    RoundPlan sp = a.scale;
    sp.stoch_len = 0;
    return __uint_as_float(round_params_elem_rt<LOF_F32>(a.scale_mode, __float_as_uint(ratio), sp, w));
}

// the generic routine out of line: the cold tail of mx_scale_quick
static __device__ __noinline__ float mx_block_scale_generic(float amax, const MxArgs& a, const RngArgs& rng_scale, uint64_t gblock)
{
    return mx_block_scale(amax, a, rng_scale, gblock);
}

// virtual_round(amax / pmn, scale format) on the branch-free routines of lof_fast_f32.cuh where they apply (the same
// selection lof_mx_fast.cu makes; every one of them is proven equal to the exact routine by the exhaustive sweeps),
// the generic 15-mode routine otherwise
struct MxScaleQuick {
    MagicPlan magic;
    FastPlan fast;
    int32_t quick_mode_ok;          // scale mode in {RNE, RTZ, RAZ, RU, RD, RoundToOdd}
};

__device__ __forceinline__ float mx_scale_quick(float amax, const MxArgs& a, const MxScaleQuick& q, const RngArgs& rng_scale, uint64_t gblock)
{
    if (q.quick_mode_ok) {
        const uint32_t x = __float_as_uint(__fdiv_rn(amax, a.pmn));          // >= 0, never NaN (fmaxf drops NaNs)
        const MagicPlan& p = q.magic;
        if (p.valid && !(x >= p.limit) && !((x - 1u) < p.lo_rare)) {
            uint32_t mag, rs;
            switch (a.scale_mode) {
            case LOF_RoundToNearestEven: rs = magic_round_core<LOF_RoundToNearestEven>(x, p.c_add, p.c_floor, mag); break;
            case LOF_RoundTowardsZero:   rs = magic_round_core<LOF_RoundTowardsZero>(x, p.c_add, p.c_floor, mag); break;
            case LOF_RoundAwayFromZero:  rs = magic_round_core<LOF_RoundAwayFromZero>(x, p.c_add, p.c_floor, mag); break;
            case LOF_RoundUp:            rs = magic_round_core<LOF_RoundUp>(x, p.c_add, p.c_floor, mag); break;
            case LOF_RoundDown:          rs = magic_round_core<LOF_RoundDown>(x, p.c_add, p.c_floor, mag); break;
            default:                     rs = magic_round_core<LOF_RoundToOdd>(x, p.c_add, p.c_floor, mag); break;
            }
            return __uint_as_float((mag >= p.ovf_thresh) ? (a.scale.saturating ? x : 0x7F800000u) : rs);
        }
        if (q.fast.valid) {
            bool rare = false;
            uint32_t rb;
            const bool sat = a.scale.saturating != 0;
            switch (a.scale_mode) {
            case LOF_RoundToNearestEven: rb = round_fast_f32_rt<LOF_RoundToNearestEven>(x, q.fast, sat, 0u, rare); break;
            case LOF_RoundTowardsZero:   rb = round_fast_f32_rt<LOF_RoundTowardsZero>(x, q.fast, sat, 0u, rare); break;
            case LOF_RoundAwayFromZero:  rb = round_fast_f32_rt<LOF_RoundAwayFromZero>(x, q.fast, sat, 0u, rare); break;
            case LOF_RoundUp:            rb = round_fast_f32_rt<LOF_RoundUp>(x, q.fast, sat, 0u, rare); break;
            case LOF_RoundDown:          rb = round_fast_f32_rt<LOF_RoundDown>(x, q.fast, sat, 0u, rare); break;
            default:                     rb = round_fast_f32_rt<LOF_RoundToOdd>(x, q.fast, sat, 0u, rare); break;
            }
            if (!rare) return __uint_as_float(rb);
        }
    }
    return mx_block_scale_generic(amax, a, rng_scale, gblock);
}

// Per-block constants for the element pass.  x / scale is evaluated as x * (1/scale) when
// scale is a power of two whose reciprocal is representable: both are the correctly
// rounded value of the same real number, so the result is bit-identical to the division.
struct MxBlockCtx {
    float scale;
    float recip;
    bool  live;      // scale > 0
    bool  pow2;
};

__device__ __forceinline__ MxBlockCtx mx_make_ctx(float scale)
{
    MxBlockCtx c;
    c.scale = scale;
    c.live = scale > 0.0f;
    const uint32_t sb = __float_as_uint(scale);
    const uint32_t e = (sb >> 23) & 0xFFu;
    c.pow2 = ((sb & 0x7FFFFFu) == 0u) && e >= 1u && e <= 254u;
    c.recip = __fdiv_rn(1.0f, scale);     // exact for every power of two in that range (2^-127 is a representable subnormal)
    return c;
}

template <int MODE>
__device__ __forceinline__ float mx_quant(float x, const MxBlockCtx& c, const MxArgs& a, uint32_t w)
{
    const float q = c.pow2 ? __fmul_rn(x, c.recip) : __fdiv_rn(x, c.scale);
    uint32_t rb;
    if (a.fast.valid) {
        bool rare = false;
        rb = round_fast_f32_rt<MODE>(__float_as_uint(q), a.fast, a.elem.saturating != 0, w, rare);
        if (rare) rb = exact_params_f32<MODE>(__float_as_uint(q), a.elem, w);
    } else {
        rb = exact_params_f32<MODE>(__float_as_uint(q), a.elem, w);
    }
    float r = __fmul_rn(__uint_as_float(rb), c.scale);
    // NaN bit patterns of the reference's CPU arithmetic (x86 SSE): a NaN operand propagates quieted (virtual_round
    // only ever produces +qNaN), an invalid operation (0 * inf with an infinite scale) yields the negative "indefinite"
    if (r != r) r = __uint_as_float((rb & 0x7FFFFFFFu) > 0x7F800000u ? 0x7FC00000u : 0xFFC00000u);
    return c.live ? r : 0.0f;
}

// Element quantisation on the magic-add routine (lof_fast_f32.cuh) for a block whose scale is a normal power of two
// 2^k: x / scale is the exact product x * 2^-k, the rounding is one magic add, the result is scaled back exactly.
// Same sequence as lof_mx_fast.cu's quant_fast (pinned there by the bit-exact MX tests); the caller guarantees that the
// element (format, mode) is magic-eligible without the checked variants.
template <int MODE>
__device__ __forceinline__ float mx_quant_magic(float x, float recip, float scale, const MagicPlan& em, bool sat)
{
    const float q = __fmul_rn(x, recip);
    uint32_t mag;
    const uint32_t rs = magic_round_core<MODE>(__float_as_uint(q), em.c_add, em.c_floor, mag);
    const uint32_t rb = (mag >= em.ovf_thresh) ? (sat ? __float_as_uint(q) : 0x7F800000u) : rs;
    const float o = __fmul_rn(__uint_as_float(rb), scale);
    return (x != x) ? __uint_as_float(0x7FC00000u) : o;                     // NaN element -> +qNaN
}

__device__ __forceinline__ float amax4(const float4& v)
{
    return fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
}

__device__ __forceinline__ float4 ld_stream_f4(const float4* p)
{
    float4 v;
    asm volatile("ld.global.cs.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream_f4(float4* p, const float4& v)
{
    asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ---------------------------------------------------------------------------------------
// shuffle variant.  BS in {16, 32, 64, 128, 256}.  A "row" is 128 consecutive floats = one
// float4 per lane; ROWS rows are in flight per warp iteration.  nrows = n / 128 full rows;
// the remainder (< 128 elements, possibly a partial MX block) goes to mx_tail_kernel.
// For BS = 256 two consecutive rows form one MX block (ROWS is even).
// ---------------------------------------------------------------------------------------
template <int BS, int MODE, int ROWS>
__global__ void __launch_bounds__(kMxThreads)
mx_shuffle_kernel(const float4* in, float4* out, int64_t nrows, MxArgs a, RngArgs rng, RngArgs rng_scale,
                  uint64_t first_index)
{
    static_assert(ROWS % 2 == 0, "ROWS must be even (BS = 256 pairs rows)");
    constexpr int kWarps = kMxThreads / 32;
    constexpr int GROUP = (BS >= 128) ? 32 : BS / 4;        // lanes that share one MX block within a row
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int64_t nsteps = (nrows + ROWS - 1) / ROWS;       // ROWS consecutive rows per warp step

    for (int64_t step = (int64_t)blockIdx.x * kWarps + warp; step < nsteps; step += (int64_t)gridDim.x * kWarps) {
        const int64_t row0 = step * ROWS;
        float4 v[ROWS];
        float m[ROWS];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            if (row0 + r < nrows) v[r] = ld_stream_f4(in + (row0 + r) * 32 + lane);
            else v[r] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            float t = amax4(v[r]);
#pragma unroll
            for (int off = GROUP / 2; off > 0; off >>= 1) t = fmaxf(t, __shfl_xor_sync(0xFFFFFFFFu, t, off));
            m[r] = t;
        }
        if constexpr (BS == 256) {
#pragma unroll
            for (int r = 0; r < ROWS; r += 2) { const float t = fmaxf(m[r], m[r + 1]); m[r] = t; m[r + 1] = t; }
        }
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            if (row0 + r >= nrows) continue;                 // warp-uniform
            const uint64_t g0 = first_index + (uint64_t)(row0 + r) * 128u + (uint64_t)lane * 4u;
            const float scale = mx_block_scale(m[r], a, rng_scale, g0 / (uint64_t)BS);
            const MxBlockCtx c = mx_make_ctx(scale);
            uint4 w = make_uint4(0, 0, 0, 0);
            if constexpr (ModeTraits<MODE>::kUsesRng) w = philox_group4(rng, g0 >> 2);
            float4 o;
            o.x = mx_quant<MODE>(v[r].x, c, a, w.x);
            o.y = mx_quant<MODE>(v[r].y, c, a, w.y);
            o.z = mx_quant<MODE>(v[r].z, c, a, w.z);
            o.w = mx_quant<MODE>(v[r].w, c, a, w.w);
            st_stream_f4(out + (row0 + r) * 32 + lane, o);
        }
    }
}

// Tail / unaligned fallback: one CTA-wide loop, one thread per MX block, scalar accesses.
// Handles n not a multiple of 128, a trailing partial block, and unaligned pointers.
template <int MODE>
__global__ void __launch_bounds__(kMxThreads)
mx_block_per_thread_kernel(const float* in, float* out, int64_t elem0, int64_t n, MxArgs a, RngArgs rng,
                           RngArgs rng_scale, uint64_t first_index)
{
    const int bs = a.block_size;
    const int64_t nblocks = (n - elem0 + bs - 1) / bs;
    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < nblocks; b += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = elem0 + b * bs;
        const int len = (int)((n - s) < bs ? (n - s) : bs);
        float amax = 0.0f;
        for (int i = 0; i < len; ++i) amax = fmaxf(amax, fabsf(in[s + i]));
        const float scale = mx_block_scale(amax, a, rng_scale, (first_index + (uint64_t)s) / (uint64_t)bs);
        const MxBlockCtx c = mx_make_ctx(scale);
        for (int i = 0; i < len; ++i) {
            uint32_t w = 0;
            if constexpr (ModeTraits<MODE>::kUsesRng) w = philox_event(rng, first_index + (uint64_t)(s + i));
            out[s + i] = mx_quant<MODE>(in[s + i], c, a, w);
        }
    }
}

struct MxLaunchCfg {
    int variant;        // 0 = shuffle, 1 = TMA-staged
    int ctas_per_sm;
    int num_sms;
};

cudaError_t launch_mx_shuffle(int mode, const float* in, float* out, int64_t n, const MxArgs& a, const RngArgs& rng,
                              const RngArgs& rng_scale, uint64_t first_index, const MxLaunchCfg& cfg,
                              cudaStream_t stream, int* launches);

}  // namespace lof
