// lof_mx_fast.cu -- second-generation MX block-rounding kernel (sm_100a, fp32): the HBM-bound fast path.
//
// Same results as lof_mx_kernels.cuh / the reference's CPU definition (LoPy_bind.cpp:197-215 ==
// lo_float.h:4311-4383), restricted to what MXFP8 / MXFP6 / MXFP4 actually use so that the per-element work
// fits under the HBM roofline (ncu of the first-generation kernels: 60 instructions / element, ALU pipe at 65 %,
// DRAM at 39 %, profiles/r01_*):
//   * element mode in {RNE, RTZ, RAZ, RoundUp, RoundDown, RoundToOdd} and a magic-add eligible element format
//     (lof_fast_f32.cuh): ~10 instructions per rounding, no regime branches;
//   * scale format without mantissa bits (E8M0-like), so a live scale is a power of two: x / scale is the exact
//     multiplication by 2^-k, built from the exponent field (no division, no reciprocal);
//   * amax / priv_max_normal by Markstein's correctly rounded reciprocal-FMA sequence (3 FMAs) in the range where
//     its residual is exact, IEEE division elsewhere;
//   * one lane holds 8 elements (two float4) of ONE MX block, so the block scale is computed once per 8
//     elements instead of once per 4, and BS/8 lanes share a block (BS = 32: 2 shuffle steps for the amax).
// Blocks whose scale is not a normal power of two (all-zero block, scale 0 / inf / 2^127) take a block-uniform
// slow path through the generic element routine.  Everything else (other modes, stochastic rounding, scale
// formats with mantissa bits, unaligned pointers, tails) stays on the first-generation kernels.
#include <cstring>

#include <atomic>
#include "lof_mx_kernels.cuh"

namespace lof {

namespace {

constexpr int kFastThreads = 256;
constexpr int kStepFloats = 256;            // floats per warp step: 32 lanes x 2 float4

struct MxFastArgs {
    MagicPlan elem_magic;
    MagicPlan scale_magic;
    FastPlan scale_fast;                    // used when the scale (format, mode) is not magic-eligible (E8M0 under RNE)
    float pmn, rcp_pmn;
    float div_lo, div_hi;                   // amax range in which the 3-FMA quotient is exact
    int32_t scale_mode;
    int32_t elem_sat;
};

// amax / pmn, correctly rounded (== __fdiv_rn).  q0 = a * r, rem = a - q0 * pmn (exact in an FMA while q0 and the
// residual stay normal), q = q0 + rem * r is the correctly rounded quotient (Markstein 1990) -- verified
// exhaustively over every positive fp32 amax for every MX element format by tests/test_gpu_parity.py.
__device__ __forceinline__ float div_pmn(float a, const MxFastArgs& k)
{
    if (a >= k.div_lo && a <= k.div_hi) {
        const float q0 = __fmul_rn(a, k.rcp_pmn);
        const float rem = __fmaf_rn(-q0, k.pmn, a);
        return __fmaf_rn(rem, k.rcp_pmn, q0);
    }
    return __fdiv_rn(a, k.pmn);
}

__device__ __noinline__ uint32_t scale_round_exact(int mode, uint32_t x, const MxArgs& g)
{
    RoundPlan sp = g.scale;
    sp.stoch_len = 0;
    return round_params_elem_rt<LOF_F32>(mode, x, sp, 0u);
}

// virtual_round(ratio, scale_fmt) for ratio >= 0 (never NaN: fmaxf drops NaNs), runtime mode
__device__ __forceinline__ uint32_t scale_round(float ratio, const MxFastArgs& k, const MxArgs& g)
{
    const uint32_t x = __float_as_uint(ratio);
    const MagicPlan& p = k.scale_magic;
    if (p.valid && !(x >= p.limit) && !(p.check_denorm_in && (x - 1u) < 0x007FFFFFu)) {
        uint32_t mag, rs;
        switch (k.scale_mode) {
        case LOF_RoundToNearestEven: rs = magic_round_core<LOF_RoundToNearestEven>(x, p.c_add, p.c_floor, mag); break;
        case LOF_RoundTowardsZero:   rs = magic_round_core<LOF_RoundTowardsZero>(x, p.c_add, p.c_floor, mag); break;
        case LOF_RoundAwayFromZero:  rs = magic_round_core<LOF_RoundAwayFromZero>(x, p.c_add, p.c_floor, mag); break;
        case LOF_RoundUp:            rs = magic_round_core<LOF_RoundUp>(x, p.c_add, p.c_floor, mag); break;
        case LOF_RoundDown:          rs = magic_round_core<LOF_RoundDown>(x, p.c_add, p.c_floor, mag); break;
        default:                     rs = magic_round_core<LOF_RoundToOdd>(x, p.c_add, p.c_floor, mag); break;
        }
        return (mag >= p.ovf_thresh) ? (g.scale.saturating ? x : 0x7F800000u) : rs;
    }
    if (k.scale_fast.valid) {                                    // first-generation branch-free path (integer normal regime)
        bool rare = false;
        uint32_t rb;
        const bool sat = g.scale.saturating != 0;
        switch (k.scale_mode) {
        case LOF_RoundToNearestEven: rb = round_fast_f32_rt<LOF_RoundToNearestEven>(x, k.scale_fast, sat, 0u, rare); break;
        case LOF_RoundTowardsZero:   rb = round_fast_f32_rt<LOF_RoundTowardsZero>(x, k.scale_fast, sat, 0u, rare); break;
        case LOF_RoundAwayFromZero:  rb = round_fast_f32_rt<LOF_RoundAwayFromZero>(x, k.scale_fast, sat, 0u, rare); break;
        case LOF_RoundUp:            rb = round_fast_f32_rt<LOF_RoundUp>(x, k.scale_fast, sat, 0u, rare); break;
        case LOF_RoundDown:          rb = round_fast_f32_rt<LOF_RoundDown>(x, k.scale_fast, sat, 0u, rare); break;
        default:                     rb = round_fast_f32_rt<LOF_RoundToOdd>(x, k.scale_fast, sat, 0u, rare); break;
        }
        if (!rare) return rb;
    }
    return scale_round_exact(k.scale_mode, x, g);
}

template <int MODE>
__device__ __noinline__ float mx_quant_generic(float x, float scale, const MxArgs& g)
{
    const MxBlockCtx c = mx_make_ctx(scale);
    return mx_quant<MODE>(x, c, g, 0u);
}

template <int MODE>
__device__ __forceinline__ float quant_fast(float x, float recip, float scale, const MxFastArgs& k)
{
    const float q = __fmul_rn(x, recip);                                    // exact: recip is a power of two
    uint32_t mag;
    const uint32_t rs = magic_round_core<MODE>(__float_as_uint(q), k.elem_magic.c_add, k.elem_magic.c_floor, mag);
    const uint32_t rb = (mag >= k.elem_magic.ovf_thresh) ? (k.elem_sat ? __float_as_uint(q) : 0x7F800000u) : rs;
    const float o = __fmul_rn(__uint_as_float(rb), scale);
    return (x != x) ? __uint_as_float(0x7FC00000u) : o;                     // NaN element -> +qNaN (x86 keeps the quieted operand)
}

// stochastic element modes (SR-A / SR-D): the same with the magic-add stochastic step (lof_fast_f32.cuh magic_round_sr), event word w
template <int MODE>
__device__ __forceinline__ float quant_fast_sr(float x, float recip, float scale, const MxFastArgs& k, uint32_t w)
{
    const float q = __fmul_rn(x, recip);
    uint32_t mag;
    const uint32_t rs = magic_round_sr<MODE>(__float_as_uint(q), k.elem_magic, w, mag);
    const uint32_t rb = (mag >= k.elem_magic.ovf_thresh) ? (k.elem_sat ? __float_as_uint(q) : 0x7F800000u) : rs;
    const float o = __fmul_rn(__uint_as_float(rb), scale);
    return (x != x) ? __uint_as_float(0x7FC00000u) : o;
}

// Stochastic element rounding (class-0 events: <= 8 random bits each) on the fast path.  The event stream gives the 4-element group v
// of a 4096-element run the word (v >> 8) of Philox block (run, v & 255) (lof_core.cuh philox_group4), so a lane that takes the SAME
// float4 position of the steps s, s + 4, s + 8, s + 12 of a 16-step run uses all four words of one block: a warp group walks a run with
// stride 4 and draws 2 blocks per lane for its 64 elements -- one block per 16 elements, like the elementwise kernels.  Scale rounding stays
// deterministic; first_index is a multiple of 4096 (host-checked); the steps beyond the last whole run go to the generic kernel.
template <int BS, int MODE>
__global__ void __launch_bounds__(kFastThreads)
mx_fast_sr_kernel(const float4* __restrict__ in, float4* __restrict__ out, int64_t nruns, MxArgs g, MxFastArgs k, RngArgs rng, uint64_t first_run)
{
    constexpr int G = BS / 8;
    constexpr int kWarps = kFastThreads / 32;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int blk = lane / G, j = lane % G;
    const int f0 = blk * (2 * G) + j;
    const int64_t ngroups = nruns * 4;
    for (int64_t grp = (int64_t)blockIdx.x * kWarps + warp; grp < ngroups; grp += (int64_t)gridDim.x * kWarps) {
        const int64_t run = grp >> 2;
        const int sub = (int)(grp & 3);
        float4 v[4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int64_t step = run * 16 + sub + 4 * i;
            v[i][0] = ld_stream_f4(in + step * 64 + f0);
            v[i][1] = ld_stream_f4(in + step * 64 + f0 + G);
        }
        const uint64_t blk0 = ((first_run + (uint64_t)run) << 8) | (uint32_t)(sub * 64 + f0);
        const uint4 ba = philox_block(rng, blk0), bb = philox_block(rng, blk0 + (uint32_t)G);
        const uint32_t wa[4] = {ba.x, ba.y, ba.z, ba.w}, wb[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int64_t step = run * 16 + sub + 4 * i;
            float m = fmaxf(amax4(v[i][0]), amax4(v[i][1]));
#pragma unroll
            for (int off = G / 2; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, off));
            const uint32_t sb = scale_round(div_pmn(m, k), k, g);
            const uint4 e0 = bytes_of(wa[i]), e1 = bytes_of(wb[i]);
            float4 o0, o1;
            if ((sb - 0x00800000u) < 0x7E800000u && (sb & 0x007FFFFFu) == 0u) {
                const float scale = __uint_as_float(sb);
                const float recip = __uint_as_float(0x7F000000u - sb);
                o0.x = quant_fast_sr<MODE>(v[i][0].x, recip, scale, k, e0.x); o0.y = quant_fast_sr<MODE>(v[i][0].y, recip, scale, k, e0.y);
                o0.z = quant_fast_sr<MODE>(v[i][0].z, recip, scale, k, e0.z); o0.w = quant_fast_sr<MODE>(v[i][0].w, recip, scale, k, e0.w);
                o1.x = quant_fast_sr<MODE>(v[i][1].x, recip, scale, k, e1.x); o1.y = quant_fast_sr<MODE>(v[i][1].y, recip, scale, k, e1.y);
                o1.z = quant_fast_sr<MODE>(v[i][1].z, recip, scale, k, e1.z); o1.w = quant_fast_sr<MODE>(v[i][1].w, recip, scale, k, e1.w);
            } else {
                const MxBlockCtx c = mx_make_ctx(__uint_as_float(sb));
                o0.x = mx_quant<MODE>(v[i][0].x, c, g, e0.x); o0.y = mx_quant<MODE>(v[i][0].y, c, g, e0.y);
                o0.z = mx_quant<MODE>(v[i][0].z, c, g, e0.z); o0.w = mx_quant<MODE>(v[i][0].w, c, g, e0.w);
                o1.x = mx_quant<MODE>(v[i][1].x, c, g, e1.x); o1.y = mx_quant<MODE>(v[i][1].y, c, g, e1.y);
                o1.z = mx_quant<MODE>(v[i][1].z, c, g, e1.z); o1.w = mx_quant<MODE>(v[i][1].w, c, g, e1.w);
            }
            st_stream_f4(out + step * 64 + f0, o0);
            st_stream_f4(out + step * 64 + f0 + G, o1);
        }
    }
}

template <int BS, int MODE>
cudaError_t launch_bs_sr(const float* in, float* out, int64_t nruns, const MxArgs& g, const MxFastArgs& k, const RngArgs& rng, uint64_t first_run,
                         const MxLaunchCfg& cfg, cudaStream_t stream)
{
    constexpr int kWarps = kFastThreads / 32;
    int64_t grid = (nruns * 4 + kWarps - 1) / kWarps;
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, mx_fast_sr_kernel<BS, MODE>, kFastThreads, 0) != cudaSuccess || nb < 1) nb = 1;
    const int64_t max_ctas = (int64_t)cfg.num_sms * (nb < cfg.ctas_per_sm ? nb : cfg.ctas_per_sm);
    if (grid > max_ctas) grid = max_ctas;
    mx_fast_sr_kernel<BS, MODE><<<(unsigned)grid, kFastThreads, 0, stream>>>(reinterpret_cast<const float4*>(in), reinterpret_cast<float4*>(out),
                                                                              nruns, g, k, rng, first_run);
    return cudaGetLastError();
}
template <int MODE>
cudaError_t launch_sr(const float* in, float* out, int64_t nruns, const MxArgs& g, const MxFastArgs& k, const RngArgs& rng, uint64_t first_run,
                      const MxLaunchCfg& cfg, cudaStream_t stream)
{
    switch (g.block_size) {
    case 16:  return launch_bs_sr<16, MODE>(in, out, nruns, g, k, rng, first_run, cfg, stream);
    case 32:  return launch_bs_sr<32, MODE>(in, out, nruns, g, k, rng, first_run, cfg, stream);
    case 64:  return launch_bs_sr<64, MODE>(in, out, nruns, g, k, rng, first_run, cfg, stream);
    case 128: return launch_bs_sr<128, MODE>(in, out, nruns, g, k, rng, first_run, cfg, stream);
    case 256: return launch_bs_sr<256, MODE>(in, out, nruns, g, k, rng, first_run, cfg, stream);
    default:  return cudaErrorInvalidValue;
    }
}

// BS in {16, 32, 64, 128, 256}.  A warp step covers 256 consecutive floats = 256 / BS blocks; G = BS / 8 lanes
// share a block and each lane holds the block's float4 number j and j + G (j = lane % G).
template <int BS, int MODE, int STEPS>
__global__ void __launch_bounds__(kFastThreads)
mx_fast_kernel(const float4* __restrict__ in, float4* __restrict__ out, int64_t nsteps_total, MxArgs g, MxFastArgs k)
{
    constexpr int G = BS / 8;
    constexpr int kWarps = kFastThreads / 32;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int blk = lane / G, j = lane % G;
    const int f0 = blk * (2 * G) + j;                  // float4 index inside the 64-float4 step
    const int64_t ngroups = (nsteps_total + STEPS - 1) / STEPS;

    for (int64_t grp = (int64_t)blockIdx.x * kWarps + warp; grp < ngroups; grp += (int64_t)gridDim.x * kWarps) {
        float4 v[STEPS][2];
#pragma unroll
        for (int s = 0; s < STEPS; ++s) {
            const int64_t step = grp * STEPS + s;
            if (step < nsteps_total) {
                v[s][0] = ld_stream_f4(in + step * 64 + f0);
                v[s][1] = ld_stream_f4(in + step * 64 + f0 + G);
            } else {
                v[s][0] = v[s][1] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int s = 0; s < STEPS; ++s) {
            const int64_t step = grp * STEPS + s;
            float m = fmaxf(amax4(v[s][0]), amax4(v[s][1]));
#pragma unroll
            for (int off = G / 2; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, off));
            if (step >= nsteps_total) continue;                                   // warp-uniform
            const uint32_t sb = scale_round(div_pmn(m, k), k, g);
            float4 o0, o1;
            // normal power of two, 2^-126 .. 2^126.  The mantissa test matters for m = 0 scale formats with a narrow exponent
            // range and Saturating overflow: virtual_round then returns the UNROUNDED quotient, which is
            // not a power of two -- such blocks take the generic routine (a real division).
            if ((sb - 0x00800000u) < 0x7E800000u && (sb & 0x007FFFFFu) == 0u) {
                const float scale = __uint_as_float(sb);
                const float recip = __uint_as_float(0x7F000000u - sb);            // 2^-k, exact
                o0.x = quant_fast<MODE>(v[s][0].x, recip, scale, k); o0.y = quant_fast<MODE>(v[s][0].y, recip, scale, k);
                o0.z = quant_fast<MODE>(v[s][0].z, recip, scale, k); o0.w = quant_fast<MODE>(v[s][0].w, recip, scale, k);
                o1.x = quant_fast<MODE>(v[s][1].x, recip, scale, k); o1.y = quant_fast<MODE>(v[s][1].y, recip, scale, k);
                o1.z = quant_fast<MODE>(v[s][1].z, recip, scale, k); o1.w = quant_fast<MODE>(v[s][1].w, recip, scale, k);
            } else {                                                              // zero / inf / 2^127 / subnormal scale: generic routine
                const float scale = __uint_as_float(sb);
                o0.x = mx_quant_generic<MODE>(v[s][0].x, scale, g); o0.y = mx_quant_generic<MODE>(v[s][0].y, scale, g);
                o0.z = mx_quant_generic<MODE>(v[s][0].z, scale, g); o0.w = mx_quant_generic<MODE>(v[s][0].w, scale, g);
                o1.x = mx_quant_generic<MODE>(v[s][1].x, scale, g); o1.y = mx_quant_generic<MODE>(v[s][1].y, scale, g);
                o1.z = mx_quant_generic<MODE>(v[s][1].z, scale, g); o1.w = mx_quant_generic<MODE>(v[s][1].w, scale, g);
            }
            st_stream_f4(out + step * 64 + f0, o0);
            st_stream_f4(out + step * 64 + f0 + G, o1);
        }
    }
}

template <int BS, int MODE>
cudaError_t launch_bs_mode(const float* in, float* out, int64_t nsteps, const MxArgs& g, const MxFastArgs& k,
                           const MxLaunchCfg& cfg, cudaStream_t stream)
{
    constexpr int STEPS = 2;
    constexpr int kWarps = kFastThreads / 32;
    const int64_t ngroups = (nsteps + STEPS - 1) / STEPS;
    int64_t grid = (ngroups + kWarps - 1) / kWarps;
    static std::atomic<int> resident_cache{0};                 // CTAs of this instantiation that fit on one SM (same for every B200)
    int resident = resident_cache.load(std::memory_order_relaxed);
    if (resident == 0) {
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, mx_fast_kernel<BS, MODE, STEPS>, kFastThreads, 0) != cudaSuccess || nb < 1) nb = 1;
        resident = nb;
        resident_cache.store(nb, std::memory_order_relaxed);
    }
    const int64_t max_ctas = (int64_t)cfg.num_sms * (resident < cfg.ctas_per_sm ? resident : cfg.ctas_per_sm);
    if (grid > max_ctas) grid = max_ctas;                      // persistent: never more CTAs than are resident at once
    mx_fast_kernel<BS, MODE, STEPS><<<(unsigned)grid, kFastThreads, 0, stream>>>(
        reinterpret_cast<const float4*>(in), reinterpret_cast<float4*>(out), nsteps, g, k);
    return cudaGetLastError();
}

template <int MODE>
cudaError_t launch_mode(const float* in, float* out, int64_t nsteps, const MxArgs& g, const MxFastArgs& k,
                        const MxLaunchCfg& cfg, cudaStream_t stream)
{
    switch (g.block_size) {
    case 16:  return launch_bs_mode<16, MODE>(in, out, nsteps, g, k, cfg, stream);
    case 32:  return launch_bs_mode<32, MODE>(in, out, nsteps, g, k, cfg, stream);
    case 64:  return launch_bs_mode<64, MODE>(in, out, nsteps, g, k, cfg, stream);
    case 128: return launch_bs_mode<128, MODE>(in, out, nsteps, g, k, cfg, stream);
    case 256: return launch_bs_mode<256, MODE>(in, out, nsteps, g, k, cfg, stream);
    default:  return cudaErrorInvalidValue;
    }
}

MxFastArgs make_fast_args(const MxArgs& g, const MagicPlan& elem_magic, const MagicPlan& scale_magic,
                          const FastPlan& scale_fast, int scale_mode)
{
    MxFastArgs k;
    k.elem_magic = elem_magic;
    k.scale_magic = scale_magic;
    k.scale_fast = scale_fast;
    k.pmn = g.pmn;
    k.rcp_pmn = 1.0f / g.pmn;
    // 3-FMA quotient: keep q0, the residual and the quotient comfortably inside the normal range
    k.div_lo = g.pmn * 1.0e-30f > 1.0e-30f ? g.pmn * 1.0e-30f : 1.0e-30f;
    k.div_hi = g.pmn * 1.0e30f < 1.0e30f ? g.pmn * 1.0e30f : 1.0e30f;
    k.scale_mode = scale_mode;
    k.elem_sat = g.elem.saturating;
    return k;
}

// self-test: div_pmn(a) == a / pmn (IEEE) for EVERY non-negative finite fp32 a
__global__ void mx_div_selftest_kernel(MxFastArgs k, unsigned long long* mismatches)
{
    unsigned long long bad = 0;
    for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < 0x7F800000ull; b += (uint64_t)gridDim.x * blockDim.x) {
        const float a = __uint_as_float((uint32_t)b);
        bad += __float_as_uint(div_pmn(a, k)) != __float_as_uint(__fdiv_rn(a, k.pmn));
    }
    if (bad) atomicAdd(mismatches, bad);
}

}  // namespace

cudaError_t run_mx_div_selftest(float pmn, unsigned long long* host_mismatches, cudaStream_t stream)
{
    MxArgs g;
    memset(&g, 0, sizeof g);
    g.pmn = pmn;
    MagicPlan none;
    memset(&none, 0, sizeof none);
    FastPlan nofast;
    memset(&nofast, 0, sizeof nofast);
    const MxFastArgs k = make_fast_args(g, none, none, nofast, 0);
    unsigned long long* d = nullptr;
    cudaError_t e = cudaMalloc(&d, sizeof *d);
    if (e != cudaSuccess) return e;
    cudaMemsetAsync(d, 0, sizeof *d, stream);
    mx_div_selftest_kernel<<<148 * 16, 256, 0, stream>>>(k, d);
    e = cudaMemcpyAsync(host_mismatches, d, sizeof *d, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(d);
    return e;
}

// Handles the first (n / 256) * 256 elements when the request is eligible; *main_elems tells the caller how many
// elements were covered (0 = not eligible: use the generic kernels for everything).
cudaError_t launch_mx_fast(int mode, int scale_mode, const float* in, float* out, int64_t n, const MxArgs& g,
                           const MagicPlan& elem_magic, const MagicPlan& scale_magic, const FastPlan& scale_fast,
                           int scale_mantissa_bits, const MxLaunchCfg& cfg, cudaStream_t stream, int64_t* main_elems,
                           int* launches, const RngArgs* rng_elem, uint64_t first_index)
{
    *main_elems = 0;
    const bool aligned = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0;
    const bool scale_mode_ok = scale_mode == LOF_RoundToNearestEven || scale_mode == LOF_RoundTowardsZero ||
                               scale_mode == LOF_RoundAwayFromZero || scale_mode == LOF_RoundUp ||
                               scale_mode == LOF_RoundDown || scale_mode == LOF_RoundToOdd;
    if (!aligned || !elem_magic.valid || elem_magic.limit != 0xFFFFFFFFu || elem_magic.check_denorm_in ||
        scale_mantissa_bits != 0 || !scale_mode_ok || n < kStepFloats)
        return cudaSuccess;
    const MxFastArgs k = make_fast_args(g, elem_magic, scale_magic, scale_fast, scale_mode);
    const int64_t nsteps = n / kStepFloats;
    cudaError_t e;
    if (mode == LOF_StochasticRoundingA || mode == LOF_StochasticRoundingD) {
        // class-0 events only (<= 8 random bits), whole 4096-element runs aligned with the event stream, no per-lane range conditions
        const int64_t nruns = n / 4096;
        if (!rng_elem || rng_elem->cls != 0u || (first_index & 4095u) != 0u || elem_magic.lo_rare != 0u || nruns < 1) return cudaSuccess;
        e = (mode == LOF_StochasticRoundingA) ? launch_sr<LOF_StochasticRoundingA>(in, out, nruns, g, k, *rng_elem, first_index >> 12, cfg, stream)
                                              : launch_sr<LOF_StochasticRoundingD>(in, out, nruns, g, k, *rng_elem, first_index >> 12, cfg, stream);
        if (e != cudaSuccess) return e;
        ++*launches;
        *main_elems = nruns * 4096;
        return cudaSuccess;
    }
    switch (mode) {
#define LOF_CASE(M) case M: e = launch_mode<M>(in, out, nsteps, g, k, cfg, stream); break;
        LOF_CASE(LOF_RoundToNearestEven) LOF_CASE(LOF_RoundTowardsZero) LOF_CASE(LOF_RoundAwayFromZero)
        LOF_CASE(LOF_RoundDown) LOF_CASE(LOF_RoundUp) LOF_CASE(LOF_RoundToOdd)
#undef LOF_CASE
    default: return cudaSuccess;                 // other modes: generic kernels
    }
    if (e != cudaSuccess) return e;
    ++*launches;
    *main_elems = nsteps * kStepFloats;
    return cudaSuccess;
}

}  // namespace lof
