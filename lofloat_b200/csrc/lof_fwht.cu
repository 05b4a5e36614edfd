// lof_fwht.cu -- callers fused around the path: the block-diagonal orthonormal
// Walsh-Hadamard rotation LoF_Linear applies to its activations (/root/reference/LoFloat/layers.py:182-216, 363-365)
// fused with the rounding that follows it (STEMXRound :366-369 or STERound :378), one pass over HBM instead of the
// reference's log2(B) + 2 passes (its _fwht materialises a new tensor per butterfly stage).
//
//     y = FWHT_B(x) * fl32(1 / sqrt(B))          per contiguous block of B elements (B a power of two, 4 .. 1024)
//     out = y | virtual_round(scale * y, fmt) | virtual_mx_round(y, mx_block, ...)
//
// Bit-exactness with the reference's PyTorch formulation: its stages run h = 1, 2, 4, ... with new[i] = a + b,
// new[i + h] = a - b for the pair (a, b) = (x[i], x[i + h]); every add / subtract here is the same fp32 operation on the
// same operands in the same stage order (in-lane for h = 1, 2; shfl.xor across lanes for h = 4 .. 64; across the rows a
// lane holds for h = 128 .. 512), then one fp32 multiply by the rounded constant.
// Layout: a warp owns rows of 128 consecutive floats, lane l holds elements 4l .. 4l + 3 of each row (one 128-bit load).
#include <cuda_runtime.h>

#include "lof_core.cuh"
#include "lof_mx_kernels.cuh"

namespace lof {

namespace {

constexpr int kFwhtThreads = 256;
enum : int { kFwhtPlain = 0, kFwhtRound = 1, kFwhtMx = 2 };

struct FwhtArgs {
    int log2_b;             // Hadamard block = 2^log2_b elements
    float inv_sqrt_b;
    RoundPlan plan;         // kFwhtRound
    float pre_scale;        // kFwhtRound: fused scale-then-round (Lof_kernel.cu:41,53); 1 = none
    MxArgs mx;              // kFwhtMx
    MxScaleQuick sq;        // kFwhtMx: branch-free scale rounding where eligible
    MagicPlan elem_magic;   // kFwhtMx: magic-add element rounding for power-of-two scales when elem_fast_ok
    int elem_fast_ok;
};

__device__ __forceinline__ void butterfly(float& a, float& b)
{
    const float s = __fadd_rn(a, b), d = __fsub_rn(a, b);
    a = s; b = d;
}

// generic element path of one lane's four elements (any scale, any element format / mode): out of line, the streaming loop
// carries one call instead of four inlined copies of the division + exact rounding per row
template <int MODE>
__device__ __noinline__ float4 mx_quant_row_generic(float4 v, float scale, const MxArgs& mx, uint4 w)
{
    const MxBlockCtx c = mx_make_ctx(scale);
    float4 o;
    o.x = mx_quant<MODE>(v.x, c, mx, w.x); o.y = mx_quant<MODE>(v.y, c, mx, w.y);
    o.z = mx_quant<MODE>(v.z, c, mx, w.z); o.w = mx_quant<MODE>(v.w, c, mx, w.w);
    return o;
}

// ROWS = rows (of 128 floats) one warp step holds per lane; a Hadamard block spans max(1, B / 128) of them
// LG = log2(lanes per MX block) as a compile-time constant (3: the OCP block of 32 elements), -1 = taken from the arguments
// FULL = the Hadamard block spans at least one row (log2_b >= 7): all two in-lane and five cross-lane stages, unrolled
template <int ROWS, int OP, int MODE, int LG, bool FULL>
__global__ void __launch_bounds__(kFwhtThreads, ROWS == 4 ? 4 : 2)
fwht_kernel(const float4* __restrict__ in, float4* __restrict__ out, int64_t nrows, FwhtArgs a, RngArgs rng, RngArgs rng_scale,
            uint64_t first_index)
{
    constexpr int kWarps = kFwhtThreads / 32;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int64_t nsteps = (nrows + ROWS - 1) / ROWS;
    const int lane_stages = FULL ? 5 : min(a.log2_b, 7) - 2; // shuffle stages h = 4 .. min(B, 128) / 2
    const int row_stages = max(a.log2_b - 7, 0);             // register stages across rows, h = 128 ..

    for (int64_t step = (int64_t)blockIdx.x * kWarps + warp; step < nsteps; step += (int64_t)gridDim.x * kWarps) {
        const int64_t row0 = step * ROWS;
        float4 v[ROWS];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            if (row0 + r < nrows) v[r] = ld_stream_f4(in + (row0 + r) * 32 + lane);
            else v[r] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        // h = 1, 2: inside the lane's four elements
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            if (FULL || a.log2_b >= 1) { butterfly(v[r].x, v[r].y); butterfly(v[r].z, v[r].w); }
            if (FULL || a.log2_b >= 2) { butterfly(v[r].x, v[r].z); butterfly(v[r].y, v[r].w); }
        }
        // h = 4 .. 64: across lanes.  lower lane: v + p, upper lane: p - v -- one fused multiply-add with sg = +-1 (the product is exact, so
        // the single rounding is that of the sum / difference: bit-identical to fadd / fsub) instead of two adds and a select per element
        auto lane_stage = [&](int m) {
            const float sg = (lane & m) != 0 ? -1.0f : 1.0f;
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                const float px = __shfl_xor_sync(0xFFFFFFFFu, v[r].x, m), py = __shfl_xor_sync(0xFFFFFFFFu, v[r].y, m);
                const float pz = __shfl_xor_sync(0xFFFFFFFFu, v[r].z, m), pw = __shfl_xor_sync(0xFFFFFFFFu, v[r].w, m);
                v[r].x = __fmaf_rn(v[r].x, sg, px);
                v[r].y = __fmaf_rn(v[r].y, sg, py);
                v[r].z = __fmaf_rn(v[r].z, sg, pz);
                v[r].w = __fmaf_rn(v[r].w, sg, pw);
            }
        };
        if constexpr (FULL) {
            lane_stage(1); lane_stage(2); lane_stage(4); lane_stage(8); lane_stage(16);
        } else {
            for (int s = 0; s < lane_stages; ++s) lane_stage(1 << s);
        }
        // h = 128 ..: across the rows this lane holds (rows of one block are consecutive and aligned: nrows-tail rows are zero)
#pragma unroll
        for (int s = 0; s < 3; ++s) {
            if (s < row_stages) {
                const int hr = 1 << s;
#pragma unroll
                for (int r = 0; r < ROWS; ++r) {
                    if ((r & hr) == 0 && r + hr < ROWS) {
                        butterfly(v[r].x, v[r + hr].x); butterfly(v[r].y, v[r + hr].y);
                        butterfly(v[r].z, v[r + hr].z); butterfly(v[r].w, v[r + hr].w);
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            v[r].x = __fmul_rn(v[r].x, a.inv_sqrt_b); v[r].y = __fmul_rn(v[r].y, a.inv_sqrt_b);
            v[r].z = __fmul_rn(v[r].z, a.inv_sqrt_b); v[r].w = __fmul_rn(v[r].w, a.inv_sqrt_b);
        }
        if constexpr (OP == kFwhtMx) {
            const int log2_mx = 31 - __clz(a.mx.block_size);
            const int log2_group = LG >= 0 ? LG : log2_mx - 2;        // lanes per MX block: 4 .. 32, a power of two
            const int group = 1 << log2_group;
            float m[ROWS];
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                float t = amax4(v[r]);
                if constexpr (LG >= 0) {
#pragma unroll
                    for (int off = (1 << LG) >> 1; off > 0; off >>= 1) t = fmaxf(t, __shfl_xor_sync(0xFFFFFFFFu, t, off));
                } else {
                    for (int off = group >> 1; off > 0; off >>= 1) t = fmaxf(t, __shfl_xor_sync(0xFFFFFFFFu, t, off));
                }
                m[r] = t;
            }
            // scales: the ROWS * (32 / group) block maxima are transposed onto the lanes so that every lane rounds ONE scale
            // (no divergent per-block path); when that is more than 32 blocks the first lane of each block does it
            const int log2_bpr = 5 - log2_group, bpr = 1 << log2_bpr;    // blocks per row
            const bool transposed = ROWS * bpr <= 32;
            float scales_r[ROWS];
            if (transposed) {
                float mine = 0.0f;
                const int src_lane = (lane & (bpr - 1)) << log2_group, my_row = lane >> log2_bpr, blk_in_row = lane >> log2_group;
#pragma unroll
                for (int r = 0; r < ROWS; ++r) {
                    const float t = __shfl_sync(0xFFFFFFFFu, m[r], src_lane);
                    if (my_row == r) mine = t;
                }
                const uint64_t gb0 = (first_index + (uint64_t)row0 * 128u) >> log2_mx;
                float sc = 0.0f;
                if (lane < ROWS * bpr) sc = mx_scale_quick(mine, a.mx, a.sq, rng_scale, gb0 + (uint64_t)lane);
#pragma unroll
                for (int r = 0; r < ROWS; ++r) scales_r[r] = __shfl_sync(0xFFFFFFFFu, sc, (r << log2_bpr) + blk_in_row);
            } else {
#pragma unroll
                for (int r = 0; r < ROWS; ++r) {
                    const uint64_t g0 = first_index + (uint64_t)(row0 + r) * 128u + (uint64_t)lane * 4u;
                    float sc = 0.0f;
                    if ((lane & (group - 1)) == 0) sc = mx_scale_quick(m[r], a.mx, a.sq, rng_scale, g0 >> log2_mx);
                    scales_r[r] = __shfl_sync(0xFFFFFFFFu, sc, lane & ~(group - 1));
                }
            }
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                if (row0 + r >= nrows) continue;
                const uint64_t g0 = first_index + (uint64_t)(row0 + r) * 128u + (uint64_t)lane * 4u;
                const float scale = scales_r[r];
                uint4 w = make_uint4(0, 0, 0, 0);
                if constexpr (ModeTraits<MODE>::kUsesRng) w = philox_group4(rng, g0 >> 2);
                float4 o;
                const uint32_t sb = __float_as_uint(scale);
                bool done = false;
                if constexpr (FastTraits<MODE>::kFaddSub) {
                    if (a.elem_fast_ok && (sb & 0x007FFFFFu) == 0u && (sb - 0x00800000u) < 0x7E800000u) {   // normal power of two, 2^-126 .. 2^126
                        const float recip = __uint_as_float(0x7F000000u - sb);                               // 2^-k, exact
                        const bool sat = a.mx.elem.saturating != 0;
                        o.x = mx_quant_magic<MODE>(v[r].x, recip, scale, a.elem_magic, sat); o.y = mx_quant_magic<MODE>(v[r].y, recip, scale, a.elem_magic, sat);
                        o.z = mx_quant_magic<MODE>(v[r].z, recip, scale, a.elem_magic, sat); o.w = mx_quant_magic<MODE>(v[r].w, recip, scale, a.elem_magic, sat);
                        done = true;
                    }
                }
                if (!done) o = mx_quant_row_generic<MODE>(v[r], scale, a.mx, w);
                st_stream_f4(out + (row0 + r) * 32 + lane, o);
            }
        } else {
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                if (row0 + r >= nrows) continue;
                float4 o = v[r];
                if constexpr (OP == kFwhtRound) {
                    const uint64_t g0 = first_index + (uint64_t)(row0 + r) * 128u + (uint64_t)lane * 4u;
                    uint4 w = make_uint4(0, 0, 0, 0);
                    if constexpr (ModeTraits<MODE>::kUsesRng) w = philox_group4(rng, g0 >> 2);
                    if (a.pre_scale != 1.0f) {
                        o.x = __fmul_rn(a.pre_scale, o.x); o.y = __fmul_rn(a.pre_scale, o.y);
                        o.z = __fmul_rn(a.pre_scale, o.z); o.w = __fmul_rn(a.pre_scale, o.w);
                    }
                    bool done = false;
                    if constexpr (MagicTraits<MODE>::kOk) {
                        // the magic-add rounding of the elementwise kernels (lof_fast_f32.cuh, checked variant: lanes it cannot take set `rare`
                        // and the vector is redone by the exact routine) instead of the ~35-instruction integer routine per element
                        if (a.elem_fast_ok) {
                            const uint32_t q[4] = {__float_as_uint(o.x), __float_as_uint(o.y), __float_as_uint(o.z), __float_as_uint(o.w)};
                            const uint32_t wq[4] = {w.x, w.y, w.z, w.w};
                            uint32_t rr[4];
                            bool rare = false;
                            if (a.plan.saturating) {
#pragma unroll
                                for (int e = 0; e < 4; ++e) rr[e] = round_magic_f32<MODE, true, true>(q[e], a.elem_magic, wq[e], rare);
                            } else {
#pragma unroll
                                for (int e = 0; e < 4; ++e) rr[e] = round_magic_f32<MODE, false, true>(q[e], a.elem_magic, wq[e], rare);
                            }
                            if (rare) {
#pragma unroll
                                for (int e = 0; e < 4; ++e) rr[e] = exact_params_f32<MODE>(q[e], a.plan, wq[e]);
                            }
                            o = make_float4(__uint_as_float(rr[0]), __uint_as_float(rr[1]), __uint_as_float(rr[2]), __uint_as_float(rr[3]));
                            done = true;
                        }
                    }
                    if (!done) {
                        o.x = __uint_as_float(round_params_elem<LOF_F32, MODE>(__float_as_uint(o.x), a.plan, w.x));
                        o.y = __uint_as_float(round_params_elem<LOF_F32, MODE>(__float_as_uint(o.y), a.plan, w.y));
                        o.z = __uint_as_float(round_params_elem<LOF_F32, MODE>(__float_as_uint(o.z), a.plan, w.z));
                        o.w = __uint_as_float(round_params_elem<LOF_F32, MODE>(__float_as_uint(o.w), a.plan, w.w));
                    }
                }
                st_stream_f4(out + (row0 + r) * 32 + lane, o);
            }
        }
    }
}

template <int OP, int MODE>
cudaError_t launch_rows(int rows_per_block, const float* in, float* out, int64_t nrows, const FwhtArgs& a, const RngArgs& rng,
                        const RngArgs& rng_scale, uint64_t first_index, int num_sms, cudaStream_t stream)
{
    const float4* vin = reinterpret_cast<const float4*>(in);
    float4* vout = reinterpret_cast<float4*>(out);
    constexpr int kWarps = kFwhtThreads / 32;
    auto grid_for = [&](int rows) {
        const int64_t nsteps = (nrows + rows - 1) / rows;
        const int64_t want = (nsteps + kWarps - 1) / kWarps, cap = (int64_t)num_sms /* = SMs x resident CTAs per SM, chosen by the caller */;
        return (unsigned)(want < 1 ? 1 : want < cap ? want : cap);
    };
    if constexpr (OP == kFwhtMx) {
        if (a.mx.block_size == 32) {                              // the OCP MX block: lanes per block as a compile-time constant
            if (rows_per_block == 8) fwht_kernel<8, OP, MODE, 3, true><<<grid_for(8), kFwhtThreads, 0, stream>>>(vin, vout, nrows, a, rng, rng_scale, first_index);
            else if (a.log2_b >= 7) fwht_kernel<4, OP, MODE, 3, true><<<grid_for(4), kFwhtThreads, 0, stream>>>(vin, vout, nrows, a, rng, rng_scale, first_index);
            else fwht_kernel<4, OP, MODE, 3, false><<<grid_for(4), kFwhtThreads, 0, stream>>>(vin, vout, nrows, a, rng, rng_scale, first_index);
            return cudaGetLastError();
        }
    }
    if constexpr (OP == kFwhtRound) {
        if (a.log2_b >= 7) {                                      // Hadamard blocks >= 128: the stage structure compile-time
            if (rows_per_block == 8) fwht_kernel<8, OP, MODE, -1, true><<<grid_for(8), kFwhtThreads, 0, stream>>>(vin, vout, nrows, a, rng, rng_scale, first_index);
            else fwht_kernel<4, OP, MODE, -1, true><<<grid_for(4), kFwhtThreads, 0, stream>>>(vin, vout, nrows, a, rng, rng_scale, first_index);
            return cudaGetLastError();
        }
    }
    switch (rows_per_block) {
    case 8: fwht_kernel<8, OP, MODE, -1, false><<<grid_for(8), kFwhtThreads, 0, stream>>>(vin, vout, nrows, a, rng, rng_scale, first_index); break;
    default: fwht_kernel<4, OP, MODE, -1, false><<<grid_for(4), kFwhtThreads, 0, stream>>>(vin, vout, nrows, a, rng, rng_scale, first_index); break;   // B <= 512
    }
    return cudaGetLastError();
}

template <int OP>
cudaError_t launch_mode(int mode, int rows_per_block, const float* in, float* out, int64_t nrows, const FwhtArgs& a, const RngArgs& rng,
                        const RngArgs& rng_scale, uint64_t first_index, int num_sms, cudaStream_t stream)
{
    switch (mode) {
#define LOF_CASE(M) case M: return launch_rows<OP, M>(rows_per_block, in, out, nrows, a, rng, rng_scale, first_index, num_sms, stream);
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

// Hadamard blocks of G x 1024 elements (G = 2, 4, 8, 16, 32): the stages h = 1 .. 512 run inside the 1024-element chunks (fwht_kernel,
// unnormalised), this kernel runs the remaining log2(G) stages h = 1024 .. B / 2 ACROSS the chunks -- a thread owns one float4
// column of the G chunks, so every pair (i, i + h) sits in its registers -- and applies fl32(1 / sqrt(B)) once at the end.
// Same adds and subtracts on the same operands in the same stage order as the reference's formulation: bit-identical.
// In place (every thread reads and writes the same addresses).
template <int G>
__global__ void __launch_bounds__(256)
fwht_cross_kernel(float4* __restrict__ data, int64_t nblocks, float inv_sqrt_b)
{
    const int64_t total = nblocks * 256;                       // 256 float4 columns per block
    for (int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; id < total; id += (int64_t)gridDim.x * blockDim.x) {
        float4* base = data + (id >> 8) * (int64_t)(G * 256) + (id & 255);
        float4 v[G];
#pragma unroll
        for (int g = 0; g < G; ++g) v[g] = __ldcs(base + g * 256);
#pragma unroll
        for (int h = 1; h < G; h <<= 1)
#pragma unroll
            for (int g = 0; g < G; ++g)
                if ((g & h) == 0) {
                    butterfly(v[g].x, v[g + h].x); butterfly(v[g].y, v[g + h].y);
                    butterfly(v[g].z, v[g + h].z); butterfly(v[g].w, v[g + h].w);
                }
#pragma unroll
        for (int g = 0; g < G; ++g) {
            v[g].x = __fmul_rn(v[g].x, inv_sqrt_b); v[g].y = __fmul_rn(v[g].y, inv_sqrt_b);
            v[g].z = __fmul_rn(v[g].z, inv_sqrt_b); v[g].w = __fmul_rn(v[g].w, inv_sqrt_b);
            __stcs(base + g * 256, v[g]);
        }
    }
}

}  // namespace

// Plain transform for Hadamard blocks of 2048 ... 32768 elements: two passes (16 B / element) instead of the reference's
// log2(B) stage tensors.
cudaError_t launch_fwht_large(const float* in, float* out, int64_t n, int hadamard_block, int num_sms, cudaStream_t stream)
{
    FwhtArgs a;
    memset(&a, 0, sizeof a);
    a.log2_b = 10;
    a.inv_sqrt_b = 1.0f;                                        // x * 1.0f is the identity: normalisation happens after the last stage
    const RngArgs none{};
    cudaError_t e = launch_rows<kFwhtPlain, LOF_RoundToNearestEven>(8, in, out, n / 128, a, none, none, 0, num_sms, stream);
    if (e != cudaSuccess) return e;
    const int G = hadamard_block / 1024;
    const int64_t nblocks = n / hadamard_block;
    const float norm = (float)(1.0 / sqrt((double)hadamard_block));
    const int64_t want = (nblocks * 256 + 255) / 256;
    const unsigned grid = (unsigned)(want < 1 ? 1 : want < num_sms ? want : num_sms);
    float4* v = reinterpret_cast<float4*>(out);
    switch (G) {
    case 2: fwht_cross_kernel<2><<<grid, 256, 0, stream>>>(v, nblocks, norm); break;
    case 4: fwht_cross_kernel<4><<<grid, 256, 0, stream>>>(v, nblocks, norm); break;
    case 8: fwht_cross_kernel<8><<<grid, 256, 0, stream>>>(v, nblocks, norm); break;
    case 16: fwht_cross_kernel<16><<<grid, 256, 0, stream>>>(v, nblocks, norm); break;       // 16384: an LLM-sized in_features with hadamard_block_size=None
    case 32: fwht_cross_kernel<32><<<grid, 256, 0, stream>>>(v, nblocks, norm); break;       // 32768 (32 float4 = 128 registers per thread)
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

// op: 0 plain FWHT, 1 + virtual_round(params), 2 + virtual_mx_round.  n % 128 == 0, block in {4 .. 1024}, 16-byte aligned.
cudaError_t launch_fwht(int op, int mode, const float* in, float* out, int64_t n, int hadamard_block, const RoundPlan* plan,
                        float pre_scale, const MxArgs* mx, const MagicPlan* scale_magic, const FastPlan* scale_fast,
                        const MagicPlan* elem_magic, const RngArgs& rng, const RngArgs& rng_scale, uint64_t first_index, int num_sms,
                        cudaStream_t stream)
{
    FwhtArgs a;
    memset(&a, 0, sizeof a);
    int lg = 0;
    while ((1 << lg) < hadamard_block) ++lg;
    a.log2_b = lg;
    a.inv_sqrt_b = (float)(1.0 / sqrt((double)hadamard_block));       // layers.py:215: flat * (1.0 / math.sqrt(block_size))
    if (plan) a.plan = *plan;
    a.pre_scale = pre_scale;
    if (mx) {
        a.mx = *mx;
        if (scale_magic) a.sq.magic = *scale_magic;
        if (scale_fast) a.sq.fast = *scale_fast;
        if (elem_magic) { a.elem_magic = *elem_magic; a.elem_fast_ok = 1; }
        const int sm = mx->scale_mode;
        a.sq.quick_mode_ok = sm == LOF_RoundToNearestEven || sm == LOF_RoundTowardsZero || sm == LOF_RoundAwayFromZero ||
                             sm == LOF_RoundUp || sm == LOF_RoundDown || sm == LOF_RoundToOdd;
    } else if (elem_magic && elem_magic->valid) {               // op 1: magic-add plan of the fused virtual_round
        a.elem_magic = *elem_magic;
        a.elem_fast_ok = 1;
    }
    const int rows_per_block = hadamard_block > 512 ? 8 : 4;
    const int64_t nrows = n / 128;
    switch (op) {
    case kFwhtRound: return launch_mode<kFwhtRound>(mode, rows_per_block, in, out, nrows, a, rng, rng_scale, first_index, num_sms, stream);
    case kFwhtMx:    return launch_mode<kFwhtMx>(mode, rows_per_block, in, out, nrows, a, rng, rng_scale, first_index, num_sms, stream);
    default:         return launch_rows<kFwhtPlain, LOF_RoundToNearestEven>(rows_per_block, in, out, nrows, a, rng, rng_scale, first_index, num_sms, stream);
    }
}

}  // namespace lof
