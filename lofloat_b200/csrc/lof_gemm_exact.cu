// lof_gemm_exact.cu -- LoFloat-accumulation GEMM, exact per-FMA mode (SIMT, sm_100a).
//
// Reproduces, bit for bit, the reference's accumulation semantics
//     acc <- virtual_round( round_to_odd_fma(a, b, acc), accum_mant_bits, ps )   for every k
// CPU reference: lof_sgemm_ukr, /root/reference/src/blis_lof_gemm.hpp:147-262 (accumulator
// restarts every BLIS KC = 256 panel; panels are summed in plain fp32, :244-261);
// CUDA reference leaf: /root/reference/src/cutlass_gemms.cuh:97-168 (one accumulator over
// all K; == kc <= 0 here).  The per-MAC rounding chain is strictly sequential in k, so this
// path cannot use the tensor cores; they are used by the chunked mode (lof_gemm_tc.cu).
//
// The kernel is ALU-bound (~14 integer/FMA instructions per MAC); operands are staged in
// shared memory in 64x16 / 16x64 tiles, each thread owns a 4x4 block of a 64x64 C tile.
#include "lof_core.cuh"
#include "lof_gemm_common.cuh"

namespace lof {

namespace {

constexpr int TM = 64, TN = 64, TK = 16;
constexpr int kGemmThreads = 256;

template <int MODE>
__device__ __forceinline__ float accum_round(float t, int r, int len, uint32_t w)
{
    return __uint_as_float(round_mantissa_elem<LOF_F32, MODE>(__float_as_uint(t), r, len, w));
}

// Optimistic round-to-odd FMA, no branch: of the two directed results rd <= exact <= ru, take the one whose LSB is odd.
// Inexact: rd and ru are adjacent floats, exactly one is odd, and it is the reference's "toward-zero neighbour with the LSB
// set" (rz itself if odd, else rz | 1 = the other neighbour).  Exact: rd == ru, either will do -- except the exact zero, where
// rd = -0 (even) selects ru = +0 like the reference.  Equal to rto_fma whenever |result| < 2^127; overflow, infinities and
// NaNs (where the reference switches to RN) land at |result| >= 2^127 or NaN, which the caller detects and recomputes with
// the exact routine.  Two FFMAs + LOP3 (-> predicate) + SEL.
__device__ __forceinline__ float rto_fma_optimistic(float a, float b, float c)
{
    const float rd = __fmaf_rd(a, b, c), ru = __fmaf_ru(a, b, c);
    return (__float_as_uint(rd) & 1u) ? rd : ru;
}

// The same with two FMAs: the toward-zero result is picked by the sign of rd (an exact zero has rd = -0 and takes ru = +0).
__device__ __forceinline__ float rto_fma_optimistic2(float a, float b, float c)
{
    const float rd = __fmaf_rd(a, b, c), ru = __fmaf_ru(a, b, c);
    const uint32_t rz = (int32_t)__float_as_uint(rd) >= 0 ? __float_as_uint(rd) : __float_as_uint(ru);
    return __uint_as_float(rz | (rd != ru ? 1u : 0u));
}

// Mantissa-only rounding on the signed bit pattern (no |x| / sign re-attach): the increments of every mode only look at the
// low r bits (+ the sign for RoundUp / RoundDown), and for |t| < 2^127 the carry cannot reach the sign bit.
template <int MODE>
__device__ __forceinline__ float accum_round_signed(float t, int r, uint32_t keep, int len, uint32_t w)
{
    const uint32_t b = __float_as_uint(t);
    return __uint_as_float(round_bits<MODE, uint32_t>(b, r, len, (int32_t)b >= 0, w) & keep);
}

// RNE to (23 - r) mantissa bits in the float domain:  K = sign(t) * 2^(e + r)  (e = exponent of t), y = (t + K) - K.
// t + K lies in [2^(e+r) + 2^e, 2^(e+r) + 2^(e+1)), one binade whose ulp is 2^(e + r - 23) -- the target grid -- so the
// hardware's RNE add performs the rounding (ties to even on that grid, i.e. on bit r of t's pattern) and the subtraction is
// exact.  Equal to round_bits<RNE> + mask for every NORMAL t with e + r < 127 and r <= 22 (lof_selftest_gemm_rne_round checks
// all 2^32 patterns x 22 widths on the device; at r = 23 the integer routine's "even" is the exponent's low bit, ours is not); larger t, infinities and NaNs give a non-finite y (K = inf -> NaN; at e + r = 127 the
// sum can round up to inf), which sticks to the accumulator until the panel ends, where the caller looks for it.  One alu-pipe and two fma-pipe instructions instead of
// four alu-pipe ones: the chain is bound by the (half-rate) alu pipe, see profiles/r01k_gemm_exact_summary.txt.
__device__ __forceinline__ float accum_round_rne_float(float t, float two_r)
{
    // K = m * 2^r is an exact product, so both steps are single FMAs: s = fma(m, 2^r, t) = RN(t + K), y = fma(m, -2^r, s) = s - K
    return round_mant_float<LOF_RoundToNearestEven>(t, two_r);      // lof_gemm_common.cuh
}

// One 64 x 64 tile of C into `crun` (not stored).  VARIANT 0: the exact routines (rto_fma with its non-finite branch,
// round_mantissa_elem); VARIANT 1 / 2: the optimistic chain (r > 0 / r == 0) -- returns true if any intermediate left
// the range where it is proven equal, in which case the caller discards `crun` and runs VARIANT 0.  VARIANT 3 (RNE, r > 0):
// the same chain with the rounding in the float domain; its range conditions are checked outside the MAC loop -- operands
// are 0 or >= 2^-40 in magnitude (then every product and every accumulator is a multiple of 2^-126: no subnormal
// intermediate can occur), and no accumulator is NaN / inf at the end of a panel.
template <int MODE, int VARIANT>
__device__ __forceinline__ bool
gemm_exact_tile(const float* __restrict__ A, const float* __restrict__ B, const float* C, int M, int N, int K,
                int r, int len, int kc, float alpha, float beta, bool custom, const RngArgs& rng, uint64_t rng_base,
                int bias_off, const ExactView& v, float (*As)[TM + 4], float (*Bs)[TN + 4], float (&crun)[4][4])
{
    bool flagged = false;
    const uint32_t keep = r > 0 ? ~((1u << r) - 1u) : 0xFFFFFFFFu;
    const float two_r = __uint_as_float((uint32_t)(127 + r) << 23);
    constexpr float kTinyOperand = 9.094947017729282e-13f;      // 2^-40

    const int tx = threadIdx.x & 15;     // n direction
    const int ty = threadIdx.x >> 4;     // m direction
    const int m0 = (v.swap_xy ? blockIdx.x : blockIdx.y) * TM, n0 = (v.swap_xy ? blockIdx.y : blockIdx.x) * TN;

#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) crun[i][j] = 0.0f;

    // conv: the output position this thread stages (column n0 + (tid & 63) of the im2col matrix) never changes
    int ih0 = 0, iw0 = 0;
    const int RS = v.R * v.S;
    if (v.b_kind == 2) {
        const int n = n0 + (threadIdx.x & 63);
        const int p = n / v.Q, q = n - p * v.Q;
        ih0 = p * v.sh - v.ph;
        iw0 = q * v.sw - v.pw;
    }

    bool first = true;
    for (int p0 = 0; p0 < K; p0 += kc) {
        const int p1 = min(p0 + kc, K);
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

        for (int k0 = p0; k0 < p1; k0 += TK) {
            const int kt = min(TK, p1 - k0);
            // stage A (64 x kt) transposed and B (kt x 64)
            {
                const int ak = threadIdx.x & 15, am = threadIdx.x >> 4;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int m = am + 16 * i;
                    float x = 0.0f;
                    if (ak < kt && m0 + m < M) {
                        int kp = k0 + ak;                       // physical column of the filter row (stored (c, r, s))
                        if (v.b_kind == 2 && v.k_order == 1) { const int rs = kp / v.Cg, c = kp - rs * v.Cg; kp = c * RS + rs; }
                        x = A[(int64_t)(m0 + m) * K + kp];
                    }
                    if constexpr (VARIANT >= 3) flagged |= (x != 0.0f) & (fabsf(x) < kTinyOperand);
                    As[ak][m] = x;
                }
                if (v.b_kind == 1) {                            // B stored N x K: same access pattern as A
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int n = am + 16 * i;
                        float x = 0.0f;
                        if (ak < kt && n0 + n < N) x = B[(int64_t)(n0 + n) * K + k0 + ak];
                        if constexpr (VARIANT >= 3) flagged |= (x != 0.0f) & (fabsf(x) < kTinyOperand);
                        Bs[ak][n] = x;
                    }
                } else {
                    const int bn = threadIdx.x & 63, bk = threadIdx.x >> 6;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int k = bk + 4 * i;
                        float x = 0.0f;
                        if (k < kt && n0 + bn < N) {
                            if (v.b_kind == 0) {
                                x = B[(int64_t)(k0 + k) * N + n0 + bn];
                            } else {
                                const int kk = k0 + k;
                                int c, r, s2;
                                if (v.k_order == 0) { c = kk / RS; const int rs = kk - c * RS; r = rs / v.S; s2 = rs - r * v.S; }
                                else { const int rs = kk / v.Cg; c = kk - rs * v.Cg; r = rs / v.S; s2 = rs - r * v.S; }
                                const int ih = ih0 + r * v.dh, iw = iw0 + s2 * v.dw;
                                if (ih >= 0 && ih < v.H && iw >= 0 && iw < v.W) x = B[((int64_t)c * v.H + ih) * v.W + iw];   // zero padding otherwise
                            }
                        }
                        if constexpr (VARIANT >= 3) flagged |= (x != 0.0f) & (fabsf(x) < kTinyOperand);
                        Bs[k][bn] = x;
                    }
                }
            }
            __syncthreads();
            // one multiply-accumulate of chain (i, j): w = the rounding event's random word (stochastic modes)
            auto mac = [&](float a, float b, float& c, uint32_t w) {
                if constexpr (VARIANT != 0) {
                    // truncation composes: RTZ to 23 - r bits (r > 0) of the exact sum == mask(RZ fma); no sticky bit needed
                    const float t = (MODE == LOF_RoundTowardsZero && VARIANT == 1) ? __fmaf_rz(a, b, c)
                                  : VARIANT == 4 ? rto_fma_optimistic2(a, b, c) : rto_fma_optimistic(a, b, c);
                    if constexpr (VARIANT >= 3) {
                        c = accum_round_rne_float(t, two_r);
                    } else {
                        flagged |= !(fabsf(t) < 1.7014118346046923e38f);   // 2^127: also inf / NaN
                        if constexpr (VARIANT == 1) c = accum_round_signed<MODE>(t, r, keep, len, w);
                        else c = t;
                    }
                } else if (custom) {
                    c = accum_round<MODE>(rto_fma(a, b, c), r, len, w);
                } else {
                    c = __fadd_rn(c, __fmul_rn(a, b));   // blis_lof_gemm.hpp:233-239
                }
            };
            if constexpr (ModeTraits<MODE>::kUsesRng) {
                // stochastic modes: event index = ((row * N + col) * K + k); four consecutive k of one chain share a Philox
                // block when that index is a multiple of 4 -- one generator call per 4 MACs instead of 4 (same words)
                for (int kq = 0; kq < kt; kq += 4) {
                    float a4[4][4], b4[4][4];
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const float4 av = *reinterpret_cast<const float4*>(&As[kq + kk][ty * 4]);   // rows >= kt hold stale data, unused
                        const float4 bv = *reinterpret_cast<const float4*>(&Bs[kq + kk][tx * 4]);
                        a4[kk][0] = av.x; a4[kk][1] = av.y; a4[kk][2] = av.z; a4[kk][3] = av.w;
                        b4[kk][0] = bv.x; b4[kk][1] = bv.y; b4[kk][2] = bv.z; b4[kk][3] = bv.w;
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const uint64_t gi = (uint64_t)(m0 + ty * 4 + i), gj = (uint64_t)(n0 + tx * 4 + j);
                            const uint64_t e0 = rng_base + (gi * (uint64_t)N + gj) * (uint64_t)K + (uint64_t)(k0 + kq);
                            uint32_t w[4];
                            if ((e0 & 3u) == 0) {
                                const uint4 blk = philox_block(rng, e0 >> 2);
                                w[0] = blk.x; w[1] = blk.y; w[2] = blk.z; w[3] = blk.w;
                            } else {
#pragma unroll
                                for (int kk = 0; kk < 4; ++kk) w[kk] = philox_word(rng, e0 + kk);
                            }
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk)
                                if (kq + kk < kt) mac(a4[kk][i], b4[kk][j], acc[i][j], w[kk]);
                        }
                }
            } else {
                for (int k = 0; k < kt; ++k) {
                    const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
                    const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
                    const float a4[4] = {av.x, av.y, av.z, av.w};
                    const float b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                        for (int i = 0; i < 4; ++i) mac(a4[i], b4[j], acc[i][j], 0u);
                }
            }
            __syncthreads();
        }
        // fold the panel: ab *= alpha; C = first ? (beta == 0 ? ab : beta*C + ab) : C + ab
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if constexpr (VARIANT >= 3) flagged |= !(fabsf(acc[i][j]) < __int_as_float(0x7F800000));
                const float ab = __fmul_rn(acc[i][j], alpha);
                if (first) {
                    if (beta == 0.0f) crun[i][j] = ab;
                    else {
                        const int gm = m0 + ty * 4 + i, gn = n0 + tx * 4 + j;
                        const float c0 = (gm < M && gn < N) ? C[(int64_t)gm * N + gn] : 0.0f;
                        crun[i][j] = __fadd_rn(__fmul_rn(beta, c0), ab);
                    }
                } else {
                    crun[i][j] = __fadd_rn(crun[i][j], ab);
                }
            }
        first = false;
    }
    if (K == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int gm = m0 + ty * 4 + i, gn = n0 + tx * 4 + j;
                crun[i][j] = (beta == 0.0f || gm >= M || gn >= N) ? 0.0f : __fmul_rn(beta, C[(int64_t)gm * N + gn]);
            }
    }
    return flagged;
}

template <int MODE>
__device__ __forceinline__ void
gemm_exact_body(const float* __restrict__ A, const float* __restrict__ B, float* C, int M, int N, int K,
                  int r, int len, int kc, float alpha, float beta, bool custom, int fast, RngArgs rng,
                  int64_t stride_a, int64_t stride_b, int64_t stride_c, ExactView v)
{
    // batched call (blockIdx.z = batch index): the (b, h) loop of LoFloat/layers.py:142-149 as one launch; the stochastic
    // event index continues across the batch ((batch * M + i) * N + j) * K + k
    int bias_off = 0;
    if (v.b_kind == 2) {                                  // conv: z = image * groups + group; M = output channels per group, N = P * Q
        const int img = blockIdx.z / v.groups, grp = blockIdx.z - img * v.groups;
        A += (int64_t)grp * M * K;
        B += ((int64_t)img * v.C + (int64_t)grp * v.Cg) * v.H * v.W;
        C += (int64_t)blockIdx.z * M * N;
        bias_off = grp * M;
    } else {
        A += (int64_t)blockIdx.z * stride_a;
        B += (int64_t)blockIdx.z * stride_b;
        C += (int64_t)blockIdx.z * stride_c;
    }
    const uint64_t rng_base = (uint64_t)blockIdx.z * (uint64_t)M * (uint64_t)N * (uint64_t)K;
    __shared__ float As[TK][TM + 4];
    __shared__ float Bs[TK][TN + 4];
    const int tx = threadIdx.x & 15;     // n direction
    const int ty = threadIdx.x >> 4;     // m direction
    const int m0 = (v.swap_xy ? blockIdx.x : blockIdx.y) * TM, n0 = (v.swap_xy ? blockIdx.y : blockIdx.x) * TN;

    float crun[4][4];
    bool redo = true;
    if (fast && custom) {
        constexpr int kRounding = MODE == LOF_RoundToNearestEven ? 3 : 1;
        constexpr int kRounding2 = MODE == LOF_RoundToNearestEven ? 4 : 1;
        // r = 23 (no mantissa bit kept) stays on the integer rounding: there "even" refers to the exponent's low bit
        const bool f = (r > 0 && r < 23 && fast == 2) ? gemm_exact_tile<MODE, kRounding2>(A, B, C, M, N, K, r, len, kc, alpha, beta, custom, rng, rng_base, bias_off, v, As, Bs, crun)
                     : (r > 0 && r < 23) ? gemm_exact_tile<MODE, kRounding>(A, B, C, M, N, K, r, len, kc, alpha, beta, custom, rng, rng_base, bias_off, v, As, Bs, crun)
                     : r > 0 ? gemm_exact_tile<MODE, 1>(A, B, C, M, N, K, r, len, kc, alpha, beta, custom, rng, rng_base, bias_off, v, As, Bs, crun)
                             : gemm_exact_tile<MODE, 2>(A, B, C, M, N, K, r, len, kc, alpha, beta, custom, rng, rng_base, bias_off, v, As, Bs, crun);
        redo = __syncthreads_or(f) != 0;
    }
    if (redo) gemm_exact_tile<MODE, 0>(A, B, C, M, N, K, r, len, kc, alpha, beta, custom, rng, rng_base, bias_off, v, As, Bs, crun);

    if (v.bias != nullptr) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int gm = m0 + ty * 4 + i, gn = n0 + tx * 4 + j;
                if (gm < M && gn < N) crun[i][j] = __fadd_rn(crun[i][j], v.bias[v.bias_axis == 0 ? bias_off + gm : gn]);
            }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gm = m0 + ty * 4 + i;
        if (gm >= M) continue;
        const int gn = n0 + tx * 4;
        if (gn + 3 < N && (N & 3) == 0 && (reinterpret_cast<uintptr_t>(C) & 15u) == 0) {
            *reinterpret_cast<float4*>(&C[(int64_t)gm * N + gn]) = make_float4(crun[i][0], crun[i][1], crun[i][2], crun[i][3]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (gn + j < N) C[(int64_t)gm * N + gn + j] = crun[i][j];
        }
    }
}

// Two entry points that differ only in their register budget: the deterministic modes take what they need (~120 registers,
// 2 CTAs per SM); the stochastic modes would take 242 (16 interleaved Philox blocks, 1 CTA per SM) and run 35 % faster when
// capped at 128 even though they then spill.
#define LOF_EXACT_PARAMS const float* __restrict__ A, const float* __restrict__ B, float* C, int M, int N, int K, int r, int len, int kc, \
                         float alpha, float beta, bool custom, int fast, RngArgs rng, int64_t stride_a, int64_t stride_b,          \
                         int64_t stride_c, ExactView v
template <int MODE>
__global__ void __launch_bounds__(kGemmThreads) gemm_exact_kernel(LOF_EXACT_PARAMS)
{
    gemm_exact_body<MODE>(A, B, C, M, N, K, r, len, kc, alpha, beta, custom, fast, rng, stride_a, stride_b, stride_c, v);
}
template <int MODE>
__global__ void __launch_bounds__(kGemmThreads, 2) gemm_exact_rng_kernel(LOF_EXACT_PARAMS)
{
    gemm_exact_body<MODE>(A, B, C, M, N, K, r, len, kc, alpha, beta, custom, fast, rng, stride_a, stride_b, stride_c, v);
}
#undef LOF_EXACT_PARAMS
template <int MODE>
constexpr auto exact_kernel_for()          // instantiates only the entry point the mode uses
{
    if constexpr (ModeTraits<MODE>::kUsesRng) return &gemm_exact_rng_kernel<MODE>;
    else return &gemm_exact_kernel<MODE>;
}

// every fp32 pattern x every width: the float-domain RNE of the optimistic chain against round_bits<RNE> + mask
__global__ void selftest_rne_round_kernel(unsigned long long* mismatches)
{
    unsigned long long bad = 0;
    for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < (1ull << 32); p += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t b = (uint32_t)p, field = (b >> 23) & 0xFFu;
        if (field == 0) continue;                                              // zeros / subnormals: excluded by the operand check
        for (int r = 1; r <= 22; ++r) {
            const float y = accum_round_rne_float(__uint_as_float(b), __uint_as_float((uint32_t)(127 + r) << 23));
            const bool nonfinite = !(fabsf(y) < __int_as_float(0x7F800000));
            if (field == 255 || (int)field + r > 254) { bad += !nonfinite; continue; }     // out of range must surface as inf / NaN
            const uint32_t want = round_bits<LOF_RoundToNearestEven, uint32_t>(b, r, 0, true, 0) & ~((1u << r) - 1u);
            bad += __float_as_uint(y) != want && !((int)field + r == 254 && nonfinite);    // top binade: t + K may round up to 2^128
            // the directed variants of the same step (chunked GEMM fold): RoundTowardsZero, RoundDown, RoundUp
            const bool pos = (b >> 31) == 0u;
            const uint32_t a = b & 0x7FFFFFFFu, sgn = b & 0x80000000u, keep = ~((1u << r) - 1u);
            const float two_r = __uint_as_float((uint32_t)(127 + r) << 23), x = __uint_as_float(b);
            const uint32_t w_rz = (round_bits<LOF_RoundTowardsZero, uint32_t>(a, r, 0, pos, 0) & keep) | sgn;
            const uint32_t w_rd = (round_bits<LOF_RoundDown, uint32_t>(a, r, 0, pos, 0) & keep) | sgn;
            const uint32_t w_ru = (round_bits<LOF_RoundUp, uint32_t>(a, r, 0, pos, 0) & keep) | sgn;
            const float y_rz = round_mant_float<LOF_RoundTowardsZero>(x, two_r), y_rd = round_mant_float<LOF_RoundDown>(x, two_r),
                        y_ru = round_mant_float<LOF_RoundUp>(x, two_r);
            const bool top = (int)field + r == 254;                              // top binade: a round-up may reach 2^128 = inf
            bad += __float_as_uint(y_rz) != w_rz;
            bad += __float_as_uint(y_rd) != w_rd && !(top && !(fabsf(y_rd) < __int_as_float(0x7F800000)));
            bad += __float_as_uint(y_ru) != w_ru && !(top && !(fabsf(y_ru) < __int_as_float(0x7F800000)));
        }
    }
    if (bad) atomicAdd(mismatches, bad);
}

}  // namespace

cudaError_t launch_selftest_rne_round(unsigned long long* mismatches, int num_sms, cudaStream_t stream)
{
    selftest_rne_round_kernel<<<num_sms * 8, 256, 0, stream>>>(mismatches);
    return cudaGetLastError();
}

cudaError_t launch_gemm_exact(int mode, const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K,
                              int accum_mant_bits, int len, int64_t kc, float alpha, float beta, const RngArgs& rng,
                              cudaStream_t stream, int64_t batch, int64_t stride_a, int64_t stride_b, int64_t stride_c, const ExactView* view,
                              int fast)
{
    ExactView v = view ? *view : ExactView{};
    if (M <= 0 || N <= 0 || batch <= 0) return cudaSuccess;
    const bool custom = accum_mant_bits >= 0;
    const int r = custom ? 23 - accum_mant_bits : 0;
    int kci = (kc <= 0 || kc > K) ? (int)(K > 0 ? K : 1) : (int)kc;
    const int64_t mt = (M + TM - 1) / TM, nt = (N + TN - 1) / TN;
    v.swap_xy = mt > 65535;                                   // grid.y / grid.z hold at most 65535 blocks
    if ((v.swap_xy ? nt : mt) > 65535 || batch > 65535) return cudaErrorInvalidConfiguration;
    dim3 grid((unsigned)(v.swap_xy ? mt : nt), (unsigned)(v.swap_xy ? nt : mt), (unsigned)batch);
    switch (mode) {
#define LOF_CASE(Mo) case Mo: exact_kernel_for<Mo>()<<<grid, kGemmThreads, 0, stream>>>(A, B, C, (int)M, (int)N, (int)K, r, len, kci, alpha, beta, custom, fast, rng, stride_a, stride_b, stride_c, v); break;
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

}  // namespace lof
