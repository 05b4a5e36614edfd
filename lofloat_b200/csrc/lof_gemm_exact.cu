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

template <int MODE>
__global__ void __launch_bounds__(kGemmThreads)
gemm_exact_kernel(const float* __restrict__ A, const float* __restrict__ B, float* C, int M, int N, int K,
                  int r, int len, int kc, float alpha, float beta, bool custom, RngArgs rng,
                  int64_t stride_a, int64_t stride_b, int64_t stride_c)
{
    // batched call (blockIdx.z = batch index): the (b, h) loop of LoFloat/layers.py:142-149 as one launch; the stochastic
    // event index continues across the batch ((batch * M + i) * N + j) * K + k
    A += (int64_t)blockIdx.z * stride_a;
    B += (int64_t)blockIdx.z * stride_b;
    C += (int64_t)blockIdx.z * stride_c;
    const uint64_t rng_base = (uint64_t)blockIdx.z * (uint64_t)M * (uint64_t)N * (uint64_t)K;
    __shared__ float As[TK][TM + 4];
    __shared__ float Bs[TK][TN + 4];

    const int tx = threadIdx.x & 15;     // n direction
    const int ty = threadIdx.x >> 4;     // m direction
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;

    float crun[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) crun[i][j] = 0.0f;

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
                    float v = 0.0f;
                    if (ak < kt && m0 + m < M) v = A[(int64_t)(m0 + m) * K + k0 + ak];
                    As[ak][m] = v;
                }
                const int bn = threadIdx.x & 63, bk = threadIdx.x >> 6;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int k = bk + 4 * i;
                    float v = 0.0f;
                    if (k < kt && n0 + bn < N) v = B[(int64_t)(k0 + k) * N + n0 + bn];
                    Bs[k][bn] = v;
                }
            }
            __syncthreads();
            for (int k = 0; k < kt; ++k) {
                const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
                const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
                const float a4[4] = {av.x, av.y, av.z, av.w};
                const float b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if (custom) {
                            uint32_t w = 0;
                            if constexpr (ModeTraits<MODE>::kUsesRng) {
                                const uint64_t gi = (uint64_t)(m0 + ty * 4 + i), gj = (uint64_t)(n0 + tx * 4 + j);
                                w = philox_word(rng, rng_base + (gi * (uint64_t)N + gj) * (uint64_t)K + (uint64_t)(k0 + k));
                            }
                            acc[i][j] = accum_round<MODE>(rto_fma(a4[i], b4[j], acc[i][j]), r, len, w);
                        } else {
                            acc[i][j] = __fadd_rn(acc[i][j], __fmul_rn(a4[i], b4[j]));   // blis_lof_gemm.hpp:233-239
                        }
                    }
            }
            __syncthreads();
        }
        // fold the panel: ab *= alpha; C = first ? (beta == 0 ? ab : beta*C + ab) : C + ab
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
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

}  // namespace

cudaError_t launch_gemm_exact(int mode, const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K,
                              int accum_mant_bits, int len, int64_t kc, float alpha, float beta, const RngArgs& rng,
                              cudaStream_t stream, int64_t batch, int64_t stride_a, int64_t stride_b, int64_t stride_c)
{
    if (M <= 0 || N <= 0 || batch <= 0) return cudaSuccess;
    const bool custom = accum_mant_bits >= 0;
    const int r = custom ? 23 - accum_mant_bits : 0;
    int kci = (kc <= 0 || kc > K) ? (int)(K > 0 ? K : 1) : (int)kc;
    dim3 grid((unsigned)((N + TN - 1) / TN), (unsigned)((M + TM - 1) / TM), (unsigned)batch);
    switch (mode) {
#define LOF_CASE(Mo) case Mo: gemm_exact_kernel<Mo><<<grid, kGemmThreads, 0, stream>>>(A, B, C, (int)M, (int)N, (int)K, r, len, kci, alpha, beta, custom, rng, stride_a, stride_b, stride_c); break;
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
