// lof_gemm_tc.cu -- chunked tensor-core LoFloat-accumulation GEMM for sm_100a (tcgen05 + TMEM + TMA).
//
// Semantics (the design intent of /root/reference/src/gemms.hpp:430-552 -- a k-chunk micro-accumulator
// folded into a low-precision running accumulator -- which has no runnable reference, SURVEY.md 2.1 #9;
// the CPU statement of these semantics is oracle/lof_oracle.c: lofo_gemm_chunked_f32):
//
//     acc = 0
//     for every k-chunk c:   P_c = sum_{k in chunk} a_ik * b_kj          (tensor core, fp32 accumulate in TMEM)
//         S1 (default):      acc = round( acc + P_c )                     one rounding of the exact sum
//         S2 (PARTIAL_ROUND) acc = round( acc + round(P_c) )              gemms.hpp's typed micro-accumulator
//     C = alpha * acc
//
// where round() is LoFloat's virtual_round into the accumulator format -- mantissa-only
// (lo_float.h:2715-2756, what the reference GEMM leaf uses, cutlass_gemms.cuh:122) or a full
// FloatingPointParams format such as E5M10 (lo_float.h:2843-2939) -- under any deterministic rounding mode.
//
// Data path per CTA (one 128 x 128 output tile, 384 threads):
//   warp 0 / lane 0 : TMA producer.  A and B^T operand tiles [128 rows x 128 bytes of K] arrive in shared
//                     memory with the 128-byte swizzle through a 4-stage mbarrier ring.
//   warp 1 / lane 0 : MMA issuer.  One tcgen05.mma consumes 32 bytes of K per row (K = 32 for 8-bit operands
//                     with kind::f8f6f4, K = 16 for bf16 with kind::f16); the first MMA of a chunk overwrites
//                     its TMEM stage (accumulate = 0), the last one is followed by a tcgen05.commit on the
//                     stage's "full" mbarrier.  4 TMEM stages x 128 columns = all 512 columns, so the tensor
//                     core runs up to 3 chunks ahead of the rounding warps.
//   warp 2          : TMEM allocation / deallocation.
//   warps 4..11     : accumulator owners.  Warp w reads TMEM lanes 32*(w%4).. (hardware lane-quarter rule) and
//                     columns 64*((w-4)/4)..; every thread keeps its 64 running accumulators in REGISTERS for
//                     the whole K loop.  Per chunk: tcgen05.ld the partial, release the TMEM stage, then
//                     acc = round(acc + P) on the CUDA cores while the tensor core works on the next chunks.
//
// The k-chunk re-rounding makes this kernel CUDA-core bound, not tensor bound: at k_chunk = 32 every MMA
// (64 clk at FP8 rate) is paired with 16384 accumulator roundings (~5-13 instructions each).  Two mitigations:
//   * the rounding epilogue for fp16-like (E5M10) / bf16-like accumulators under RNE + S2 runs on packed
//     half2 / bfloat162 hardware: cvt.rn.{f16x2,bf16x2}.f32 + HADD2 = 1 instruction per accumulator per chunk.
//     It is bit-identical to the generic routine whenever no accumulator leaves the finite fp16 range (the
//     kernel checks, reports and re-runs the generic routine otherwise) -- see packed_ok() below.
//   * larger k-chunks amortise the rounding (k_chunk = 512: 16x fewer roundings).
//
// Operands: A (M x K) and B (K x N) are fp32 containers whose values must be exactly representable in the
// tensor-core operand type (e4m3 / e5m2 / bf16); tc_convert_* quantise them into a workspace (B transposed to
// K-major), verify the round trip and raise status bits otherwise.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>

#include <algorithm>
#include <cstring>
#include <string>
#include <type_traits>

#include "lof_core.cuh"
#include "lof_fast_f32.cuh"
#include "lof_gemm_common.cuh"
#include "lof_ptx.cuh"

namespace lof {

namespace {

using namespace ptx;

constexpr int BM = 128, BN = 128;
constexpr int kRowBytes = 128;              // bytes of K per shared-memory row (one 128B swizzle atom row)
constexpr int kStepBytes = 32;              // bytes of K per tcgen05.mma
constexpr int kStepsPerStage = kRowBytes / kStepBytes;
constexpr int kStages = 4;                  // shared-memory ring
constexpr int kTmemStages = 4;              // x BN fp32 columns = 512 = all of TMEM
constexpr int kStageBytesA = BM * kRowBytes;
constexpr int kStageBytesB = BN * kRowBytes;
constexpr int kSmemBytes = kStages * (kStageBytesA + kStageBytesB) + 1024;   // + alignment slack
constexpr int kTcThreads = 384;
constexpr int kEpiWarps = 8;
constexpr int kColsPerThread = BN / 2;      // 64

enum : int { kEpiMant = 0, kEpiFmt = 1, kEpiHalf2 = 2, kEpiBf162 = 3,
             kEpiHw16 = 4,      // the tensor core itself keeps the accumulator in fp16 (c_format = F16): no per-chunk read-back
             kEpiMantFloat = 5 };
struct TcArgs {
    int M, N, K;
    int k_steps;            // ceil(K * esize / 32): number of tcgen05.mma per output tile
    int steps_per_chunk;    // k_chunk * esize / 32
    uint32_t idesc;
    int kind_f8;            // 1: kind::f8f6f4, 0: kind::f16
    float alpha;
    int r;                  // mantissa-only: fp32 mantissa bits to drop (23 - accum_mant_bits)
    float two_r;            // 2^r (float-domain fold)
    RoundPlan plan;         // full-format accumulator
    FastPlan fast;
    int* status;            // device status word (never NULL)
    int run_mask, run_value;    // the kernel runs iff (*status & run_mask) == run_value
    int tiles_m, tiles_n;
    // stochastic accumulator rounding (persistent fold kernel, mantissa-only accumulators): one dense Philox word per rounding event,
    // event index = ((which * n_chunks + chunk) * M + row) * N + col, which = 0 for the partial's rounding (PARTIAL_ROUND), 1 for the accumulator's
    RngArgs rng;
    int stoch_len;
    int n_chunks;
};

// ------------------------------------------------------------------------------------------------
// accumulator update  acc <- round(acc + P)
// ------------------------------------------------------------------------------------------------
template <int MODE, int EPI>
__device__ __forceinline__ float round_acc(float t, const TcArgs& a, uint32_t w = 0u)
{
    if constexpr (EPI == kEpiMant) {
        return __uint_as_float(round_mantissa_elem<LOF_F32, MODE>(__float_as_uint(t), a.r, ModeTraits<MODE>::kUsesRng ? a.stoch_len : 0, w));
    } else {
        uint32_t rb;
        if (a.fast.valid) {
            bool rare = false;
            rb = round_fast_f32_rt<MODE>(__float_as_uint(t), a.fast, a.plan.saturating != 0, w, rare);
            if (rare) rb = exact_params_f32<MODE>(__float_as_uint(t), a.plan, w);
        } else {
            rb = exact_params_f32<MODE>(__float_as_uint(t), a.plan, w);
        }
        return __uint_as_float(rb);
    }
}

template <int MODE, int EPI, bool PARTIAL>
__device__ __forceinline__ float fold(float acc, float p, const TcArgs& a, uint32_t w_part = 0u, uint32_t w_acc = 0u)
{
    if constexpr (EPI == kEpiMantFloat) {
        // mantissa-only accumulator, float-domain rounding (round_mant_float, lof_gemm_common.cuh): RNE / RTZ / RD / RU.
        // The directed modes compose with the directed fp32 add (truncations / floors of truncations / floors); RNE needs the
        // round-to-odd add so that the format rounding is the only rounding -- except under PARTIAL with a mantissa of <= 10 bits,
        // where both summands are (m+1)-bit numbers and fp32 (24 >= 2 (m+1) + 2 bits) makes the double rounding innocuous.
        const float two_r = a.two_r;
        if constexpr (PARTIAL) p = round_mant_float<MODE>(p, two_r);
        float t;
        if constexpr (MODE == LOF_RoundTowardsZero) t = __fadd_rz(acc, p);
        else if constexpr (MODE == LOF_RoundDown) t = __fadd_rd(acc, p);
        else if constexpr (MODE == LOF_RoundUp) t = __fadd_ru(acc, p);
        else if constexpr (PARTIAL) t = a.r >= 13 ? __fadd_rn(acc, p) : rto_add(acc, p);
        else t = rto_add(acc, p);
        return round_mant_float<MODE>(t, two_r);
    }
    if constexpr (PARTIAL) p = round_acc<MODE, EPI>(p, a, w_part);
    return round_acc<MODE, EPI>(rto_add(acc, p), a, w_acc);    // round-to-odd add: the format rounding is the only rounding
}

// ------------------------------------------------------------------------------------------------
// the GEMM kernel
// ------------------------------------------------------------------------------------------------
template <int MODE, int EPI, bool PARTIAL>
__global__ void __launch_bounds__(kTcThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, float* __restrict__ C,
               const TcArgs args)
{
    if ((*args.status & args.run_mask) != args.run_value) return;      // device-side conditional launch (uniform)

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + kStages * kStageBytesA;
    __shared__ __align__(8) uint64_t full_bar[kStages], empty_bar[kStages];
    __shared__ __align__(8) uint64_t tmem_full_bar[kTmemStages], tmem_empty_bar[kTmemStages];
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // output tile of this CTA: groups of 16 tile-rows are walked column by column so that concurrently
    // resident CTAs share A / B panels in L2
    int tile_m, tile_n;
    {
        constexpr int GM = 16;
        const int pid = blockIdx.x;
        const int per_group = GM * args.tiles_n;
        const int group = pid / per_group;
        const int first_m = group * GM;
        const int gsz = min(args.tiles_m - first_m, GM);
        tile_m = first_m + (pid % per_group) % gsz;
        tile_n = (pid % per_group) / gsz;
    }
    const int m0 = tile_m * BM, n0 = tile_n * BN;
    const int n_chunks = (args.k_steps + args.steps_per_chunk - 1) / args.steps_per_chunk;
    const int n_stage_iters = (args.k_steps + kStepsPerStage - 1) / kStepsPerStage;

    if (threadIdx.x == 0) {
        prefetch_tensormap(&map_a);
        prefetch_tensormap(&map_b);
#pragma unroll
        for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
#pragma unroll
        for (int s = 0; s < kTmemStages; ++s) { mbar_init(&tmem_full_bar[s], 1); mbar_init(&tmem_empty_bar[s], kEpiWarps); }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc(&tmem_base_slot, kTmemStages * BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 0) {
        // ===================================== TMA producer =====================================
        if (elect_one()) {
            for (int it = 0; it < n_stage_iters; ++it) {
                const int s = it % kStages;
                mbar_wait(&empty_bar[s], (uint32_t)(((it / kStages) & 1) ^ 1));
                mbar_expect_tx(&full_bar[s], kStageBytesA + kStageBytesB);
                tma_load_2d(smem_a + s * kStageBytesA, &map_a, it * kRowBytes, m0, &full_bar[s]);   // coords in BYTES of K (uint8 map)
                tma_load_2d(smem_b + s * kStageBytesB, &map_b, it * kRowBytes, n0, &full_bar[s]);
            }
        }
    } else if (warp == 1) {
        // ====================================== MMA issuer ======================================
        if (elect_one()) {
            int step = 0;
            for (int c = 0; c < n_chunks; ++c) {
                const int ts = c % kTmemStages;
                mbar_wait(&tmem_empty_bar[ts], (uint32_t)(((c / kTmemStages) & 1) ^ 1));
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(ts * BN);
                const int steps_here = min(args.steps_per_chunk, args.k_steps - step);
                for (int j = 0; j < steps_here; ++j, ++step) {
                    const int it = step / kStepsPerStage, sub = step % kStepsPerStage, s = it % kStages;
                    if (sub == 0) {
                        mbar_wait(&full_bar[s], (uint32_t)((it / kStages) & 1));
                        tc_fence_after();
                    }
                    const uint64_t adesc = umma_desc_k128(smem_u32(smem_a + s * kStageBytesA) + sub * kStepBytes);
                    const uint64_t bdesc = umma_desc_k128(smem_u32(smem_b + s * kStageBytesB) + sub * kStepBytes);
                    umma(tmem_d, adesc, bdesc, args.idesc, j > 0 ? 1u : 0u, args.kind_f8);
                    if (sub == kStepsPerStage - 1 || step == args.k_steps - 1) umma_commit(&empty_bar[s]);   // stage consumed
                }
                umma_commit(&tmem_full_bar[ts]);                                                             // chunk ready
            }
        }
    } else if (warp >= 4) {
        // ============================ accumulator owners (rounding epilogue) ============================
        const int q = warp & 3;                     // TMEM lane quarter this warp may access
        const int h = (warp - 4) >> 2;              // column half
        const uint32_t t_lane = (uint32_t)(q * 32) << 16;
        const int row = m0 + q * 32 + lane;
        const int col0 = n0 + h * kColsPerThread;
        constexpr bool kPacked = (EPI == kEpiHalf2 || EPI == kEpiBf162);

        float acc[kPacked ? 1 : kColsPerThread];
        uint32_t accp[kPacked ? kColsPerThread / 2 : 1];
        if constexpr (kPacked) {
#pragma unroll
            for (int i = 0; i < kColsPerThread / 2; ++i) accp[i] = 0u;
        } else {
#pragma unroll
            for (int i = 0; i < kColsPerThread; ++i) acc[i] = 0.0f;
        }

        // The partials are pulled out of TMEM in UNITs of columns through two register buffers: the
        // tcgen05.ld of unit u+1 is in flight while unit u is folded into the accumulators (tcgen05.wait::ld
        // waits for every outstanding load of the thread, so the prefetch distance is one unit).  The packed
        // epilogue has few accumulator registers and takes a whole chunk (64 columns) per unit; the generic one
        // takes half a chunk.
        constexpr int UNIT = kPacked ? kColsPerThread : kColsPerThread / 2;
        constexpr int UPC = kColsPerThread / UNIT;               // units per chunk (1 or 2)
        uint32_t buf0[UNIT], buf1[UNIT];
        const int total_units = n_chunks * UPC;

        auto issue = [&](int u, uint32_t* buf) {                 // start loading unit u
            const int c = u / UPC, part = u % UPC, ts = c % kTmemStages;
            if (part == 0) {
                mbar_wait(&tmem_full_bar[ts], (uint32_t)((c / kTmemStages) & 1));
                tc_fence_after();
            }
            const uint32_t taddr = tmem_base + t_lane + (uint32_t)(ts * BN + h * kColsPerThread + part * UNIT);
#pragma unroll
            for (int k = 0; k < UNIT / 32; ++k) tmem_ld32(taddr + 32 * k, buf + 32 * k);
        };
        auto landed = [&](int u, uint32_t* buf) {                // unit u is in registers; release its TMEM stage when it was the chunk's last
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < UNIT; ++i) asm volatile("" : "+r"(buf[i]));     // order every later use after the wait
            if (u % UPC == UPC - 1) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty_bar[(u / UPC) % kTmemStages]);
            }
        };
        auto fold_unit = [&](auto part_tag, const uint32_t* p) {
            constexpr int PART = decltype(part_tag)::value;
            if constexpr (EPI == kEpiHalf2) {
#pragma unroll
                for (int i = 0; i < UNIT / 2; ++i) {
                    const __half2 a0 = *reinterpret_cast<const __half2*>(&accp[i]);
                    const __half2 r0 = __hadd2(a0, __floats2half2_rn(__uint_as_float(p[2 * i]), __uint_as_float(p[2 * i + 1])));
                    accp[i] = *reinterpret_cast<const uint32_t*>(&r0);
                }
            } else if constexpr (EPI == kEpiBf162) {
#pragma unroll
                for (int i = 0; i < UNIT / 2; ++i) {
                    const __nv_bfloat162 a0 = *reinterpret_cast<const __nv_bfloat162*>(&accp[i]);
                    const __nv_bfloat162 r0 = __hadd2(a0, __floats2bfloat162_rn(__uint_as_float(p[2 * i]), __uint_as_float(p[2 * i + 1])));
                    accp[i] = *reinterpret_cast<const uint32_t*>(&r0);
                }
            } else if constexpr (EPI == kEpiHw16) {
#pragma unroll
                for (int i = 0; i < UNIT; ++i)                    // fp16 accumulator: one element per 32-bit TMEM column, low half
                    acc[PART * UNIT + i] = __half2float(__ushort_as_half((unsigned short)(p[i] & 0xFFFFu)));
            } else {
#pragma unroll
                for (int i = 0; i < UNIT; ++i)
                    acc[PART * UNIT + i] = fold<MODE, EPI, PARTIAL>(acc[PART * UNIT + i], __uint_as_float(p[i]), args);
            }
        };
        using Part0 = std::integral_constant<int, 0>;
        using Part1 = std::integral_constant<int, UPC - 1>;      // second half-chunk (== Part0 when a unit is a whole chunk)

        issue(0, buf0);
        for (int u = 0; u < total_units; u += 2) {
            landed(u, buf0);
            if (u + 1 < total_units) issue(u + 1, buf1);
            fold_unit(Part0{}, buf0);
            if (u + 1 >= total_units) break;
            landed(u + 1, buf1);
            if (u + 2 < total_units) issue(u + 2, buf0);
            fold_unit(Part1{}, buf1);
        }

        // C = alpha * acc
        auto acc_at = [&](int i) -> float {
            if constexpr (EPI == kEpiHalf2) {
                const __half2 hv = *reinterpret_cast<const __half2*>(&accp[i >> 1]);
                return (i & 1) ? __high2float(hv) : __low2float(hv);
            } else if constexpr (EPI == kEpiBf162) {
                return (i & 1) ? __uint_as_float(accp[i >> 1] & 0xFFFF0000u) : __uint_as_float(accp[i >> 1] << 16);
            } else {
                return acc[i];
            }
        };
        if constexpr (EPI == kEpiHalf2) {
            bool bad = false;
#pragma unroll
            for (int i = 0; i < kColsPerThread / 2; ++i)          // inf / NaN: an accumulator left the finite fp16 range
                bad |= ((accp[i] & 0x7C00u) == 0x7C00u) || ((accp[i] & 0x7C000000u) == 0x7C000000u);
            if (bad) atomicOr(args.status, LOF_GEMM_STATUS_PACKED_RANGE);
        }
        if constexpr (EPI == kEpiHw16) {
            bool bad = false;
#pragma unroll
            for (int i = 0; i < kColsPerThread; ++i) bad |= (__float_as_uint(acc[i]) & 0x7F800000u) == 0x7F800000u;
            if (bad) atomicOr(args.status, LOF_GEMM_STATUS_PACKED_RANGE);
        }
        if (row < args.M) {
            float* crow = C + (int64_t)row * args.N + col0;
            const bool vec_ok = ((args.N & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15u) == 0);
#pragma unroll
            for (int i = 0; i < kColsPerThread; i += 4) {
                const float4 v = make_float4(__fmul_rn(acc_at(i), args.alpha), __fmul_rn(acc_at(i + 1), args.alpha),
                                             __fmul_rn(acc_at(i + 2), args.alpha), __fmul_rn(acc_at(i + 3), args.alpha));
                if (vec_ok && col0 + i + 3 < args.N) {
                    *reinterpret_cast<float4*>(crow + i) = v;
                } else {
                    if (col0 + i < args.N) crow[i] = v.x;
                    if (col0 + i + 1 < args.N) crow[i + 1] = v.y;
                    if (col0 + i + 2 < args.N) crow[i + 2] = v.z;
                    if (col0 + i + 3 < args.N) crow[i + 3] = v.w;
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, kTmemStages * BN);
}

#include "lof_gemm_tc2.cuh"
#include "lof_gemm_tc3.cuh"

// ------------------------------------------------------------------------------------------------
// operand preparation: fp32 -> tensor-core operand bytes (+ exact-representability check)
// ------------------------------------------------------------------------------------------------
enum : int { kOpE4M3 = 0, kOpE5M2 = 1, kOpBF16 = 2 };

template <int OP> struct OpT;
template <> struct OpT<kOpE4M3> {
    using T = uint8_t;
    __device__ static __forceinline__ T cvt(float x, bool& bad)
    {
        const __nv_fp8_storage_t b = __nv_cvt_float_to_fp8(x, __NV_SATFINITE, __NV_E4M3);
        const float back = __half2float(__half(__nv_cvt_fp8_to_halfraw(b, __NV_E4M3)));
        bad |= !(back == x);
        return b;
    }
};
template <> struct OpT<kOpE5M2> {
    using T = uint8_t;
    __device__ static __forceinline__ T cvt(float x, bool& bad)
    {
        const __nv_fp8_storage_t b = __nv_cvt_float_to_fp8(x, __NV_SATFINITE, __NV_E5M2);
        const float back = __half2float(__half(__nv_cvt_fp8_to_halfraw(b, __NV_E5M2)));
        bad |= !(back == x);
        return b;
    }
};
template <> struct OpT<kOpBF16> {
    using T = uint16_t;
    __device__ static __forceinline__ T cvt(float x, bool& bad)
    {
        const uint32_t u = __float_as_uint(x);
        bad |= (u & 0xFFFFu) != 0u || (u & 0x7F800000u) == 0x7F800000u;     // low bits set, or inf / NaN
        return (uint16_t)(u >> 16);
    }
};

// A: [rows, cols] fp32 -> [rows, cols] operand type, same (K-major) layout
template <int OP>
__global__ void __launch_bounds__(256)
tc_convert_rows_kernel(const float* __restrict__ X, typename OpT<OP>::T* __restrict__ out, int64_t n, int* status, int bad_bit)
{
    using T = typename OpT<OP>::T;
    bool bad = false;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
        if (i + 3 < n && ((reinterpret_cast<uintptr_t>(X) & 15u) == 0)) {
            const float4 v = *reinterpret_cast<const float4*>(X + i);
            T o[4] = {OpT<OP>::cvt(v.x, bad), OpT<OP>::cvt(v.y, bad), OpT<OP>::cvt(v.z, bad), OpT<OP>::cvt(v.w, bad)};
            if constexpr (sizeof(T) == 1) *reinterpret_cast<uchar4*>(out + i) = make_uchar4(o[0], o[1], o[2], o[3]);
            else *reinterpret_cast<ushort4*>(out + i) = make_ushort4(o[0], o[1], o[2], o[3]);
        } else {
            for (int e = 0; e < 4 && i + e < n; ++e) out[i + e] = OpT<OP>::cvt(X[i + e], bad);
        }
    }
    if (__any_sync(0xFFFFFFFFu, bad) && (threadIdx.x & 31) == 0) atomicOr(status, bad_bit);
}

// B: [K, N] fp32 row-major -> B^T [N, K] operand type (K-major), through a shared-memory tile
template <int OP>
__global__ void __launch_bounds__(256)
tc_convert_transpose_kernel(const float* __restrict__ B, typename OpT<OP>::T* __restrict__ out, int K, int N, int* status, int bad_bit)
{
    using T = typename OpT<OP>::T;
    constexpr int TK = 128 / (int)sizeof(T);      // k values per tile (128 bytes of output row)
    constexpr int TN = 64;
    constexpr int PITCH = TK + 4 / (int)sizeof(T);          // 132-byte rows: transposed writes hit 32 distinct banks
    __shared__ __align__(16) T tile[TN][PITCH];
    const int k0 = blockIdx.y * TK, n0 = blockIdx.x * TN;
    bool bad = false;
    for (int idx = threadIdx.x; idx < TK * TN; idx += 256) {
        const int k = idx / TN, n = idx % TN;                  // consecutive threads -> consecutive n: coalesced reads
        float v = 0.0f;
        if (k0 + k < K && n0 + n < N) v = B[(int64_t)(k0 + k) * N + n0 + n];
        tile[n][k] = OpT<OP>::cvt(v, bad);
    }
    __syncthreads();
    constexpr int SEGS = 128 / 16;                            // 16-byte segments per output row
    constexpr int EPS = 16 / (int)sizeof(T);                  // elements per segment
    for (int idx = threadIdx.x; idx < TN * SEGS; idx += 256) {
        const int n = idx / SEGS, seg = idx % SEGS;
        const int k = k0 + seg * EPS;
        if (n0 + n >= N || k >= K) continue;
        T* dst = out + (int64_t)(n0 + n) * K + k;
        if (k + EPS <= K && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
            const uint32_t* w = reinterpret_cast<const uint32_t*>(&tile[n][seg * EPS]);     // 4-byte aligned (132-byte pitch)
            *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
        } else {
            for (int e = 0; e < EPS && k + e < K; ++e) dst[e] = tile[n][seg * EPS + e];
        }
    }
    if (__any_sync(0xFFFFFFFFu, bad) && (threadIdx.x & 31) == 0) atomicOr(status, bad_bit);
}

__global__ void tc_clear_status_kernel(int* status) { *status = 0; }
__global__ void tc_or_status_kernel(int* dst, const int* src) { const int v = *src; if (v) atomicOr(dst, v); }

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn get_encode()
{
    static EncodeFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeFn>(p);
    }
    return fn;
}

// operand matrix [rows, k_bytes] as a uint8 tensor: box = 128 bytes x 128 rows, 128B swizzle, zero fill out of bounds
bool make_operand_map(CUtensorMap* map, const void* base, int64_t rows, int64_t k_bytes)
{
    EncodeFn enc = get_encode();
    if (!enc) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)k_bytes, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)k_bytes};
    const cuuint32_t box[2] = {(cuuint32_t)kRowBytes, (cuuint32_t)BM};
    const cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

uint32_t make_idesc(int op, bool f16_accum = false, int mma_m = BM, int mma_n = BN)
{
    // tcgen05 instruction descriptor: c_format F32 (bits 4-5 = 1), a/b format (bits 7-9 / 10-12), K-major A and B
    // (bits 15, 16 = 0), N >> 3 (bits 17-22), M >> 4 (bits 24-28)
    const uint32_t fmt = (op == kOpE4M3) ? 0u : (op == kOpE5M2) ? 1u : 1u /* kind::f16: 1 = BF16 */;
    return ((f16_accum ? 0u : 1u) << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(mma_n >> 3) << 17) | ((uint32_t)(mma_m >> 4) << 24);
}

template <int MODE, int EPI, bool PARTIAL>
cudaError_t launch_one(const CUtensorMap& ma, const CUtensorMap& mb, float* C, const TcArgs& a, cudaStream_t stream)
{
    auto kern = gemm_tc_kernel<MODE, EPI, PARTIAL>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return e;
    kern<<<(unsigned)(a.tiles_m * a.tiles_n), kTcThreads, kSmemBytes, stream>>>(ma, mb, C, a);
    return cudaGetLastError();
}

template <int EPI, bool PARTIAL>
cudaError_t launch_mode(int mode, const CUtensorMap& ma, const CUtensorMap& mb, float* C, const TcArgs& a, cudaStream_t stream)
{
    switch (mode) {
#define LOF_CASE(Mo) case Mo: return launch_one<Mo, EPI, PARTIAL>(ma, mb, C, a, stream);
        LOF_CASE(LOF_RoundToNearestEven) LOF_CASE(LOF_RoundTowardsZero) LOF_CASE(LOF_RoundAwayFromZero)
        LOF_CASE(LOF_RoundToNearestOdd) LOF_CASE(LOF_RoundDown) LOF_CASE(LOF_RoundUp) LOF_CASE(LOF_RoundTiesToAway)
        LOF_CASE(LOF_RoundToOdd) LOF_CASE(LOF_RoundTiesTowardsZero)
#undef LOF_CASE
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_generic(int mode, bool fmt, bool partial, const CUtensorMap& ma, const CUtensorMap& mb, float* C,
                           const TcArgs& a, cudaStream_t stream)
{
    if (fmt) return partial ? launch_mode<kEpiFmt, true>(mode, ma, mb, C, a, stream) : launch_mode<kEpiFmt, false>(mode, ma, mb, C, a, stream);
    return partial ? launch_mode<kEpiMant, true>(mode, ma, mb, C, a, stream) : launch_mode<kEpiMant, false>(mode, ma, mb, C, a, stream);
}

// third-generation persistent fold kernel (lof_gemm_tc3.cuh)
template <int EPI, bool PARTIAL>
cudaError_t launch_fold_mode(int mode, const CUtensorMap& ma, const CUtensorMap& mb, float* C, const TcArgs& a, int sms, cudaStream_t stream)
{
    switch (mode) {
#define LOF_CASE(Mo) case Mo: return launch_fold_one<Mo, EPI, PARTIAL, 128, 1>(ma, mb, C, a, sms, stream);
        LOF_CASE(LOF_RoundToNearestEven) LOF_CASE(LOF_RoundTowardsZero) LOF_CASE(LOF_RoundAwayFromZero)
        LOF_CASE(LOF_RoundToNearestOdd) LOF_CASE(LOF_RoundDown) LOF_CASE(LOF_RoundUp) LOF_CASE(LOF_RoundTiesToAway)
        LOF_CASE(LOF_RoundToOdd) LOF_CASE(LOF_RoundTiesTowardsZero)
#undef LOF_CASE
    default: return cudaErrorInvalidValue;
    }
}

template <bool PARTIAL>
cudaError_t launch_fold_float(int mode, const CUtensorMap& ma, const CUtensorMap& mb, float* C, const TcArgs& a, int sms, cudaStream_t stream)
{
    switch (mode) {
    case LOF_RoundToNearestEven: return launch_fold_one<LOF_RoundToNearestEven, kEpiMantFloat, PARTIAL, 128, 1>(ma, mb, C, a, sms, stream);
    case LOF_RoundTowardsZero:   return launch_fold_one<LOF_RoundTowardsZero, kEpiMantFloat, PARTIAL, 128, 1>(ma, mb, C, a, sms, stream);
    case LOF_RoundDown:          return launch_fold_one<LOF_RoundDown, kEpiMantFloat, PARTIAL, 128, 1>(ma, mb, C, a, sms, stream);
    case LOF_RoundUp:            return launch_fold_one<LOF_RoundUp, kEpiMantFloat, PARTIAL, 128, 1>(ma, mb, C, a, sms, stream);
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_fold_generic(int mode, bool fmt, bool partial, const CUtensorMap& ma, const CUtensorMap& mb, float* C,
                                const TcArgs& a, int sms, cudaStream_t stream, bool float_ok = false)
{
    // float-domain fold: mantissa-only accumulator, 1 <= r <= 22, RNE / RTZ / RD / RU, e4m3 operands (float_ok: every partial and
    // accumulator is a multiple of 2^-18 below 2^50 -- normal or zero, far from the overflow of K = 2^(e+r))
    if (float_ok && !fmt && a.r >= 1 && a.r <= 22 && mode_has_float_round(mode))
        return partial ? launch_fold_float<true>(mode, ma, mb, C, a, sms, stream) : launch_fold_float<false>(mode, ma, mb, C, a, sms, stream);
    if (fmt) return partial ? launch_fold_mode<kEpiFmt, true>(mode, ma, mb, C, a, sms, stream) : launch_fold_mode<kEpiFmt, false>(mode, ma, mb, C, a, sms, stream);
    return partial ? launch_fold_mode<kEpiMant, true>(mode, ma, mb, C, a, sms, stream) : launch_fold_mode<kEpiMant, false>(mode, ma, mb, C, a, sms, stream);
}

template <bool PARTIAL>
cudaError_t launch_fold_stochastic(int mode, const CUtensorMap& ma, const CUtensorMap& mb, float* C, const TcArgs& a, int sms, cudaStream_t stream)
{
    switch (mode) {
#define LOF_CASE(Mo) case Mo: return launch_fold_one<Mo, kEpiMant, PARTIAL, 128, 1>(ma, mb, C, a, sms, stream);
        LOF_CASE(LOF_StochasticRoundingA) LOF_CASE(LOF_StochasticRoundingB) LOF_CASE(LOF_StochasticRoundingC)
        LOF_CASE(LOF_StochasticRoundingD) LOF_CASE(LOF_True_StochasticRounding) LOF_CASE(LOF_ProbabilisticRounding)
#undef LOF_CASE
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_fold_packed(int packed, bool split2, const CUtensorMap& ma, const CUtensorMap& mb, float* C, const TcArgs& a, int sms,
                               cudaStream_t stream)
{
    constexpr int RNE = LOF_RoundToNearestEven;
    if (split2) {
        if (packed == kEpiHalf2) return launch_fold_one<RNE, kEpiHalf2, true, 256, 2>(ma, mb, C, a, sms, stream);
        return launch_fold_one<RNE, kEpiBf162, true, 256, 2>(ma, mb, C, a, sms, stream);
    }
    if (packed == kEpiHalf2) return launch_fold_one<RNE, kEpiHalf2, true, 256, 1>(ma, mb, C, a, sms, stream);
    return launch_fold_one<RNE, kEpiBf162, true, 256, 1>(ma, mb, C, a, sms, stream);
}

// Is the packed half2 / bfloat162 epilogue bit-identical to the generic routine for this request?
//   fp16 hardware == LoFloat accumulator (mantissa 10) iff every value that is ever rounded is either a normal
//   fp16 number or a multiple of 2^-24 (fp16's subnormal quantum), and nothing reaches 65520.  8-bit operands
//   guarantee the first part (every product of two e4m3 / e5m2 numbers is a multiple of 2^-18 resp. 2^-32 ...
//   only e4m3 qualifies); the second part is checked by the kernel (inf / NaN in the final accumulators).
//   bf16 hardware == mantissa-only 7-bit rounding of the fp32 container for every finite value.
int packed_epi(int op, int mode, bool partial, bool allow, int accum_mant_bits, const lof_format_t* fmt)
{
    if (!allow || !partial || mode != LOF_RoundToNearestEven) return -1;
    if (!fmt) {
        if (accum_mant_bits == 10 && op == kOpE4M3) return kEpiHalf2;
        if (accum_mant_bits == 7) return kEpiBf162;
        return -1;
    }
    const int ebits = fmt->bitwidth - fmt->mantissa_bits - (fmt->is_signed == LOF_Signed ? 1 : 0);
    if (fmt->mantissa_bits == 10 && ebits == 5 && fmt->bias == 15 && op == kOpE4M3) return kEpiHalf2;   // E5M10: IEEE half below 65520
    return -1;
}

// Can the tensor core keep the accumulator itself?  With c_format = F16 every tcgen05.mma performs
//     D_fp16 <- RNE_fp16( D_fp16 + sum_k a_k * b_k )          (one rounding of the exact sum; scripts/probe_tmem_f16.cu,
//                                                              profiles/r01_probe_tmem_f16.txt; fp16 incl. subnormals)
// which IS the single-rounding fold with an E5M10 accumulator under RNE when the k-chunk is exactly one MMA
// (K = 32 for 8-bit operands, 16 for bf16).  Valid while the accumulator stays finite (checked, generic recompute
// otherwise).  Mantissa-only 10-bit accumulators qualify for e4m3 operands only (see packed_epi).
bool hw_f16_accum_ok(int op, int mode, bool partial, bool allow, int accum_mant_bits, const lof_format_t* fmt, int64_t k_chunk, int kstep)
{
    if (!allow || partial || mode != LOF_RoundToNearestEven || k_chunk != kstep) return false;
    if (op == kOpBF16) return false;        // kind::f16 with bf16 operands has no F16 accumulator (illegal instruction on sm_100a)
    if (!fmt) return accum_mant_bits == 10 && op == kOpE4M3;
    const int ebits = fmt->bitwidth - fmt->mantissa_bits - (fmt->is_signed == LOF_Signed ? 1 : 0);
    return fmt->mantissa_bits == 10 && ebits == 5 && fmt->bias == 15;
}

}  // namespace

// operand_dtype: LOF_BF16, LOF_E4M3 or LOF_E5M2.  *err is set for unsupported requests (no launch happened).
cudaError_t launch_gemm_tc(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K, int mode,
                           int accum_mant_bits, const lof_format_t* acc_fmt, const RoundPlan* acc_plan, const FastPlan* acc_fast,
                           int64_t k_chunk, float alpha, int operand_dtype, uint32_t flags, int* dev_status,
                           cudaStream_t stream, int* launches, std::string* err, const RngArgs* rng, int stoch_len)
{
    const int op = operand_dtype == LOF_E4M3 ? kOpE4M3 : operand_dtype == LOF_E5M2 ? kOpE5M2 : operand_dtype == LOF_BF16 ? kOpBF16 : -1;
    if (op < 0) { *err = "chunked GEMM: operand_dtype must be LOF_E4M3, LOF_E5M2 or LOF_BF16"; return cudaSuccess; }
    const int esize = (op == kOpBF16) ? 2 : 1;
    const int kstep = kStepBytes / esize;                                   // K per tcgen05.mma
    if (k_chunk <= 0 || k_chunk > K) k_chunk = ((K + kstep - 1) / kstep) * kstep;
    if (k_chunk % kstep != 0) { *err = "chunked GEMM: k_chunk must be a multiple of " + std::to_string(kstep) + " for this operand type"; return cudaSuccess; }
    if ((K * esize) % 16 != 0) { *err = "chunked GEMM: K * sizeof(operand) must be a multiple of 16 bytes (TMA row pitch)"; return cudaSuccess; }
    if (M > (1 << 24) || N > (1 << 24) || K > (1ll << 30)) { *err = "chunked GEMM: dimension too large"; return cudaSuccess; }
    const bool stochastic = mode_uses_rng(mode);
    if (stochastic && acc_fmt) { *err = "chunked GEMM: stochastic accumulator rounding needs a mantissa-only accumulator (accum_format = NULL)"; return cudaSuccess; }

    // operands already stored in the tensor core's operand type (lof_gemm_chunked_packed): A = M x K, B = Bt = N x K, e4m3 / e5m2
    // bytes or bf16 words
    const bool packed_in = (flags & LOF_GEMM_INTERNAL_PACKED_OPERANDS) != 0;
    if (packed_in) {
        if (((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15u) != 0) {
            *err = "chunked GEMM (packed operands): A and Bt must be 16-byte aligned (TMA)"; return cudaSuccess;
        }
    }
    // workspace: operand copies + status word, stream-ordered
    const size_t bytes_a = packed_in ? 0 : (size_t)M * K * esize, bytes_b = packed_in ? 0 : (size_t)N * K * esize;
    const size_t off_b = (bytes_a + 1023) & ~size_t(1023), off_s = off_b + ((bytes_b + 1023) & ~size_t(1023));
    uint8_t* ws = nullptr;
    {   // keep freed workspace blocks cached in the device's default pool (the default threshold of 0 returns them
        // to the driver at every synchronisation, which costs a real cudaMalloc per call)
        static bool pool_ready[64] = {false};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev >= 0 && dev < 64 && !pool_ready[dev]) {
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
                uint64_t keep = ~0ull;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            }
            pool_ready[dev] = true;
        }
    }
    cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&ws), off_s + 256, stream);
    if (e != cudaSuccess) return e;
    int* status = reinterpret_cast<int*>(ws + off_s);
    auto fail = [&](cudaError_t ce) { cudaFreeAsync(ws, stream); return ce; };

    tc_clear_status_kernel<<<1, 1, 0, stream>>>(status);
    ++*launches;
    if (!packed_in) {
        const int64_t na = M * K;
        const unsigned ga = (unsigned)std::min<int64_t>((na / 4 + 255) / 256 + 1, 148 * 16);
        const dim3 gb((unsigned)((N + 63) / 64), (unsigned)((K + (128 / esize) - 1) / (128 / esize)));
        switch (op) {
        case kOpE4M3:
            tc_convert_rows_kernel<kOpE4M3><<<ga, 256, 0, stream>>>(A, ws, na, status, LOF_GEMM_STATUS_A_INEXACT);
            tc_convert_transpose_kernel<kOpE4M3><<<gb, 256, 0, stream>>>(B, ws + off_b, (int)K, (int)N, status, LOF_GEMM_STATUS_B_INEXACT);
            break;
        case kOpE5M2:
            tc_convert_rows_kernel<kOpE5M2><<<ga, 256, 0, stream>>>(A, ws, na, status, LOF_GEMM_STATUS_A_INEXACT);
            tc_convert_transpose_kernel<kOpE5M2><<<gb, 256, 0, stream>>>(B, ws + off_b, (int)K, (int)N, status, LOF_GEMM_STATUS_B_INEXACT);
            break;
        default:
            tc_convert_rows_kernel<kOpBF16><<<ga, 256, 0, stream>>>(A, reinterpret_cast<uint16_t*>(ws), na, status, LOF_GEMM_STATUS_A_INEXACT);
            tc_convert_transpose_kernel<kOpBF16><<<gb, 256, 0, stream>>>(B, reinterpret_cast<uint16_t*>(ws + off_b), (int)K, (int)N, status, LOF_GEMM_STATUS_B_INEXACT);
            break;
        }
        *launches += 2;
        e = cudaGetLastError();
        if (e != cudaSuccess) return fail(e);
    }

    CUtensorMap ma, mb;
    const void* opa = packed_in ? static_cast<const void*>(A) : static_cast<const void*>(ws);
    const void* opb = packed_in ? static_cast<const void*>(B) : static_cast<const void*>(ws + off_b);
    if (!make_operand_map(&ma, opa, M, K * esize) || !make_operand_map(&mb, opb, N, K * esize)) {
        cudaFreeAsync(ws, stream);
        *err = "chunked GEMM: cuTensorMapEncodeTiled failed";
        return cudaSuccess;
    }

    TcArgs a;
    memset(&a, 0, sizeof a);
    a.M = (int)M; a.N = (int)N; a.K = (int)K;
    a.k_steps = (int)((K * esize + kStepBytes - 1) / kStepBytes);
    a.steps_per_chunk = (int)(k_chunk / kstep);
    a.idesc = make_idesc(op);
    a.kind_f8 = (op != kOpBF16);
    a.alpha = alpha;
    a.r = acc_fmt ? 0 : 23 - accum_mant_bits;
    {
        const uint32_t bits = (uint32_t)(127 + (a.r > 0 && a.r < 100 ? a.r : 0)) << 23;
        memcpy(&a.two_r, &bits, 4);
    }
    if (acc_plan) a.plan = *acc_plan;
    if (acc_fast) a.fast = *acc_fast;
    a.status = status;
    a.tiles_m = (int)((M + BM - 1) / BM);
    a.tiles_n = (int)((N + BN - 1) / BN);
    if (rng) a.rng = *rng;
    a.stoch_len = stoch_len;
    a.n_chunks = (a.k_steps + a.steps_per_chunk - 1) / a.steps_per_chunk;

    const bool partial = (flags & LOF_GEMM_PARTIAL_ROUND) != 0;
    const int packed = packed_epi(op, mode, partial, (flags & LOF_GEMM_NO_PACKED) == 0, accum_mant_bits, acc_fmt);
    const bool hw16 = hw_f16_accum_ok(op, mode, partial, (flags & (LOF_GEMM_NO_PACKED | LOF_GEMM_NO_HW_ACCUM)) == 0,
                                      accum_mant_bits, acc_fmt, k_chunk, kstep);
    // second-generation main loop (persistent CTA pairs, lof_gemm_tc2.cuh) for the accumulator semantics it serves
    const bool one_chunk = a.steps_per_chunk >= a.k_steps;
    int sms = 0, dev_id = 0;
    cudaGetDevice(&dev_id);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev_id);
    if (stochastic) {       // persistent fold kernel, generic fold with one Philox word per rounding event
        a.run_mask = 0; a.run_value = 0;
        e = partial ? launch_fold_stochastic<true>(mode, ma, mb, C, a, sms, stream) : launch_fold_stochastic<false>(mode, ma, mb, C, a, sms, stream);
        ++*launches;
        if (e != cudaSuccess) return fail(e);
        if (dev_status) {
            tc_or_status_kernel<<<1, 1, 0, stream>>>(dev_status, status);
            ++*launches;
        }
        return cudaFreeAsync(ws, stream);
    }
    const bool pair_ok = (flags & LOF_GEMM_NO_PAIR) == 0 && sms >= 2 && K > 0 &&
                         (hw16 || (one_chunk && mode == LOF_RoundToNearestEven));
    if (pair_ok) {
        PairArgs pa;
        memset(&pa, 0, sizeof pa);
        pa.t = a;
        pa.t.idesc = make_idesc(op, hw16, 256, P_BN);
        pa.tiles_m2 = (int)((M + 255) / 256);
        pa.tiles_n2 = (int)((N + P_BN - 1) / P_BN);
        pa.n_iters = (int)((K * esize + kRowBytes - 1) / kRowBytes);
        const bool f8 = op != kOpBF16;
        if (hw16) e = launch_pair<kEpiHw16, true>(ma, mb, C, pa, sms, stream);
        else if (acc_fmt) e = f8 ? launch_pair<kEpiFmt, true>(ma, mb, C, pa, sms, stream) : launch_pair<kEpiFmt, false>(ma, mb, C, pa, sms, stream);
        else e = f8 ? launch_pair<kEpiMant, true>(ma, mb, C, pa, sms, stream) : launch_pair<kEpiMant, false>(ma, mb, C, pa, sms, stream);
        ++*launches;
        if (e != cudaSuccess) return fail(e);
        if (hw16) {         // left the fp16 range somewhere: the generic fold recomputes C (device-side conditional launch; the persistent
                            // kernel makes the untaken gate one wave of 148 CTAs instead of a 4096-CTA launch: ~15 us of a 0.41 ms call)
            a.run_mask = LOF_GEMM_STATUS_PACKED_RANGE; a.run_value = LOF_GEMM_STATUS_PACKED_RANGE;
            if ((flags & LOF_GEMM_NO_PERSISTENT) == 0)
                e = launch_fold_generic(mode, acc_fmt != nullptr, partial, ma, mb, C, a, sms, stream,
                                        op == kOpE4M3 && (flags & LOF_GEMM_NO_FLOAT_FOLD) == 0);
            else
                e = launch_generic(mode, acc_fmt != nullptr, partial, ma, mb, C, a, stream);
            ++*launches;
            if (e != cudaSuccess) return fail(e);
        }
        if (dev_status) {
            tc_or_status_kernel<<<1, 1, 0, stream>>>(dev_status, status);
            ++*launches;
        }
        return cudaFreeAsync(ws, stream);
    }
    // third-generation persistent fold kernel for everything that needs the per-chunk software fold
    if ((flags & LOF_GEMM_NO_PERSISTENT) == 0 && !hw16 && sms >= 1 && K > 0) {
        if (packed >= 0) {
            a.run_mask = 0; a.run_value = 0;
            const bool split2 = (flags & LOF_GEMM_INTERNAL_FOLD_SPLIT2) != 0;
            TcArgs ap = a;
            if (!split2) ap.idesc = make_idesc(op, false, F_BM, 256);       // one N = 256 MMA per k-step (else two of N = 128: a.idesc as it is)
            e = launch_fold_packed(packed, split2, ma, mb, C, ap, sms, stream);
            ++*launches;
            if (e != cudaSuccess) return fail(e);
            a.run_mask = LOF_GEMM_STATUS_PACKED_RANGE; a.run_value = LOF_GEMM_STATUS_PACKED_RANGE;   // left the fp16 range: generic recompute
        } else {
            a.run_mask = 0; a.run_value = 0;
        }
        if (packed != kEpiBf162) {          // the bfloat162 fold has no range condition (packed_epi): nothing to gate
            e = launch_fold_generic(mode, acc_fmt != nullptr, partial, ma, mb, C, a, sms, stream,
                                    op == kOpE4M3 && (flags & LOF_GEMM_NO_FLOAT_FOLD) == 0);
            ++*launches;
            if (e != cudaSuccess) return fail(e);
        }
        if (dev_status) {
            tc_or_status_kernel<<<1, 1, 0, stream>>>(dev_status, status);
            ++*launches;
        }
        return cudaFreeAsync(ws, stream);
    }
    if (hw16) {
        // fp16 accumulation inside the tensor core: one "chunk" as far as the kernel's pipeline is concerned
        TcArgs h = a;
        h.idesc = make_idesc(op, true);
        h.steps_per_chunk = h.k_steps;
        h.run_mask = 0; h.run_value = 0;
        e = launch_one<LOF_RoundToNearestEven, kEpiHw16, false>(ma, mb, C, h, stream);
        ++*launches;
        if (e != cudaSuccess) return fail(e);
        a.run_mask = LOF_GEMM_STATUS_PACKED_RANGE; a.run_value = LOF_GEMM_STATUS_PACKED_RANGE;
    } else if (packed >= 0) {
        // packed epilogue; if it reports LOF_GEMM_STATUS_PACKED_RANGE the generic kernel below recomputes C
        a.run_mask = 0; a.run_value = 0;
        e = (packed == kEpiHalf2) ? launch_one<LOF_RoundToNearestEven, kEpiHalf2, true>(ma, mb, C, a, stream)
                                  : launch_one<LOF_RoundToNearestEven, kEpiBf162, true>(ma, mb, C, a, stream);
        ++*launches;
        if (e != cudaSuccess) return fail(e);
        a.run_mask = LOF_GEMM_STATUS_PACKED_RANGE; a.run_value = LOF_GEMM_STATUS_PACKED_RANGE;
    } else {
        a.run_mask = 0; a.run_value = 0;
    }
    e = launch_generic(mode, acc_fmt != nullptr, partial, ma, mb, C, a, stream);
    ++*launches;
    if (e != cudaSuccess) return fail(e);
    if (dev_status) {
        tc_or_status_kernel<<<1, 1, 0, stream>>>(dev_status, status);
        ++*launches;
    }
    e = cudaFreeAsync(ws, stream);
    return e;
}

}  // namespace lof
