// lof_gemm_mx.cuh -- kernel template of the block-scaled (MX) chunked tensor-core GEMM; see lof_gemm_mx.cu for the description.
// Included by lof_gemm_mx.cu (packed bfloat162 fold, host side) and lof_gemm_mx_gen.cu (the generic per-mode folds), so that the
// two sets of instantiations compile in parallel.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>

#include <algorithm>
#include <cstring>
#include <string>

#include "lof_core.cuh"
#include "lof_fast_f32.cuh"
#include "lof_gemm_common.cuh"
#include "lof_ptx.cuh"

namespace lof {

namespace {

using namespace ptx;

constexpr int X_BM = 128;
constexpr int X_BN = 224;                       // 2 x 224 accumulator columns + 2 x 32 scale-factor columns = 512 = all of TMEM
constexpr int X_ROW_BYTES = 128;                // bytes of K per shared-memory row (one 128B swizzle atom row) = 4 MX blocks
constexpr int X_STEP_BYTES = 32;                // bytes of K per tcgen05.mma = one MX block
constexpr int X_STEPS_PER_STAGE = X_ROW_BYTES / X_STEP_BYTES;
constexpr int X_STAGE_A = X_BM * X_ROW_BYTES;   // 16 KB
constexpr int X_STAGE_B = X_BN * X_ROW_BYTES;   // 28 KB
constexpr int X_SFA_BYTES = 512;                // 128 rows x 4 blocks
constexpr int X_SFB_BYTES = 1024;               // 2 chunks of 128 rows (224 used) x 4 blocks
constexpr int X_STAGES = 4;
constexpr int X_SMEM = X_STAGES * (X_STAGE_A + X_STAGE_B + X_SFA_BYTES + X_SFB_BYTES) + 1024;
constexpr int X_ACC_STAGES = 2;
constexpr int X_ACC_STRIDE = 256;               // TMEM column of accumulator stage i = 256 i; the 32 columns behind each stage hold one
constexpr int X_SF_COL0 = X_BN;                 // scale-factor slot (SFA at +0, 4 columns; SFB at +4, 8 columns)
constexpr int X_SF_STRIDE = X_ACC_STRIDE;

enum : int { kXBf162 = 0, kXMant = 1, kXFmt = 2 };

struct MxGemmArgs {
    int M, N, K;
    int k_steps;            // K / 32
    int steps_per_chunk;    // k_chunk / 32
    uint32_t idesc;
    float alpha;
    int r;                  // mantissa-only accumulator: fp32 mantissa bits to drop
    RoundPlan plan;         // full-format accumulator
    FastPlan fast;
    int tiles_m, tiles_n, n_tiles;
    int kb4;                // 128-byte K groups = ceil(K / 128)
    const uint8_t* sfa;     // [tiles_m][kb4][512]   pre-arranged scale bytes of A
    const uint8_t* sfb;     // [tiles_n][kb4][1024]  ... of B^T
};

__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s_a(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint32_t bar_addr)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(bar_addr) : "memory");
}
// shared memory -> TMEM: 32 rows x 16 bytes, replicated into the four 32-lane quarters (the scale-factor layout of block-scaled MMAs)
__device__ __forceinline__ void tmem_cp_32x128b_warpx4(uint32_t taddr, uint64_t sdesc)
{
    asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
// no-swizzle K-major descriptor of a [32 rows x 16 bytes] block stored contiguously: 8-row core matrices 128 bytes apart (SBO)
__device__ __forceinline__ uint64_t sf_desc(uint32_t smem_addr)
{
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void umma_mx(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate,
                                        uint32_t tmem_sfa, uint32_t tmem_sfb)
{
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale [%0], %1, %2, %3, [%5], [%6], p;\n}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(tmem_sfa), "r"(tmem_sfb) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* v)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v)
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}
// W columns (a sum of 64 / 32 / 16 / 8) of this warp's 32 lanes -> v[0 .. W): all loads in flight at once, ONE wait by the caller
template <int W>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t* v)
{
    if constexpr (W >= 32) { tmem_ld32(taddr, v); tmem_ld_cols<W - 32>(taddr + 32, v + 32); }
    else if constexpr (W >= 16) { tmem_ld16(taddr, v); tmem_ld_cols<W - 16>(taddr + 16, v + 16); }
    else if constexpr (W >= 8) { tmem_ld8(taddr, v); tmem_ld_cols<W - 8>(taddr + 8, v + 8); }
    else static_assert(W == 0, "column counts are multiples of 8");
}

template <int MODE, int EPI>
__device__ __forceinline__ float x_round_acc(float t, const MxGemmArgs& a)
{
    if constexpr (EPI == kXMant) {
        return __uint_as_float(round_mantissa_elem<LOF_F32, MODE>(__float_as_uint(t), a.r, 0, 0u));
    } else {
        uint32_t rb;
        if (a.fast.valid) {
            bool rare = false;
            rb = round_fast_f32_rt<MODE>(__float_as_uint(t), a.fast, a.plan.saturating != 0, 0u, rare);
            if (rare) rb = exact_params_f32<MODE>(__float_as_uint(t), a.plan, 0u);
        } else {
            rb = exact_params_f32<MODE>(__float_as_uint(t), a.plan, 0u);
        }
        return __uint_as_float(rb);
    }
}
template <int MODE, int EPI, bool PARTIAL>
__device__ __forceinline__ float x_fold(float acc, float p, const MxGemmArgs& a)
{
    if constexpr (PARTIAL) p = x_round_acc<MODE, EPI>(p, a);
    return x_round_acc<MODE, EPI>(rto_add(acc, p), a);            // round-to-odd add: the format rounding is the only rounding
}

__device__ __forceinline__ void x_tile_coords(int tile, const MxGemmArgs& a, int& tm, int& tn)
{
    constexpr int GM = 16;      // 16 tile-rows are walked column by column: the resident CTAs share operand panels in L2
    const int per_group = GM * a.tiles_n;
    const int group = tile / per_group;
    const int first_m = group * GM;
    const int gsz = min(a.tiles_m - first_m, GM);
    tm = first_m + (tile % per_group) % gsz;
    tn = (tile % per_group) / gsz;
}

// FW fold warps (8 or 16) = warps 0 .. FW-1 (warp % 4 = the TMEM lane quarter it may read); then one warp each for TMA, MMA issue, TMEM allocation.
// A fold warp owns 32 rows x CPT = 224 / (FW / 4) columns (112 or 56) of the tile.
// cycle breakdown of CTA 0 (option gemm.mx_prof): [0..3] MMA thread: total, wait smem full, wait tmem empty, chunks; [8..13] fold warp 0:
// total, wait tmem full, read-back (ld + wait), fold arithmetic, C store, chunks
__device__ unsigned long long g_mx_prof[16];

template <int MODE, int EPI, bool PARTIAL, int FW, bool PROF = false>
__global__ void __launch_bounds__((FW + 3) * 32, 1)
gemm_mx_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, float* __restrict__ C,
               const MxGemmArgs args)
{
    constexpr int CPT = X_BN / (FW / 4);
    constexpr bool kPacked = (EPI == kXBf162);
    // read-back granularity: the packed fold pulls all of a chunk's columns at once (one wait, the TMEM stage goes back to the tensor core
    // immediately); the generic fold (fp32 accumulators: 2x the registers) goes through a 32- or 8-column buffer
    constexpr int UB = kPacked ? CPT : (CPT % 32 == 0 ? 32 : (CPT % 16 == 0 ? 16 : 8));
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + X_STAGES * X_STAGE_A;
    uint8_t* smem_sfa = smem_b + X_STAGES * X_STAGE_B;
    uint8_t* smem_sfb = smem_sfa + X_STAGES * X_SFA_BYTES;
    __shared__ __align__(8) uint64_t full_bar[X_STAGES], empty_bar[X_STAGES];
    __shared__ __align__(8) uint64_t tmem_full_bar[X_ACC_STAGES], tmem_empty_bar[X_ACC_STAGES];
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_chunks = (args.k_steps + args.steps_per_chunk - 1) / args.steps_per_chunk;
    const int n_stage_iters = args.kb4;

    if (threadIdx.x == 0) {
        prefetch_tensormap(&map_a);
        prefetch_tensormap(&map_b);
#pragma unroll
        for (int s = 0; s < X_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
#pragma unroll
        for (int s = 0; s < X_ACC_STAGES; ++s) { mbar_init(&tmem_full_bar[s], 1); mbar_init(&tmem_empty_bar[s], FW); }
        fence_mbar_init();
    }
    if (warp == FW + 2) tmem_alloc(&tmem_base_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == FW) {
        // ===================================== TMA producer =====================================
        if (elect_one()) {
            uint32_t stage = 0, phase = 0;
            const uint32_t empty_a = smem_addr_once(&empty_bar[0]), full_a = smem_addr_once(&full_bar[0]);
            const uint32_t sa_a = smem_addr_once(smem_a), sb_a = smem_addr_once(smem_b), sfa_a = smem_addr_once(smem_sfa), sfb_a = smem_addr_once(smem_sfb);
            for (int tile = blockIdx.x; tile < args.n_tiles; tile += gridDim.x) {
                int tm, tn;
                x_tile_coords(tile, args, tm, tn);
                const int m0 = tm * X_BM, n0 = tn * X_BN;
                const uint8_t* sfa = args.sfa + (size_t)tm * args.kb4 * X_SFA_BYTES;
                const uint8_t* sfb = args.sfb + (size_t)tn * args.kb4 * X_SFB_BYTES;
                for (int it = 0; it < n_stage_iters; ++it) {
                    mbar_wait_a(empty_a + stage * 8u, phase ^ 1u);
                    const uint32_t fb = full_a + stage * 8u;
                    mbar_expect_tx_a(fb, X_STAGE_A + X_STAGE_B + X_SFA_BYTES + X_SFB_BYTES);
                    tma_load_2d_a(sa_a + stage * X_STAGE_A, &map_a, it * X_ROW_BYTES, m0, fb);
                    tma_load_2d_a(sb_a + stage * X_STAGE_B, &map_b, it * X_ROW_BYTES, n0, fb);                        // box = 224 rows
                    bulk_g2s_a(sfa_a + stage * X_SFA_BYTES, sfa + (size_t)it * X_SFA_BYTES, X_SFA_BYTES, fb);
                    bulk_g2s_a(sfb_a + stage * X_SFB_BYTES, sfb + (size_t)it * X_SFB_BYTES, X_SFB_BYTES, fb);
                    if (++stage == (uint32_t)X_STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == FW + 1) {
        // ====================================== MMA issuer ======================================
        if (elect_one()) {
            uint32_t stage = 0, phase = 0, ts = 0, tphase = 0, sf_slot = 0;
            const uint64_t desc0 = umma_desc_k128(0);
            const uint32_t desc_hi = (uint32_t)(desc0 >> 32), desc_lo0 = (uint32_t)desc0;
            const uint32_t a_lo_base = desc_lo0 + (smem_addr_once(smem_a) >> 4), b_lo_base = desc_lo0 + (smem_addr_once(smem_b) >> 4);
            const uint32_t idesc = args.idesc;
            const int spc = args.steps_per_chunk;
            auto mk = [&](uint32_t lo) { return ((uint64_t)desc_hi << 32) | lo; };
            const uint32_t full_a = smem_addr_once(&full_bar[0]), empty_a = smem_addr_once(&empty_bar[0]);
            const uint32_t tfull_a = smem_addr_once(&tmem_full_bar[0]), tempty_a = smem_addr_once(&tmem_empty_bar[0]);
            const uint32_t sfa_a = smem_addr_once(smem_sfa), sfb_a = smem_addr_once(smem_sfb);
            long long p_t0 = 0, p_full = 0, p_empty = 0, p_chunks = 0, p_x = 0;
            if constexpr (PROF) p_t0 = clock64();
            for (int tile = blockIdx.x; tile < args.n_tiles; tile += gridDim.x) {
                int k_left = args.k_steps;
                int in_chunk = 0;
                for (int it = 0; it < n_stage_iters; ++it) {
                    if constexpr (PROF) p_x = clock64();
                    mbar_wait_a(full_a + stage * 8u, phase);
                    if constexpr (PROF) p_full += clock64() - p_x;
                    tc_fence_after();
                    // this stage's scale factors -> TMEM.  tcgen05.cp and tcgen05.mma execute in issue order, so a slot may be rewritten as
                    // soon as the MMAs that read it have been ISSUED; two slots keep a whole stage of MMAs between a copy and the
                    // previous readers of its slot anyway
                    const uint32_t sf_col = tmem_base + X_SF_COL0 + sf_slot * X_SF_STRIDE;
                    sf_slot ^= 1u;
                    tmem_cp_32x128b_warpx4(sf_col, sf_desc(sfa_a + stage * X_SFA_BYTES));
                    tmem_cp_32x128b_warpx4(sf_col + 4, sf_desc(sfb_a + stage * X_SFB_BYTES));
                    tmem_cp_32x128b_warpx4(sf_col + 8, sf_desc(sfb_a + stage * X_SFB_BYTES + 512));
                    const uint32_t a_lo = a_lo_base + stage * (X_STAGE_A >> 4), b_lo = b_lo_base + stage * (X_STAGE_B >> 4);
#pragma unroll
                    for (int sub = 0; sub < X_STEPS_PER_STAGE; ++sub) {
                        if (k_left > 0) {
                            const bool first = (in_chunk == 0);
                            if (first) {
                                in_chunk = min(spc, k_left);
                                if constexpr (PROF) { p_x = clock64(); ++p_chunks; }
                                mbar_wait_a(tempty_a + ts * 8u, tphase ^ 1u);       // the accumulator stage must have been drained
                                if constexpr (PROF) p_empty += clock64() - p_x;
                                tc_fence_after();
                            }
                            const bool last = (in_chunk == 1);
                            const uint32_t id = idesc | ((uint32_t)sub << 29) | ((uint32_t)sub << 4);      // a_sf_id / b_sf_id = block in stage
                            umma_mx(tmem_base + ts * X_ACC_STRIDE, mk(a_lo + sub * (X_STEP_BYTES >> 4)), mk(b_lo + sub * (X_STEP_BYTES >> 4)), id,
                                    first ? 0u : 1u, sf_col, sf_col + 4);
                            if (last) umma_commit_a(tfull_a + ts * 8u);
                            --k_left;
                            if (--in_chunk == 0) {
                                if (++ts == (uint32_t)X_ACC_STAGES) { ts = 0; tphase ^= 1u; }
                            }
                        }
                    }
                    umma_commit_a(empty_a + stage * 8u);                            // operands and scales of this stage consumed
                    if (++stage == (uint32_t)X_STAGES) { stage = 0; phase ^= 1u; }
                }
            }
            if constexpr (PROF) {
                if (blockIdx.x == 0) {
                    g_mx_prof[0] = (unsigned long long)(clock64() - p_t0); g_mx_prof[1] = (unsigned long long)p_full;
                    g_mx_prof[2] = (unsigned long long)p_empty; g_mx_prof[3] = (unsigned long long)p_chunks;
                }
            }
        }
    } else if (warp < FW) {
        // ============================ accumulator owners (rounding fold) ============================
        const int q = warp & 3;                         // TMEM lane quarter this warp may access
        const int h = warp >> 2;                        // column group
        const uint32_t t_lane = (uint32_t)(q * 32) << 16;
        const bool vec_ok = ((args.N & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15u) == 0);

        float acc[kPacked ? 1 : CPT];
        uint32_t accp[kPacked ? CPT / 2 : 1];
        uint32_t buf[UB];
        uint32_t cs = 0, tphase = 0;

        const uint32_t tfull_a = smem_addr_once(&tmem_full_bar[0]), tempty_a = smem_addr_once(&tmem_empty_bar[0]);
        auto release_stage = [&]() {                     // every load of this warp from stage cs has landed
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_a(tempty_a + cs * 8u);
            if (++cs == (uint32_t)X_ACC_STAGES) { cs = 0; tphase ^= 1u; }
        };

        long long p_t0 = 0, p_wait = 0, p_ld = 0, p_fold = 0, p_store = 0, p_chunks = 0, p_x = 0;
        if constexpr (PROF) p_t0 = clock64();
        for (int tile = blockIdx.x; tile < args.n_tiles; tile += gridDim.x) {
            int tm, tn;
            x_tile_coords(tile, args, tm, tn);
            const int row = tm * X_BM + q * 32 + lane;
            const int col0 = tn * X_BN + h * CPT;
            if constexpr (kPacked) {
#pragma unroll
                for (int i = 0; i < CPT / 2; ++i) accp[i] = 0u;
            } else {
#pragma unroll
                for (int i = 0; i < CPT; ++i) acc[i] = 0.0f;
            }

            for (int c = 0; c < n_chunks; ++c) {
                if constexpr (PROF) { p_x = clock64(); ++p_chunks; }
                mbar_wait_a(tfull_a + cs * 8u, tphase);
                if constexpr (PROF) { const long long t = clock64(); p_wait += t - p_x; p_x = t; }
                tc_fence_after();
                const uint32_t ta = tmem_base + t_lane + cs * X_ACC_STRIDE + (uint32_t)(h * CPT);
                if constexpr (kPacked) {
                    tmem_ld_cols<CPT>(ta, buf);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < CPT; ++i) asm volatile("" : "+r"(buf[i]));      // order every use after the wait
                    release_stage();
                    if constexpr (PROF) { const long long t = clock64(); p_ld += t - p_x; p_x = t; }
#pragma unroll
                    for (int i = 0; i < CPT / 2; ++i) {
                        const __nv_bfloat162 a0 = *reinterpret_cast<const __nv_bfloat162*>(&accp[i]);
                        const __nv_bfloat162 r0 = __hadd2(a0, __floats2bfloat162_rn(__uint_as_float(buf[2 * i]), __uint_as_float(buf[2 * i + 1])));
                        accp[i] = *reinterpret_cast<const uint32_t*>(&r0);
                    }
                    if constexpr (PROF) {
#pragma unroll
                        for (int i = 0; i < CPT / 2; ++i) asm volatile("" : "+r"(accp[i]));
                        p_fold += clock64() - p_x;
                    }
                } else {
#pragma unroll
                    for (int u = 0; u < CPT / UB; ++u) {
                        tmem_ld_cols<UB>(ta + u * UB, buf);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < UB; ++i) asm volatile("" : "+r"(buf[i]));
                        if (u + 1 == CPT / UB) release_stage();
#pragma unroll
                        for (int i = 0; i < UB; ++i) acc[u * UB + i] = x_fold<MODE, EPI, PARTIAL>(acc[u * UB + i], __uint_as_float(buf[i]), args);
                    }
                }
            }

            auto acc_at = [&](int i) -> float {
                if constexpr (kPacked) return (i & 1) ? __uint_as_float(accp[i >> 1] & 0xFFFF0000u) : __uint_as_float(accp[i >> 1] << 16);
                else return acc[i];
            };
            if constexpr (PROF) p_x = clock64();
            if (row < args.M) {
                float* crow = C + (int64_t)row * args.N + col0;
#pragma unroll
                for (int i = 0; i < CPT; i += 4) {
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
            if constexpr (PROF) p_store += clock64() - p_x;
        }
        if constexpr (PROF) {
            if (blockIdx.x == 0 && threadIdx.x == 0) {
                g_mx_prof[8] = (unsigned long long)(clock64() - p_t0); g_mx_prof[9] = (unsigned long long)p_wait; g_mx_prof[10] = (unsigned long long)p_ld;
                g_mx_prof[11] = (unsigned long long)p_fold; g_mx_prof[12] = (unsigned long long)p_store; g_mx_prof[13] = (unsigned long long)p_chunks;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == FW + 2) tmem_dealloc(tmem_base, 512);
}

template <int MODE, int EPI, bool PARTIAL, int FW = 8, bool PROF = false>
cudaError_t x_launch_one(const CUtensorMap& ma, const CUtensorMap& mb, float* C, const MxGemmArgs& a, int grid, cudaStream_t stream)
{
    auto kern = gemm_mx_kernel<MODE, EPI, PARTIAL, FW, PROF>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, X_SMEM);
    if (e != cudaSuccess) return e;
    kern<<<(unsigned)grid, (FW + 3) * 32, X_SMEM, stream>>>(ma, mb, C, a);
    return cudaGetLastError();
}
}  // namespace

// lof_gemm_mx_gen.cu: generic fold (fp32 accumulators, any deterministic mode, mantissa-only or full-format accumulator)
cudaError_t launch_gemm_mx_generic(int mode, bool fmt, bool partial, const CUtensorMap& ma, const CUtensorMap& mb, float* C, const void* args,
                                   int grid, cudaStream_t stream);

}  // namespace lof
