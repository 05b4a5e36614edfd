// lof_gemm_tc3.cuh -- third-generation chunked GEMM: a PERSISTENT kernel for the per-chunk software fold
//     acc = round(acc + P_c)   /   acc = round(acc + round(P_c))
// i.e. every accumulator format / rounding mode the tensor core cannot keep by itself (bf16-like accumulators, non-RNE modes,
// partial rounding, custom formats).  Included by lof_gemm_tc.cu inside its anonymous namespace.
//
// Why a new kernel: the first-generation gemm_tc_kernel ran at ~710 clk per 128 x 128 x 32 chunk with the packed fold
// (416 TFLOP/s at 8192^3) and DESIGN.md blamed the TMEM read-back.  scripts/probe_tmem_ld.cu (profiles/r02a_probe_tmem_ld.txt)
// measures tcgen05.ld at 700-900 B/clk/SM with 8-16 warps -- a 64 KB partial tile is ~80 clk, about the time of the FP8 MMA
// that produced it -- so the read-back was an order of magnitude away from its bound: the old kernel's per-chunk
// hand-shake (4 TMEM stages of one 128-column chunk each, a full barrier round trip per chunk and per warp, 64-register
// staging per prefetch) was latency-bound.  Here:
//   * one CTA per SM walks a static tile schedule; the smem ring, the TMEM ring and their phases run on across tiles, so the
//     TMA producer and the MMA issuer are always ahead of the fold (they start the next tile while C is being written);
//   * packed accumulators (half2 / bfloat162): tile 128 x 256, thread = 1 row x 128 columns = 64 registers of accumulator;
//     fp32 accumulators (generic fold): tile 128 x 128, 64 registers;
//   * TMEM is a ring of FOUR 128-column stages in both cases.  A 256-wide tile issues TWO MMAs of N = 128 per k-step (column
//     halves j = 0, 1 into stages 2i + j) and each half is drained by the four warps that own those columns.  Why not one
//     N = 256 MMA into two 256-column stages: scripts/probe_fold_pipeline.cu measures the issue -> commit-visible LATENCY of one
//     128 x 256 x 32 MMA at ~420 clk (throughput 128 clk) and a barrier hop at ~60 clk, so the tensor core has to run ~550 clk
//     ahead of the fold; two stages allow one chunk (~280 clk of fold) of look-ahead and the loop ran at RTT / 1.25 = ~710 clk per
//     chunk with the fold and the MMA latency in series (profiles/r02d_probe_fold_pipeline.txt, ncu r02c: 34 % of the fold
//     warps' time on the next chunk's full barrier, tensor pipe 15 %).  Four half-width stages give 1.5 chunks of look-ahead;
//   * the partial is pulled out in 32-column units through two 32-register buffers, the tcgen05.ld of unit u + 1 in flight
//     while unit u is folded; a TMEM stage is handed back as soon as its last unit has landed, BEFORE that unit is folded,
//     and the full-barrier of the next chunk is only polled (try_wait) ahead of the fold -- the blocking wait comes after it.
// Per-chunk cost model (per SM, packed fold, 128 x 256 x 32 chunk): MMA 128 clk (FP8) / 256 clk (bf16 operands), read-back
// 128 KB / ~800 B/clk = 160 clk, fold 8192 cvt.rn.x2 + 8192 HADD2 per SMSP-pair of warps = ~280 clk (both half-rate, on
// different pipes: scripts/probe_cvt.cu).  The fold's conversion / add instructions bound the kernel.

constexpr int F_BM = 128;
constexpr int F_THREADS = 384;                  // warp 0 TMA, warp 1 MMA, warp 2 TMEM alloc, warp 3 idle, warps 4..11 fold
constexpr int F_EPI_WARPS = 8;
constexpr int F_SMEM_BUDGET = 192 * 1024;       // operand ring

// BN: columns of the CTA tile (accumulators in registers).  SPLIT: MMAs per k-step -- each covers SW = BN / SPLIT columns and
// writes its own TMEM stage, so TMEM is a ring of 512 / SW stages:
//   <128, 1>  fp32 accumulators:   4 stages x 128 columns, one N = 128 MMA per k-step, a stage is drained by all 8 fold warps
//   <256, 1>  packed accumulators: 2 stages x 256 columns, one N = 256 MMA per k-step (fewest instructions on the issuing thread)
//   <256, 2>  packed accumulators: 4 stages x 128 columns, two N = 128 MMAs per k-step, a stage is drained by the 4 warps of its
//             column half (1.5 instead of 1 chunk of look-ahead for the tensor core, twice the issue work)
template <int BN, int SPLIT> struct FoldGeom {
    static constexpr int kStageA = F_BM * kRowBytes;                 // 16 KB
    static constexpr int kStageB = BN * kRowBytes;                   // 16 / 32 KB
    static constexpr int kStages = F_SMEM_BUDGET / (kStageA + kStageB);   // 6 (BN = 128) / 4 (BN = 256)
    static constexpr int kSmem = kStages * (kStageA + kStageB) + 1024;
    static constexpr int kSplit = SPLIT;
    static constexpr int kSW = BN / SPLIT;                           // columns per TMEM stage = N of one MMA
    static constexpr int kTmemStages = 512 / kSW;
    static constexpr int kColsPerThread = BN / 2;                    // two column halves x four lane quarters = 8 warps
    static constexpr int kWarpsPerStage = F_EPI_WARPS / SPLIT;       // fold warps that drain one TMEM stage
};

struct FoldArgs {
    TcArgs t;
    int n_tiles;
};

__device__ __forceinline__ void fold_tile_coords(int tile, const TcArgs& a, int& tm, int& tn)
{
    constexpr int GM = 16;      // 16 tile-rows (2048 rows of A) are walked column by column: the resident CTAs share panels in L2
    const int per_group = GM * a.tiles_n;
    const int group = tile / per_group;
    const int first_m = group * GM;
    const int gsz = min(a.tiles_m - first_m, GM);
    tm = first_m + (tile % per_group) % gsz;
    tn = (tile % per_group) / gsz;
}

template <int MODE, int EPI, bool PARTIAL, int BN, int SPLIT>
__global__ void __launch_bounds__(F_THREADS, 1)
gemm_tc_fold_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, float* __restrict__ C,
                    const FoldArgs fa)
{
    using G = FoldGeom<BN, SPLIT>;
    const TcArgs& args = fa.t;
    if ((*args.status & args.run_mask) != args.run_value) return;      // device-side conditional launch (uniform)

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + G::kStages * G::kStageA;
    __shared__ __align__(8) uint64_t full_bar[G::kStages], empty_bar[G::kStages];
    __shared__ __align__(8) uint64_t tmem_full_bar[G::kTmemStages], tmem_empty_bar[G::kTmemStages];
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_chunks = (args.k_steps + args.steps_per_chunk - 1) / args.steps_per_chunk;
    const int n_stage_iters = (args.k_steps + kStepsPerStage - 1) / kStepsPerStage;

    if (threadIdx.x == 0) {
        prefetch_tensormap(&map_a);
        prefetch_tensormap(&map_b);
#pragma unroll
        for (int s = 0; s < G::kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
#pragma unroll
        for (int s = 0; s < G::kTmemStages; ++s) { mbar_init(&tmem_full_bar[s], 1); mbar_init(&tmem_empty_bar[s], G::kWarpsPerStage); }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc(&tmem_base_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 0) {
        // ===================================== TMA producer =====================================
        if (elect_one()) {
            uint32_t stage = 0, phase = 0;
            // shared-window addresses as opaque values (rebuilt at every use otherwise: S2UR SR_CgaCtaId inside the loops, ncu r02p)
            const uint32_t full_a = smem_addr_once(&full_bar[0]), empty_a = smem_addr_once(&empty_bar[0]);
            const uint32_t sa_a = smem_addr_once(smem_a), sb_a = smem_addr_once(smem_b);
            for (int tile = blockIdx.x; tile < fa.n_tiles; tile += gridDim.x) {
                int tm, tn;
                fold_tile_coords(tile, args, tm, tn);
                const int m0 = tm * F_BM, n0 = tn * BN;
                for (int it = 0; it < n_stage_iters; ++it) {
                    mbar_wait_a(empty_a + stage * 8u, phase ^ 1u);
                    mbar_expect_tx_a(full_a + stage * 8u, G::kStageA + G::kStageB);
                    tma_load_2d_a(sa_a + stage * G::kStageA, &map_a, it * kRowBytes, m0, full_a + stage * 8u);   // coords in BYTES of K (uint8 map)
#pragma unroll
                    for (int h = 0; h < BN / 128; ++h)                                                       // the operand maps' box is 128 rows
                        tma_load_2d_a(sb_a + stage * G::kStageB + h * (128 * kRowBytes), &map_b, it * kRowBytes, n0 + h * 128, full_a + stage * 8u);
                    if (++stage == (uint32_t)G::kStages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ====================================== MMA issuer ======================================
        if (elect_one()) {
            // This thread's instruction stream is the tensor core's supply line (every instruction of a single thread costs several
            // clocks): descriptors are kept as running low words (the high word of umma_desc_k128 is a constant, and adding
            // bytes >> 4 to the low word never carries out of its 14-bit address field), the k-steps of a shared-memory stage
            // are unrolled, and the chunk bookkeeping is one counter.
            uint32_t stage = 0, phase = 0, ts = 0, tphase = 0;      // ts: first TMEM stage of the current chunk
            const uint64_t desc0 = umma_desc_k128(0);
            const uint32_t desc_hi = (uint32_t)(desc0 >> 32), desc_lo0 = (uint32_t)desc0;
            const uint32_t a_lo_base = desc_lo0 + (smem_addr_once(smem_a) >> 4), b_lo_base = desc_lo0 + (smem_addr_once(smem_b) >> 4);
            const uint32_t full_a = smem_addr_once(&full_bar[0]), empty_a = smem_addr_once(&empty_bar[0]);
            const uint32_t tfull_a = smem_addr_once(&tmem_full_bar[0]), tempty_a = smem_addr_once(&tmem_empty_bar[0]);
            const uint32_t idesc = args.idesc;
            const bool f8 = args.kind_f8 != 0;
            const int spc = args.steps_per_chunk;
            auto mk = [&](uint32_t lo) { return ((uint64_t)desc_hi << 32) | lo; };
            for (int tile = blockIdx.x; tile < fa.n_tiles; tile += gridDim.x) {
                int k_left = args.k_steps;              // k-steps of this tile still to issue
                int in_chunk = 0;                       // k-steps of the current chunk still to issue (0: the next MMA opens a chunk)
                for (int it = 0; it < n_stage_iters; ++it) {
                    mbar_wait_a(full_a + stage * 8u, phase);
                    tc_fence_after();
                    const uint32_t a_lo = a_lo_base + stage * (G::kStageA >> 4), b_lo = b_lo_base + stage * (G::kStageB >> 4);
#pragma unroll
                    for (int sub = 0; sub < kStepsPerStage; ++sub) {
                        if (k_left > 0) {
                            const bool first = (in_chunk == 0);
                            if (first) in_chunk = min(spc, k_left);
                            const bool last = (in_chunk == 1);
#pragma unroll
                            for (int hh = 0; hh < SPLIT; ++hh) {                  // one MMA of N = kSW per TMEM stage of the chunk
                                if (first) {                                     // first MMA into this TMEM stage: it must have been drained
                                    mbar_wait_a(tempty_a + (ts + hh) * 8u, tphase ^ 1u);
                                    tc_fence_after();
                                }
                                const uint64_t adesc = mk(a_lo + sub * (kStepBytes >> 4));
                                const uint64_t bdesc = mk(b_lo + hh * ((G::kSW * kRowBytes) >> 4) + sub * (kStepBytes >> 4));
                                const uint32_t tmem_d = tmem_base + (ts + hh) * G::kSW;
                                if (f8) umma_t<true>(tmem_d, adesc, bdesc, idesc, first ? 0u : 1u);        // uniform branch (one elected thread)
                                else umma_t<false>(tmem_d, adesc, bdesc, idesc, first ? 0u : 1u);
                                if (last) umma_commit_a(tfull_a + (ts + hh) * 8u);                       // this stage of the chunk is ready
                            }
                            --k_left;
                            if (--in_chunk == 0) {
                                ts += SPLIT;
                                if (ts == (uint32_t)G::kTmemStages) { ts = 0; tphase ^= 1u; }
                            }
                        }
                    }
                    umma_commit_a(empty_a + stage * 8u);                                                  // smem stage consumed
                    if (++stage == (uint32_t)G::kStages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp >= 4) {
        // ============================ accumulator owners (rounding fold) ============================
        constexpr bool kPacked = (EPI == kEpiHalf2 || EPI == kEpiBf162);
        constexpr int CPT = G::kColsPerThread;          // 128 (packed) / 64
        constexpr int UNITS = CPT / 32;                 // 32-column units per chunk and thread
        const int q = warp & 3;                         // TMEM lane quarter this warp may access
        const int h = (warp - 4) >> 2;                  // column half
        const uint32_t t_lane = (uint32_t)(q * 32) << 16;
        const bool vec_ok = ((args.N & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15u) == 0);

        float acc[kPacked ? 1 : CPT];
        uint32_t accp[kPacked ? CPT / 2 : 1];
        uint32_t buf0[32], buf1[32];
        // TMEM stage of this warp's columns of the current chunk: SPLIT = 2 -> stage cs + h holds exactly this warp's 128 columns;
        // SPLIT = 1 -> stage cs holds all of the tile's columns, this warp's start at column h * CPT
        uint32_t cs = 0, tphase = 0;
        constexpr uint32_t kStageOfHalf = (G::kSplit == 2) ? 1u : 0u, kColOfHalf = (G::kSplit == 2) ? 0u : (uint32_t)CPT;
        constexpr uint32_t kSW = (uint32_t)G::kSW;
        bool bad = false;
        int chunk_now = 0, row_now = 0, col_now = 0;    // position of the fold in the event stream (stochastic accumulator rounding)

        auto issue_unit = [&](int u, uint32_t* buf) {   // start loading unit u of this warp's columns of the chunk in stage cs (+ h)
            tmem_ld32(tmem_base + t_lane + (cs + kStageOfHalf * h) * kSW + kColOfHalf * h + (uint32_t)(u * 32), buf);
        };
        const uint32_t tfull_a = smem_addr_once(&tmem_full_bar[0]), tempty_a = smem_addr_once(&tmem_empty_bar[0]);
        auto full_bar_of = [&]() { return tfull_a + (cs + kStageOfHalf * h) * 8u; };
        auto fold_unit = [&](int u, const uint32_t* p) {
            if constexpr (EPI == kEpiHalf2) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const __half2 a0 = *reinterpret_cast<const __half2*>(&accp[u * 16 + i]);
                    const __half2 r0 = __hadd2(a0, __floats2half2_rn(__uint_as_float(p[2 * i]), __uint_as_float(p[2 * i + 1])));
                    accp[u * 16 + i] = *reinterpret_cast<const uint32_t*>(&r0);
                }
            } else if constexpr (EPI == kEpiBf162) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const __nv_bfloat162 a0 = *reinterpret_cast<const __nv_bfloat162*>(&accp[u * 16 + i]);
                    const __nv_bfloat162 r0 = __hadd2(a0, __floats2bfloat162_rn(__uint_as_float(p[2 * i]), __uint_as_float(p[2 * i + 1])));
                    accp[u * 16 + i] = *reinterpret_cast<const uint32_t*>(&r0);
                }
            } else if constexpr (ModeTraits<MODE>::kUsesRng) {
                // stochastic accumulator rounding: one dense Philox word per event, index = ((which * n_chunks + chunk) * M + row) * N + col;
                // four consecutive columns share a block when N % 4 == 0 (col0 and the unit offsets are multiples of 4)
                const uint64_t e_acc = (((uint64_t)args.n_chunks + (uint64_t)chunk_now) * (uint64_t)args.M + (uint64_t)row_now) * (uint64_t)args.N + (uint64_t)(col_now + u * 32);
                const uint64_t e_part = ((uint64_t)chunk_now * (uint64_t)args.M + (uint64_t)row_now) * (uint64_t)args.N + (uint64_t)(col_now + u * 32);
                const bool grp = (args.N & 3) == 0;
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    uint32_t wa[4], wp[4] = {0u, 0u, 0u, 0u};
                    if (grp) {
                        const uint4 b = philox_block(args.rng, (e_acc + i) >> 2);
                        wa[0] = b.x; wa[1] = b.y; wa[2] = b.z; wa[3] = b.w;
                        if constexpr (PARTIAL) { const uint4 bp = philox_block(args.rng, (e_part + i) >> 2); wp[0] = bp.x; wp[1] = bp.y; wp[2] = bp.z; wp[3] = bp.w; }
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            wa[e] = philox_word(args.rng, e_acc + i + e);
                            if constexpr (PARTIAL) wp[e] = philox_word(args.rng, e_part + i + e);
                        }
                    }
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        acc[u * 32 + i + e] = fold<MODE, EPI, PARTIAL>(acc[u * 32 + i + e], __uint_as_float(p[i + e]), args, wp[e], wa[e]);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    acc[u * 32 + i] = fold<MODE, EPI, PARTIAL>(acc[u * 32 + i], __uint_as_float(p[i]), args);
            }
        };
        auto pin = [&](uint32_t* buf) {                 // order every later use of the buffer after the tcgen05.wait::ld
#pragma unroll
            for (int i = 0; i < 32; ++i) asm volatile("" : "+r"(buf[i]));
        };

        for (int tile = blockIdx.x; tile < fa.n_tiles; tile += gridDim.x) {
            int tm, tn;
            fold_tile_coords(tile, args, tm, tn);
            const int row = tm * F_BM + q * 32 + lane;
            const int col0 = tn * BN + h * CPT;
            row_now = row; col_now = col0;
            if constexpr (kPacked) {
#pragma unroll
                for (int i = 0; i < CPT / 2; ++i) accp[i] = 0u;
            } else {
#pragma unroll
                for (int i = 0; i < CPT; ++i) acc[i] = 0.0f;
            }

            // first unit of the tile
            mbar_wait_a(full_bar_of(), tphase);
            tc_fence_after();
            issue_unit(0, buf0);
            for (int c = 0; c < n_chunks; ++c) {
                chunk_now = c;
#pragma unroll
                for (int u = 0; u < UNITS; ++u) {
                    uint32_t* cur = (u & 1) ? buf1 : buf0;
                    uint32_t* nxt = (u & 1) ? buf0 : buf1;
                    tmem_ld_wait();                                     // unit (c, u) is in registers
                    pin(cur);
                    bool issued = true;
                    if (u + 1 < UNITS) {
                        issue_unit(u + 1, nxt);                         // same chunk: no barrier
                    } else {
                        // last unit of chunk c has landed: hand its TMEM stage back to the MMA issuer before folding it
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_a(tempty_a + (cs + kStageOfHalf * h) * 8u);
                        cs += G::kSplit;
                        if (cs == (uint32_t)G::kTmemStages) { cs = 0; tphase ^= 1u; }
                        if (c + 1 < n_chunks) {
                            issued = mbar_try_wait_a(full_bar_of(), tphase);          // poll only: fold first if the next chunk is not ready
                            if (issued) { tc_fence_after(); issue_unit(0, nxt); }
                        }
                    }
                    fold_unit(u, cur);
                    if (!issued) {                                      // (UNITS is even: nxt == buf0 here, as the next chunk expects)
                        mbar_wait_a(full_bar_of(), tphase);
                        tc_fence_after();
                        issue_unit(0, nxt);
                    }
                }
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
#pragma unroll
                for (int i = 0; i < CPT / 2; ++i)                 // inf / NaN: an accumulator left the finite fp16 range
                    bad |= ((accp[i] & 0x7C00u) == 0x7C00u) || ((accp[i] & 0x7C000000u) == 0x7C000000u);
            }
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
        }
        if constexpr (EPI == kEpiHalf2) {
            if (bad) atomicOr(args.status, LOF_GEMM_STATUS_PACKED_RANGE);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

template <int MODE, int EPI, bool PARTIAL, int BN, int SPLIT>
cudaError_t launch_fold_one(const CUtensorMap& ma, const CUtensorMap& mb, float* C, const TcArgs& a, int num_sms, cudaStream_t stream)
{
    using G = FoldGeom<BN, SPLIT>;
    auto kern = gemm_tc_fold_kernel<MODE, EPI, PARTIAL, BN, SPLIT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G::kSmem);
    if (e != cudaSuccess) return e;
    FoldArgs fa;
    fa.t = a;
    fa.t.tiles_m = (a.M + F_BM - 1) / F_BM;
    fa.t.tiles_n = (a.N + BN - 1) / BN;
    fa.n_tiles = fa.t.tiles_m * fa.t.tiles_n;
    const int grid = fa.n_tiles < num_sms ? fa.n_tiles : num_sms;
    kern<<<(unsigned)(grid < 1 ? 1 : grid), F_THREADS, G::kSmem, stream>>>(ma, mb, C, fa);
    return cudaGetLastError();
}
