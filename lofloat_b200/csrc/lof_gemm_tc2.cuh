// lof_gemm_tc2.cuh -- second-generation tensor-core GEMM main loop: persistent CTA PAIRS (cta_group::2).
// Included by lof_gemm_tc.cu inside its anonymous namespace (shares TcArgs, round_acc and the operand maps).
//
// Why: ncu of the first kernel (profiles/r01e_gemm_*) shows a 128 x 128 tile per CTA pulling 8.6 GB through L2 for
// 8192^3 (tensor pipe 12 % busy) and paying the prologue / C write-back of every tile serially.  Here
//   * two CTAs of a cluster (one TPC) execute ONE tcgen05.mma of M = 256, N = 256: each CTA loads its own 128 rows
//     of A and HALF of the B tile (128 of the 256 N-rows), the tensor cores read both halves of B through the pair's
//     shared memory -> operand traffic per flop is half that of a 128 x 128 tile;
//   * the CTA pair is persistent and walks a static tile schedule; the accumulator is double-buffered in TMEM
//     (2 x 256 columns), so the epilogue of tile t (TMEM -> registers -> C) overlaps the main loop of tile t + 1;
//   * a 6-stage TMA ring (32 KB per stage and CTA).
//
// Accumulator semantics served by this kernel (everything else stays on gemm_tc_kernel):
//   kEpiHw16   fp16 accumulation inside the tensor core, re-rounded by the hardware after every K = 32 MMA
//              (== LoFloat E5M10 / RNE single-rounding fold with k_chunk = 32; see hw_f16_accum_ok)
//   kEpiMant / kEpiFmt with one chunk (k_chunk >= K): fp32 accumulation, one LoFloat rounding at the end (RNE).
//
// Roles per CTA (256 threads): warp 0 TMA producer, warp 1 MMA issuer (leader CTA only), warp 2 TMEM allocation,
// warps 4..7 epilogue (TMEM lane quarter = warp % 4).

constexpr int P_BN = 256;                       // N of the pair's tile; M = 2 x 128
constexpr int P_STAGES = 6;
constexpr int P_STAGE_A = 128 * kRowBytes;      // 16 KB: this CTA's 128 rows of A
constexpr int P_STAGE_B = 128 * kRowBytes;      // 16 KB: this CTA's half of the B tile
constexpr int P_SMEM = P_STAGES * (P_STAGE_A + P_STAGE_B) + 1024;
constexpr int P_THREADS = 256;
constexpr int P_ACC_STAGES = 2;                 // x 256 columns = all of TMEM

struct PairArgs {
    TcArgs t;
    int tiles_m2, tiles_n2;     // 256 x 256 tiles
    int n_iters;                // ceil(K bytes / 128): shared-memory stages per tile
};

__device__ __forceinline__ void pair_tile_coords(int tile, const PairArgs& a, int& tm, int& tn)
{
    constexpr int GM = 8;       // 8 tile-rows (2048 rows of A) are walked column by column: the resident pairs share panels in L2
    const int per_group = GM * a.tiles_n2;
    const int group = tile / per_group;
    const int first_m = group * GM;
    const int gsz = min(a.tiles_m2 - first_m, GM);
    tm = first_m + (tile % per_group) % gsz;
    tn = (tile % per_group) / gsz;
}

template <int EPI, bool F8>
__global__ void __launch_bounds__(P_THREADS, 1)
gemm_tc_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, float* __restrict__ C,
                    const PairArgs args)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + P_STAGES * P_STAGE_A;
    __shared__ __align__(8) uint64_t full_bar[P_STAGES], empty_bar[P_STAGES];
    __shared__ __align__(8) uint64_t acc_full_bar[P_ACC_STAGES], acc_empty_bar[P_ACC_STAGES];
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();            // 0 = leader
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int n_tiles = args.tiles_m2 * args.tiles_n2;

    if (threadIdx.x == 0) {
        prefetch_tensormap(&map_a);
        prefetch_tensormap(&map_b);
#pragma unroll
        for (int s = 0; s < P_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
#pragma unroll
        for (int s = 0; s < P_ACC_STAGES; ++s) { mbar_init(&acc_full_bar[s], 1); mbar_init(&acc_empty_bar[s], 8); }   // 4 epilogue warps x 2 CTAs
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc_pair(&tmem_base_slot, P_ACC_STAGES * P_BN);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 0) {
        // ===================================== TMA producer (both CTAs) =====================================
        if (elect_one()) {
            uint32_t stage = 0, phase = 0;
            // shared-window addresses as opaque values: rebuilt at every use otherwise (S2UR SR_CgaCtaId + address arithmetic inside the loops, ncu r02p)
            const uint32_t full_a = smem_addr_once(&full_bar[0]), empty_a = smem_addr_once(&empty_bar[0]);
            const uint32_t sa_a = smem_addr_once(smem_a), sb_a = smem_addr_once(smem_b);
            for (int tile = pair; tile < n_tiles; tile += n_pairs) {
                int tm, tn;
                pair_tile_coords(tile, args, tm, tn);
                const int row_a = tm * 256 + (int)rank * 128, row_b = tn * P_BN + (int)rank * 128;
                for (int it = 0; it < args.n_iters; ++it) {
                    mbar_wait_a(empty_a + stage * 8u, phase ^ 1u);                 // this CTA's slot is free (multicast commit)
                    if (rank == 0) mbar_expect_tx_a(full_a + stage * 8u, 2 * (P_STAGE_A + P_STAGE_B));   // both CTAs' bytes land on the leader's barrier
                    tma_load_2d_pair_a(sa_a + stage * P_STAGE_A, &map_a, it * kRowBytes, row_a, full_a + stage * 8u);
                    tma_load_2d_pair_a(sb_a + stage * P_STAGE_B, &map_b, it * kRowBytes, row_b, full_a + stage * 8u);
                    if (++stage == P_STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ====================================== MMA issuer (leader) ======================================
        if (rank == 0 && elect_one()) {
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
            const uint32_t a_base = smem_addr_once(smem_a), b_base = smem_addr_once(smem_b);
            const uint32_t full_a = smem_addr_once(&full_bar[0]), empty_a = smem_addr_once(&empty_bar[0]);
            const uint32_t accf_a = smem_addr_once(&acc_full_bar[0]), acce_a = smem_addr_once(&acc_empty_bar[0]);
            const uint32_t idesc = args.t.idesc;
            for (int tile = pair; tile < n_tiles; tile += n_pairs) {
                mbar_wait_a(acce_a + acc * 8u, acc_phase ^ 1u);                     // both CTAs drained this accumulator stage
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * P_BN;
                for (int it = 0; it < args.n_iters; ++it) {
                    mbar_wait_a(full_a + stage * 8u, phase);
                    tc_fence_after();
                    const uint32_t a_addr = a_base + stage * P_STAGE_A, b_addr = b_base + stage * P_STAGE_B;
#pragma unroll
                    for (int sub = 0; sub < kStepsPerStage; ++sub)
                        umma_pair<F8>(tmem_d, umma_desc_k128(a_addr + sub * kStepBytes), umma_desc_k128(b_addr + sub * kStepBytes),
                                      idesc, (it | sub) != 0 ? 1u : 0u);
                    umma_commit_pair_a(empty_a + stage * 8u);                       // frees the slot in BOTH CTAs
                    if (++stage == P_STAGES) { stage = 0; phase ^= 1u; }
                }
                umma_commit_pair_a(accf_a + acc * 8u);                              // tile done: wake both epilogues
                acc ^= 1u;
                if (acc == 0) acc_phase ^= 1u;
            }
        }
    } else if (warp >= 4) {
        // ============================================ epilogue ============================================
        const int q = warp & 3;
        const uint32_t t_lane = (uint32_t)(q * 32) << 16;
        uint32_t acc = 0, acc_phase = 0;
        bool bad = false;
        const bool vec_ok = ((args.t.N & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15u) == 0);
        const uint32_t accf_a = smem_addr_once(&acc_full_bar[0]), acce_a = smem_addr_once(&acc_empty_bar[0]);
        for (int tile = pair; tile < n_tiles; tile += n_pairs) {
            int tm, tn;
            pair_tile_coords(tile, args, tm, tn);
            const int row = tm * 256 + (int)rank * 128 + q * 32 + lane;
            const int col0 = tn * P_BN;
            mbar_wait_a(accf_a + acc * 8u, acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + t_lane + acc * P_BN;
            float* crow = C + (int64_t)row * args.t.N + col0;

            uint32_t buf0[32], buf1[32];
            auto emit = [&](const uint32_t* p, int g) {
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    float v[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        if constexpr (EPI == kEpiHw16) {
                            v[e] = __half2float(__ushort_as_half((unsigned short)(p[i + e] & 0xFFFFu)));
                            bad |= (p[i + e] & 0x7C00u) == 0x7C00u;                 // inf / NaN: the accumulator left the fp16 range
                        } else {
                            v[e] = round_acc<LOF_RoundToNearestEven, EPI>(__uint_as_float(p[i + e]), args.t);
                        }
                        v[e] = __fmul_rn(v[e], args.t.alpha);
                    }
                    const int c = col0 + g * 32 + i;
                    if (row < args.t.M) {
                        if (vec_ok && c + 3 < args.t.N) {
                            *reinterpret_cast<float4*>(crow + g * 32 + i) = make_float4(v[0], v[1], v[2], v[3]);
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                if (c + e < args.t.N) crow[g * 32 + i + e] = v[e];
                        }
                    }
                }
            };
            auto landed = [&](uint32_t* buf) {
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) asm volatile("" : "+r"(buf[i]));
            };
            tmem_ld32(taddr, buf0);
#pragma unroll 1
            for (int g = 0; g < P_BN / 32; g += 2) {
                landed(buf0);
                tmem_ld32(taddr + 32 * (g + 1), buf1);
                emit(buf0, g);
                landed(buf1);
                if (g + 2 < P_BN / 32) tmem_ld32(taddr + 32 * (g + 2), buf0);
                else {                                                              // everything is in registers: hand the stage back
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_leader_a(acce_a + acc * 8u);
                }
                emit(buf1, g + 1);
            }
            acc ^= 1u;
            if (acc == 0) acc_phase ^= 1u;
        }
        if constexpr (EPI == kEpiHw16) {
            if (bad) atomicOr(args.t.status, LOF_GEMM_STATUS_PACKED_RANGE);
        }
    }

    tc_fence_before();
    cluster_sync_all();                 // the peer may still signal barriers in this CTA's shared memory until it is done
    if (warp == 2) tmem_dealloc_pair(tmem_base, P_ACC_STAGES * P_BN);
}

template <int EPI, bool F8>
cudaError_t launch_pair(const CUtensorMap& ma, const CUtensorMap& mb, float* C, const PairArgs& a, int num_sms, cudaStream_t stream)
{
    auto kern = gemm_tc_pair_kernel<EPI, F8>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM);
    if (e != cudaSuccess) return e;
    const int n_tiles = a.tiles_m2 * a.tiles_n2;
    int pairs = num_sms / 2;
    if (pairs > n_tiles) pairs = n_tiles;
    if (pairs < 1) pairs = 1;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3((unsigned)(2 * pairs), 1, 1);
    cfg.blockDim = dim3(P_THREADS, 1, 1);
    cfg.dynamicSmemBytes = P_SMEM;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, ma, mb, C, a);
}
