// probe_fold_pipeline.cu -- isolates the MMA -> TMEM -> fold hand-shake of the chunked GEMM's persistent fold kernel
// (lofloat_b200/csrc/lof_gemm_tc3.cuh) from its operand traffic: one CTA per SM runs `n_chunks` k-chunks of a 128 x 256
// tile on operands that already sit in shared memory (no TMA, no L2), with the same barrier protocol, tcgen05.ld units and
// packed bfloat162 fold as the product kernel.  Switches: MMA on / off, fold arithmetic on / off, number of fold warps,
// unit width, MMAs per chunk -- to see which part of the per-chunk time is the tensor core, the TMEM read-back, the fold's
// ALU work, or the synchronisation between them.  Reports clk per chunk (SM clock64).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I lofloat_b200/csrc scripts/probe_fold_pipeline.cu -o scripts/build/probe_fold_pipeline.bin
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "lof_ptx.cuh"

using namespace lof::ptx;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int BN = 256;

__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t* v)
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}

// FW fold warps (8 or 16), UNIT columns per tcgen05.ld (16 or 32), DO_MMA / DO_LD / DO_FOLD switches.
// SPLIT = 1: one N = 256 MMA per k-step into a ring of two 256-column TMEM stages (first version of the product kernel);
// SPLIT = 2: two N = 128 MMAs per k-step into a ring of four 128-column stages, each drained by the warps of its column half.
template <int FW, int UNIT, bool DO_MMA, bool DO_LD, bool DO_FOLD, int SPLIT = 1>
__global__ void __launch_bounds__(128 + FW * 32, 1)
probe(int n_chunks, int mmas_per_chunk, uint32_t idesc, long long* clk_out, uint32_t* sink, int mma_mode)
{
    constexpr int kTmemStages = 2 * SPLIT;
    constexpr int SW = BN / SPLIT;                  // columns per TMEM stage
    static_assert(SPLIT == 1 || FW == 8, "the split layout is written for 8 fold warps");
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;                         // 128 rows x 128 B
    uint8_t* smem_b = smem + 128 * 128;             // 256 rows x 128 B
    __shared__ __align__(8) uint64_t tmem_full_bar[kTmemStages], tmem_empty_bar[kTmemStages];
    __shared__ uint32_t tmem_base_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (128 + 256) * 128 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x38383838u;   // e4m3 1.0
    if (threadIdx.x == 0) {
        for (int s = 0; s < kTmemStages; ++s) { mbar_init(&tmem_full_bar[s], 1); mbar_init(&tmem_empty_bar[s], FW / SPLIT); }
        fence_mbar_init();
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 2) tmem_alloc(&tmem_base_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;
    const long long t0 = clock64();

    if (warp == 1) {
        if (elect_one()) {
            uint32_t ts = 0, tphase = 0;
            const uint32_t a_base = smem_u32(smem_a), b_base = smem_u32(smem_b);
            for (int c = 0; c < n_chunks; ++c) {
                for (int j = 0; j < mmas_per_chunk; ++j) {
#pragma unroll
                    for (int hh = 0; hh < SPLIT; ++hh) {
                        if (j == 0 && mma_mode != 1) {
                            mbar_wait(&tmem_empty_bar[ts + hh], tphase ^ 1u);
                            if (mma_mode != 2) tc_fence_after();
                        }
                        if constexpr (DO_MMA) {
                            umma_t<true>(tmem_base + (ts + hh) * SW, umma_desc_k128(a_base + (j & 3) * 32),
                                         umma_desc_k128(b_base + hh * (128 * 128) + (j & 3) * 32), idesc, j > 0 ? 1u : 0u);
                            if (j == mmas_per_chunk - 1) umma_commit(&tmem_full_bar[ts + hh]);
                        } else {
                            if (j == mmas_per_chunk - 1) mbar_arrive(&tmem_full_bar[ts + hh]);
                        }
                    }
                }
                ts += SPLIT;
                if (ts == kTmemStages) { ts = 0; tphase ^= 1u; }
            }
            if (blockIdx.x == 0) { sink[1] = (uint32_t)n_chunks; sink[3] = (uint32_t)(clock64() - t0); }
        }
    } else if (warp >= 4) {
        constexpr int CPT = 128 * BN / (FW * 32);           // columns per thread: 128 (8 warps) / 64 (16 warps)
        constexpr int UNITS = CPT / UNIT;
        const int q = warp & 3, hcol = (warp - 4) >> 2;
        const uint32_t t_lane = (uint32_t)(q * 32) << 16;
        uint32_t accp[CPT / 2];
#pragma unroll
        for (int i = 0; i < CPT / 2; ++i) accp[i] = 0u;
        uint32_t buf0[UNIT], buf1[UNIT];
        uint32_t cs = 0, tphase = 0;
        constexpr uint32_t kStageOfHalf = (SPLIT == 2) ? 1u : 0u, kColOfHalf = (SPLIT == 2) ? 0u : (uint32_t)CPT;
        auto issue_unit = [&](int u, uint32_t* buf) {
            if constexpr (DO_LD) {
                const uint32_t a = tmem_base + t_lane + (cs + kStageOfHalf * hcol) * SW + kColOfHalf * hcol + (uint32_t)(u * UNIT);
                if constexpr (UNIT == 32) tmem_ld32(a, buf); else ld16(a, buf);
            }
        };
        auto fold_unit = [&](int u, const uint32_t* p) {
            if constexpr (DO_FOLD) {
#pragma unroll
                for (int i = 0; i < UNIT / 2; ++i) {
                    const __nv_bfloat162 a0 = *reinterpret_cast<const __nv_bfloat162*>(&accp[u * (UNIT / 2) + i]);
                    const __nv_bfloat162 r0 = __hadd2(a0, __floats2bfloat162_rn(__uint_as_float(p[2 * i]), __uint_as_float(p[2 * i + 1])));
                    accp[u * (UNIT / 2) + i] = *reinterpret_cast<const uint32_t*>(&r0);
                }
            } else {
                accp[u] ^= p[0] ^ p[UNIT - 1];
            }
        };
        auto pin = [&](uint32_t* buf) {
#pragma unroll
            for (int i = 0; i < UNIT; ++i) asm volatile("" : "+r"(buf[i]));
        };
        mbar_wait(&tmem_full_bar[cs + kStageOfHalf * hcol], tphase);
        tc_fence_after();
        issue_unit(0, buf0);
        for (int c = 0; c < n_chunks; ++c) {
#pragma unroll
            for (int u = 0; u < UNITS; ++u) {
                uint32_t* cur = (u & 1) ? buf1 : buf0;
                uint32_t* nxt = (u & 1) ? buf0 : buf1;
                if constexpr (DO_LD) tmem_ld_wait();
                pin(cur);
                bool issued = true;
                if (u + 1 < UNITS) {
                    issue_unit(u + 1, nxt);
                } else {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty_bar[cs + kStageOfHalf * hcol]);
                    cs += SPLIT;
                    if (cs == kTmemStages) { cs = 0; tphase ^= 1u; }
                    if (c + 1 < n_chunks) {
                        issued = mbar_try_wait(&tmem_full_bar[cs + kStageOfHalf * hcol], tphase);
                        if (issued) { tc_fence_after(); issue_unit(0, nxt); }
                    }
                }
                fold_unit(u, cur);
                if (!issued) {
                    mbar_wait(&tmem_full_bar[cs + kStageOfHalf * hcol], tphase);
                    tc_fence_after();
                    issue_unit(0, nxt);
                }
            }
        }
        uint32_t s = 0;
#pragma unroll
        for (int i = 0; i < CPT / 2; ++i) s ^= accp[i];
        if (s == 0x1234567u) sink[0] = s;
        if (blockIdx.x == 0 && threadIdx.x == 128) sink[2] = (uint32_t)n_chunks;
    }
    tc_fence_before();
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) clk_out[blockIdx.x] = t1 - t0;
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}


// ---- (1) tcgen05.mma latency / throughput: N back-to-back MMAs (128 x 256 x 32, e4m3) into one TMEM stage, one commit, one wait
template <int COMMIT_EVERY, int ACC, int NDST>
__global__ void __launch_bounds__(128, 1) probe_mma_rate(int n_mma, int reps, uint32_t idesc, uint32_t* out)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) uint64_t bar, side_bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (128 + 256) * 128 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x38383838u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&side_bar, 1); fence_mbar_init(); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 2) tmem_alloc(&slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1 && elect_one()) {
        const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + 128 * 128);
        uint32_t phase = 0;
        long long total = 0;
        for (int r = 0; r < reps; ++r) {
            const long long t0 = clock64();
            for (int j = 0; j < n_mma; ++j) {
                umma_t<true>(slot + (j & (NDST - 1)) * 256, umma_desc_k128(a_base + (j & 3) * 32), umma_desc_k128(b_base + (j & 3) * 32), idesc, (uint32_t)ACC);
                if constexpr (COMMIT_EVERY > 0) {
                    if ((j & (COMMIT_EVERY - 1)) == COMMIT_EVERY - 1) umma_commit(&side_bar);   // nobody waits on it
                }
            }
            umma_commit(&bar);
            while (!mbar_try_wait(&bar, phase)) {}
            total += clock64() - t0;
            phase ^= 1u;
        }
        out[0] = (uint32_t)(total / reps);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(slot, 512);
}

// ---- (2) mbarrier hop latency: warp 1 (one thread) and warp 4 (32 threads, one arrives) ping-pong through two mbarriers
// WAIT: 0 = mbarrier.try_wait (may suspend), 1 = mbarrier.test_wait spin, 2 = try_wait with a 0 ns suspend-time hint
template <int WAIT>
__device__ __forceinline__ void wait_mode(uint64_t* bar, uint32_t parity)
{
    if constexpr (WAIT == 0) {
        while (!mbar_try_wait(bar, parity)) {}
    } else if constexpr (WAIT == 1) {
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                         : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        }
    } else {
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}\n"
                         : "=r"(done) : "r"(smem_u32(bar)), "r"(parity), "r"(0u) : "memory");
        }
    }
}
template <int WAIT>
__global__ void __launch_bounds__(256, 1) probe_hop(int n, uint32_t* out)
{
    __shared__ __align__(8) uint64_t ping, pong;
    if (threadIdx.x == 0) { mbar_init(&ping, 1); mbar_init(&pong, 1); fence_mbar_init(); }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 1) {
        if (elect_one()) {
            const long long t0 = clock64();
            uint32_t phase = 0;
            for (int i = 0; i < n; ++i) {
                mbar_arrive(&ping);
                wait_mode<WAIT>(&pong, phase);
                phase ^= 1u;
            }
            out[0] = (uint32_t)((clock64() - t0) / n);
        }
    } else if (warp == 4) {
        uint32_t phase = 0;
        for (int i = 0; i < n; ++i) {
            wait_mode<WAIT>(&ping, phase);
            phase ^= 1u;
            __syncwarp();
            if (lane == 0) mbar_arrive(&pong);
        }
    }
}

template <int FW, int UNIT, bool DO_MMA, bool DO_LD, bool DO_FOLD, int SPLIT = 1>
void run(const char* what, int grid, int mmas_per_chunk, long long* d_clk, uint32_t* d_sink, int mma_mode = 0)
{
    const int n_chunks = 2048;
    const uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)((BN / SPLIT) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // f32 accum, e4m3 x e4m3
    auto kern = probe<FW, UNIT, DO_MMA, DO_LD, DO_FOLD, SPLIT>;
    const int smem = (128 + 256) * 128 + 1024;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kern<<<grid, 128 + FW * 32, smem>>>(64, mmas_per_chunk, idesc, d_clk, d_sink, mma_mode);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    kern<<<grid, 128 + FW * 32, smem>>>(n_chunks, mmas_per_chunk, idesc, d_clk, d_sink, mma_mode);
    cudaEventRecord(e1);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> clk(grid);
    CK(cudaMemcpy(clk.data(), d_clk, grid * sizeof(long long), cudaMemcpyDeviceToHost));
    long long mx = 0;
    for (auto c : clk) mx = c > mx ? c : mx;
    uint32_t dbg[4] = {0, 0, 0, 0};
    CK(cudaMemcpy(dbg, d_sink, sizeof dbg, cudaMemcpyDeviceToHost));
    printf("%-58s split %d grid %3d  fold warps %2d  unit x%d  MMAs/chunk %d : %7.1f clk / chunk  (event time %.1f ns / chunk; mma thread saw %u chunks, fold warp %u)\n", what, SPLIT, grid, FW, UNIT, mmas_per_chunk, (double)mx / n_chunks, ms * 1e6 / n_chunks, dbg[1], dbg[2]);
    printf("    mma thread: %.1f clk / chunk\n", (double)dbg[3] / n_chunks);
}

int main()
{
    long long* d_clk; uint32_t* d_sink;
    CK(cudaMalloc(&d_clk, 1024 * sizeof(long long)));
    CK(cudaMalloc(&d_sink, 64));
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const int smem = (128 + 256) * 128 + 1024;
        auto rate = [&](auto kern, const char* what) {
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            printf("tcgen05.mma 128x256x32 e4m3, %s\n", what);
            for (int n : {1, 4, 16, 256}) {
                kern<<<1, 128, smem>>>(n, 64, idesc, d_sink);
                CK(cudaGetLastError());
                CK(cudaDeviceSynchronize());
                uint32_t c = 0;
                CK(cudaMemcpy(&c, d_sink, 4, cudaMemcpyDeviceToHost));
                printf("    %3d MMAs + final commit + wait: %6u clk total, %7.1f clk / MMA\n", n, c, (double)c / n);
            }
        };
        rate(probe_mma_rate<0, 1, 2>, "accumulate, 2 destinations, one commit at the end");
        rate(probe_mma_rate<0, 0, 2>, "OVERWRITE (accumulate = 0), 2 destinations, one commit at the end");
        rate(probe_mma_rate<1, 1, 2>, "accumulate, 2 destinations, a commit after EVERY MMA");
        rate(probe_mma_rate<1, 0, 2>, "overwrite, 2 destinations, a commit after every MMA  (== the chunked GEMM, k_chunk = one MMA)");
        rate(probe_mma_rate<1, 1, 1>, "accumulate, 1 destination, a commit after every MMA");
        rate(probe_mma_rate<4, 1, 2>, "accumulate, 2 destinations, a commit after every 4th MMA");
        uint32_t c = 0;
        probe_hop<0><<<1, 256>>>(4096, d_sink); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&c, d_sink, 4, cudaMemcpyDeviceToHost));
        printf("mbarrier ping-pong round trip (2 hops), try_wait            : %u clk\n", c);
        probe_hop<1><<<1, 256>>>(4096, d_sink); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&c, d_sink, 4, cudaMemcpyDeviceToHost));
        printf("mbarrier ping-pong round trip (2 hops), test_wait spin      : %u clk\n", c);
        probe_hop<2><<<1, 256>>>(4096, d_sink); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&c, d_sink, 4, cudaMemcpyDeviceToHost));
        printf("mbarrier ping-pong round trip (2 hops), try_wait hint 0 ns  : %u clk\n", c);
    }
    for (int grid : {sms}) {
        run<8, 32, true, false, false>("MMA only (fold warps just hand the stage back)", grid, 1, d_clk, d_sink);
        run<8, 32, true, false, false>("MMA only, 4 MMAs per chunk", grid, 4, d_clk, d_sink);
        run<8, 32, true, false, false>("MMA only, 4 MMAs per chunk, issuer never waits for empty", grid, 4, d_clk, d_sink, 1);
        run<8, 32, true, false, false>("MMA only, 4 MMAs per chunk, issuer waits, no tcgen05 fence", grid, 4, d_clk, d_sink, 2);
        run<8, 32, true, false, false>("MMA only, 1 MMA per chunk, issuer never waits for empty", grid, 1, d_clk, d_sink, 1);
        run<8, 32, true, false, false>("MMA only, 1 MMA per chunk, issuer waits, no tcgen05 fence", grid, 1, d_clk, d_sink, 2);
        run<8, 32, false, true, false>("tcgen05.ld only (no MMA, no fold arithmetic)", grid, 1, d_clk, d_sink);
        run<8, 32, false, false, true>("fold arithmetic only (no MMA, no tcgen05.ld)", grid, 1, d_clk, d_sink);
        run<8, 32, false, true, true>("tcgen05.ld + fold (no MMA)", grid, 1, d_clk, d_sink);
        run<8, 32, true, true, false>("MMA + tcgen05.ld (no fold arithmetic)", grid, 1, d_clk, d_sink);
        run<8, 32, true, true, true>("MMA + tcgen05.ld + fold  == the product kernel's loop", grid, 1, d_clk, d_sink);
        run<8, 32, true, true, true>("MMA + tcgen05.ld + fold, 2 MMAs per chunk (bf16 operands)", grid, 2, d_clk, d_sink);
        run<8, 32, true, true, true>("MMA + tcgen05.ld + fold, 4 MMAs per chunk (k_chunk 128)", grid, 4, d_clk, d_sink);
        run<16, 32, false, true, true>("16 fold warps: tcgen05.ld + fold (no MMA)", grid, 1, d_clk, d_sink);
        run<16, 32, true, true, true>("16 fold warps: MMA + tcgen05.ld + fold", grid, 1, d_clk, d_sink);
        run<16, 16, true, true, true>("16 fold warps, x16 units: MMA + tcgen05.ld + fold", grid, 1, d_clk, d_sink);
        run<8, 16, true, true, true>("8 fold warps, x16 units: MMA + tcgen05.ld + fold", grid, 1, d_clk, d_sink);
        run<8, 32, true, false, false, 2>("4 x 128-column stages: MMA only", grid, 1, d_clk, d_sink);
        run<8, 32, true, true, false, 2>("4 x 128-column stages: MMA + tcgen05.ld", grid, 1, d_clk, d_sink);
        run<8, 32, true, true, true, 2>("4 x 128-column stages: MMA + tcgen05.ld + fold", grid, 1, d_clk, d_sink);
        run<8, 32, true, true, true, 2>("4 x 128-column stages: ... 2 k-steps per chunk (bf16 operands)", grid, 2, d_clk, d_sink);
        run<8, 32, true, true, true, 2>("4 x 128-column stages: ... 4 k-steps per chunk (k_chunk 128)", grid, 4, d_clk, d_sink);
        run<8, 16, true, true, true, 2>("4 x 128-column stages, x16 units: MMA + tcgen05.ld + fold", grid, 1, d_clk, d_sink);
    }
    return 0;
}
