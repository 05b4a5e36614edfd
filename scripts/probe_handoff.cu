// probe_handoff.cu -- timeline of the MMA <-> fold-warp hand-off of the chunked GEMM's per-chunk pipeline (k_chunk = one MMA):
// one CTA per SM, operands resident in shared memory (no TMA), 2 TMEM stages of 256 columns, kind::f8f6f4 128 x 256 x 32 MMAs.
// For chunks [T0, T0 + 32) of CTA 0 it records (SM clock64): MMA issue, full barrier seen by the consumer, stage released, empty
// barrier seen by the issuer -- the four edges of the cycle -- for several consumer protocols.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I lofloat_b200/csrc scripts/probe_handoff.cu -o scripts/build/probe_handoff.bin
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "lof_ptx.cuh"
using namespace lof::ptx;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int T0 = 512, TN = 32;

// FW consumer warps (warps 4 ..); POLL: 0 = every lane waits on the barrier, 1 = lane 0 waits, then __syncwarp; LD: 0 none, 1 = each
// consumer warp reads 32 columns before releasing, 2 = reads its whole share (256 / (FW / 4) columns, x64 pieces) before releasing
template <int FW, int POLL, int LD, int NSTAGE, int IMODE = 0>
__global__ void __launch_bounds__(128 + FW * 32, 1)
probe(int n_chunks, uint32_t idesc, long long* tl, long long* clk_out)
{
    constexpr int SW = 512 / NSTAGE;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) uint64_t full_bar[NSTAGE], empty_bar[NSTAGE], side_bar;
    __shared__ uint32_t tmem_base_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (128 + 256) * 128 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x38383838u;
    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], FW); }
        mbar_init(&side_bar, 1);
        fence_mbar_init();
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 2) tmem_alloc(&tmem_base_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;
    const long long t_start = clock64();
    if (warp == 1) {
        if (elect_one()) {
            uint32_t ts = 0, tphase = 0;
            const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + 128 * 128);
            for (int c = 0; c < n_chunks; ++c) {
                if constexpr (IMODE != 1) mbar_wait(&empty_bar[ts], tphase ^ 1u);   // IMODE 1: the issuer never waits (timing only)
                const bool rec = blockIdx.x == 0 && c >= T0 && c < T0 + TN;
                if (rec) tl[(c - T0) * 4 + 3] = clock64();                       // empty seen (for the chunk that used this stage before)
                tc_fence_after();
                if (rec) tl[(c - T0) * 4 + 0] = clock64();                       // MMA issue
                umma_t<true>(tmem_base + ts * SW, umma_desc_k128(a_base + (c & 3) * 32), umma_desc_k128(b_base + (c & 3) * 32), idesc, 0u);
                if constexpr (IMODE == 2) { umma_commit(&side_bar); mbar_arrive(&full_bar[ts]); }   // IMODE 2: plain arrive instead of the commit's (timing only)
                else umma_commit(&full_bar[ts]);
                if (++ts == NSTAGE) { ts = 0; tphase ^= 1u; }
            }
        }
    } else if (warp >= 4) {
        const int q = warp & 3, grp = (warp - 4) >> 2;
        constexpr int CPW = SW / (FW / 4);
        uint32_t cs = 0, tphase = 0, sink = 0;
        uint32_t buf[64];
        for (int c = 0; c < n_chunks; ++c) {
            if constexpr (POLL == 0) mbar_wait(&full_bar[cs], tphase);
            else if constexpr (POLL == 1) { if (lane == 0) mbar_wait(&full_bar[cs], tphase); __syncwarp(); }
            else if constexpr (POLL == 2) {                                    // ONE warp polls the mbarrier, the others sleep in a named barrier
                if (warp == 4) mbar_wait(&full_bar[cs], tphase);
                asm volatile("bar.sync 1, %0;" ::"r"(FW * 32) : "memory");
            } else {                                                           // every lane polls, with a back-off between polls
                while (!mbar_try_wait(&full_bar[cs], tphase)) __nanosleep(POLL);
            }
            const bool rec = blockIdx.x == 0 && warp == 4 && lane == 0 && c >= T0 && c < T0 + TN;
            if (rec) tl[(c - T0) * 4 + 1] = clock64();                           // full seen
            tc_fence_after();
            if constexpr (LD >= 1) {
                const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + cs * SW + grp * CPW;
                constexpr int NLD = LD == 1 ? 1 : (CPW / 32);
#pragma unroll
                for (int u = 0; u < NLD && u < 2; ++u) tmem_ld32(ta + u * 32, buf + u * 32);
                if constexpr (NLD > 2) {
#pragma unroll
                    for (int u = 2; u < NLD; ++u) { tmem_ld_wait(); sink ^= buf[0] ^ buf[33]; tmem_ld32(ta + u * 32, buf + (u & 1) * 32); }
                }
                tmem_ld_wait();
                sink ^= buf[0] ^ buf[31];
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[cs]);
            if (rec) tl[(c - T0) * 4 + 2] = clock64();                           // released
            if (++cs == NSTAGE) { cs = 0; tphase ^= 1u; }
        }
        if (sink == 0x12345u) clk_out[1] = sink;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) clk_out[blockIdx.x + 2] = clock64() - t_start;
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

template <int FW, int POLL, int LD, int NSTAGE, int IMODE = 0>
void run(const char* what, int grid, long long* d_tl, long long* d_clk)
{
    const int n_chunks = 2048;
    constexpr int SW = 512 / NSTAGE;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(SW >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    auto kern = probe<FW, POLL, LD, NSTAGE, IMODE>;
    const int smem = (128 + 256) * 128 + 1024;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaMemset(d_tl, 0, TN * 4 * sizeof(long long)));
    kern<<<grid, 128 + FW * 32, smem>>>(n_chunks, idesc, d_tl, d_clk);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<long long> tl(TN * 4), clk(grid + 2);
    CK(cudaMemcpy(tl.data(), d_tl, TN * 4 * sizeof(long long), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(clk.data(), d_clk, (grid + 2) * sizeof(long long), cudaMemcpyDeviceToHost));
    long long mx = 0;
    for (int i = 0; i < grid; ++i) mx = clk[i + 2] > mx ? clk[i + 2] : mx;
    double issue_to_full = 0, full_to_rel = 0, rel_to_empty = 0, period = 0;
    int n = 0;
    for (int i = 0; i + NSTAGE < TN; ++i) {
        issue_to_full += (double)(tl[i * 4 + 1] - tl[i * 4 + 0]);
        full_to_rel += (double)(tl[i * 4 + 2] - tl[i * 4 + 1]);
        rel_to_empty += (double)(tl[(i + NSTAGE) * 4 + 3] - tl[i * 4 + 2]);       // the issuer sees the stage free when it comes back for chunk i + NSTAGE
        period += (double)(tl[(i + 1) * 4 + 0] - tl[i * 4 + 0]);
        ++n;
    }
    printf("%-72s stages %d x %3d cols, %2d consumer warps: %6.1f clk / chunk | issue->full seen %6.1f, full seen->released %6.1f, released->empty seen %6.1f, issue period %6.1f\n",
           what, NSTAGE, SW, FW, (double)mx / n_chunks, issue_to_full / n, full_to_rel / n, rel_to_empty / n, period / n);
}

int main()
{
    long long *d_tl, *d_clk;
    CK(cudaMalloc(&d_tl, TN * 4 * sizeof(long long)));
    CK(cudaMalloc(&d_clk, 1024 * sizeof(long long)));
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int grid : {sms}) {
        printf("grid %d\n", grid);
        run<8, 0, 0, 2>("8 warps, every lane waits, no read-back", grid, d_tl, d_clk);
        run<8, 1, 0, 2>("8 warps, lane 0 waits, no read-back", grid, d_tl, d_clk);
        run<4, 0, 0, 2>("4 warps, every lane waits, no read-back", grid, d_tl, d_clk);
        run<4, 1, 0, 2>("4 warps, lane 0 waits, no read-back", grid, d_tl, d_clk);
        run<8, 0, 1, 2>("8 warps, every lane waits, one x32 read-back", grid, d_tl, d_clk);
        run<8, 1, 1, 2>("8 warps, lane 0 waits, one x32 read-back", grid, d_tl, d_clk);
        run<8, 0, 2, 2>("8 warps, every lane waits, full read-back (4 x32 per warp)", grid, d_tl, d_clk);
        run<16, 0, 2, 2>("16 warps, every lane waits, full read-back (2 x32 per warp)", grid, d_tl, d_clk);
        run<8, 0, 0, 4>("8 warps, every lane waits, no read-back", grid, d_tl, d_clk);
        run<8, 0, 2, 4>("8 warps, every lane waits, full read-back (2 x32 per warp)", grid, d_tl, d_clk);
        run<8, 0, 0, 4, 2>("8 warps, no read-back, commit to a side barrier + plain arrive on full", grid, d_tl, d_clk);
        run<4, 0, 0, 4, 0>("4 warps, no read-back", grid, d_tl, d_clk);
        run<8, 2, 0, 4, 0>("8 warps, ONE polls + named barrier, no read-back", grid, d_tl, d_clk);
        run<8, 2, 0, 2, 0>("8 warps, ONE polls + named barrier, no read-back", grid, d_tl, d_clk);
        run<8, 2, 2, 2, 0>("8 warps, ONE polls + named barrier, full read-back", grid, d_tl, d_clk);
        run<8, 20, 0, 4, 0>("8 warps, poll with nanosleep(20), no read-back", grid, d_tl, d_clk);
        run<8, 20, 0, 2, 0>("8 warps, poll with nanosleep(20), no read-back", grid, d_tl, d_clk);
        run<8, 100, 0, 2, 0>("8 warps, poll with nanosleep(100), no read-back", grid, d_tl, d_clk);
        run<8, 100, 2, 2, 0>("8 warps, poll with nanosleep(100), full read-back", grid, d_tl, d_clk);
        run<8, 0, 0, 2, 2>("8 warps, no read-back, commit to a side barrier + plain arrive on full", grid, d_tl, d_clk);
    }
    return 0;
}
