// probe_handoff3.cu -- who should issue the next MMA?  In the per-chunk pipeline (k_chunk = one MMA, 2 TMEM stages) the accumulator stage travels
// MMA -> fold warps (read-back) -> issuing thread -> MMA; probe_handoff.cu measured ~340 clk issue -> full seen and ~200 clk released -> empty seen.
// Variant B lets the LAST fold warp that finishes reading a stage issue the next MMA into it (shared-memory arrival counter, no mbarrier hop, no
// separate issuing thread).  Operands resident in shared memory, 128 x 256 x 32 e4m3 MMAs, 8 consumer warps x 128 columns (x32 units, double-buffered,
// fold of the previous unit overlaps the next load) so that the register budget matches the product kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I lofloat_b200/csrc scripts/probe_handoff3.cu -o scripts/build/probe_handoff3.bin
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

// SELF = false: warp 1 issues, fold warps arrive on empty mbarriers.  SELF = true: the last fold warp to finish a stage's read-back issues.
template <bool SELF, bool FOLD, int NISSUE = 1>
__global__ void __launch_bounds__(128 + 8 * 32, 1)
probe(int n_chunks, uint32_t idesc, long long* clk_out)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) uint64_t full_bar[2], empty_bar[2];
    __shared__ uint32_t tmem_base_slot;
    __shared__ unsigned int done_cnt[2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (128 + 256) * 128 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x38383838u;
    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 8); done_cnt[s] = 0; }
        fence_mbar_init();
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 2) tmem_alloc(&tmem_base_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;
    const uint32_t a_base = smem_addr_once(smem), b_base = smem_addr_once(smem + 128 * 128);
    const uint32_t full_a = smem_addr_once(&full_bar[0]), empty_a = smem_addr_once(&empty_bar[0]);
    const long long t_start = clock64();
    if (NISSUE == 2 && (warp == 1 || warp == 3)) {
        if (elect_one()) {                              // one issuing thread PER TMEM stage (hoisted addresses)
            const uint32_t ts = (warp == 1) ? 0u : 1u;
            uint32_t tphase = 0;
            for (int c = (int)ts; c < n_chunks; c += 2) {
                mbar_wait_a(empty_a + ts * 8u, tphase ^ 1u);
                tc_fence_after();
                umma_t<true>(tmem_base + ts * 256, umma_desc_k128(a_base + (c & 3) * 32), umma_desc_k128(b_base + (c & 3) * 32), idesc, 0u);
                umma_commit_a(full_a + ts * 8u);
                tphase ^= 1u;
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            if constexpr (SELF) {                       // prime the two stages, the fold warps take over
                for (int c = 0; c < 2 && c < n_chunks; ++c) {
                    umma_t<true>(tmem_base + c * 256, umma_desc_k128(a_base + (c & 3) * 32), umma_desc_k128(b_base + (c & 3) * 32), idesc, 0u);
                    umma_commit_a(full_a + c * 8u);
                }
            } else {
                uint32_t ts = 0, tphase = 0;
                for (int c = 0; c < n_chunks; ++c) {
                    mbar_wait_a(empty_a + ts * 8u, tphase ^ 1u);
                    tc_fence_after();
                    umma_t<true>(tmem_base + ts * 256, umma_desc_k128(a_base + (c & 3) * 32), umma_desc_k128(b_base + (c & 3) * 32), idesc, 0u);
                    umma_commit_a(full_a + ts * 8u);
                    if (++ts == 2) { ts = 0; tphase ^= 1u; }
                }
            }
        }
    } else if (warp >= 4) {
        const int q = warp & 3, grp = (warp - 4) >> 2;
        uint32_t cs = 0, tphase = 0;
        uint32_t buf[128];
        uint32_t accp[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) accp[i] = 0u;
        for (int c = 0; c < n_chunks; ++c) {
            mbar_wait_a(full_a + cs * 8u, tphase);
            tc_fence_after();
            const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + cs * 256 + grp * 128;
#pragma unroll
            for (int u = 0; u < 4; ++u) tmem_ld32(ta + u * 32, buf + u * 32);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 128; ++i) asm volatile("" : "+r"(buf[i]));
            tc_fence_before();
            __syncwarp();
            if constexpr (SELF) {
                unsigned int old = 0;
                if (lane == 0) old = atomicAdd(&done_cnt[cs], 1u);
                old = __shfl_sync(0xFFFFFFFFu, old, 0);
                if (old == 7u) {                        // this warp finished the stage last: it issues the MMA of chunk c + 2 into it
                    if (lane == 0) done_cnt[cs] = 0;
                    __syncwarp();
                    if (c + 2 < n_chunks && elect_one()) {
                        tc_fence_after();
                        umma_t<true>(tmem_base + cs * 256, umma_desc_k128(a_base + ((c + 2) & 3) * 32), umma_desc_k128(b_base + ((c + 2) & 3) * 32), idesc, 0u);
                        umma_commit_a(full_a + cs * 8u);
                    }
                    __syncwarp();
                }
            } else {
                if (lane == 0) mbar_arrive_a(empty_a + cs * 8u);
            }
            if (++cs == 2) { cs = 0; tphase ^= 1u; }
            if constexpr (FOLD) {
#pragma unroll
                for (int i = 0; i < 64; ++i) {
                    const __nv_bfloat162 a0 = *reinterpret_cast<const __nv_bfloat162*>(&accp[i]);
                    const __nv_bfloat162 r0 = __hadd2(a0, __floats2bfloat162_rn(__uint_as_float(buf[2 * i]), __uint_as_float(buf[2 * i + 1])));
                    accp[i] = *reinterpret_cast<const uint32_t*>(&r0);
                }
            } else {
                accp[0] ^= buf[0] ^ buf[127];
            }
        }
        uint32_t s = 0;
#pragma unroll
        for (int i = 0; i < 64; ++i) s ^= accp[i];
        if (s == 0x12345u) clk_out[1] = s;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) clk_out[blockIdx.x + 2] = clock64() - t_start;
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

template <bool SELF, bool FOLD, int NISSUE = 1>
void run(const char* what, int grid, long long* d_clk)
{
    const int n_chunks = 4096;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    auto kern = probe<SELF, FOLD, NISSUE>;
    const int smem = (128 + 256) * 128 + 1024;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kern<<<grid, 128 + 8 * 32, smem>>>(n_chunks, idesc, d_clk);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<long long> clk(grid + 2);
    CK(cudaMemcpy(clk.data(), d_clk, (grid + 2) * sizeof(long long), cudaMemcpyDeviceToHost));
    long long mx = 0;
    for (int i = 0; i < grid; ++i) mx = clk[i + 2] > mx ? clk[i + 2] : mx;
    printf("%-86s %6.1f clk / chunk\n", what, (double)mx / n_chunks);
}

int main()
{
    long long* d_clk;
    CK(cudaMalloc(&d_clk, 1024 * sizeof(long long)));
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    run<false, false>("issuing thread + empty mbarriers, read-back only", sms, d_clk);
    run<false, false, 2>("TWO issuing threads (one per stage) + empty mbarriers, read-back only", sms, d_clk);
    run<false, true, 2>("TWO issuing threads (one per stage), read-back + bfloat162 fold (spills)", sms, d_clk);
    run<true, false>("last fold warp issues the next MMA, read-back only", sms, d_clk);
    run<false, true>("issuing thread + empty mbarriers, read-back + bfloat162 fold (spills: 192 registers wanted)", sms, d_clk);
    run<true, true>("last fold warp issues the next MMA, read-back + bfloat162 fold (spills)", sms, d_clk);
    return 0;
}
