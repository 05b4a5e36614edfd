// probe_handoff2.cu -- the per-chunk pipeline (k_chunk = one MMA) with ONE or TWO issuing threads.
// Finding of probe_issue_rate.cu: any value-returning instruction (mbarrier.try_wait / test_wait, ld.shared) issued by the thread
// that has just issued a tcgen05.mma does not return before that MMA has run -- an issuer that must poll an "accumulator stage free"
// barrier once per MMA therefore never has two MMAs in flight, and the issue -> complete latency (~280-340 clk) is exposed on every chunk.
// With one issuing thread PER TMEM STAGE (each polls only its own stage's barrier, whose release needs that MMA's completion anyway)
// the two stages' cycles overlap.  Operands resident in shared memory (no TMA), 128 x 256 x 32 e4m3 MMAs, 2 TMEM stages, 8 consumer warps
// that read their 128 columns back at once (4 x tcgen05.ld.x32, one wait), release the stage, then optionally run the bfloat162 fold.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I lofloat_b200/csrc scripts/probe_handoff2.cu -o scripts/build/probe_handoff2.bin
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

template <int WM>
__device__ __forceinline__ void wait_bar(uint64_t* bar, uint32_t parity)
{
    if constexpr (WM == 0) { mbar_wait(bar, parity); }
    else {
        uint32_t done = 0;
        while (!done) asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
}

template <int NISSUE, int FOLD, int NLD, int WM = 0>
__global__ void __launch_bounds__(128 + 8 * 32, 1)
probe(int n_chunks, uint32_t idesc, long long* clk_out)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) uint64_t full_bar[2], empty_bar[2];
    __shared__ uint32_t tmem_base_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (128 + 256) * 128 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x38383838u;
    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 8); }
        fence_mbar_init();
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 2) tmem_alloc(&tmem_base_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;
    const long long t_start = clock64();
    const bool issuer = (NISSUE == 1) ? (warp == 1) : (warp == 1 || warp == 3);
    if (issuer) {
        if (elect_one()) {
            const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + 128 * 128);
            if constexpr (NISSUE == 1) {
                uint32_t ts = 0, tphase = 0;
                for (int c = 0; c < n_chunks; ++c) {
                    wait_bar<WM>(&empty_bar[ts], tphase ^ 1u);
                    tc_fence_after();
                    umma_t<true>(tmem_base + ts * 256, umma_desc_k128(a_base + (c & 3) * 32), umma_desc_k128(b_base + (c & 3) * 32), idesc, 0u);
                    umma_commit(&full_bar[ts]);
                    if (++ts == 2) { ts = 0; tphase ^= 1u; }
                }
            } else {
                const uint32_t ts = (warp == 1) ? 0u : 1u;       // this thread's TMEM stage: chunks c = ts, ts + 2, ...
                uint32_t tphase = 0;
                for (int c = (int)ts; c < n_chunks; c += 2) {
                    wait_bar<WM>(&empty_bar[ts], tphase ^ 1u);
                    tc_fence_after();
                    umma_t<true>(tmem_base + ts * 256, umma_desc_k128(a_base + (c & 3) * 32), umma_desc_k128(b_base + (c & 3) * 32), idesc, 0u);
                    umma_commit(&full_bar[ts]);
                    tphase ^= 1u;
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
            wait_bar<WM>(&full_bar[cs], tphase);
            tc_fence_after();
            const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + cs * 256 + grp * 128;
#pragma unroll
            for (int u = 0; u < NLD; ++u) tmem_ld32(ta + u * 32, buf + u * 32);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < NLD * 32; ++i) asm volatile("" : "+r"(buf[i]));
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[cs]);
            if (++cs == 2) { cs = 0; tphase ^= 1u; }
            if constexpr (FOLD > 0) {
#pragma unroll
                for (int rep = 0; rep < (FOLD + 63) / 64; ++rep)
#pragma unroll
                for (int i = 0; i < (FOLD < 64 ? FOLD : 64); ++i) {
                    const __nv_bfloat162 a0 = *reinterpret_cast<const __nv_bfloat162*>(&accp[i]);
                    const __nv_bfloat162 r0 = __hadd2(a0, __floats2bfloat162_rn(__uint_as_float(buf[2 * i]), __uint_as_float(buf[2 * i + 1])));
                    accp[i] = *reinterpret_cast<const uint32_t*>(&r0);
                }
            } else {
                accp[0] ^= buf[0] ^ buf[NLD * 32 - 1];
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

template <int NISSUE, int FOLD, int NLD, int WM = 0>
void run(const char* what, int grid, long long* d_clk)
{
    const int n_chunks = 4096;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    auto kern = probe<NISSUE, FOLD, NLD, WM>;
    const int smem = (128 + 256) * 128 + 1024;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kern<<<grid, 128 + 8 * 32, smem>>>(n_chunks, idesc, d_clk);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<long long> clk(grid + 2);
    CK(cudaMemcpy(clk.data(), d_clk, (grid + 2) * sizeof(long long), cudaMemcpyDeviceToHost));
    long long mx = 0;
    for (int i = 0; i < grid; ++i) mx = clk[i + 2] > mx ? clk[i + 2] : mx;
    printf("%-70s issuing threads %d: %6.1f clk / chunk\n", what, NISSUE, (double)mx / n_chunks);
}

int main()
{
    long long* d_clk;
    CK(cudaMalloc(&d_clk, 1024 * sizeof(long long)));
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    run<1, 0, 4>("full read-back, no fold", sms, d_clk);
    run<1, 16, 4>("full read-back + 16 of 64 fold pairs", sms, d_clk);
    run<1, 32, 4>("full read-back + 32 of 64 fold pairs", sms, d_clk);
    run<1, 64, 4>("full read-back + the 64 fold pairs", sms, d_clk);
    run<1, 128, 4>("full read-back + 128 fold pairs (2x)", sms, d_clk);
    return 0;
}
