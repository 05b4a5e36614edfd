// probe_issue_rate.cu -- what does each instruction of the single-thread MMA issue loop cost?  One CTA, one issuing thread, operands in shared
// memory, 128 x 256 x 32 e4m3 MMAs alternating between two TMEM stages, with optional per-MMA extras:
//   1 = mbarrier.try_wait on an already completed barrier   2 = tcgen05.fence::after_thread_sync   4 = tcgen05.commit to a side barrier
//   8 = mbarrier.arrive on a second barrier                16 = descriptors rebuilt from a runtime stage index (R2UR traffic)
//  32 = mbarrier.test_wait (non-suspending) on a completed barrier   64 = ld.acquire.cta.shared poll of a flag word that is already set
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I lofloat_b200/csrc scripts/probe_issue_rate.cu -o scripts/build/probe_issue_rate.bin
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include "lof_ptx.cuh"
using namespace lof::ptx;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

template <int F>
__global__ void __launch_bounds__(128, 1) probe(int n_mma, uint32_t idesc, uint32_t* out, int nwarps_spin)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) uint64_t bar, side_bar, done_bar, arr_bar;
    __shared__ uint32_t slot;
    __shared__ uint32_t flag;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (128 + 256) * 128 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x38383838u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&side_bar, 1); mbar_init(&done_bar, 1); mbar_init(&arr_bar, 1); fence_mbar_init(); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 2) tmem_alloc(&slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x == 0) flag = 1000000u;
    if (threadIdx.x == 0) mbar_arrive(&done_bar);                 // phase 0 of done_bar is complete: try_wait(done_bar, 0) succeeds at once
    __syncthreads();
    if (warp == 1 && elect_one()) {
        const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + 128 * 128);
        const long long t0 = clock64();
        uint32_t st = 0;
        for (int j = 0; j < n_mma; ++j) {
            if constexpr ((F & 1) != 0) mbar_wait(&done_bar, 0u);
            if constexpr ((F & 32) != 0) {
                uint32_t done = 0;
                while (!done) asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(smem_u32(&done_bar)), "r"(0u) : "memory");
            }
            if constexpr ((F & 64) != 0) {
                uint32_t v = 0;
                do { asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(&flag)) : "memory"); } while (v < (uint32_t)j);
            }
            if constexpr ((F & 2) != 0) tc_fence_after();
            if constexpr ((F & 16) != 0) {
                st = (st + 1) & 3;
                umma_t<true>(slot + (j & 1) * 256, umma_desc_k128(a_base + st * 32), umma_desc_k128(b_base + st * 32), idesc, 0u);
            } else {
                umma_t<true>(slot + (j & 1) * 256, umma_desc_k128(a_base + (j & 3) * 32), umma_desc_k128(b_base + (j & 3) * 32), idesc, 0u);
            }
            if constexpr ((F & 4) != 0) umma_commit(&side_bar);
            if constexpr ((F & 8) != 0) mbar_arrive(&arr_bar);
        }
        umma_commit(&bar);
        while (!mbar_try_wait(&bar, 0u)) {}
        out[0] = (uint32_t)((clock64() - t0) / n_mma);
    } else if (warp == 3 && nwarps_spin > 0) {
        // a warp spinning on a barrier that never completes until the end (same SMSP as nobody: warp 3)
        while (!mbar_try_wait(&bar, 0u)) {}
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(slot, 512);
}

template <int F> void run(const char* what, uint32_t* d_out, int spin = 0)
{
    const uint32_t idesc = (1u << 4) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const int smem = (128 + 256) * 128 + 1024;
    auto kern = probe<F>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kern<<<1, 128, smem>>>(1024, idesc, d_out, spin);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    uint32_t c = 0;
    CK(cudaMemcpy(&c, d_out, 4, cudaMemcpyDeviceToHost));
    printf("%-90s %5u clk / MMA\n", what, c);
}

int main()
{
    uint32_t* d_out;
    CK(cudaMalloc(&d_out, 64));
    run<0>("MMA only", d_out);
    run<4>("MMA + commit (side barrier)", d_out);
    run<2>("MMA + tcgen05.fence::after_thread_sync", d_out);
    run<1>("MMA + try_wait on a completed barrier", d_out);
    run<8>("MMA + mbarrier.arrive", d_out);
    run<16>("MMA with runtime-indexed descriptors", d_out);
    run<4 | 2>("MMA + commit + fence", d_out);
    run<4 | 1>("MMA + commit + try_wait", d_out);
    run<4 | 2 | 1>("MMA + commit + fence + try_wait  (the pipeline's issue loop)", d_out);
    run<4 | 2 | 1 | 8>("MMA + commit + fence + try_wait + arrive", d_out);
    run<4 | 2 | 1>("MMA + commit + fence + try_wait, another warp spinning on a barrier", d_out, 1);
    run<32>("MMA + test_wait on a completed barrier", d_out);
    run<64>("MMA + ld.acquire.cta.shared poll of a set flag", d_out);
    run<4 | 2 | 64>("MMA + commit + fence + ld.acquire poll", d_out);
    run<4 | 2 | 64 | 8>("MMA + commit + fence + ld.acquire poll + mbarrier.arrive", d_out);
    return 0;
}
