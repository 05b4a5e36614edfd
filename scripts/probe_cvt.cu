// probe_cvt.cu -- issue throughput of the instructions the packed GEMM fold is made of (per SM, all 4 SMSPs busy):
// cvt.rn.bf16x2.f32 / cvt.rn.f16x2.f32 (F2FP.*.PACK_AB), add.bf16x2 / add.f16x2 (HADD2), and the fold pair (cvt + add) on
// independent registers.  Tells whether the fold is bound by the conversion pipe.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 scripts/probe_cvt.cu -o scripts/build/probe_cvt.bin
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

template <int OP>
__global__ void __launch_bounds__(1024) probe(int iters, long long* clk, uint32_t* sink, float seed)
{
    constexpr int N = 16;
    float a[N], b[N];
    uint32_t acc[N];
#pragma unroll
    for (int i = 0; i < N; ++i) { a[i] = seed * (i + 1) + threadIdx.x; b[i] = seed * (i + 3); acc[i] = threadIdx.x + i; }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if constexpr (OP == 0) {            // cvt.rn.bf16x2.f32 only
                asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(acc[i]) : "f"(a[i]), "f"(b[i]));
            } else if constexpr (OP == 1) {     // cvt.rn.f16x2.f32 only
                asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(acc[i]) : "f"(a[i]), "f"(b[i]));
            } else if constexpr (OP == 2) {     // add.bf16x2 only
                asm volatile("add.rn.bf16x2 %0, %0, %1;" : "+r"(acc[i]) : "r"(0x3F803F80u));
            } else if constexpr (OP == 3) {     // add.f16x2 only
                asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(acc[i]) : "r"(0x3C003C00u));
            } else if constexpr (OP == 4) {     // the bf16 fold: cvt + add
                uint32_t t;
                asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(t) : "f"(a[i]), "f"(b[i]));
                asm volatile("add.rn.bf16x2 %0, %0, %1;" : "+r"(acc[i]) : "r"(t));
            } else if constexpr (OP == 5) {     // fp32 add (reference: full-rate fma pipe)
                asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i]));
            } else if constexpr (OP == 6) {     // integer RNE-to-bf16 of two floats + PRMT pack (alu pipe alternative to F2FP)
                uint32_t x = __float_as_uint(a[i]), y = __float_as_uint(b[i]);
                x += 0x7FFFu + ((x >> 16) & 1u);
                y += 0x7FFFu + ((y >> 16) & 1u);
                acc[i] ^= __byte_perm(x, y, 0x7632);
            }
        }
    }
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) s ^= acc[i] ^ __float_as_uint(a[i]);
    if (s == 0x12345u) sink[0] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int threads, long long* d_clk, uint32_t* d_sink)
{
    const int iters = 4096;
    probe<OP><<<1, threads>>>(16, d_clk, d_sink, 1.25f);
    probe<OP><<<1, threads>>>(iters, d_clk, d_sink, 1.25f);
    cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, d_clk, sizeof c, cudaMemcpyDeviceToHost);
    const double warp_instr = (double)iters * 16 * (threads / 32);
    printf("%-34s threads %4d : %6.2f clk per warp-instruction-group per SMSP  (%.1f thread-ops / clk / SM)\n", name, threads,
           (double)c / (warp_instr / 4.0), warp_instr * 32.0 / (double)c);
}

int main()
{
    long long* d_clk; uint32_t* d_sink;
    cudaMalloc(&d_clk, 64); cudaMalloc(&d_sink, 64);
    for (int threads : {128, 256, 512, 1024}) {
        run<0>("cvt.rn.bf16x2.f32", threads, d_clk, d_sink);
        run<1>("cvt.rn.f16x2.f32", threads, d_clk, d_sink);
        run<2>("add.rn.bf16x2", threads, d_clk, d_sink);
        run<3>("add.rn.f16x2", threads, d_clk, d_sink);
        run<4>("cvt.bf16x2 + add.bf16x2 (fold)", threads, d_clk, d_sink);
        run<5>("add.rn.f32", threads, d_clk, d_sink);
        run<6>("integer RNE x2 + PRMT", threads, d_clk, d_sink);
    }
    return 0;
}
