// probe_fold_ops.cu -- which pipes do the packed fold's instructions share?  Issue cost per SMSP of
//   F2FP.BF16.F32.PACK_AB (cvt.rn.bf16x2.f32), HADD2.BF16_V2 (add.rn.bf16x2), HFMA2.BF16_V2 (fma.rn.bf16x2), and the
//   fold pairs cvt + add / cvt + fma / cvt + alternating, with operands re-defined every iteration by an empty volatile asm
//   (nothing can be hoisted or folded), 16 independent chains per thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 scripts/probe_fold_ops.cu -o scripts/build/probe_fold_ops.bin
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

template <int OP>
__global__ void __launch_bounds__(1024) probe(int iters, long long* clk, uint32_t* sink, float seed)
{
    constexpr int N = 16;
    float a[N], b[N];
    uint32_t acc[N];
#pragma unroll
    for (int i = 0; i < N; ++i) { a[i] = seed * (i + 1) + threadIdx.x; b[i] = seed * (i + 3); acc[i] = 0x3F803F80u; }
    const uint32_t one = 0x3F803F80u;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < N; ++i) asm volatile("" : "+f"(a[i]), "+f"(b[i]));       // "new data arrived"
#pragma unroll
        for (int i = 0; i < N; ++i) {
            uint32_t t = 0;
            if constexpr (OP == 0 || (OP >= 3 && OP <= 5)) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(t) : "f"(a[i]), "f"(b[i]));
            if constexpr (OP == 0) acc[i] ^= t;                                           // LOP3 (alu) keeps the result alive
            if constexpr (OP == 1) asm volatile("add.rn.bf16x2 %0, %0, %1;" : "+r"(acc[i]) : "r"(__float_as_uint(a[i])));
            if constexpr (OP == 2) asm volatile("fma.rn.bf16x2 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(__float_as_uint(a[i])), "r"(one));
            if constexpr (OP == 3) asm volatile("add.rn.bf16x2 %0, %0, %1;" : "+r"(acc[i]) : "r"(t));
            if constexpr (OP == 4) asm volatile("fma.rn.bf16x2 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(t), "r"(one));
            if constexpr (OP == 5) {
                if (i & 1) asm volatile("fma.rn.bf16x2 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(t), "r"(one));
                else asm volatile("add.rn.bf16x2 %0, %0, %1;" : "+r"(acc[i]) : "r"(t));
            }
            if constexpr (OP == 6) {        // cvt to f16x2 + HADD2.F16
                asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(t) : "f"(a[i]), "f"(b[i]));
                asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(acc[i]) : "r"(t));
            }
            if constexpr (OP == 7) {        // integer RNE to bf16 x 2 + PRMT + add
                uint32_t x = __float_as_uint(a[i]), y = __float_as_uint(b[i]);
                x += 0x7FFFu + ((x >> 16) & 1u);
                y += 0x7FFFu + ((y >> 16) & 1u);
                t = __byte_perm(y, x, 0x7632);
                asm volatile("fma.rn.bf16x2 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(t), "r"(one));
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
    const double groups = (double)iters * 16 * (threads / 32) / 4.0;      // per SMSP
    printf("%-44s threads %4d : %6.2f clk per element-pair group (warp-wide) per SMSP\n", name, threads, (double)c / groups);
}

int main()
{
    long long* d_clk; uint32_t* d_sink;
    cudaMalloc(&d_clk, 64); cudaMalloc(&d_sink, 64);
    for (int threads : {256, 512}) {
        run<0>("cvt.rn.bf16x2.f32 (+ LOP3)", threads, d_clk, d_sink);
        run<1>("add.rn.bf16x2", threads, d_clk, d_sink);
        run<2>("fma.rn.bf16x2", threads, d_clk, d_sink);
        run<3>("cvt + add.bf16x2", threads, d_clk, d_sink);
        run<4>("cvt + fma.bf16x2", threads, d_clk, d_sink);
        run<5>("cvt + alternating add / fma", threads, d_clk, d_sink);
        run<6>("cvt.f16x2 + add.f16x2", threads, d_clk, d_sink);
        run<7>("integer RNE x2 + PRMT + fma.bf16x2", threads, d_clk, d_sink);
    }
    return 0;
}
