// probe_tmem_ld.cu -- measures the TMEM -> register read-back throughput (tcgen05.ld) of one B200 SM, the bound of
// the chunked LoFloat GEMM's per-chunk software fold: at k_chunk = 32 every tcgen05.mma is followed
// by a read of the whole fp32 partial tile.  Reports bytes / clk / SM for different instruction shapes, warps per
// lane quarter and loads in flight per wait, on a single SM and on all SMs at once.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I lofloat_b200/csrc scripts/probe_tmem_ld.cu -o scripts/build/probe_tmem_ld.bin
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "lof_ptx.cuh"

using namespace lof::ptx;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

template <int N> struct Regs { uint32_t v[N]; };

__device__ __forceinline__ void ld_x16(uint32_t taddr, uint32_t* v)
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ld_x64(uint32_t taddr, uint32_t* v)
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
        "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
        "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]),
          "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]),
          "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]),
          "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]),
          "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
        : "r"(taddr) : "memory");
}
// 16 lanes x 256 bits, repeated 16 times along the columns: 64 registers per thread, 8 KB per warp instruction
__device__ __forceinline__ void ld_16x256_x16(uint32_t taddr, uint32_t* v)
{
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
        "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
        "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]),
          "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]),
          "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]),
          "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]),
          "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
        : "r"(taddr) : "memory");
}

// SHAPE: 0 = 32x32b.x16 (2 KB / warp instr), 1 = 32x32b.x32 (4 KB), 2 = 32x32b.x64 (8 KB), 3 = 16x256b.x16 (8 KB)
// INFLIGHT: loads issued back to back before one tcgen05.wait::ld
template <int SHAPE, int INFLIGHT>
__global__ void __launch_bounds__(512, 1) probe(int iters, long long* clk_out, uint32_t* sink)
{
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc(&slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
    constexpr int COLS = SHAPE == 0 ? 16 : SHAPE == 1 ? 32 : SHAPE == 2 ? 64 : 128;   // columns covered by one instruction (16x256b.x16: 16 lanes x 128 columns)
    constexpr int REGS = SHAPE == 0 ? 16 : SHAPE == 1 ? 32 : 64;
    uint32_t v[INFLIGHT][REGS];
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < INFLIGHT; ++j) {
            const uint32_t col = (uint32_t)(((it * INFLIGHT + j) * COLS + (warp >> 2) * 128) & 511) & ~(uint32_t)(COLS - 1);
            const uint32_t a = base + (col <= 512 - COLS ? col : 0);
            if constexpr (SHAPE == 0) ld_x16(a, v[j]);
            else if constexpr (SHAPE == 1) tmem_ld32(a, v[j]);
            else if constexpr (SHAPE == 2) ld_x64(a, v[j]);
            else ld_16x256_x16(a, v[j]);
        }
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < INFLIGHT; ++j) acc ^= v[j][0] ^ v[j][REGS - 1];
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) clk_out[blockIdx.x] = t1 - t0;
    if (acc == 0x12345678u) sink[0] = acc;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(slot, 512);
}

template <int SHAPE, int INFLIGHT>
void run(const char* name, int warps, int grid, long long* d_clk, uint32_t* d_sink)
{
    const int iters = 4096;
    probe<SHAPE, INFLIGHT><<<grid, warps * 32>>>(64, d_clk, d_sink);     // warm-up
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<SHAPE, INFLIGHT><<<grid, warps * 32>>>(iters, d_clk, d_sink);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> clk(grid);
    CK(cudaMemcpy(clk.data(), d_clk, grid * sizeof(long long), cudaMemcpyDeviceToHost));
    long long mx = 0;
    for (auto c : clk) mx = c > mx ? c : mx;
    const double bytes_per_instr = (SHAPE == 0 ? 2048.0 : SHAPE == 1 ? 4096.0 : 8192.0);
    const double bytes = bytes_per_instr * INFLIGHT * iters * warps;        // per SM
    printf("%-14s inflight %d  warps %2d  grid %3d : %7.1f B/clk/SM  (%lld clk, %.3f ms, %.2f TB/s chip-wide at this grid)\n",
           name, INFLIGHT, warps, grid, bytes / (double)mx, mx, ms, bytes * grid / (ms * 1e-3) / 1e12);
}

int main()
{
    long long* d_clk; uint32_t* d_sink;
    CK(cudaMalloc(&d_clk, 1024 * sizeof(long long)));
    CK(cudaMalloc(&d_sink, 64));
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("SMs %d\n", sms);
    for (int grid : {1, sms}) {
        for (int warps : {4, 8, 16}) {
            run<0, 1>("32x32b.x16", warps, grid, d_clk, d_sink);
            run<0, 4>("32x32b.x16", warps, grid, d_clk, d_sink);
            run<1, 1>("32x32b.x32", warps, grid, d_clk, d_sink);
            run<1, 2>("32x32b.x32", warps, grid, d_clk, d_sink);
            run<2, 1>("32x32b.x64", warps, grid, d_clk, d_sink);
            run<2, 2>("32x32b.x64", warps, grid, d_clk, d_sink);
            run<3, 1>("16x256b.x16", warps, grid, d_clk, d_sink);
        }
    }
    return 0;
}
