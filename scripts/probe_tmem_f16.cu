// probe_tmem_f16.cu -- characterises tcgen05.mma with an FP16 accumulator (instruction-descriptor c_format = F16)
// on B200: (1) where D[m][n] lives in TMEM, (2) how the tensor core rounds the accumulator when it writes it
// (within one MMA and across chained MMAs).  Input for DESIGN.md section 8 (fp16 accumulation inside the tensor
// core would remove the TMEM read-back that bounds the chunked GEMM).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I lofloat_b200/csrc scripts/probe_tmem_f16.cu -lcuda -o scripts/build/probe_tmem_f16.bin
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "lof_ptx.cuh"

using namespace lof::ptx;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

// one CTA, 128 threads: TMA-load A and B tiles [128 rows x 128 bytes], run `nsteps` MMAs (K = 32 each) into TMEM
// column 0.., first with accumulate = 0, then dump 128 lanes x 128 columns of raw TMEM words
__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                                                     uint32_t idesc, int nsteps, uint32_t* out)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sa = smem;
    uint8_t* sb = smem + 128 * 128;
    __shared__ __align__(8) uint64_t full_bar, done_bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(&full_bar, 1); mbar_init(&done_bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(&tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        mbar_expect_tx(&full_bar, 2 * 128 * 128);
        tma_load_2d(sa, &map_a, 0, 0, &full_bar);
        tma_load_2d(sb, &map_b, 0, 0, &full_bar);
        mbar_wait(&full_bar, 0);
        tc_fence_after();
        for (int s = 0; s < nsteps; ++s)
            umma(tmem, umma_desc_k128(smem_u32(sa) + s * 32), umma_desc_k128(smem_u32(sb) + s * 32), idesc, s > 0 ? 1u : 0u, 1);
        umma_commit(&done_bar);
    }
    mbar_wait(&done_bar, 0);
    tc_fence_after();
    uint32_t v[32];
    for (int c = 0; c < 128; c += 32) {
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) out[(warp * 32 + lane) * 128 + c + i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

static uint8_t enc_e4m3(double v)          // exact encoder for representable values (OCP e4m3fn)
{
    if (v == 0) return 0;
    uint8_t s = v < 0 ? 0x80 : 0;
    double a = std::fabs(v);
    int e; double m = std::frexp(a, &e);   // a = m * 2^e, m in [0.5, 1)
    int E = e - 1 + 7;                      // biased exponent for 1.xxx * 2^(e-1)
    if (E >= 1) {
        double frac = m * 2 - 1;            // [0,1)
        int mant = (int)std::lround(frac * 8);
        if (std::fabs(frac * 8 - mant) > 1e-12 || E > 15) { printf("not representable: %g\n", v); exit(1); }
        return s | (uint8_t)(E << 3) | (uint8_t)mant;
    }
    double k = a / std::ldexp(1.0, -9);     // subnormal: k * 2^-9
    int ki = (int)std::lround(k);
    if (std::fabs(k - ki) > 1e-12 || ki > 7) { printf("not representable (subnormal): %g\n", v); exit(1); }
    return s | (uint8_t)ki;
}

static bool make_map(CUtensorMap* map, void* base)
{
    const cuuint64_t dims[2] = {128, 128};
    const cuuint64_t strides[1] = {128};
    const cuuint32_t box[2] = {128, 128};
    const cuuint32_t estr[2] = {1, 1};
    return cuTensorMapEncodeTiled(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static double half_bits_to_double(uint16_t h)
{
    __half_raw r; r.x = h;
    return (double)__half2float(__half(r));
}

struct Scenario { const char* name; double b[8]; };   // beta for k = 0,1,2,3 (step 0) and k = 32,33,34,35 (step 1)

int main()
{
    CK(cudaSetDevice(0));
    CK(cudaFree(0));
    // alpha_k (row-independent): k0 = 1.875, k1 = 2^-9, k2 = 2^-9, k3 = 3 * 2^-9; same in step 1 (k = 32..35)
    const double alpha[4] = {1.875, std::ldexp(1.0, -9), std::ldexp(1.0, -9), std::ldexp(3.0, -9)};
    const double T = 1.875;       // base product 1.875^2 = 3.515625 = 1800 * 2^-9 (fp16 ulp in [2,4) is 2^-9... actually 2^-9 = ulp)
    std::vector<Scenario> sc = {
        {"even base (exact)",                    {T, 0, 0, 0, 0, 0, 0, 0}},
        {"odd base (exact)",                     {T, 1, 0, 0, 0, 0, 0, 0}},
        {"odd + 1/2 ulp   in-MMA  (tie)",        {T, 1, 0.5, 0, 0, 0, 0, 0}},
        {"even + 1/2 ulp  in-MMA  (tie)",        {T, 0, 0.5, 0, 0, 0, 0, 0}},
        {"odd + 1/4 ulp   in-MMA",               {T, 1, 0.25, 0, 0, 0, 0, 0}},
        {"odd + 3/4 ulp   in-MMA",               {T, 1, 0, 0.25, 0, 0, 0, 0}},
        {"-(odd + 1/2 ulp) in-MMA (tie)",        {-T, -1, -0.5, 0, 0, 0, 0, 0}},
        {"-(odd + 1/4 ulp) in-MMA",              {-T, -1, -0.25, 0, 0, 0, 0, 0}},
        {"-(odd + 3/4 ulp) in-MMA",              {-T, -1, 0, -0.25, 0, 0, 0, 0}},
        {"odd, +1/2 ulp   next MMA (tie)",       {T, 1, 0, 0, 0, 0, 0.5, 0}},
        {"even, +1/2 ulp  next MMA (tie)",       {T, 0, 0, 0, 0, 0, 0.5, 0}},
        {"odd, +1/4 ulp   next MMA",             {T, 1, 0, 0, 0, 0, 0.25, 0}},
        {"odd, +3/4 ulp   next MMA",             {T, 1, 0, 0, 0, 0, 0, 0.25}},
        {"-odd, -1/2 ulp  next MMA (tie)",       {-T, -1, 0, 0, 0, 0, -0.5, 0}},
        {"-odd, -1/4 ulp  next MMA",             {-T, -1, 0, 0, 0, 0, -0.25, 0}},
        {"-odd, -3/4 ulp  next MMA",             {-T, -1, 0, 0, 0, 0, 0, -0.25}},
        {"odd, two 1/4 ulp terms next MMA",      {T, 1, 0, 0, 0, 0.25, 0.25, 0}},
        {"two steps of 1.875*448 = 840",        {448, 0, 0, 0, 448, 0, 0, 0}},
    };
    std::vector<uint8_t> hA(128 * 128, 0), hB(128 * 128, 0);
    for (int m = 0; m < 128; ++m)
        for (int s = 0; s < 2; ++s)
            for (int k = 0; k < 4; ++k) hA[m * 128 + 32 * s + k] = enc_e4m3(alpha[k]);
    for (size_t n = 0; n < sc.size(); ++n) {
        for (int j = 0; j < 8; ++j) {
            double beta = sc[n].b[j];
            if (n == sc.size() - 1) {                          // "big": put 240 in A too for this column only -> use beta^2 / alpha trick
                beta = (j == 0 || j == 4) ? 448.0 : 0.0;       // 1.875 * 448 = 840 per step; two steps = 1680 (fits) -- replaced below
            }
            hB[n * 128 + (j < 4 ? j : 32 + (j - 4))] = enc_e4m3(beta);
        }
    }
    // layout probe: columns 64.. carry b[n][0] = 1 + (n % 8) * 0.125 (distinct per column), rows all equal
    for (int n = 64; n < 128; ++n) hB[n * 128 + 0] = enc_e4m3(1.0 + (n % 8) * 0.125);

    uint8_t *dA, *dB; uint32_t* dOut;
    CK(cudaMalloc(&dA, hA.size())); CK(cudaMalloc(&dB, hB.size())); CK(cudaMalloc(&dOut, 128 * 128 * 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice));
    CUtensorMap ma, mb;
    if (!make_map(&ma, dA) || !make_map(&mb, dB)) { printf("tensor map failed\n"); return 1; }
    const int smem = 2 * 128 * 128 + 1024;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    std::vector<uint32_t> out32(128 * 128), out16(128 * 128);
    for (int cfmt = 1; cfmt >= 0; --cfmt) {                    // 1 = F32 accumulator, 0 = F16 accumulator
        const uint32_t idesc = ((uint32_t)cfmt << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        CK(cudaMemset(dOut, 0xCD, 128 * 128 * 4));
        probe_kernel<<<1, 128, smem>>>(ma, mb, idesc, 2, dOut);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy((cfmt ? out32 : out16).data(), dOut, 128 * 128 * 4, cudaMemcpyDeviceToHost));
    }
    printf("== F32 accumulator, lane 0, first scenarios (float):\n");
    for (size_t n = 0; n < sc.size(); ++n) { float f; memcpy(&f, &out32[n], 4); printf("  col %2zu %-38s %.10g\n", n, sc[n].name, f); }
    printf("== F16 accumulator: raw words lane 0, TMEM columns 0..15 and 32..39:\n  ");
    for (int c = 0; c < 16; ++c) printf("%08x ", out16[c]);
    printf("\n  ");
    for (int c = 32; c < 40; ++c) printf("%08x ", out16[c]);
    printf("\n  lane 1, columns 0..7: ");
    for (int c = 0; c < 8; ++c) printf("%08x ", out16[128 + c]);
    printf("\n== layout probe: expected D[.][n] = 1.875 * (1 + (n%%8)/8) for n >= 64 -> which (column, half) holds it?\n");
    for (int n = 64; n < 72; ++n) {
        const double want = 1.875 * (1.0 + (n % 8) * 0.125);
        printf("  n=%d want %.6f :", n, want);
        for (int c = 0; c < 128; ++c)
            for (int h = 0; h < 2; ++h)
                if (half_bits_to_double((uint16_t)(out16[c] >> (16 * h))) == want) printf(" (col %d, half %d)", c, h);
        printf("\n");
    }
    printf("== F16 accumulator scenarios: decoded assuming packed (col n/2, half n%%2) and unpacked (col n, low half):\n");
    for (size_t n = 0; n < sc.size(); ++n) {
        double exact = 0;
        for (int j = 0; j < 8; ++j) exact += alpha[j % 4] * (n == sc.size() - 1 ? ((j == 0 || j == 4) ? 448.0 : 0.0) : sc[n].b[j]);
        const double packed = half_bits_to_double((uint16_t)(out16[n / 2] >> (16 * (n % 2))));
        const double unpacked = half_bits_to_double((uint16_t)(out16[n] & 0xFFFF));
        const double q = std::ldexp(1.0, -9);
        printf("  %-38s exact %.9f (%.3f ulp)  packed %.9f (%.2f)  unpacked %.9f (%.2f)\n", sc[n].name, exact, exact / q,
               packed, packed / q, unpacked, unpacked / q);
    }
    return 0;
}
