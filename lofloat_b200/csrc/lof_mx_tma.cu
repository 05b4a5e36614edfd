// lof_mx_tma.cu -- TMA-staged MX block-rounding kernel (sm_100a), fp32.
//
// Same arithmetic as lof_mx_kernels.cuh (reference: mx_round.cuh:27-214, CPU definition
// LoPy_bind.cpp:197-215), different data movement:
//
//   HBM --TMA (cp.async.bulk.tensor.2d, 128B swizzle)--> smem tile [256 rows x 32 floats]
//   one THREAD per 32-float row: 8 conflict-free LDS.128, the whole MX block lives in
//   registers (no shuffles for BS <= 32, log2(BS/32) shuffles above), the block scale is
//   computed once per block instead of once per lane, results are written back into the
//   same smem row and leave through a TMA store.
//
// All global traffic is issued by the TMA unit (one elected thread): full 128-byte lines,
// kStages tiles (kStages x 32 KB) in flight per CTA, no LSU address arithmetic per
// element.  Loads complete on per-stage mbarriers; a stage is refilled once the bulk store
// that drains it has finished reading shared memory (cp.async.bulk.wait_group.read).
#include <cuda.h>

#include "lof_mx_kernels.cuh"

namespace lof {

namespace {

constexpr int kTileRows = 256;                         // TMA box: 256 rows x 32 floats = 32 KB
constexpr int kTileBytes = kTileRows * 128;
constexpr int kTmaThreads = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int c0, int c1, const void* smem_src)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                 ::"l"(map), "r"(c0), "r"(c1), "r"(smem_u32(smem_src)) : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int BS, int MODE, int STAGES>
__global__ void __launch_bounds__(kTmaThreads)
mx_tma_kernel(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_out,
              int64_t ntiles, MxArgs a, RngArgs rng, RngArgs rng_scale, uint64_t first_index)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte alignment for the 128B swizzle atom (8 rows x 128 B)
    uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) uint64_t full_bar[STAGES];

    const int t = threadIdx.x;
    if (t == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(&full_bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_in) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_out) : "memory");
    }
    __syncthreads();

    const int64_t stride = gridDim.x;
    const int64_t my_first = blockIdx.x;
    const int64_t my_count = (my_first < ntiles) ? (ntiles - my_first + stride - 1) / stride : 0;

    if (t == 0) {                                       // prologue: fill the ring
        for (int s = 0; s < STAGES && s < my_count; ++s) {
            mbar_expect_tx(&full_bar[s], kTileBytes);
            tma_load_2d(tiles + (size_t)s * kTileBytes, &map_in, 0, (int)((my_first + s * stride) * kTileRows), &full_bar[s]);
        }
    }

    const int swz = t & 7;                              // 128B swizzle: chunk c of row r sits at chunk c ^ (r & 7)
    for (int64_t it = 0; it < my_count; ++it) {
        const int stage = (int)(it % STAGES);
        const uint32_t parity = (uint32_t)((it / STAGES) & 1);
        const int64_t tile = my_first + it * stride;
        uint8_t* row = tiles + (size_t)stage * kTileBytes + (size_t)t * 128;

        mbar_wait(&full_bar[stage], parity);

        float4 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = *reinterpret_cast<const float4*>(row + ((j ^ swz) << 4));

        const uint64_t g0 = first_index + ((uint64_t)tile * kTileRows + (uint64_t)t) * 32u;   // global index of this row
        if constexpr (BS == 16) {
            float m0 = fmaxf(fmaxf(amax4(v[0]), amax4(v[1])), fmaxf(amax4(v[2]), amax4(v[3])));
            float m1 = fmaxf(fmaxf(amax4(v[4]), amax4(v[5])), fmaxf(amax4(v[6]), amax4(v[7])));
            const MxBlockCtx c0 = mx_make_ctx(mx_block_scale(m0, a, rng_scale, g0 / 16u));
            const MxBlockCtx c1 = mx_make_ctx(mx_block_scale(m1, a, rng_scale, g0 / 16u + 1u));
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const MxBlockCtx& c = (j < 4) ? c0 : c1;
                uint4 w = make_uint4(0, 0, 0, 0);
                if constexpr (ModeTraits<MODE>::kUsesRng) w = philox_group4(rng, (g0 >> 2) + j);
                v[j].x = mx_quant<MODE>(v[j].x, c, a, w.x);
                v[j].y = mx_quant<MODE>(v[j].y, c, a, w.y);
                v[j].z = mx_quant<MODE>(v[j].z, c, a, w.z);
                v[j].w = mx_quant<MODE>(v[j].w, c, a, w.w);
            }
        } else {
            float m = fmaxf(fmaxf(fmaxf(amax4(v[0]), amax4(v[1])), fmaxf(amax4(v[2]), amax4(v[3]))),
                            fmaxf(fmaxf(amax4(v[4]), amax4(v[5])), fmaxf(amax4(v[6]), amax4(v[7]))));
            constexpr int GROUP = BS / 32;              // consecutive rows (threads) of one MX block
#pragma unroll
            for (int off = GROUP / 2; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, off));
            const MxBlockCtx c = mx_make_ctx(mx_block_scale(m, a, rng_scale, g0 / (uint64_t)BS));
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                uint4 w = make_uint4(0, 0, 0, 0);
                if constexpr (ModeTraits<MODE>::kUsesRng) w = philox_group4(rng, (g0 >> 2) + j);
                v[j].x = mx_quant<MODE>(v[j].x, c, a, w.x);
                v[j].y = mx_quant<MODE>(v[j].y, c, a, w.y);
                v[j].z = mx_quant<MODE>(v[j].z, c, a, w.z);
                v[j].w = mx_quant<MODE>(v[j].w, c, a, w.w);
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(row + ((j ^ swz) << 4)) = v[j];

        fence_async_smem();                             // generic-proxy writes -> visible to the TMA store
        __syncthreads();
        if (t == 0) {
            tma_store_2d(&map_out, 0, (int)(tile * kTileRows), tiles + (size_t)stage * kTileBytes);
            tma_commit();
            // refill the stage drained by the PREVIOUS store (its smem reads must be done)
            if (it >= 1) {
                tma_wait_read<1>();
                const int64_t nxt = it - 1 + STAGES;
                if (nxt < my_count) {
                    const int ps = (int)((it - 1) % STAGES);
                    mbar_expect_tx(&full_bar[ps], kTileBytes);
                    tma_load_2d(tiles + (size_t)ps * kTileBytes, &map_in, 0, (int)((my_first + nxt * stride) * kTileRows),
                                &full_bar[ps]);
                }
            }
        }
    }
    if (t == 0) tma_wait_all();
}

using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn get_encode()
{
    static EncodeFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeFn>(p);
    }
    return fn;
}

bool make_row_map(CUtensorMap* map, const float* base, int64_t rows)
{
    EncodeFn enc = get_encode();
    if (!enc) return false;
    const cuuint64_t dims[2] = {32, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {128};                 // bytes between rows
    const cuuint32_t box[2] = {32, (cuuint32_t)kTileRows};
    const cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int MODE>
cudaError_t launch_tail(const float* in, float* out, int64_t elem0, int64_t n, const MxArgs& a, const RngArgs& rng,
                        const RngArgs& rng_scale, uint64_t first_index, int num_sms, cudaStream_t stream, int* launches)
{
    const int64_t nblocks = (n - elem0 + a.block_size - 1) / a.block_size;
    int64_t grid = (nblocks + kMxThreads - 1) / kMxThreads;
    if (grid > (int64_t)num_sms * 8) grid = (int64_t)num_sms * 8;
    mx_block_per_thread_kernel<MODE><<<(unsigned)grid, kMxThreads, 0, stream>>>(in, out, elem0, n, a, rng, rng_scale,
                                                                               first_index);
    ++*launches;
    return cudaGetLastError();
}

template <int BS, int MODE>
cudaError_t launch_bs_mode(const float* in, float* out, int64_t n, const MxArgs& a, const RngArgs& rng,
                           const RngArgs& rng_scale, uint64_t first_index, const MxLaunchCfg& cfg, cudaStream_t stream,
                           int* launches, bool* handled)
{
    constexpr int STAGES = 3;
    constexpr int64_t UNIT = BS > 32 ? BS : 32;
    const bool aligned = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0;
    const bool rng_ok = !ModeTraits<MODE>::kUsesRng || (first_index % 4) == 0;
    const int64_t main_elems = (n / UNIT) * UNIT;
    const int64_t rows = main_elems / 32;
    if (!aligned || !rng_ok || rows < kTileRows || rows > 0x7FFFFF00ll) return cudaSuccess;   // not handled: generic path
    CUtensorMap map_in, map_out;
    if (!make_row_map(&map_in, in, rows) || !make_row_map(&map_out, out, rows)) return cudaSuccess;
    const int64_t ntiles = (rows + kTileRows - 1) / kTileRows;
    const size_t smem = (size_t)STAGES * kTileBytes + 1024;
    auto kern = mx_tma_kernel<BS, MODE, STAGES>;
    {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    int64_t grid = ntiles;
    const int ctas_per_sm = cfg.ctas_per_sm > 2 ? 2 : cfg.ctas_per_sm;       // 2 x 97 KB fits in 227 KB
    if (grid > (int64_t)cfg.num_sms * ctas_per_sm) grid = (int64_t)cfg.num_sms * ctas_per_sm;
    kern<<<(unsigned)grid, kTmaThreads, smem, stream>>>(map_in, map_out, ntiles, a, rng, rng_scale, first_index);
    ++*launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    *handled = true;
    if (main_elems < n) return launch_tail<MODE>(in, out, main_elems, n, a, rng, rng_scale, first_index, cfg.num_sms, stream, launches);
    return cudaSuccess;
}

template <int MODE>
cudaError_t launch_mode(const float* in, float* out, int64_t n, const MxArgs& a, const RngArgs& rng,
                        const RngArgs& rng_scale, uint64_t first_index, const MxLaunchCfg& cfg, cudaStream_t stream,
                        int* launches, bool* handled)
{
    switch (a.block_size) {
    case 16:  return launch_bs_mode<16, MODE>(in, out, n, a, rng, rng_scale, first_index, cfg, stream, launches, handled);
    case 32:  return launch_bs_mode<32, MODE>(in, out, n, a, rng, rng_scale, first_index, cfg, stream, launches, handled);
    case 64:  return launch_bs_mode<64, MODE>(in, out, n, a, rng, rng_scale, first_index, cfg, stream, launches, handled);
    case 128: return launch_bs_mode<128, MODE>(in, out, n, a, rng, rng_scale, first_index, cfg, stream, launches, handled);
    case 256: return launch_bs_mode<256, MODE>(in, out, n, a, rng, rng_scale, first_index, cfg, stream, launches, handled);
    default:  return cudaErrorInvalidValue;
    }
}

}  // namespace

// *handled = false means the shape/alignment is not eligible and the caller must use the
// generic (shuffle) path.
cudaError_t launch_mx_tma(int mode, const float* in, float* out, int64_t n, const MxArgs& a, const RngArgs& rng,
                          const RngArgs& rng_scale, uint64_t first_index, const MxLaunchCfg& cfg, cudaStream_t stream,
                          int* launches, bool* handled)
{
    *handled = false;
    switch (mode) {
#define LOF_CASE(M) case M: return launch_mode<M>(in, out, n, a, rng, rng_scale, first_index, cfg, stream, launches, handled);
        LOF_CASE(LOF_RoundToNearestEven) LOF_CASE(LOF_RoundTowardsZero) LOF_CASE(LOF_RoundAwayFromZero)
        LOF_CASE(LOF_StochasticRoundingA) LOF_CASE(LOF_RoundToNearestOdd) LOF_CASE(LOF_RoundDown)
        LOF_CASE(LOF_RoundUp) LOF_CASE(LOF_RoundTiesToAway) LOF_CASE(LOF_StochasticRoundingB)
        LOF_CASE(LOF_StochasticRoundingC) LOF_CASE(LOF_True_StochasticRounding)
        LOF_CASE(LOF_ProbabilisticRounding) LOF_CASE(LOF_RoundToOdd) LOF_CASE(LOF_RoundTiesTowardsZero)
        LOF_CASE(LOF_StochasticRoundingD)
#undef LOF_CASE
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace lof
