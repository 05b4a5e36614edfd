// lof_mx.cu -- launchers for the MX block-rounding kernels (shuffle variant + tail).
#include "lof_mx_kernels.cuh"

namespace lof {

namespace {

template <int BS, int MODE>
cudaError_t launch_bs_mode(const float* in, float* out, int64_t n, const MxArgs& a, const RngArgs& rng,
                           const RngArgs& rng_scale, uint64_t first_index, const MxLaunchCfg& cfg,
                           cudaStream_t stream, int* launches)
{
    constexpr int ROWS = 4;
    constexpr int64_t UNIT = (BS > 128) ? BS : 128;             // whole rows AND whole blocks
    const bool aligned = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0;
    const bool rng_ok = !ModeTraits<MODE>::kUsesRng || (first_index % 4) == 0;
    int64_t main_elems = (aligned && rng_ok) ? (n / UNIT) * UNIT : 0;
    if (main_elems > 0) {
        const int64_t nrows = main_elems / 128;
        const int64_t nsteps = (nrows + ROWS - 1) / ROWS;
        constexpr int kWarps = kMxThreads / 32;
        int64_t grid = (nsteps + kWarps - 1) / kWarps;
        const int64_t max_ctas = (int64_t)cfg.num_sms * cfg.ctas_per_sm;
        if (grid > max_ctas) grid = max_ctas;
        mx_shuffle_kernel<BS, MODE, ROWS><<<(unsigned)grid, kMxThreads, 0, stream>>>(
            reinterpret_cast<const float4*>(in), reinterpret_cast<float4*>(out), nrows, a, rng, rng_scale, first_index);
        ++*launches;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    if (main_elems < n) {
        const int64_t nblocks = (n - main_elems + BS - 1) / BS;
        int64_t grid = (nblocks + kMxThreads - 1) / kMxThreads;
        const int64_t max_ctas = (int64_t)cfg.num_sms * 8;
        if (grid > max_ctas) grid = max_ctas;
        mx_block_per_thread_kernel<MODE><<<(unsigned)grid, kMxThreads, 0, stream>>>(in, out, main_elems, n, a, rng,
                                                                                   rng_scale, first_index);
        ++*launches;
        return cudaGetLastError();
    }
    return cudaSuccess;
}

template <int MODE>
cudaError_t launch_mode(const float* in, float* out, int64_t n, const MxArgs& a, const RngArgs& rng,
                        const RngArgs& rng_scale, uint64_t first_index, const MxLaunchCfg& cfg, cudaStream_t stream,
                        int* launches)
{
    switch (a.block_size) {
    case 16:  return launch_bs_mode<16, MODE>(in, out, n, a, rng, rng_scale, first_index, cfg, stream, launches);
    case 32:  return launch_bs_mode<32, MODE>(in, out, n, a, rng, rng_scale, first_index, cfg, stream, launches);
    case 64:  return launch_bs_mode<64, MODE>(in, out, n, a, rng, rng_scale, first_index, cfg, stream, launches);
    case 128: return launch_bs_mode<128, MODE>(in, out, n, a, rng, rng_scale, first_index, cfg, stream, launches);
    case 256: return launch_bs_mode<256, MODE>(in, out, n, a, rng, rng_scale, first_index, cfg, stream, launches);
    default:  return cudaErrorInvalidValue;
    }
}

}  // namespace

cudaError_t launch_mx_shuffle(int mode, const float* in, float* out, int64_t n, const MxArgs& a, const RngArgs& rng,
                              const RngArgs& rng_scale, uint64_t first_index, const MxLaunchCfg& cfg,
                              cudaStream_t stream, int* launches)
{
    switch (mode) {
#define LOF_CASE(M) case M: return launch_mode<M>(in, out, n, a, rng, rng_scale, first_index, cfg, stream, launches);
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
