// lof_gemm_mx_gen.cu -- generic-fold instantiations of the MX GEMM kernel (lof_gemm_mx.cuh): 9 deterministic modes x
// {mantissa-only, full-format accumulator} x {single rounding, partial rounding}.
#include "lof_gemm_mx.cuh"

namespace lof {

namespace {

template <int EPI, bool PARTIAL>
cudaError_t x_launch_mode(int mode, const CUtensorMap& ma, const CUtensorMap& mb, float* C, const MxGemmArgs& a, int grid, cudaStream_t stream)
{
    switch (mode) {
#define LOF_CASE(Mo) case Mo: return x_launch_one<Mo, EPI, PARTIAL>(ma, mb, C, a, grid, stream);
        LOF_CASE(LOF_RoundToNearestEven) LOF_CASE(LOF_RoundTowardsZero) LOF_CASE(LOF_RoundAwayFromZero)
        LOF_CASE(LOF_RoundToNearestOdd) LOF_CASE(LOF_RoundDown) LOF_CASE(LOF_RoundUp) LOF_CASE(LOF_RoundTiesToAway)
        LOF_CASE(LOF_RoundToOdd) LOF_CASE(LOF_RoundTiesTowardsZero)
#undef LOF_CASE
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace

cudaError_t launch_gemm_mx_generic(int mode, bool fmt, bool partial, const CUtensorMap& ma, const CUtensorMap& mb, float* C, const void* args,
                                   int grid, cudaStream_t stream)
{
    const MxGemmArgs& a = *static_cast<const MxGemmArgs*>(args);
    if (fmt) return partial ? x_launch_mode<kXFmt, true>(mode, ma, mb, C, a, grid, stream) : x_launch_mode<kXFmt, false>(mode, ma, mb, C, a, grid, stream);
    return partial ? x_launch_mode<kXMant, true>(mode, ma, mb, C, a, grid, stream) : x_launch_mode<kXMant, false>(mode, ma, mb, C, a, grid, stream);
}

}  // namespace lof
