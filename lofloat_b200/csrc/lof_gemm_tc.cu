// lof_gemm_tc.cu -- chunked tensor-core (tcgen05) LoFloat-accumulation GEMM (placeholder).
#include <string>

#include "lof_core.cuh"

namespace lof {

cudaError_t launch_gemm_tc(const float*, const float*, float*, int64_t, int64_t, int64_t, int, int, const RoundPlan*,
                           int64_t, float, int, cudaStream_t, int*, std::string* err)
{
    *err = "chunked tensor-core GEMM is not built yet";
    return cudaSuccess;
}

}  // namespace lof
