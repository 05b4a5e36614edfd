// lof_round_f64.cu -- instantiates the elementwise virtual_round kernels for the f64 container.
#include "lof_round_kernels.cuh"

namespace lof {

cudaError_t launch_round_f64(bool params, int mode, const void* in, void* out, int64_t n, const RoundPlan& plan,
                            int r, int len, double scale, const RngArgs& rng, uint64_t first_index,
                            const RoundLaunchCfg& cfg, cudaStream_t stream)
{
    if (params) return launch_round<LOF_F64, true>(mode, in, out, n, plan, r, len, scale, rng, first_index, cfg, stream);
    return launch_round<LOF_F64, false>(mode, in, out, n, plan, r, len, scale, rng, first_index, cfg, stream);
}

}  // namespace lof
