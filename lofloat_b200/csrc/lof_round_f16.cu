// lof_round_f16.cu -- instantiates the elementwise virtual_round kernels for the f16 container.
#include "lof_round_kernels.cuh"

namespace lof {

cudaError_t launch_round_f16(int kind, int mode, const void* in, void* out, int64_t n, const RoundPlan& plan,
                            const FastPlan& fast, const MagicPlan& magic, int r, int len, double scale, const RngArgs& rng, uint64_t first_index,
                            const RoundLaunchCfg& cfg, cudaStream_t stream)
{
    // fp32 magic-add routine on the widened values (lof_api.cu decides eligibility): deterministic IEEE-like modes, no fused scale
    if (kind == kKindMagicSat) return launch_round<LOF_F16, kKindMagicSat>(mode, in, out, n, plan, fast, magic, r, len, scale, rng, first_index, cfg, stream);
    if (kind == kKindMagicExt) return launch_round<LOF_F16, kKindMagicExt>(mode, in, out, n, plan, fast, magic, r, len, scale, rng, first_index, cfg, stream);
    if (kind != kKindMantissa)
        return launch_round<LOF_F16, kKindParams>(mode, in, out, n, plan, fast, magic, r, len, scale, rng, first_index, cfg, stream);
    return launch_round<LOF_F16, kKindMantissa>(mode, in, out, n, plan, fast, magic, r, len, scale, rng, first_index, cfg, stream);
}

}  // namespace lof
