// lof_round_bf16.cu -- instantiates the elementwise virtual_round kernels for the bf16 container.
#include "lof_round_kernels.cuh"

namespace lof {

cudaError_t launch_round_bf16(int kind, int mode, const void* in, void* out, int64_t n, const RoundPlan& plan,
                            const FastPlan& fast, const MagicPlan& magic, int r, int len, double scale, const RngArgs& rng, uint64_t first_index,
                            const RoundLaunchCfg& cfg, cudaStream_t stream)
{
    // fp32 branch-free routines on (bits << 16): deterministic modes, no fused scale (lof_api.cu decides)
    switch (kind) {
    case kKindFastSatChk:  return launch_round<LOF_BF16, kKindFastSatChk>(mode, in, out, n, plan, fast, magic, r, len, scale, rng, first_index, cfg, stream);
    case kKindFastExtChk:  return launch_round<LOF_BF16, kKindFastExtChk>(mode, in, out, n, plan, fast, magic, r, len, scale, rng, first_index, cfg, stream);
    case kKindMagicSat:    return launch_round<LOF_BF16, kKindMagicSat>(mode, in, out, n, plan, fast, magic, r, len, scale, rng, first_index, cfg, stream);
    case kKindMagicExt:    return launch_round<LOF_BF16, kKindMagicExt>(mode, in, out, n, plan, fast, magic, r, len, scale, rng, first_index, cfg, stream);
    case kKindMagicSatChk: return launch_round<LOF_BF16, kKindMagicSatChk>(mode, in, out, n, plan, fast, magic, r, len, scale, rng, first_index, cfg, stream);
    case kKindMagicExtChk: return launch_round<LOF_BF16, kKindMagicExtChk>(mode, in, out, n, plan, fast, magic, r, len, scale, rng, first_index, cfg, stream);
    default: break;
    }
    if (kind != kKindMantissa)
        return launch_round<LOF_BF16, kKindParams>(mode, in, out, n, plan, fast, magic, r, len, scale, rng, first_index, cfg, stream);
    return launch_round<LOF_BF16, kKindMantissa>(mode, in, out, n, plan, fast, magic, r, len, scale, rng, first_index, cfg, stream);
}

}  // namespace lof
