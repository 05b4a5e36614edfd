// lof_round_f32.cu -- instantiates the elementwise virtual_round kernels for the f32 container.
#include "lof_round_kernels.cuh"

namespace lof {

cudaError_t launch_round_f32(int kind, int mode, const void* in, void* out, int64_t n, const RoundPlan& plan,
                            const FastPlan& fast, const MagicPlan& magic, int r, int len, double scale, const RngArgs& rng, uint64_t first_index,
                            const RoundLaunchCfg& cfg, cudaStream_t stream)
{
    switch (kind) {
    case kKindFastSat:    return launch_round<LOF_F32, kKindFastSat>(mode, in, out, n, plan, fast, magic, r, len, scale, rng, first_index, cfg, stream);
    case kKindFastExt:    return launch_round<LOF_F32, kKindFastExt>(mode, in, out, n, plan, fast, magic, r, len, scale, rng, first_index, cfg, stream);
    case kKindFastSatChk: return launch_round<LOF_F32, kKindFastSatChk>(mode, in, out, n, plan, fast, magic, r, len, scale, rng, first_index, cfg, stream);
    case kKindFastExtChk: return launch_round<LOF_F32, kKindFastExtChk>(mode, in, out, n, plan, fast, magic, r, len, scale, rng, first_index, cfg, stream);
    case kKindMagicSat:    return launch_round<LOF_F32, kKindMagicSat>(mode, in, out, n, plan, fast, magic, r, len, scale, rng, first_index, cfg, stream);
    case kKindMagicExt:    return launch_round<LOF_F32, kKindMagicExt>(mode, in, out, n, plan, fast, magic, r, len, scale, rng, first_index, cfg, stream);
    case kKindMagicSatChk: return launch_round<LOF_F32, kKindMagicSatChk>(mode, in, out, n, plan, fast, magic, r, len, scale, rng, first_index, cfg, stream);
    case kKindMagicExtChk: return launch_round<LOF_F32, kKindMagicExtChk>(mode, in, out, n, plan, fast, magic, r, len, scale, rng, first_index, cfg, stream);
    default: break;
    }
    if (kind != kKindMantissa)
        return launch_round<LOF_F32, kKindParams>(mode, in, out, n, plan, fast, magic, r, len, scale, rng, first_index, cfg, stream);
    return launch_round<LOF_F32, kKindMantissa>(mode, in, out, n, plan, fast, magic, r, len, scale, rng, first_index, cfg, stream);
}

}  // namespace lof
