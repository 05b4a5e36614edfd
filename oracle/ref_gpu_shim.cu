// ref_gpu_shim.cu -- the reference's OWN device code as a same-box GPU baseline.
//
// TEST / BENCH INFRASTRUCTURE ONLY.  Compiled by `make -C oracle refgpu` from the sources where they lie under /root/reference
// (include paths only -- nothing is copied) into oracle/_ref/libref_gpu.so for sm_100a.  The reference's kernel TU
// (src/Lof_kernel.cu) cannot be built at the pinned submodule commits (its CUTLASS GEMM does not compile),
// but the two rounding paths can: the element function lo_float::virtual_round (lo_float.h:2843-2939, device-callable with
// -DUSE_CUDA) and the MX kernel template lo_float::virtual_mx_round (src/mx_round.cuh:27-214, header-only).  This shim wraps
// them in the launch shapes the reference uses:
//   * elementwise: one thread per element, 256-thread blocks, in place, parameter structs by value -- the launch of
//     round_fp_params<T> (src/Lof_kernel.cu:49-60; that 5-line kernel is restated here because its TU is unbuildable);
//   * MX: virtual_mx_round<float, 256, MX, NX><<<min(65535, n / block), 256>>> -- the dispatch of run_virtual_mx_round
//     (src/Lof_kernel.cu:70-94), kernel body unmodified from mx_round.cuh.
// bench.py times these next to lofloat_b200's kernels on the same GPU (`other.reference_gpu_*`); tests never need it.
#include <cuda_runtime.h>

#include <cstdint>

#include "lo_float.h"
#include "mx_round.cuh"

using namespace lo_float;

namespace {
using Params = FloatingPointParams<DeviceInfChecker, DeviceNaNChecker>;

Params make_params(int bw, int mb, int bias, int ov, int na, int sgn)
{
    return Params(bw, mb, bias, static_cast<Inf_Behaviors>(ov), static_cast<NaN_Behaviors>(na), static_cast<Signedness>(sgn),
                  DeviceInfChecker{0, 0, 0, 0}, DeviceNaNChecker{0, 0, 0, 0}, Unsigned_behavior::NegtoZero);
}

// the reference's elementwise launch shape (src/Lof_kernel.cu:50-56): io[i] = virtual_round(scale * io[i], params, ps)
__global__ void ref_round_fp_params_kernel(float* inout, int64_t n, Params params, ProjSpec ps, float scale)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < n) {
        const float v = (scale == 1.0f) ? inout[idx] : scale * inout[idx];
        inout[idx] = lo_float::virtual_round(v, params, ps);
    }
}
}  // namespace

extern "C" {

int refgpu_round_params_f32(float* dev_inout, int64_t n, int bw, int mb, int bias, int ov, int na, int sgn, int mode, int stoch_len,
                            float scale, void* stream)
{
    const Params p = make_params(bw, mb, bias, ov, na, sgn);
    const ProjSpec ps{static_cast<Rounding_Mode>(mode), Saturation_Mode::OvfInf, stoch_len};
    ref_round_fp_params_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(dev_inout, n, p, ps, scale);
    return (int)cudaGetLastError();
}

int refgpu_mx_round_f32(float* dev_inout, int64_t n, int block_size, int s_bw, int s_mb, int s_bias, int s_ov, int s_sgn, int e_bw,
                        int e_mb, int e_bias, int e_ov, int e_sgn, int mode, int scale_mode, void* stream)
{
    const Params pub = make_params(s_bw, s_mb, s_bias, s_ov, 0, s_sgn), priv = make_params(e_bw, e_mb, e_bias, e_ov, 0, e_sgn);
    const ProjSpec ps_pub{static_cast<Rounding_Mode>(scale_mode), Saturation_Mode::OvfInf, 0};
    const ProjSpec ps_priv{static_cast<Rounding_Mode>(mode), Saturation_Mode::OvfInf, 0};
    constexpr int NTPC = 256;
    const int64_t nb = (n + block_size - 1) / block_size;
    const int grid = (int)(nb < 1 ? 1 : (nb > 65535 ? 65535 : nb));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    switch (block_size) {           // the reference's dispatch table (Lof_kernel.cu:83-92)
    case 16:  lo_float::virtual_mx_round<float, NTPC, 1, 16><<<grid, NTPC, 0, s>>>(dev_inout, n, nullptr, pub, priv, ps_pub, ps_priv); break;
    case 32:  lo_float::virtual_mx_round<float, NTPC, 1, 32><<<grid, NTPC, 0, s>>>(dev_inout, n, nullptr, pub, priv, ps_pub, ps_priv); break;
    case 64:  lo_float::virtual_mx_round<float, NTPC, 2, 32><<<grid, NTPC, 0, s>>>(dev_inout, n, nullptr, pub, priv, ps_pub, ps_priv); break;
    case 128: lo_float::virtual_mx_round<float, NTPC, 4, 32><<<grid, NTPC, 0, s>>>(dev_inout, n, nullptr, pub, priv, ps_pub, ps_priv); break;
    case 256: lo_float::virtual_mx_round<float, NTPC, 8, 32><<<grid, NTPC, 0, s>>>(dev_inout, n, nullptr, pub, priv, ps_pub, ps_priv); break;
    default: return -1;
    }
    return (int)cudaGetLastError();
}

}  // extern "C"
