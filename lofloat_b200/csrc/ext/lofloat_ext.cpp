// lofloat_ext.cpp -- the compiled PyTorch extension of lofloat-b200: a pybind11 module with the SAME Python surface as
// the reference's `LoFloat.LoFloat` (/root/reference/src/LoPy_bind.cpp:281-393: enums, FloatFormatDescriptor,
// virtual_round x 2 overloads, virtual_mx_round, lof_gemm, lof_conv2d, pwl_silu), so that the reference's unmodified
// LoFloat/{__init__,_custom_ops,formats,layers,utils}.py run on top of it.  Every op is a thin call into the C ABI
// (include/lofloat_b200.h) on the CURRENT torch stream; torch supplies tensors and streams only.  No CPU compute: host
// tensors go through the library's staged host-buffer entry points, a missing GPU is an error.
//
// Deliberate differences from the reference binding: kernels run on the current stream instead
// of the legacy default stream (Lof_kernel.cu:47,59,80); no clone() + in-place launch (LoPy_bind.cpp:134-135) but one
// out-of-place launch; bf16 tensors are accepted; stochastic modes draw from a Philox stream keyed by manual_seed().
#include <torch/extension.h>

#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include <atomic>
#include <cstdint>
#include <stdexcept>
#include <string>

#include "lofloat_b200.h"

namespace py = pybind11;

namespace {

// ------------------------------------------------------------------------------------------------ enums
// Same enumerator names and integer values as the reference's C++ enums (fp_tools.hpp:13-50), which is what travels
// through torch.ops.lofloat.* as plain ints (LoFloat/_custom_ops.py:34-108).
enum class RoundingMode : int {
    RoundToNearestEven = LOF_RoundToNearestEven, RoundTowardsZero = LOF_RoundTowardsZero, RoundAwayFromZero = LOF_RoundAwayFromZero,
    StochasticRoundingA = LOF_StochasticRoundingA, RoundToNearestOdd = LOF_RoundToNearestOdd, RoundDown = LOF_RoundDown,
    RoundUp = LOF_RoundUp, RoundTiesToAway = LOF_RoundTiesToAway, StochasticRoundingB = LOF_StochasticRoundingB,
    StochasticRoundingC = LOF_StochasticRoundingC, True_StochasticRounding = LOF_True_StochasticRounding,
    ProbabilisticRounding = LOF_ProbabilisticRounding, RoundToOdd = LOF_RoundToOdd, RoundTiesTowardsZero = LOF_RoundTiesTowardsZero,
    StochasticRoundingD = LOF_StochasticRoundingD
};
enum class Signedness : int { Signed = LOF_Signed, Unsigned = LOF_Unsigned };
enum class InfBehavior : int { Extended = LOF_Extended, Saturating = LOF_Saturating };
enum class NaNBehavior : int { _3109 = LOF_NaN_3109, _754 = LOF_NaN_754 };

bool is_stochastic(RoundingMode m)
{
    switch (m) {
    case RoundingMode::StochasticRoundingA: case RoundingMode::StochasticRoundingB: case RoundingMode::StochasticRoundingC:
    case RoundingMode::StochasticRoundingD: case RoundingMode::True_StochasticRounding: case RoundingMode::ProbabilisticRounding:
        return true;
    default: return false;
    }
}

// ------------------------------------------------------------------------------------------------ format descriptor
// Python-visible twin of FloatingPointParamsPy (LoPy_bind.cpp:45-68): plain fields + the two checker objects, which the
// virtual-rounding path never evaluates but which formats.py constructs and layers.py copies around.
struct FormatDescriptor {
    int total_bits, mantissa_bits, bias;
    InfBehavior inf_behavior;
    NaNBehavior nan_behavior;
    Signedness signedness;
    py::object is_inf_checker, is_nan_checker;

    FormatDescriptor(int tb, int mb, int b, InfBehavior ib, NaNBehavior nb, Signedness sg, py::object inf_c, py::object nan_c)
        : total_bits(tb), mantissa_bits(mb), bias(b), inf_behavior(ib), nan_behavior(nb), signedness(sg),
          is_inf_checker(std::move(inf_c)), is_nan_checker(std::move(nan_c))
    {
        // same acceptance rules and exception type (std::invalid_argument -> ValueError) as LoPy_bind.cpp:53-62
        auto need = [](const py::object& o, const char* attr, const char* msg) {
            if (!py::hasattr(o, attr)) throw std::invalid_argument(msg);
        };
        need(is_inf_checker, "__call__", "inf_checker must be callable");
        need(is_inf_checker, "minNegInf", "inf_checker must have minNegInf()");
        need(is_inf_checker, "minPosInf", "inf_checker must have minPosInf()");
        need(is_nan_checker, "__call__", "nan_checker must be callable");
        if (!py::hasattr(is_nan_checker, "qNanBitPattern") && !py::hasattr(is_nan_checker, "sNanBitPattern"))
            throw std::invalid_argument("nan_checker must have qNanBitPattern() or sNanBitPattern()");
    }
    lof_format_t c() const
    {
        lof_format_t f;
        f.bitwidth = total_bits; f.mantissa_bits = mantissa_bits; f.bias = bias;
        f.ov_behavior = (int)inf_behavior; f.na_behavior = (int)nan_behavior; f.is_signed = (int)signedness;
        f.unsigned_behavior = LOF_NegtoZero;
        return f;
    }
};

// ------------------------------------------------------------------------------------------------ helpers
#define LOFX_CHECK(call, what) do { const int st_ = (call); TORCH_CHECK(st_ == LOF_OK, what, ": ", lof_status_string(st_), " -- ", lof_last_error_string()); } while (0)

int dtype_code(const torch::Tensor& t, const char* what)
{
    switch (t.scalar_type()) {
    case torch::kFloat: return LOF_F32;
    case torch::kHalf: return LOF_F16;
    case torch::kBFloat16: return LOF_BF16;
    case torch::kDouble: return LOF_F64;
    default: TORCH_CHECK(false, "\"", what, "\" not implemented for '", toString(t.scalar_type()), "'");
    }
    return -1;
}

// Philox stream of the stochastic modes: manual_seed() rewinds it, every stochastic call advances the offset (the
// reference's thread_local lof_seed, lo_float.h:58, is not bound to Python and repeats the same draw per call).
std::atomic<uint64_t> g_seed{12345}, g_offset{0};
lof_rng_t next_rng(bool stochastic)
{
    lof_rng_t r;
    r.seed = g_seed.load();
    r.offset = stochastic ? g_offset.fetch_add(1) : g_offset.load();
    return r;
}

void* stream_of(const torch::Tensor& t) { return at::cuda::getCurrentCUDAStream(t.device().index()).stream(); }

// GEMM policy behind the reference's lof_gemm(A, B, accum_mant_bits, ...) entry point.  Default: the reference semantics,
// one accumulator rounded after EVERY FMA over all of K (the CUDA leaf, cutlass_gemms.cuh:97-168) on the exact SIMT kernel.
// set_gemm_policy(k_chunk=32, ...) routes the same call to the tcgen05 chunked kernels (accumulator re-rounded per k-chunk).
struct GemmPolicy { int64_t kc = 0; int64_t k_chunk = 0; int operand_dtype = LOF_E4M3; bool partial_round = false; };
GemmPolicy g_gemm;

// ------------------------------------------------------------------------------------------------ ops
torch::Tensor round_common(const torch::Tensor& input, const FormatDescriptor* fmt, int mant_bits, RoundingMode mode, int stoch_len,
                           double scale)
{
    const char* what = fmt ? "virtual_round_params" : "virtual_round_mantissa";
    const int code = dtype_code(input, what);
    auto x = input.contiguous();
    auto out = torch::empty_like(x);
    lof_projspec_t ps = {(int)mode, LOF_OvfInf, stoch_len};
    const lof_rng_t rng = next_rng(is_stochastic(mode));
    lof_format_t f = {};
    if (fmt) f = fmt->c();
    if (x.is_cuda()) {
        c10::cuda::CUDAGuard guard(x.device());
        if (fmt) LOFX_CHECK(lof_round_params(x.data_ptr(), out.data_ptr(), x.numel(), code, &f, &ps, scale, &rng, 0, stream_of(x)), what);
        else LOFX_CHECK(lof_round_mantissa(x.data_ptr(), out.data_ptr(), x.numel(), code, mant_bits, &ps, scale, &rng, 0, stream_of(x)), what);
    } else {
        py::gil_scoped_release nogil;     // staged through the GPU (H2D, kernel, D2H pipeline); blocks until out is complete
        if (fmt) LOFX_CHECK(lof_round_params_host(x.data_ptr(), out.data_ptr(), x.numel(), code, &f, &ps, scale, &rng, 0, -1), what);
        else LOFX_CHECK(lof_round_mantissa_host(x.data_ptr(), out.data_ptr(), x.numel(), code, mant_bits, &ps, scale, &rng, 0, -1), what);
    }
    return out;
}

torch::Tensor virtual_round_mantissa(const torch::Tensor& input, int to_mantissa_bits, RoundingMode round_mode, int stoch_len, double scale)
{
    return round_common(input, nullptr, to_mantissa_bits, round_mode, stoch_len, scale);
}
torch::Tensor virtual_round_params(const torch::Tensor& input, const FormatDescriptor& params, RoundingMode round_mode, int stoch_len, double scale)
{
    return round_common(input, &params, 0, round_mode, stoch_len, scale);
}

torch::Tensor virtual_mx_round(const torch::Tensor& input, int block_size, const FormatDescriptor& element_format,
                               const FormatDescriptor& scale_format, RoundingMode round_mode, RoundingMode scale_round_mode, int stoch_len)
{
    TORCH_CHECK(input.scalar_type() == torch::kFloat, "virtual_mx_round: float32 tensors only");       // LoPy_bind.cpp:172
    TORCH_CHECK(block_size > 0, "virtual_mx_round: block_size must be positive");                      // :173
    auto x = input.contiguous();
    auto out = torch::empty_like(x);
    const lof_format_t fe = element_format.c(), fs = scale_format.c();
    const lof_projspec_t pe = {(int)round_mode, LOF_OvfInf, stoch_len}, psc = {(int)scale_round_mode, LOF_OvfInf, 0};   // :192
    const lof_rng_t rng = next_rng(is_stochastic(round_mode) || is_stochastic(scale_round_mode));
    if (x.is_cuda()) {
        c10::cuda::CUDAGuard guard(x.device());
        LOFX_CHECK(lof_mx_round_f32(x.data_ptr<float>(), out.data_ptr<float>(), x.numel(), block_size, &fs, &fe, &psc, &pe, &rng, 0,
                                    stream_of(x)), "virtual_mx_round");
    } else {
        py::gil_scoped_release nogil;
        LOFX_CHECK(lof_mx_round_f32_host(x.data_ptr<float>(), out.data_ptr<float>(), x.numel(), block_size, &fs, &fe, &psc, &pe, &rng, 0, -1),
                   "virtual_mx_round");
    }
    return out.view_as(input);
}

torch::Tensor lof_gemm(const torch::Tensor& A, const torch::Tensor& B, int accum_mant_bits, RoundingMode round_mode,
                       int stochastic_rounding_bits, double scale_a, double scale_b)
{
    // argument checks and messages of Lof_kernel.cu:106-113
    TORCH_CHECK(A.dim() == 2 && B.dim() == 2, "LoF_gemm: A and B must be 2-D");
    TORCH_CHECK(A.size(1) == B.size(0), "LoF_gemm: A.cols (", A.size(1), ") != B.rows (", B.size(0), ")");
    TORCH_CHECK(A.scalar_type() == torch::kFloat && B.scalar_type() == torch::kFloat, "LoF_gemm: only float32 inputs are supported");
    TORCH_CHECK(A.is_cuda() && B.is_cuda(), "LoF_gemm requires CUDA tensors (A on ", A.device(), ", B on ", B.device(), ")");
    auto a = A.contiguous(), b = B.contiguous();
    const int64_t M = a.size(0), K = a.size(1), N = b.size(1);
    auto out = torch::empty({M, N}, a.options());
    const double denom = scale_a * scale_b;
    const float alpha = denom == 0.0 ? 1.0f : (float)(1.0 / denom);          // cutlass_gemms.cuh:1784-1786
    const lof_projspec_t ps = {(int)round_mode, LOF_OvfInf, stochastic_rounding_bits};
    c10::cuda::CUDAGuard guard(a.device());
    if (g_gemm.k_chunk > 0) {
        auto status = torch::zeros({1}, a.options().dtype(torch::kInt));
        LOFX_CHECK(lof_gemm_chunked_f32(a.data_ptr<float>(), b.data_ptr<float>(), out.data_ptr<float>(), M, N, K, accum_mant_bits, nullptr, &ps,
                                        g_gemm.k_chunk, alpha, g_gemm.operand_dtype, g_gemm.partial_round ? LOF_GEMM_PARTIAL_ROUND : 0u,
                                        status.data_ptr<int>(), stream_of(a)), "LoF_gemm");
        const int bits = status.item<int>();                                 // the reference synchronises here too (Lof_kernel.cu:149)
        TORCH_CHECK((bits & (LOF_GEMM_STATUS_A_INEXACT | LOF_GEMM_STATUS_B_INEXACT)) == 0,
                    "LoF_gemm (chunked policy): operands are not exactly representable in the tensor-core operand type");
    } else {
        const lof_rng_t rng = next_rng(is_stochastic(round_mode));
        LOFX_CHECK(lof_gemm_f32(a.data_ptr<float>(), b.data_ptr<float>(), out.data_ptr<float>(), M, N, K, accum_mant_bits, &ps, g_gemm.kc,
                                alpha, 0.0f, &rng, stream_of(a)), "LoF_gemm");
    }
    return out;
}

torch::Tensor lof_conv2d(const torch::Tensor& input, const torch::Tensor& weight, int64_t pad_h, int64_t pad_w, int64_t stride_h,
                         int64_t stride_w, int64_t dilation_h, int64_t dilation_w, int accum_mant_bits, RoundingMode round_mode,
                         int stochastic_rounding_bits, double weight_scale, double input_scale)
{
    // checks of Lof_kernel.cu:168-190
    TORCH_CHECK(input.dim() == 4 && weight.dim() == 4, "LoF_conv2d: input and weight must be 4-D (NCHW / OIHW)");
    TORCH_CHECK(input.scalar_type() == torch::kFloat && weight.scalar_type() == torch::kFloat, "LoF_conv2d: only float32 inputs are supported");
    TORCH_CHECK(input.is_cuda() && weight.is_cuda(), "LoF_conv2d requires CUDA tensors");
    TORCH_CHECK(input.size(1) == weight.size(1), "LoF_conv2d: input channels (", input.size(1), ") != weight channels (", weight.size(1),
                "); groups != 1 is not supported by this entry point");
    auto x = input.contiguous(), w = weight.contiguous();                   // NCHW / OIHW as stored: no NHWC / KRSC copies
    lof_conv2d_t g = {};
    g.batch = x.size(0); g.c_in = x.size(1); g.h = x.size(2); g.w = x.size(3);
    g.c_out = w.size(0); g.r = w.size(2); g.s = w.size(3);
    g.pad_h = pad_h; g.pad_w = pad_w; g.stride_h = stride_h; g.stride_w = stride_w; g.dilation_h = dilation_h; g.dilation_w = dilation_w;
    g.groups = 1;
    g.k_order = LOF_CONV_K_RSC;                                             // the reference's NHWC / KRSC reduction order
    int64_t P = 0, Q = 0;
    lof_conv2d_out_size(g.h, g.r, pad_h, stride_h, dilation_h, &P);
    lof_conv2d_out_size(g.w, g.s, pad_w, stride_w, dilation_w, &Q);
    TORCH_CHECK(P > 0 && Q > 0, "LoF_conv2d: non-positive output spatial size (P=", P, ", Q=", Q, ")");
    auto out = torch::empty({g.batch, g.c_out, P, Q}, x.options());
    const lof_projspec_t ps = {(int)round_mode, LOF_OvfInf, stochastic_rounding_bits};
    const double denom = weight_scale * input_scale;
    const lof_rng_t rng = next_rng(is_stochastic(round_mode));
    c10::cuda::CUDAGuard guard(x.device());
    LOFX_CHECK(lof_conv2d_f32(x.data_ptr<float>(), w.data_ptr<float>(), nullptr, out.data_ptr<float>(), &g, accum_mant_bits, &ps, /*kc=*/0,
                              denom == 0.0 ? 1.0f : (float)(1.0 / denom), &rng, stream_of(x)), "LoF_conv2d");
    return out;
}

// Piecewise-linear SiLU (LoPy_bind.cpp:228-277; PWLSiLU.forward): NOT part of the virtual-rounding hot path (SURVEY 2.1 #20,
// out of scope).  Present only because LoFloat/__init__.py imports the name; evaluated with ATen ops on the input's device:
// N+1 uniform knots over [-R, R], linear interpolation, relu asymptote outside the range.
torch::Tensor pwl_silu(const torch::Tensor& input, const torch::Tensor& lut, double R)
{
    TORCH_CHECK(lut.dim() == 1 && lut.numel() >= 2, "pwl_silu: lut must be 1-D with >= 2 entries");
    TORCH_CHECK(R > 0.0, "pwl_silu: R must be positive");
    const int64_t N = lut.numel() - 1;
    auto table = lut.to(input.device(), torch::kFloat).contiguous();
    auto x = input.to(torch::kFloat);
    auto u = (x.clamp(-R, R) + R) * ((double)N / (2.0 * R));
    auto idx = u.floor().clamp(0, (double)(N - 1)).to(torch::kLong);
    auto frac = u - idx.to(torch::kFloat);
    auto y0 = table.index({idx}), y1 = table.index({idx + 1});
    auto y = torch::addcmul(y0, frac, y1 - y0);
    y = torch::where(x.abs() > R, torch::relu(x), y);
    return y.to(input.scalar_type());
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m)
{
    m.doc() = "lofloat-b200: B200-native virtual rounding / MX rounding / LoFloat-accumulation GEMM behind LoFloat's Python surface";

    py::enum_<RoundingMode>(m, "RoundingMode")
        .value("RoundToNearestEven", RoundingMode::RoundToNearestEven).value("RoundTowardsZero", RoundingMode::RoundTowardsZero)
        .value("RoundAwayFromZero", RoundingMode::RoundAwayFromZero).value("StochasticRoundingA", RoundingMode::StochasticRoundingA)
        .value("RoundToNearestOdd", RoundingMode::RoundToNearestOdd).value("RoundDown", RoundingMode::RoundDown)
        .value("RoundUp", RoundingMode::RoundUp).value("RoundTiesToAway", RoundingMode::RoundTiesToAway)
        .value("StochasticRoundingB", RoundingMode::StochasticRoundingB).value("StochasticRoundingC", RoundingMode::StochasticRoundingC)
        .value("True_StochasticRounding", RoundingMode::True_StochasticRounding).value("ProbabilisticRounding", RoundingMode::ProbabilisticRounding)
        .value("RoundToOdd", RoundingMode::RoundToOdd).value("RoundTiesTowardsZero", RoundingMode::RoundTiesTowardsZero)
        .value("StochasticRoundingD", RoundingMode::StochasticRoundingD)
        .export_values();
    // LoFloat/layers.py passes cached plain ints for the rounding mode (`self._gemm_round_mode_int`, layers.py:355-386) into
    // lof_gemm: accept them wherever a RoundingMode is expected.
    py::implicitly_convertible<py::int_, RoundingMode>();
    py::enum_<Signedness>(m, "Signedness").value("Signed", Signedness::Signed).value("Unsigned", Signedness::Unsigned).export_values();
    py::enum_<InfBehavior>(m, "InfBehavior").value("Extended", InfBehavior::Extended).value("Saturating", InfBehavior::Saturating).export_values();
    py::enum_<NaNBehavior>(m, "NaNBehavior").value("_3109", NaNBehavior::_3109).value("_754", NaNBehavior::_754).export_values();

    py::class_<FormatDescriptor>(m, "FloatFormatDescriptor")
        .def(py::init<int, int, int, InfBehavior, NaNBehavior, Signedness, py::object, py::object>(), py::arg("total_bits"),
             py::arg("mantissa_bits"), py::arg("bias"), py::arg("inf_behavior"), py::arg("nan_behavior"), py::arg("signedness"),
             py::arg("is_inf_checker"), py::arg("is_nan_checker"))
        .def_readonly("total_bits", &FormatDescriptor::total_bits)
        .def_readonly("mantissa_bits", &FormatDescriptor::mantissa_bits)
        .def_readonly("bias", &FormatDescriptor::bias)
        .def_readonly("inf_behavior", &FormatDescriptor::inf_behavior)
        .def_readonly("nan_behavior", &FormatDescriptor::nan_behavior)
        .def_readonly("signedness", &FormatDescriptor::signedness)
        .def_readonly("is_inf_checker", &FormatDescriptor::is_inf_checker)
        .def_readonly("is_nan_checker", &FormatDescriptor::is_nan_checker);

    const auto rne = RoundingMode::RoundToNearestEven;
    m.def("virtual_round", &virtual_round_mantissa, py::arg("input"), py::arg("to_mantissa_bits"), py::arg("round_mode") = rne,
          py::arg("stoch_len") = 0, py::arg("scale") = 1.0,
          "round(scale * input) to `to_mantissa_bits` mantissa bits (lo_float.h:2715-2756); new tensor.");
    m.def("virtual_round", &virtual_round_params, py::arg("input"), py::arg("params"), py::arg("round_mode") = rne, py::arg("stoch_len") = 0,
          py::arg("scale") = 1.0, "round(scale * input) into the format `params` (lo_float.h:2843-2939); new tensor.");
    m.def("virtual_mx_round", &virtual_mx_round, py::arg("input"), py::arg("block_size"), py::arg("element_format"), py::arg("scale_format"),
          py::arg("round_mode") = rne, py::arg("scale_round_mode") = rne, py::arg("stoch_len") = 0,
          "MX block fake-quantisation of a float32 tensor (lo_float.h:4311-4383): blocks of `block_size` consecutive elements share one "
          "scale in `scale_format`; elements are rounded into `element_format` and rescaled.");
    m.def("lof_gemm", &lof_gemm, py::arg("A"), py::arg("B"), py::arg("accum_mant_bits"), py::arg("round_mode") = rne,
          py::arg("stochastic_rounding_bits") = 0, py::arg("scale_a") = 1.0, py::arg("scale_b") = 1.0,
          "float32 GEMM whose accumulator is rounded to accum_mant_bits mantissa bits (Lof_kernel.cu:96-154); output / (scale_a * scale_b).");
    m.def("lof_conv2d", &lof_conv2d, py::arg("input"), py::arg("weight"), py::arg("pad_h"), py::arg("pad_w"), py::arg("stride_h"),
          py::arg("stride_w"), py::arg("dilation_h"), py::arg("dilation_w"), py::arg("accum_mant_bits"), py::arg("round_mode") = rne,
          py::arg("stochastic_rounding_bits") = 0, py::arg("weight_scale") = 1.0, py::arg("input_scale") = 1.0,
          "float32 Conv2d forward (NCHW / OIHW, groups = 1) with LoFloat accumulation (Lof_kernel.cu:156-228).");
    m.def("pwl_silu", &pwl_silu, py::arg("input"), py::arg("lut"), py::arg("R"), "Piecewise-linear SiLU over a uniform LUT (not on the hot path).");

    // --- additions (not in the reference module)
    m.def("manual_seed", [](uint64_t seed) { g_seed.store(seed); g_offset.store(0); }, py::arg("seed"),
          "Seed the Philox stream of the stochastic rounding modes and rewind it.");
    m.def("set_gemm_policy", [](int64_t k_chunk, const std::string& operand_dtype, int64_t kc, bool partial_round) {
              int od = operand_dtype == "e4m3" ? LOF_E4M3 : operand_dtype == "e5m2" ? LOF_E5M2 : operand_dtype == "bf16" ? LOF_BF16 : -1;
              TORCH_CHECK(od >= 0, "set_gemm_policy: operand_dtype must be 'e4m3', 'e5m2' or 'bf16'");
              g_gemm.k_chunk = k_chunk; g_gemm.operand_dtype = od; g_gemm.kc = kc; g_gemm.partial_round = partial_round;
          }, py::arg("k_chunk") = 0, py::arg("operand_dtype") = "e4m3", py::arg("kc") = 0, py::arg("partial_round") = false,
          "What lof_gemm(A, B, accum_mant_bits, ...) runs: k_chunk = 0 (default) the reference's per-FMA rounding on the exact kernel "
          "(kc = panel length, 0 = one accumulator over K like the CUDA leaf, 256 = the BLIS CPU path); k_chunk > 0 the tcgen05 chunked "
          "kernels (accumulator re-rounded every k_chunk; operands must be exactly representable in operand_dtype).");
    m.def("kernel_launch_count", []() { return lof_kernel_launch_count(); });
    m.def("abi_version", []() { return lof_abi_version(); });
}
