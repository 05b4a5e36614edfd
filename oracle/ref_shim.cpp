// ref_shim.cpp -- C-ABI shim around the UNMODIFIED reference headers.
//
// TEST INFRASTRUCTURE ONLY.  Compiled by oracle/Makefile from the sources where they lie
// under /root/reference (include paths only -- nothing is copied) into
// oracle/_ref/libref_shim.so.  It exists to (a) pin oracle/lof_oracle.c against the real
// reference and (b) serve as the `cpu_baseline.kind = "reference"` leg of bench.py.
//
// Reference entry points called:
//   lo_float::virtual_round(float/double, FloatingPointParams, ProjSpec)   lo_float.h:4272 -> 2843
//   lo_float::virtual_round(float/double, int mantissa_bits, ProjSpec)     lo_float.h:4260 -> 2715
//   lo_float::virtual_round(float*, float*, n, params, ps)  [xsimd array]  lo_float.h:4278 -> 2942
//   lo_float::virtual_mx_round(float*, n, block, pub, priv, ps, ps)        lo_float.h:4311
//   lof_round_to_odd_fma / lof_sgemm_ukr via bli_sgemm                     blis_lof_gemm.hpp:63,147,280
//   lo_pack_internal::convert_float_to_rep / convert_rep_to_float          lo_pack.hpp:109, 277
//   LoF_encode<E>(uint8_t*, const float*, params, count, ps) / LoF_decode  lo_pack.hpp:419, 452
#include <cstdint>
#include <cstring>
#include <random>

#include "lo_float.h"
#include "lo_pack.hpp"
#ifdef LOF_SHIM_WITH_BLIS
#include "blis_lof_gemm.hpp"
#endif

#ifdef _OPENMP
#include <omp.h>
#endif

using namespace lo_float;

namespace {
using Params = FloatingPointParams<DeviceInfChecker, DeviceNaNChecker>;

Params make_params(int bw, int mb, int bias, int ov, int na, int sgn)
{
    // checkers are never read by virtual_round (LoPy_bind.cpp:178-185 zero-fills them too)
    return Params(bw, mb, bias, static_cast<Inf_Behaviors>(ov), static_cast<NaN_Behaviors>(na),
                  static_cast<Signedness>(sgn), DeviceInfChecker{0, 0, 0, 0}, DeviceNaNChecker{0, 0, 0, 0},
                  Unsigned_behavior::NegtoZero);
}
ProjSpec make_ps(int mode, int stoch_len)
{
    return ProjSpec{static_cast<Rounding_Mode>(mode), Saturation_Mode::OvfInf, stoch_len};
}
}  // namespace

extern "C" {

// ---- scalar loops over the parity-defining overloads ---------------------------------
void ref_round_params_f32(const float* in, float* out, int64_t n, int bw, int mb, int bias, int ov,
                          int na, int sgn, int mode, int stoch_len, unsigned seed, int threads)
{
    const Params p = make_params(bw, mb, bias, ov, na, sgn);
    const ProjSpec ps = make_ps(mode, stoch_len);
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)
    for (int64_t i = 0; i < n; ++i) {
        lo_float::lof_seed = seed;                       // thread_local (lo_float.h:58)
        out[i] = lo_float::virtual_round(in[i], p, ps);
    }
}

void ref_round_params_f64(const double* in, double* out, int64_t n, int bw, int mb, int bias, int ov,
                          int na, int sgn, int mode, int stoch_len, unsigned seed, int threads)
{
    const Params p = make_params(bw, mb, bias, ov, na, sgn);
    const ProjSpec ps = make_ps(mode, stoch_len);
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)
    for (int64_t i = 0; i < n; ++i) {
        lo_float::lof_seed = seed;
        out[i] = lo_float::virtual_round(in[i], p, ps);
    }
}

void ref_round_mantissa_f32(const float* in, float* out, int64_t n, int mant_bits, int mode,
                            int stoch_len, unsigned seed, int threads)
{
    const ProjSpec ps = make_ps(mode, stoch_len);
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)
    for (int64_t i = 0; i < n; ++i) {
        lo_float::lof_seed = seed;
        out[i] = lo_float::virtual_round(in[i], mant_bits, ps);
    }
}

void ref_round_mantissa_f64(const double* in, double* out, int64_t n, int mant_bits, int mode,
                            int stoch_len, unsigned seed, int threads)
{
    const ProjSpec ps = make_ps(mode, stoch_len);
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)
    for (int64_t i = 0; i < n; ++i) {
        lo_float::lof_seed = seed;
        out[i] = lo_float::virtual_round(in[i], mant_bits, ps);
    }
}

// ---- the xsimd array overload (what CPU Python builds call; timing baseline only: it
//      drops the sign and skips the tail) ----------------------
void ref_round_params_array_f32(float* in, float* out, int n, int bw, int mb, int bias, int ov,
                                int na, int sgn, int mode, int stoch_len)
{
    lo_float::virtual_round(in, out, n, make_params(bw, mb, bias, ov, na, sgn), make_ps(mode, stoch_len));
}

// ---- MX ---------------------------------------------------------------------------------
void ref_mx_round_f32(float* inout, int n, int block_size,
                      int pbw, int pmb, int pbias, int pov, int psgn, int pmode,
                      int ebw, int emb, int ebias, int eov, int esgn, int emode, int stoch_len,
                      unsigned seed)
{
    lo_float::lof_seed = seed;
    lo_float::virtual_mx_round(inout, n, block_size,
                               make_params(pbw, pmb, pbias, pov, 0, psgn),
                               make_params(ebw, emb, ebias, eov, 0, esgn),
                               make_ps(pmode, 0), make_ps(emode, stoch_len));
}

// ---- the draws the reference's host stochastic helpers make from std::mt19937(seed) ------
// (lo_float.h:606,666,705,764,835,872): used to pin the stochastic bit recipes.
unsigned ref_draw_uint(unsigned seed, unsigned maxv)            // SR-A/B/C/D first draw
{
    std::mt19937 mt(seed);
    std::uniform_int_distribution<unsigned> dist(0u, maxv);
    return dist(mt);
}
unsigned ref_draw_uint_then_coin(unsigned seed, unsigned maxv)  // SR-D: dist(mt) then coin(mt)
{
    std::mt19937 mt(seed);
    std::uniform_int_distribution<unsigned> dist(0u, maxv);
    std::uniform_int_distribution<int> coin(0, 1);
    unsigned r = dist(mt);
    unsigned c = static_cast<unsigned>(coin(mt));
    return r | (c << 31);                                       // coin returned in bit 31
}
int ref_draw_01(unsigned seed)                                  // Probabilistic_Round
{
    std::mt19937 mt(seed);
    std::uniform_int_distribution<int> d01(0, 1);
    return d01(mt);
}
double ref_draw_real(unsigned seed)                             // True_Stochastic_Round
{
    std::mt19937 mt(seed);
    std::uniform_real_distribution<double> distribution(0.0, 1.0);
    return distribution(mt);
}

int ref_shim_threads()
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

#ifdef LOF_SHIM_WITH_BLIS
float ref_round_to_odd_fma(float a, float b, float c) { return lof_round_to_odd_fma(a, b, c); }

// Row-major C(MxN) = alpha * A(MxK) * B(KxN) + beta*C through bli_sgemm with the LoFloat
// micro-kernel installed (test/test_blis_lof_accum.cpp:122,160,70).  accum_mant_bits < 0
// disables the custom accumulation (plain fp32).
void ref_gemm_f32(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K,
                  int accum_mant_bits, int mode, int stoch_len, float alpha, float beta,
                  unsigned seed, int threads)
{
    static bool installed = false;
    if (!installed) { lof_blis_install_kernels(); installed = true; }
    lo_float::lof_seed = seed;
    lof_blis_set_accum(accum_mant_bits, make_ps(mode, stoch_len));
    bli_thread_set_num_threads(threads > 0 ? threads : 1);
    // row-major: rs = ld, cs = 1
    bli_sgemm(BLIS_NO_TRANSPOSE, BLIS_NO_TRANSPOSE, M, N, K, &alpha,
              const_cast<float*>(A), K, 1, const_cast<float*>(B), N, 1, &beta, C, N, 1);
}
int ref_has_blis() { return 1; }
#else
int ref_has_blis() { return 0; }
#endif


// ---- low-bit storage codec (lo_pack.hpp) with the Inf / NaN functors as data ---------------
struct ShimCodecFormat {                    // == lofo_codec_format (oracle/lof_oracle.h)
    int32_t bitwidth, mantissa_bits, bias, ov_behavior, na_behavior, is_signed, unsigned_behavior;
    uint32_t inf_mask, inf_value, nan_mask, nan_value, nan_mant_mask, min_pos_inf, qnan, snan;
};
struct ShimInf {
    uint32_t mask, value, min_pos, sign;
    constexpr bool operator()(uint64_t bits) const { return ((uint32_t)bits & mask) == value; }
    constexpr uint64_t minNegInf() const { return min_pos | sign; }
    constexpr uint64_t minPosInf() const { return min_pos; }
};
struct ShimNaN {
    uint32_t mask, value, mant_mask, q, s;
    constexpr bool operator()(uint64_t bits) const { return ((uint32_t)bits & mask) == value && (mant_mask == 0u || ((uint32_t)bits & mant_mask) != 0u); }
    constexpr uint64_t qNanBitPattern() const { return q; }
    constexpr uint64_t sNanBitPattern() const { return s; }
};
using CodecParams = FloatingPointParams<ShimInf, ShimNaN>;
static CodecParams make_codec(const ShimCodecFormat* f)
{
    const uint32_t sign = (f->is_signed == 0) ? (1u << (f->bitwidth - 1)) : 0u;
    return CodecParams(f->bitwidth, f->mantissa_bits, f->bias, static_cast<Inf_Behaviors>(f->ov_behavior),
                       static_cast<NaN_Behaviors>(f->na_behavior), static_cast<Signedness>(f->is_signed),
                       ShimInf{f->inf_mask, f->inf_value, f->min_pos_inf, sign},
                       ShimNaN{f->nan_mask, f->nan_value, f->nan_mant_mask, f->qnan, f->snan},
                       static_cast<Unsigned_behavior>(f->unsigned_behavior));
}
static ProjSpec make_ps3(int mode, int sat, int stoch_len)
{
    return ProjSpec{static_cast<Rounding_Mode>(mode), static_cast<Saturation_Mode>(sat), stoch_len};
}

void ref_float_to_rep(const float* in, uint32_t* out, int64_t n, const ShimCodecFormat* f, int mode, int sat, int stoch_len, unsigned seed)
{
    const CodecParams p = make_codec(f);
    const ProjSpec ps = make_ps3(mode, sat, stoch_len);
    for (int64_t i = 0; i < n; ++i) {
        lo_float::lof_seed = seed;
        out[i] = lo_pack_internal::convert_float_to_rep(in[i], p, ps);
    }
}
void ref_rep_to_float(const uint32_t* in, float* out, int64_t n, const ShimCodecFormat* f)
{
    const CodecParams p = make_codec(f);
    for (int64_t i = 0; i < n; ++i) out[i] = lo_pack_internal::convert_rep_to_float(in[i], p);
}
uint64_t ref_packed_size(int n, uint64_t count) { return LoF_packed_size(n, count); }
int64_t ref_encode_f32(uint8_t* dst, const float* src, uint64_t count, const ShimCodecFormat* f, int mode, int sat, int stoch_len, int endian)
{
    const CodecParams p = make_codec(f);
    const ProjSpec ps = make_ps3(mode, sat, stoch_len);
    return endian == 0 ? LoF_encode<PackEndian::Forward>(dst, src, p, count, ps) : LoF_encode<PackEndian::Reverse>(dst, src, p, count, ps);
}
int64_t ref_decode_f32(float* dst, const uint8_t* src, uint64_t count, const ShimCodecFormat* f, int endian)
{
    const CodecParams p = make_codec(f);
    return endian == 0 ? LoF_decode<PackEndian::Forward>(dst, src, p, count) : LoF_decode<PackEndian::Reverse>(dst, src, p, count);
}
// typed ground truth of test/test_pack.cpp for two named presets: rep of Templated_Float<Fp>(x) and (float)FromRep(rep)
void ref_typed_e4m3fn(const float* in, uint32_t* rep, float* back, int64_t n)
{
    using TF = lo_float_internal::Templated_Float<lo_float_internal::param_float8_e4m3fn>;
    for (int64_t i = 0; i < n; ++i) { TF t(in[i]); rep[i] = (uint32_t)t.rep(); back[i] = (float)t; }
}
void ref_typed_p3109_8_4(const float* in, uint32_t* rep, float* back, int64_t n)
{
    using TF = lo_float_internal::Templated_Float<lo_float_internal::param_float_p_3109<8, 4>>;
    for (int64_t i = 0; i < n; ++i) { TF t(in[i]); rep[i] = (uint32_t)t.rep(); back[i] = (float)t; }
}
}  // extern "C"
