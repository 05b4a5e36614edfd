// lof_api.cu -- C ABI of liblofloat_b200.so (declared in include/lofloat_b200.h).
//
// Host side of the hot path: argument validation (the reference's TORCH_CHECKs,
// LoPy_bind.cpp:172-173, Lof_kernel.cu:90,106-113 become status codes), folding a
// (format, source type) pair into a RoundPlan once per call instead of per element
// (the reference calls std::pow twice per element, lo_float.h:2861), kernel selection, and
// the host-buffer pipeline (chunked H2D / kernel / D2H on three streams).
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <utility>
#include <vector>
#include <cuda.h>

#include "lof_core.cuh"
#include "lof_gemm_common.cuh"
#include "lof_mx_kernels.cuh"
#include "lof_pack.cuh"
#include "lof_round_kernels.cuh"

namespace lof {
cudaError_t launch_gemm_exact(int mode, const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K,
                              int accum_mant_bits, int len, int64_t kc, float alpha, float beta, const RngArgs& rng,
                              cudaStream_t stream, int64_t batch = 1, int64_t stride_a = 0, int64_t stride_b = 0, int64_t stride_c = 0,
                              const ExactView* view = nullptr, int fast = 1);
cudaError_t launch_fwht_large(const float* in, float* out, int64_t n, int hadamard_block, int num_sms, cudaStream_t stream);
cudaError_t launch_selftest_rne_round(unsigned long long* mismatches, int num_sms, cudaStream_t stream);
cudaError_t launch_mx_tma(int mode, const float* in, float* out, int64_t n, const MxArgs& a, const RngArgs& rng,
                          const RngArgs& rng_scale, uint64_t first_index, const MxLaunchCfg& cfg, cudaStream_t stream,
                          int* launches, bool* handled);
cudaError_t launch_mx_fast(int mode, int scale_mode, const float* in, float* out, int64_t n, const MxArgs& g,
                           const MagicPlan& elem_magic, const MagicPlan& scale_magic, const FastPlan& scale_fast, int scale_mantissa_bits,
                           const MxLaunchCfg& cfg, cudaStream_t stream, int64_t* main_elems, int* launches, const RngArgs* rng_elem, uint64_t first_index);
cudaError_t launch_fwht(int op, int mode, const float* in, float* out, int64_t n, int hadamard_block, const RoundPlan* plan,
                        float pre_scale, const MxArgs* mx, const MagicPlan* scale_magic, const FastPlan* scale_fast,
                        const MagicPlan* elem_magic, const RngArgs& rng, const RngArgs& rng_scale, uint64_t first_index, int num_sms,
                        cudaStream_t stream);
cudaError_t run_mx_div_selftest(float pmn, unsigned long long* host_mismatches, cudaStream_t stream);
cudaError_t launch_gemm_tc(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K, int mode,
                           int accum_mant_bits, const lof_format_t* acc_fmt, const RoundPlan* acc_plan, const FastPlan* acc_fast,
                           int64_t k_chunk, float alpha, int operand_dtype, uint32_t flags, int* dev_status,
                           cudaStream_t stream, int* launches, std::string* err, const RngArgs* rng, int stoch_len);
cudaError_t launch_gemm_mx(const uint8_t* A, const uint8_t* SA, const uint8_t* Bt, const uint8_t* SB, float* C, int64_t M, int64_t N, int64_t K,
                           int elem_dtype, int mode, int accum_mant_bits, const lof_format_t* acc_fmt, const RoundPlan* acc_plan,
                           const FastPlan* acc_fast, int64_t k_chunk, float alpha, uint32_t flags, cudaStream_t stream, int* launches,
                           std::string* err);
cudaError_t launch_mx_pack_tc(const float* y, int64_t n, uint8_t* elems, uint8_t* scales, int* dev_status, int num_sms, cudaStream_t stream);
cudaError_t read_mx_prof(unsigned long long* out16);
cudaError_t launch_p2p_post(int* flags, int epoch, int expected_done, cudaStream_t stream);
cudaError_t launch_p2p_wait(const int* peer_flags, int* my_flags, int epoch, cudaStream_t stream);
cudaError_t launch_p2p_done(int* peer_flags, cudaStream_t stream);
}  // namespace lof

using namespace lof;

namespace {

thread_local std::string g_last_error;
std::atomic<uint64_t> g_launches{0};

struct Options {
    // B200 sweeps (scripts/round_tune.py, scripts/geom_tune2.py): the streaming kernels peak with ~48 KB of loads in flight per
    // SM -- 3 resident CTAs x 256 threads x 4 x 16 B -- (6.1-6.3 TB/s); more residency costs 5-10 % (DRAM read/write turnarounds)
    std::atomic<int64_t> round_unroll{0};        // 0 = auto: 4 vectors in flight per thread
    std::atomic<int64_t> round_ctas_per_sm{0};   // 0 = auto: 3 CTAs per SM
    std::atomic<int64_t> mx_variant{0};
    std::atomic<int64_t> mx_ctas_per_sm{3};
    std::atomic<int64_t> stream_ctas_per_sm{8};  // codec / strict rounding / MX storage / fused Hadamard kernels: ALU-heavier, they want full residency (scripts/stream_tune.py)
    std::atomic<int64_t> host_chunk_bytes{32ll << 20};
    std::atomic<int64_t> pack_fast{2};           // storage codec: 0 = exact integer routine only, 1 = magic-add fast path where eligible, 2 = + register-lane kernels
    std::atomic<int64_t> bf16_fast{2};           // bf16 containers on the magic path: 2 = unchecked kinds where the fp32 path is unchecked, 1 = always checked
    std::atomic<int64_t> gemm_mx_prof{0};        // 1: lof_mx_gemm records CTA 0's cycle breakdown (lof_debug_mx_prof)
    std::atomic<int64_t> gemm_fold_split{1};     // packed fold GEMM, 128 x 256 tile: 1 = one N = 256 MMA per k-step (2 TMEM stages), 2 = two N = 128 MMAs (4 stages)
    std::atomic<int64_t> gemm_fast{1};           // exact GEMM: 1 = optimistic branch-free chain with a per-tile fallback, 0 = exact routines only
    std::atomic<int64_t> round_fast{2};          // 0: exact kernels; 1: branch-free fast path; 2: + magic-add path (default)
} g_opt;

int fail(int status, const std::string& msg)
{
    g_last_error = msg;
    return status;
}

int cuda_fail(cudaError_t e, const char* what)
{
    return fail(e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? LOF_ERR_NO_DEVICE : LOF_ERR_CUDA,
                std::string(what) + ": " + cudaGetErrorString(e));
}

int num_sms()
{
    // per-device cache, filled once (atomics: the C ABI may be called from several host threads, one per GPU)
    static std::atomic<int> cache[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (dev < 0 || dev >= 64) {
        int v = 0;
        return cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess ? v : 0;
    }
    int v = cache[dev].load(std::memory_order_relaxed);
    if (v == 0) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
        cache[dev].store(v, std::memory_order_relaxed);
    }
    return v;
}

struct SrcGeom { int mant, bias; };
bool src_geom(int dtype, SrcGeom* g)
{
    switch (dtype) {
    case LOF_F32:  *g = {23, 127};  return true;
    case LOF_F16:  *g = {10, 15};   return true;
    case LOF_BF16: *g = {7, 127};   return true;
    case LOF_F64:  *g = {52, 1023}; return true;
    default: return false;
    }
}

uint16_t f32_to_f16_bits(float f) { return __half_as_ushort(__float2half_rn(f)); }
uint16_t f32_to_bf16_bits(float f) { return __bfloat16_as_ushort(__float2bfloat16_rn(f)); }

// Fold (format, source container) into the per-launch constants.
int make_plan(const lof_format_t* f, int dtype, int mode, int stoch_len, RoundPlan* p)
{
    SrcGeom g;
    if (!src_geom(dtype, &g)) return fail(LOF_ERR_INVALID_ARG, "unknown dtype");
    if (!f) return fail(LOF_ERR_INVALID_ARG, "format is NULL");
    const int ebits = f->bitwidth - f->mantissa_bits - (f->is_signed == LOF_Signed ? 1 : 0);
    if (f->mantissa_bits < 0 || f->mantissa_bits > 52 || ebits < 0 || ebits > 30)
        return fail(LOF_ERR_INVALID_ARG, "format: need 0 <= mantissa_bits <= 52 and 0 <= exponent bits <= 30");
    if (f->is_signed != LOF_Signed && f->is_signed != LOF_Unsigned) return fail(LOF_ERR_INVALID_ARG, "format: bad signedness");
    if (mode < 0 || mode >= LOF_NUM_ROUNDING_MODES) return fail(LOF_ERR_INVALID_ARG, "unknown rounding mode");
    const int to_max_exp = (1 << ebits) - 1 - f->bias;          // lo_float.h:2857-2859
    const int to_min_exp = 1 - f->bias;                          // :2860
    const float min_val = (float)(std::pow(2.0, (double)to_min_exp) * std::pow(2.0, (double)-f->mantissa_bits));  // :2861
    p->min_val = min_val;
    p->half_val = min_val * 0.5f;                                // :2894
    switch (dtype) {                                             // from_double<From>((double)To_min_val), :2905
    case LOF_F32: { uint32_t b; std::memcpy(&b, &min_val, 4); p->min_mag_bits = b; break; }
    case LOF_F64: { double d = (double)min_val; uint64_t b; std::memcpy(&b, &d, 8); p->min_mag_bits = b; break; }
    case LOF_F16: p->min_mag_bits = f32_to_f16_bits(min_val); break;
    default:      p->min_mag_bits = f32_to_bf16_bits(min_val); break;
    }
    p->min_exp_b = to_min_exp + g.bias;
    p->max_exp_b = to_max_exp + g.bias;
    p->shiftN = g.mant - f->mantissa_bits;
    p->stoch_len = stoch_len;
    p->saturating = (f->ov_behavior == LOF_Saturating);
    p->pad_ = 0;
    // Stochastic bit counts outside [0, r] make the reference shift by a negative amount
    // (undefined, lo_float.h:664,751,810,833); refuse instead of inventing a result.
    if (mode == LOF_StochasticRoundingA || mode == LOF_StochasticRoundingB || mode == LOF_StochasticRoundingC ||
        mode == LOF_StochasticRoundingD) {
        const int rmin = p->shiftN > 1 ? p->shiftN : 1;
        if (stoch_len < 0 || stoch_len > rmin || stoch_len > 31)
            return fail(LOF_ERR_INVALID_ARG, "stoch_length must be in [0, container_mantissa - target_mantissa]");
    }
    return LOF_OK;
}

// Eligibility + constants of the fp32 fast path (lof_fast_f32.cuh).  plan must come from
// make_plan(fmt, LOF_F32, ...).
FastPlan make_fast_plan(const RoundPlan& p, int mode)
{
    FastPlan f;
    std::memset(&f, 0, sizeof f);
    const int m_exp = p.min_exp_b + p.shiftN;                 // biased exponent of M = denorm_min * 2^23
    uint32_t mv, hv;
    std::memcpy(&mv, &p.min_val, 4);
    std::memcpy(&hv, &p.half_val, 4);
    bool ok = p.shiftN >= 1 && p.shiftN <= 23 && p.min_exp_b >= 1 && p.min_exp_b <= 254 && m_exp >= 1 && m_exp <= 254 &&
              mv != 0 && p.max_exp_b >= p.min_exp_b;
    // RoundToNearestOdd / RoundTiesToAway send an exact ZERO to denorm_min when denorm_min / 2 underflows to 0
    // (lo_float.h:2894-2899, the tie branch): leave that corner to the exact kernel.
    if ((mode == LOF_RoundToNearestOdd || mode == LOF_RoundTiesToAway) && hv == 0) ok = false;
    if (!ok) return f;
    f.maskN = ~((1u << p.shiftN) - 1u);
    f.min_normal_bits = (uint32_t)p.min_exp_b << 23;
    const int64_t thr = (int64_t)p.max_exp_b + 1;
    f.ovf_thresh = thr > 256 ? 0xFFFFFFFFu : (uint32_t)(thr << 23);
    const uint32_t mbits = (uint32_t)m_exp << 23;
    std::memcpy(&f.M, &mbits, 4);
    f.min_val = p.min_val;
    f.shiftN = p.shiftN;
    f.rbase = p.shiftN + p.min_exp_b;
    f.stoch_len = p.stoch_len;
    f.check_denorm_in = mv < 0x00800000u;                     // denorm_min is itself an fp32 subnormal
    f.valid = 1;
    return f;
}

// Eligibility + constants of the magic-add path (lof_fast_f32.cuh).  plan from make_plan(fmt, LOF_F32, ...).
MagicPlan make_magic_plan(const RoundPlan& p, int mode)
{
    MagicPlan m;
    std::memset(&m, 0, sizeof m);
    const bool sr = mode == LOF_StochasticRoundingA || mode == LOF_StochasticRoundingD;
    const bool mode_ok = mode == LOF_RoundToNearestEven || mode == LOF_RoundTowardsZero || mode == LOF_RoundAwayFromZero ||
                         mode == LOF_RoundUp || mode == LOF_RoundDown || mode == LOF_RoundToOdd || sr;
    // stochastic: 1 <= len <= shiftN (reference precondition), the offset's exponent c - (23 + len) must stay normal
    if (sr && (p.stoch_len < 1 || p.stoch_len > p.shiftN || p.stoch_len > 22)) return m;
    uint32_t mv;
    std::memcpy(&mv, &p.min_val, 4);
    // m = 0 (E8M0-like): the reference's RNE / round-to-odd act on the raw bits, whose "LSB" is then the exponent
    // field's LSB, not a significand bit -- the value-based magic add would break ties differently
    if (p.shiftN == 23 && (mode == LOF_RoundToNearestEven || mode == LOF_RoundToOdd)) return m;
    if (!mode_ok || p.shiftN < 1 || p.shiftN > 23 || p.min_exp_b < 1 || p.min_exp_b > 254 ||
        p.min_exp_b + p.shiftN > 254 || mv == 0 || p.max_exp_b < p.min_exp_b)
        return m;
    m.c_add = (uint32_t)p.shiftN << 23;
    m.c_floor = (uint32_t)(p.min_exp_b + p.shiftN) << 23;
    const int64_t thr = (int64_t)p.max_exp_b + 1;
    m.ovf_thresh = thr > 255 ? 0xFFFFFFFFu : (uint32_t)(thr << 23);
    // every x whose exponent field e has e + shiftN >= 255 must either be an overflow (then the garbage constant
    // still yields a magnitude >= ovf_thresh, see lof_fast_f32.cuh) or go to the exact routine
    m.limit = (p.max_exp_b + p.shiftN <= 254) ? 0xFFFFFFFFu : (uint32_t)(255 - p.shiftN) << 23;
    m.check_denorm_in = mv < 0x00800000u;
    m.lo_rare = m.check_denorm_in ? 0x007FFFFFu : 0u;
    // SR: the offset's bit pattern is c - sr_sub with c >= (e + shiftN) << 23; it is a normal fp32 number iff
    // e + shiftN >= 24 + len.  Formats whose subnormal floor already guarantees that need no check; for wide
    // formats (bf16-like) the few binades below go to the exact routine.
    if (sr && p.min_exp_b + p.shiftN < 24 + p.stoch_len) {
        const uint32_t lo_limit = (uint32_t)(24 + p.stoch_len - p.shiftN) << 23;
        if (lo_limit - 1u > m.lo_rare) m.lo_rare = lo_limit - 1u;
    }
    m.sr_len = sr ? (uint32_t)p.stoch_len : 0u;
    m.sr_mask = sr ? ((1u << p.stoch_len) - 1u) : 0u;
    m.sr_sub = sr ? (uint32_t)(23 + p.stoch_len) << 23 : 0u;
    m.min_val = p.min_val;
    m.valid = 1;
    return m;
}

RngArgs make_rng(const lof_rng_t* rng, uint32_t stream_const, int mode = LOF_True_StochasticRounding, int len = 0)
{
    RngArgs r;
    r.cls = rand_class(mode, len);            // default: one dense 32-bit word per event (GEMM, MX scale)
    r.pad_ = 0;
    const uint64_t seed = rng ? rng->seed : 0, off = rng ? rng->offset : 0;
    r.key0 = (uint32_t)seed; r.key1 = (uint32_t)(seed >> 32);
    r.off0 = (uint32_t)off;  r.off1 = (uint32_t)(off >> 32) ^ stream_const;
    return r;
}

RoundLaunchCfg round_cfg(int mode = LOF_RoundToNearestEven, bool checked_kind = false)
{
    RoundLaunchCfg c;
    c.unroll = (int)g_opt.round_unroll.load();
    c.ctas_per_sm = (int)g_opt.round_ctas_per_sm.load();
    const bool rng_mode = mode_uses_rng(mode);      // the shared-Philox-block tile layout needs 4 vectors per thread
    (void)rng_mode; (void)checked_kind;
    if (c.unroll == 0) c.unroll = 4;
    if (c.ctas_per_sm == 0) c.ctas_per_sm = 3;
    c.num_sms = num_sms();
    return c;
}

int round_common(bool params, const void* in, void* out, int64_t n, int dtype, const lof_format_t* fmt,
                 int mantissa_bits, const lof_projspec_t* ps, double scale, const lof_rng_t* rng,
                 uint64_t first_index, void* stream)
{
    if (n < 0) return fail(LOF_ERR_INVALID_ARG, "n < 0");
    if (!ps) return fail(LOF_ERR_INVALID_ARG, "projspec is NULL");
    SrcGeom g;
    if (!src_geom(dtype, &g)) return fail(LOF_ERR_INVALID_ARG, "unknown dtype");
    const int mode = ps->rounding_mode;
    if (mode < 0 || mode >= LOF_NUM_ROUNDING_MODES) return fail(LOF_ERR_INVALID_ARG, "unknown rounding mode");
    RoundPlan plan;
    std::memset(&plan, 0, sizeof plan);
    FastPlan fast;
    std::memset(&fast, 0, sizeof fast);
    MagicPlan magic;
    std::memset(&magic, 0, sizeof magic);
    int kind = params ? kKindParams : kKindMantissa;
    int r = 0;
    if (params) {
        const int st = make_plan(fmt, dtype, mode, ps->stoch_length, &plan);
        if (st != LOF_OK) return st;
        const int64_t fast_opt = g_opt.round_fast.load();        // 0: exact kernel, 1: first-generation fast path, 2: + magic add
        const bool bf16_sr = (mode == LOF_StochasticRoundingA || mode == LOF_StochasticRoundingD) && fast_opt >= 2;
        if (dtype == LOF_BF16 && fast_opt && (!mode_uses_rng(mode) || bf16_sr) && scale == 1.0 && fmt->mantissa_bits <= 7) {
            // bf16 container: the fp32 branch-free routines applied to (bits << 16) (lof_round_kernels.cuh); always the checked
            // variants.  `plan` stays the exact bf16-container plan for the rare lanes and the scalar tail.
            RoundPlan p32;
            std::memset(&p32, 0, sizeof p32);
            if (make_plan(fmt, LOF_F32, mode, ps->stoch_length, &p32) == LOF_OK) {
                fast = make_fast_plan(p32, mode);
                if (fast.valid) kind = p32.saturating ? kKindFastSatChk : kKindFastExtChk;
                if (mode_uses_rng(mode)) { kind = kKindParams; std::memset(&fast, 0, sizeof fast); }     // stochastic: magic path or nothing
                if (fast_opt >= 2) {
                    magic = make_magic_plan(p32, mode);
                    if (magic.valid) {
                        // checked kinds only where the fp32 path itself needs them (wide targets, stochastic offsets) or bf16.fast = 1
                        const bool chk = magic.lo_rare != 0u || magic.limit != 0xFFFFFFFFu || mode_uses_rng(mode) || g_opt.bf16_fast.load() < 2;
                        kind = chk ? (p32.saturating ? kKindMagicSatChk : kKindMagicExtChk) : (p32.saturating ? kKindMagicSat : kKindMagicExt);
                        if (chk) {      // the bf16 kernel tests running extrema once per vector: bf16-subnormal inputs and NaN are "rare" lanes there
                            if (magic.lo_rare < 0x007FFFFFu) magic.lo_rare = 0x007FFFFFu;
                            if (magic.limit > 0x7F800001u) magic.limit = 0x7F800001u;
                        }
                    }
                }
            }
        }
        if (dtype == LOF_F16 && fast_opt >= 2 && !mode_uses_rng(mode) && scale == 1.0 && fmt->mantissa_bits <= 10) {
            // fp16 container: widen (exact), round with the fp32 magic add, narrow (exact).  Only where every result is an fp16
            // value again and fp16-subnormal inputs sit below denorm_min: largest target exponent <= 15, denorm_min >= 2^-14,
            // and the fp32 path itself needs no per-lane checks.  `plan` stays the container plan for the scalar tail.
            RoundPlan p32;
            std::memset(&p32, 0, sizeof p32);
            if (make_plan(fmt, LOF_F32, mode, ps->stoch_length, &p32) == LOF_OK) {
                const MagicPlan m16 = make_magic_plan(p32, mode);
                if (m16.valid && m16.lo_rare == 0u && m16.limit == 0xFFFFFFFFu && p32.max_exp_b <= 127 + 15 &&
                    p32.min_val >= 6.103515625e-05f) {
                    magic = m16;
                    kind = p32.saturating ? kKindMagicSat : kKindMagicExt;
                }
            }
        }
        if (dtype == LOF_F32 && fast_opt) {
            fast = make_fast_plan(plan, mode);
            if (fast.valid) kind = fast.check_denorm_in ? (plan.saturating ? kKindFastSatChk : kKindFastExtChk)
                                                        : (plan.saturating ? kKindFastSat : kKindFastExt);
            if (fast_opt >= 2) {
                magic = make_magic_plan(plan, mode);
                if (magic.valid) {
                    const bool chk = magic.lo_rare != 0u || magic.limit != 0xFFFFFFFFu;
                    if (chk && magic.limit > 0x7F800001u) magic.limit = 0x7F800001u;      // checked kinds: NaN inputs are "rare" lanes (lof_round_kernels.cuh)
                    kind = chk ? (plan.saturating ? kKindMagicSatChk : kKindMagicExtChk)
                               : (plan.saturating ? kKindMagicSat : kKindMagicExt);
                }
            }
        }
    } else {
        r = g.mant - mantissa_bits;                                           // -kDigitShift, lo_float.h:2722
        if (mantissa_bits < 0) return fail(LOF_ERR_INVALID_ARG, "mantissa_bits < 0");
        if (r > 0 && (mode == LOF_StochasticRoundingA || mode == LOF_StochasticRoundingB ||
                      mode == LOF_StochasticRoundingC || mode == LOF_StochasticRoundingD) &&
            (ps->stoch_length < 0 || ps->stoch_length > r || ps->stoch_length > 31))
            return fail(LOF_ERR_INVALID_ARG, "stoch_length must be in [0, container_mantissa - mantissa_bits]");
    }
    if (n == 0) return LOF_OK;
    if (!in || !out) return fail(LOF_ERR_INVALID_ARG, "null tensor pointer");
    const RoundLaunchCfg cfg = round_cfg(mode, kind == kKindFastSatChk || kind == kKindFastExtChk || kind == kKindMagicSatChk || kind == kKindMagicExtChk);
    if (cfg.num_sms <= 0) return fail(LOF_ERR_NO_DEVICE, "no CUDA device (this library has no CPU fallback)");
    const RngArgs ra = make_rng(rng, kStreamElem, mode, ps->stoch_length);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    cudaError_t e;
    switch (dtype) {
    case LOF_F32:  e = launch_round_f32(kind, mode, in, out, n, plan, fast, magic, r, ps->stoch_length, scale, ra, first_index, cfg, s); break;
    case LOF_F16:  e = launch_round_f16(kind, mode, in, out, n, plan, fast, magic, r, ps->stoch_length, scale, ra, first_index, cfg, s); break;
    case LOF_BF16: e = launch_round_bf16(kind, mode, in, out, n, plan, fast, magic, r, ps->stoch_length, scale, ra, first_index, cfg, s); break;
    default:       e = launch_round_f64(kind, mode, in, out, n, plan, fast, magic, r, ps->stoch_length, scale, ra, first_index, cfg, s); break;
    }
    if (e != cudaSuccess) return cuda_fail(e, "virtual_round launch");
    g_launches.fetch_add(1);
    return LOF_OK;
}

size_t dtype_size(int dtype) { return dtype == LOF_F64 ? 8 : dtype == LOF_F32 ? 4 : 2; }

// ---------------------------------------------------------------------------------------
// host-buffer pipeline: chunk i is copied in on stream A while chunk i-1 runs on stream B
// and chunk i-2 is copied out on stream C (events order the three stages per buffer).
// ---------------------------------------------------------------------------------------
struct HostPipe {
    static constexpr int kBufs = 3;
    void* dev[kBufs] = {nullptr, nullptr, nullptr};
    size_t cap = 0;
    cudaStream_t s_in = nullptr, s_k = nullptr, s_out = nullptr;
    cudaEvent_t in_done[kBufs] = {nullptr, nullptr, nullptr}, k_done[kBufs] = {nullptr, nullptr, nullptr},
                out_done[kBufs] = {nullptr, nullptr, nullptr};
    bool init = false;
    std::mutex mu;

    // the caller has made the pipe's device current
    int ensure(size_t bytes)
    {
        if (init && cap >= bytes) return LOF_OK;
        release();
        cudaError_t e = cudaSuccess;
        for (int i = 0; i < kBufs && e == cudaSuccess; ++i) e = cudaMalloc(&dev[i], bytes);
        if (e != cudaSuccess) { release(); return fail(LOF_ERR_ALLOC, std::string("cudaMalloc staging: ") + cudaGetErrorString(e)); }
        cap = bytes;
        e = cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s_k, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking);
        for (int i = 0; i < kBufs && e == cudaSuccess; ++i) {
            e = cudaEventCreateWithFlags(&in_done[i], cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&k_done[i], cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&out_done[i], cudaEventDisableTiming);
        }
        if (e != cudaSuccess) { release(); return cuda_fail(e, "host pipeline: stream / event creation"); }
        init = true;
        return LOF_OK;
    }
    void drain()
    {
        if (s_in) cudaStreamSynchronize(s_in);
        if (s_k) cudaStreamSynchronize(s_k);
        if (s_out) cudaStreamSynchronize(s_out);
    }
    void release()
    {
        drain();
        for (int i = 0; i < kBufs; ++i) {
            if (in_done[i]) cudaEventDestroy(in_done[i]);
            if (k_done[i]) cudaEventDestroy(k_done[i]);
            if (out_done[i]) cudaEventDestroy(out_done[i]);
            in_done[i] = k_done[i] = out_done[i] = nullptr;
            if (dev[i]) cudaFree(dev[i]);
            dev[i] = nullptr;
        }
        if (s_in) cudaStreamDestroy(s_in);
        if (s_k) cudaStreamDestroy(s_k);
        if (s_out) cudaStreamDestroy(s_out);
        s_in = s_k = s_out = nullptr;
        cap = 0; init = false;
    }
};
constexpr int kMaxDevices = 64;
HostPipe g_pipes[kMaxDevices];          // one pipeline (staging buffers, streams, events, lock) PER DEVICE: calls for different GPUs run concurrently

// kernel(dev_ptr, elems, first_index_of_chunk, stream) -> status
template <typename KernelFn>
int run_host_pipeline(const void* in_host, void* out_host, int64_t n, size_t esize, int64_t granule, int device,
                      KernelFn&& kernel)
{
    if (n < 0) return fail(LOF_ERR_INVALID_ARG, "n < 0");
    if (n == 0) return LOF_OK;
    if (!in_host || !out_host) return fail(LOF_ERR_INVALID_ARG, "null host pointer");
    int prev_dev = 0;
    cudaError_t e = cudaGetDevice(&prev_dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
    const int dev_id = device < 0 ? prev_dev : device;
    if (dev_id >= kMaxDevices) return fail(LOF_ERR_INVALID_ARG, "device index out of range");
    // the pipe's device is current for the WHOLE call (kernels, occupancy queries and num_sms() act on the current device) and the
    // caller's device is restored on every exit path
    struct DeviceGuard {
        int prev, cur;
        ~DeviceGuard() { if (cur != prev) cudaSetDevice(prev); }
    } guard{prev_dev, dev_id};
    if (dev_id != prev_dev) {
        e = cudaSetDevice(dev_id);
        if (e != cudaSuccess) { guard.cur = prev_dev; return cuda_fail(e, "cudaSetDevice"); }
    }
    HostPipe& pipe = g_pipes[dev_id];
    std::lock_guard<std::mutex> lock(pipe.mu);
    int64_t chunk_elems = (int64_t)(g_opt.host_chunk_bytes.load() / (int64_t)esize);
    chunk_elems = (chunk_elems / granule) * granule;
    if (chunk_elems < granule) chunk_elems = granule;
    if (chunk_elems > n) chunk_elems = ((n + granule - 1) / granule) * granule;
    int st = pipe.ensure((size_t)chunk_elems * esize);
    if (st != LOF_OK) return st;
    const int64_t nchunks = (n + chunk_elems - 1) / chunk_elems;
    const char* src = static_cast<const char*>(in_host);
    char* dst = static_cast<char*>(out_host);
    // On a mid-loop failure copies of earlier chunks may still be writing out_host and the staging buffers: every exit path
    // waits for the three streams before the call returns and the next call reuses the buffers.
    auto bail = [&](int status) { pipe.drain(); return status; };
    for (int64_t c = 0; c < nchunks; ++c) {
        const int b = (int)(c % HostPipe::kBufs);
        const int64_t e0 = c * chunk_elems;
        const int64_t ne = (n - e0) < chunk_elems ? (n - e0) : chunk_elems;
        if (c >= HostPipe::kBufs) {
            e = cudaStreamWaitEvent(pipe.s_in, pipe.out_done[b], 0);                           // buffer drained
            if (e != cudaSuccess) return bail(cuda_fail(e, "host pipeline: wait for the staging buffer"));
        }
        e = cudaMemcpyAsync(pipe.dev[b], src + (size_t)e0 * esize, (size_t)ne * esize, cudaMemcpyHostToDevice, pipe.s_in);
        if (e != cudaSuccess) return bail(cuda_fail(e, "H2D copy"));
        e = cudaEventRecord(pipe.in_done[b], pipe.s_in);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(pipe.s_k, pipe.in_done[b], 0);
        if (e != cudaSuccess) return bail(cuda_fail(e, "host pipeline: H2D -> kernel ordering"));
        st = kernel(pipe.dev[b], ne, (uint64_t)e0, (void*)pipe.s_k);
        if (st != LOF_OK) return bail(st);
        e = cudaEventRecord(pipe.k_done[b], pipe.s_k);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(pipe.s_out, pipe.k_done[b], 0);
        if (e != cudaSuccess) return bail(cuda_fail(e, "host pipeline: kernel -> D2H ordering"));
        e = cudaMemcpyAsync(dst + (size_t)e0 * esize, pipe.dev[b], (size_t)ne * esize, cudaMemcpyDeviceToHost, pipe.s_out);
        if (e != cudaSuccess) return bail(cuda_fail(e, "D2H copy"));
        e = cudaEventRecord(pipe.out_done[b], pipe.s_out);
        if (e != cudaSuccess) return bail(cuda_fail(e, "host pipeline: D2H event"));
    }
    e = cudaStreamSynchronize(pipe.s_out);
    if (e == cudaSuccess) e = cudaStreamSynchronize(pipe.s_k);
    if (e == cudaSuccess) e = cudaStreamSynchronize(pipe.s_in);
    if (e != cudaSuccess) return cuda_fail(e, "host pipeline sync");
    return LOF_OK;
}

int mx_common(const float* in, float* out, int64_t n, int block_size, const lof_format_t* scale_fmt,
              const lof_format_t* elem_fmt, const lof_projspec_t* ps_scale, const lof_projspec_t* ps_elem,
              const lof_rng_t* rng, uint64_t first_index, void* stream)
{
    if (n < 0) return fail(LOF_ERR_INVALID_ARG, "n < 0");
    if (!ps_scale || !ps_elem) return fail(LOF_ERR_INVALID_ARG, "projspec is NULL");
    if (block_size <= 0) return fail(LOF_ERR_INVALID_ARG, "virtual_mx_round: block_size must be positive");        // LoPy_bind.cpp:173
    if (block_size != 16 && block_size != 32 && block_size != 64 && block_size != 128 && block_size != 256)
        return fail(LOF_ERR_UNSUPPORTED, "virtual_mx_round: unsupported block_size (use 16/32/64/128/256)");        // Lof_kernel.cu:90
    if (first_index % (uint64_t)block_size != 0) return fail(LOF_ERR_INVALID_ARG, "first_index must be a multiple of block_size");
    MxArgs a;
    std::memset(&a, 0, sizeof a);
    int st = make_plan(elem_fmt, LOF_F32, ps_elem->rounding_mode, ps_elem->stoch_length, &a.elem);
    if (st != LOF_OK) return st;
    st = make_plan(scale_fmt, LOF_F32, ps_scale->rounding_mode, 0, &a.scale);
    if (st != LOF_OK) return st;
    if (g_opt.round_fast.load()) a.fast = make_fast_plan(a.elem, ps_elem->rounding_mode);
    a.pmn = lof_mx_priv_max_normal(elem_fmt);
    a.scale_mode = ps_scale->rounding_mode;
    a.block_size = block_size;
    if (n == 0) return LOF_OK;
    if (!in || !out) return fail(LOF_ERR_INVALID_ARG, "null tensor pointer");
    MxLaunchCfg cfg;
    cfg.variant = (int)g_opt.mx_variant.load();
    cfg.ctas_per_sm = (int)g_opt.mx_ctas_per_sm.load();
    cfg.num_sms = num_sms();
    if (cfg.num_sms <= 0) return fail(LOF_ERR_NO_DEVICE, "no CUDA device (this library has no CPU fallback)");
    const RngArgs re = make_rng(rng, kStreamElem, ps_elem->rounding_mode, ps_elem->stoch_length), rs = make_rng(rng, kStreamScale);
    int launches = 0;
    cudaError_t e;
    bool handled = false;
    cudaStream_t cs = static_cast<cudaStream_t>(stream);
    // variant 0 (default): second-generation fast kernel for the eligible bulk, generic kernels for the rest
    // variant 1: TMA-staged kernel; variant 2: first-generation shuffle kernel only (comparison / fallback)
    int64_t done = 0;
    const bool elem_sr_fast = ps_elem->rounding_mode == LOF_StochasticRoundingA || ps_elem->rounding_mode == LOF_StochasticRoundingD;
    if (cfg.variant == 0 && g_opt.round_fast.load() >= 2 && (!mode_uses_rng(ps_elem->rounding_mode) || elem_sr_fast) &&
        !mode_uses_rng(ps_scale->rounding_mode)) {
        const MagicPlan em = make_magic_plan(a.elem, ps_elem->rounding_mode);
        const MagicPlan sm = make_magic_plan(a.scale, ps_scale->rounding_mode);
        const FastPlan sf = make_fast_plan(a.scale, ps_scale->rounding_mode);
        e = launch_mx_fast(ps_elem->rounding_mode, ps_scale->rounding_mode, in, out, n, a, em, sm, sf, scale_fmt->mantissa_bits,
                           cfg, cs, &done, &launches, &re, first_index);
        if (e != cudaSuccess) return cuda_fail(e, "virtual_mx_round (fast) launch");
        if (done == n) handled = true;
    }
    if (!handled && cfg.variant == 1 && done == 0) {
        e = launch_mx_tma(ps_elem->rounding_mode, in, out, n, a, re, rs, first_index, cfg, cs, &launches, &handled);
        if (e != cudaSuccess) return cuda_fail(e, "virtual_mx_round (tma) launch");
    }
    if (!handled) {
        // `done` is a multiple of 256 (whole blocks for every supported block size)
        e = launch_mx_shuffle(ps_elem->rounding_mode, in + done, out + done, n - done, a, re, rs, first_index + (uint64_t)done, cfg,
                              cs, &launches);
        if (e != cudaSuccess) return cuda_fail(e, "virtual_mx_round launch");
    }
    g_launches.fetch_add((uint64_t)launches);
    return LOF_OK;
}

}  // namespace

extern "C" {

int lof_abi_version(void) { return LOFLOAT_B200_ABI_VERSION; }

const char* lof_status_string(int status)
{
    switch (status) {
    case LOF_OK: return "ok";
    case LOF_ERR_INVALID_ARG: return "invalid argument";
    case LOF_ERR_UNSUPPORTED: return "unsupported configuration";
    case LOF_ERR_CUDA: return "CUDA error";
    case LOF_ERR_NO_DEVICE: return "no CUDA device (no CPU fallback)";
    case LOF_ERR_ALLOC: return "allocation failed";
    default: return "unknown status";
    }
}

const char* lof_last_error_string(void) { return g_last_error.c_str(); }

uint64_t lof_kernel_launch_count(void) { return g_launches.load(); }

float lof_mx_priv_max_normal(const lof_format_t* p)
{
    if (!p) return 0.0f;
    const int ebits = p->bitwidth - p->mantissa_bits - (p->is_signed == LOF_Signed ? 1 : 0);
    const int me = (1 << ebits) - 1 - p->bias;
    return std::ldexp(1.0f, me) * (2.0f - std::ldexp(1.0f, -p->mantissa_bits));     // LoPy_bind.cpp:199-202
}

int lof_selftest_mx_division(const lof_format_t* elem_fmt, uint64_t* mismatches, void* stream)
{
    if (!elem_fmt || !mismatches) return fail(LOF_ERR_INVALID_ARG, "null argument");
    if (num_sms() <= 0) return fail(LOF_ERR_NO_DEVICE, "no CUDA device (this library has no CPU fallback)");
    unsigned long long bad = 0;
    cudaError_t e = run_mx_div_selftest(lof_mx_priv_max_normal(elem_fmt), &bad, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "mx division self-test");
    *mismatches = bad;
    g_launches.fetch_add(1);
    return LOF_OK;
}

int lof_set_option(const char* key, int64_t value)
{
    if (!key) return fail(LOF_ERR_INVALID_ARG, "option key is NULL");
    const std::string k(key);
    if (k == "round.unroll") { if (value != 0 && value != 1 && value != 2 && value != 4 && value != 8) return fail(LOF_ERR_INVALID_ARG, "round.unroll must be 0 (auto) or 1/2/4/8"); g_opt.round_unroll = value; }
    else if (k == "round.ctas_per_sm") { if (value < 0 || value > 64) return fail(LOF_ERR_INVALID_ARG, "round.ctas_per_sm out of range"); g_opt.round_ctas_per_sm = value; }
    else if (k == "mx.variant") { if (value < 0 || value > 2) return fail(LOF_ERR_INVALID_ARG, "mx.variant must be 0, 1 or 2"); g_opt.mx_variant = value; }
    else if (k == "mx.ctas_per_sm") { if (value < 1 || value > 64) return fail(LOF_ERR_INVALID_ARG, "mx.ctas_per_sm out of range"); g_opt.mx_ctas_per_sm = value; }
    else if (k == "round.fast") { if (value < 0 || value > 2) return fail(LOF_ERR_INVALID_ARG, "round.fast must be 0, 1 or 2"); g_opt.round_fast = value; }
    else if (k == "stream.ctas_per_sm") { if (value < 1 || value > 64) return fail(LOF_ERR_INVALID_ARG, "stream.ctas_per_sm out of range"); g_opt.stream_ctas_per_sm = value; }
    else if (k == "pack.fast") { if (value < 0 || value > 3) return fail(LOF_ERR_INVALID_ARG, "pack.fast must be 0 .. 3"); g_opt.pack_fast = value; }
    else if (k == "bf16.fast") { if (value < 1 || value > 2) return fail(LOF_ERR_INVALID_ARG, "bf16.fast must be 1 or 2"); g_opt.bf16_fast = value; }
    else if (k == "gemm.mx_prof") g_opt.gemm_mx_prof = value;
    else if (k == "gemm.fold_split") { if (value < 1 || value > 2) return fail(LOF_ERR_INVALID_ARG, "gemm.fold_split must be 1 or 2"); g_opt.gemm_fold_split = value; }
    else if (k == "gemm.fast") { if (value < 0 || value > 2) return fail(LOF_ERR_INVALID_ARG, "gemm.fast must be 0, 1 or 2"); g_opt.gemm_fast = value; }
    else if (k == "host.chunk_bytes") { if (value < (1 << 16)) return fail(LOF_ERR_INVALID_ARG, "host.chunk_bytes too small"); g_opt.host_chunk_bytes = value; }
    else return fail(LOF_ERR_INVALID_ARG, "unknown option: " + k);
    return LOF_OK;
}

int lof_get_option(const char* key, int64_t* value)
{
    if (!key || !value) return fail(LOF_ERR_INVALID_ARG, "null argument");
    const std::string k(key);
    if (k == "round.unroll") *value = g_opt.round_unroll;
    else if (k == "round.ctas_per_sm") *value = g_opt.round_ctas_per_sm;
    else if (k == "mx.variant") *value = g_opt.mx_variant;
    else if (k == "mx.ctas_per_sm") *value = g_opt.mx_ctas_per_sm;
    else if (k == "round.fast") *value = g_opt.round_fast;
    else if (k == "stream.ctas_per_sm") *value = g_opt.stream_ctas_per_sm;
    else if (k == "pack.fast") *value = g_opt.pack_fast;
    else if (k == "gemm.fast") *value = g_opt.gemm_fast;
    else if (k == "gemm.mx_prof") *value = g_opt.gemm_mx_prof;
    else if (k == "gemm.fold_split") *value = g_opt.gemm_fold_split;
    else if (k == "bf16.fast") *value = g_opt.bf16_fast;
    else if (k == "host.chunk_bytes") *value = g_opt.host_chunk_bytes;
    else return fail(LOF_ERR_INVALID_ARG, "unknown option: " + k);
    return LOF_OK;
}

int lof_round_params(const void* in, void* out, int64_t n, int dtype, const lof_format_t* fmt,
                     const lof_projspec_t* ps, double scale, const lof_rng_t* rng, uint64_t first_index, void* stream)
{
    return round_common(true, in, out, n, dtype, fmt, 0, ps, scale, rng, first_index, stream);
}

int lof_round_mantissa(const void* in, void* out, int64_t n, int dtype, int mantissa_bits, const lof_projspec_t* ps,
                       double scale, const lof_rng_t* rng, uint64_t first_index, void* stream)
{
    return round_common(false, in, out, n, dtype, nullptr, mantissa_bits, ps, scale, rng, first_index, stream);
}

int lof_mx_round_f32(const float* in, float* out, int64_t n, int block_size, const lof_format_t* scale_fmt,
                     const lof_format_t* elem_fmt, const lof_projspec_t* ps_scale, const lof_projspec_t* ps_elem,
                     const lof_rng_t* rng, uint64_t first_index, void* stream)
{
    return mx_common(in, out, n, block_size, scale_fmt, elem_fmt, ps_scale, ps_elem, rng, first_index, stream);
}

int lof_gemm_batched_f32(const float* A, const float* B, float* C, int64_t batch, int64_t M, int64_t N, int64_t K,
                         int64_t stride_a, int64_t stride_b, int64_t stride_c, int accum_mant_bits, const lof_projspec_t* ps,
                         int64_t kc, float alpha, float beta, const lof_rng_t* rng, void* stream)
{
    if (batch < 0 || M < 0 || N < 0 || K < 0) return fail(LOF_ERR_INVALID_ARG, "negative GEMM dimension");
    if (M > INT32_MAX || N > INT32_MAX || K > INT32_MAX) return fail(LOF_ERR_UNSUPPORTED, "GEMM dimension exceeds 2^31-1");
    if (batch > 65535) return fail(LOF_ERR_UNSUPPORTED, "batched GEMM: at most 65535 matrices per call");
    if (!ps) return fail(LOF_ERR_INVALID_ARG, "projspec is NULL");
    const int mode = ps->rounding_mode;
    if (mode < 0 || mode >= LOF_NUM_ROUNDING_MODES) return fail(LOF_ERR_INVALID_ARG, "unknown rounding mode");
    if (accum_mant_bits > 23) return fail(LOF_ERR_INVALID_ARG, "accum_mant_bits > 23");
    if (accum_mant_bits >= 0 && accum_mant_bits < 23 &&
        (mode == LOF_StochasticRoundingA || mode == LOF_StochasticRoundingB || mode == LOF_StochasticRoundingC ||
         mode == LOF_StochasticRoundingD) &&
        (ps->stoch_length < 0 || ps->stoch_length > 23 - accum_mant_bits))
        return fail(LOF_ERR_INVALID_ARG, "stochastic_rounding_bits must be in [0, 23 - accum_mant_bits]");
    if (batch == 0 || M == 0 || N == 0) return LOF_OK;
    if (!C || (K > 0 && (!A || !B))) return fail(LOF_ERR_INVALID_ARG, "null matrix pointer");
    if (num_sms() <= 0) return fail(LOF_ERR_NO_DEVICE, "no CUDA device (this library has no CPU fallback)");
    cudaError_t e = launch_gemm_exact(mode, A, B, C, M, N, K, accum_mant_bits, ps->stoch_length, kc, alpha, beta,
                                      make_rng(rng, kStreamGemm), static_cast<cudaStream_t>(stream), batch, stride_a, stride_b, stride_c, nullptr,
                                      (int)g_opt.gemm_fast.load());
    if (e != cudaSuccess) return cuda_fail(e, "lof_gemm_batched launch");
    g_launches.fetch_add(1);
    return LOF_OK;
}

int lof_gemm_f32(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K, int accum_mant_bits,
                 const lof_projspec_t* ps, int64_t kc, float alpha, float beta, const lof_rng_t* rng, void* stream)
{
    if (M < 0 || N < 0 || K < 0) return fail(LOF_ERR_INVALID_ARG, "negative GEMM dimension");
    if (M > INT32_MAX || N > INT32_MAX || K > INT32_MAX) return fail(LOF_ERR_UNSUPPORTED, "GEMM dimension exceeds 2^31-1");
    if (!ps) return fail(LOF_ERR_INVALID_ARG, "projspec is NULL");
    const int mode = ps->rounding_mode;
    if (mode < 0 || mode >= LOF_NUM_ROUNDING_MODES) return fail(LOF_ERR_INVALID_ARG, "unknown rounding mode");
    if (accum_mant_bits > 23) return fail(LOF_ERR_INVALID_ARG, "accum_mant_bits > 23");
    if (accum_mant_bits >= 0 && accum_mant_bits < 23 &&
        (mode == LOF_StochasticRoundingA || mode == LOF_StochasticRoundingB || mode == LOF_StochasticRoundingC ||
         mode == LOF_StochasticRoundingD) &&
        (ps->stoch_length < 0 || ps->stoch_length > 23 - accum_mant_bits))
        return fail(LOF_ERR_INVALID_ARG, "stochastic_rounding_bits must be in [0, 23 - accum_mant_bits]");
    if (M == 0 || N == 0) return LOF_OK;
    if (!C || (K > 0 && (!A || !B))) return fail(LOF_ERR_INVALID_ARG, "null matrix pointer");
    if (num_sms() <= 0) return fail(LOF_ERR_NO_DEVICE, "no CUDA device (this library has no CPU fallback)");
    cudaError_t e = launch_gemm_exact(mode, A, B, C, M, N, K, accum_mant_bits, ps->stoch_length, kc, alpha, beta,
                                      make_rng(rng, kStreamGemm), static_cast<cudaStream_t>(stream), 1, 0, 0, 0, nullptr, (int)g_opt.gemm_fast.load());
    if (e != cudaSuccess) return cuda_fail(e, "lof_gemm launch");
    g_launches.fetch_add(1);
    return LOF_OK;
}

namespace {
// argument checks shared by the exact-GEMM entry points
int check_exact_args(int accum_mant_bits, const lof_projspec_t* ps)
{
    if (!ps) return fail(LOF_ERR_INVALID_ARG, "projspec is NULL");
    const int mode = ps->rounding_mode;
    if (mode < 0 || mode >= LOF_NUM_ROUNDING_MODES) return fail(LOF_ERR_INVALID_ARG, "unknown rounding mode");
    if (accum_mant_bits > 23) return fail(LOF_ERR_INVALID_ARG, "accum_mant_bits > 23");
    if (accum_mant_bits >= 0 && accum_mant_bits < 23 &&
        (mode == LOF_StochasticRoundingA || mode == LOF_StochasticRoundingB || mode == LOF_StochasticRoundingC ||
         mode == LOF_StochasticRoundingD) &&
        (ps->stoch_length < 0 || ps->stoch_length > 23 - accum_mant_bits))
        return fail(LOF_ERR_INVALID_ARG, "stochastic_rounding_bits must be in [0, 23 - accum_mant_bits]");
    return LOF_OK;
}
}  // namespace

int lof_linear_f32(const float* x, const float* weight, const float* bias, float* out, int64_t M, int64_t N, int64_t K,
                   int accum_mant_bits, const lof_projspec_t* ps, int64_t kc, float alpha, const lof_rng_t* rng, void* stream)
{
    if (M < 0 || N < 0 || K < 0) return fail(LOF_ERR_INVALID_ARG, "negative GEMM dimension");
    if (M > INT32_MAX || N > INT32_MAX || K > INT32_MAX) return fail(LOF_ERR_UNSUPPORTED, "GEMM dimension exceeds 2^31-1");
    if (const int st = check_exact_args(accum_mant_bits, ps)) return st;
    if (M == 0 || N == 0) return LOF_OK;
    if (!out || (K > 0 && (!x || !weight))) return fail(LOF_ERR_INVALID_ARG, "null matrix pointer");
    if (num_sms() <= 0) return fail(LOF_ERR_NO_DEVICE, "no CUDA device (this library has no CPU fallback)");
    ExactView v;
    v.b_kind = 1;
    v.bias = bias;
    v.bias_axis = 1;
    cudaError_t e = launch_gemm_exact(ps->rounding_mode, x, weight, out, M, N, K, accum_mant_bits, ps->stoch_length, kc, alpha, 0.0f,
                                      make_rng(rng, kStreamGemm), static_cast<cudaStream_t>(stream), 1, 0, 0, 0, &v, (int)g_opt.gemm_fast.load());
    if (e != cudaSuccess) return cuda_fail(e, "lof_linear launch");
    g_launches.fetch_add(1);
    return LOF_OK;
}

int lof_selftest_gemm_rne_round(uint64_t* dev_mismatches, void* stream)
{
    if (!dev_mismatches) return fail(LOF_ERR_INVALID_ARG, "dev_mismatches is NULL");
    const int sms = num_sms();
    if (sms <= 0) return fail(LOF_ERR_NO_DEVICE, "no CUDA device (this library has no CPU fallback)");
    cudaError_t e = launch_selftest_rne_round(reinterpret_cast<unsigned long long*>(dev_mismatches), sms, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "lof_selftest_gemm_rne_round launch");
    g_launches.fetch_add(1);
    return LOF_OK;
}

int lof_conv2d_out_size(int64_t in, int64_t kernel, int64_t pad, int64_t stride, int64_t dilation, int64_t* out)
{
    if (!out) return fail(LOF_ERR_INVALID_ARG, "out is NULL");
    if (in <= 0 || kernel <= 0 || pad < 0 || stride <= 0 || dilation <= 0) return fail(LOF_ERR_INVALID_ARG, "conv2d: bad geometry");
    const int64_t span = in + 2 * pad - dilation * (kernel - 1) - 1;
    *out = span < 0 ? 0 : span / stride + 1;
    return LOF_OK;
}

int lof_conv2d_f32(const float* input, const float* weight, const float* bias, float* out, const lof_conv2d_t* g,
                   int accum_mant_bits, const lof_projspec_t* ps, int64_t kc, float alpha, const lof_rng_t* rng, void* stream)
{
    if (!g) return fail(LOF_ERR_INVALID_ARG, "conv2d geometry is NULL");
    if (g->batch < 0 || g->c_in <= 0 || g->c_out <= 0 || g->h <= 0 || g->w <= 0 || g->r <= 0 || g->s <= 0 || g->groups <= 0 ||
        g->pad_h < 0 || g->pad_w < 0 || g->stride_h <= 0 || g->stride_w <= 0 || g->dilation_h <= 0 || g->dilation_w <= 0)
        return fail(LOF_ERR_INVALID_ARG, "conv2d: bad geometry");
    if (g->c_in % g->groups || g->c_out % g->groups) return fail(LOF_ERR_INVALID_ARG, "conv2d: groups must divide C_in and C_out");
    if (g->k_order != LOF_CONV_K_CRS && g->k_order != LOF_CONV_K_RSC) return fail(LOF_ERR_INVALID_ARG, "conv2d: unknown k_order");
    if (const int st = check_exact_args(accum_mant_bits, ps)) return st;
    int64_t P = 0, Q = 0;
    lof_conv2d_out_size(g->h, g->r, g->pad_h, g->stride_h, g->dilation_h, &P);
    lof_conv2d_out_size(g->w, g->s, g->pad_w, g->stride_w, g->dilation_w, &Q);
    if (P <= 0 || Q <= 0) return fail(LOF_ERR_INVALID_ARG, "LoF_conv2d: non-positive output spatial size");   // Lof_kernel.cu:196
    const int64_t Cg = g->c_in / g->groups, Mg = g->c_out / g->groups, K = Cg * g->r * g->s, L = P * Q;
    if (K > INT32_MAX || L > INT32_MAX || g->c_in * g->h * g->w > INT32_MAX || g->pad_h > (1 << 20) || g->pad_w > (1 << 20))
        return fail(LOF_ERR_UNSUPPORTED, "conv2d: image or reduction extent exceeds 2^31-1");
    if (g->batch * g->groups > 65535) return fail(LOF_ERR_UNSUPPORTED, "conv2d: at most 65535 (image, group) pairs per call");
    if (g->batch == 0) return LOF_OK;
    if (!input || !weight || !out) return fail(LOF_ERR_INVALID_ARG, "null tensor pointer");
    if (num_sms() <= 0) return fail(LOF_ERR_NO_DEVICE, "no CUDA device (this library has no CPU fallback)");
    ExactView v;
    v.b_kind = 2;
    v.k_order = g->k_order;
    v.bias = bias;
    v.bias_axis = 0;
    v.C = (int)g->c_in; v.H = (int)g->h; v.W = (int)g->w; v.Cg = (int)Cg; v.R = (int)g->r; v.S = (int)g->s; v.Q = (int)Q;
    v.ph = (int)g->pad_h; v.pw = (int)g->pad_w; v.sh = (int)g->stride_h; v.sw = (int)g->stride_w;
    v.dh = (int)g->dilation_h; v.dw = (int)g->dilation_w; v.groups = (int)g->groups;
    cudaError_t e = launch_gemm_exact(ps->rounding_mode, weight, input, out, Mg, L, K, accum_mant_bits, ps->stoch_length, kc, alpha, 0.0f,
                                      make_rng(rng, kStreamGemm), static_cast<cudaStream_t>(stream), g->batch * g->groups, 0, 0, 0, &v,
                                      (int)g_opt.gemm_fast.load());
    if (e != cudaSuccess) return cuda_fail(e, "lof_conv2d launch");
    g_launches.fetch_add(1);
    return LOF_OK;
}

namespace {
int gemm_chunked_common(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K,
                        int accum_mant_bits, const lof_format_t* acc_fmt, const lof_projspec_t* ps, int64_t k_chunk,
                        float alpha, int operand_dtype, uint32_t flags, const lof_rng_t* rng, int32_t* dev_status, void* stream)
{
    if (M <= 0 || N <= 0 || K <= 0) return fail(LOF_ERR_INVALID_ARG, "GEMM dimensions must be positive");
    if (!A || !B || !C || !ps) return fail(LOF_ERR_INVALID_ARG, "null argument");
    const int mode = ps->rounding_mode;
    if (mode < 0 || mode >= LOF_NUM_ROUNDING_MODES) return fail(LOF_ERR_INVALID_ARG, "unknown rounding mode");
    RoundPlan plan;
    FastPlan fast;
    std::memset(&plan, 0, sizeof plan);
    std::memset(&fast, 0, sizeof fast);
    if (acc_fmt) {
        const int st = make_plan(acc_fmt, LOF_F32, mode, 0, &plan);
        if (st != LOF_OK) return st;
        if (g_opt.round_fast.load()) fast = make_fast_plan(plan, mode);
    } else if (accum_mant_bits < 0 || accum_mant_bits > 23) {
        return fail(LOF_ERR_INVALID_ARG, "accum_mant_bits must be in [0, 23]");
    }
    if (mode_uses_rng(mode) && !acc_fmt) {
        const int r = 23 - accum_mant_bits;                 // the reference's precondition for the shifted-offset modes (lo_float.h:653-860)
        if (r > 0 && (mode == LOF_StochasticRoundingA || mode == LOF_StochasticRoundingB || mode == LOF_StochasticRoundingC ||
                      mode == LOF_StochasticRoundingD) && (ps->stoch_length < 1 || ps->stoch_length > r))
            return fail(LOF_ERR_INVALID_ARG, "stochastic rounding needs 1 <= stoch_length <= dropped mantissa bits");
    }
    if (num_sms() <= 0) return fail(LOF_ERR_NO_DEVICE, "no CUDA device (this library has no CPU fallback)");
    int launches = 0;
    std::string err;
    const RngArgs ra = make_rng(rng, kStreamGemm);
    cudaError_t e = launch_gemm_tc(A, B, C, M, N, K, mode, accum_mant_bits, acc_fmt, acc_fmt ? &plan : nullptr,
                                   acc_fmt ? &fast : nullptr, k_chunk, alpha, operand_dtype,
                                   flags | (g_opt.gemm_fold_split.load() == 2 ? LOF_GEMM_INTERNAL_FOLD_SPLIT2 : 0u), dev_status,
                                   static_cast<cudaStream_t>(stream), &launches, &err, &ra, ps->stoch_length);
    g_launches.fetch_add((uint64_t)launches);
    if (!err.empty()) return fail(LOF_ERR_UNSUPPORTED, err);
    if (e != cudaSuccess) return cuda_fail(e, "lof_gemm_chunked launch");
    return LOF_OK;
}
}  // namespace

int lof_gemm_chunked_f32(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K,
                         int accum_mant_bits, const lof_format_t* acc_fmt, const lof_projspec_t* ps, int64_t k_chunk,
                         float alpha, int operand_dtype, uint32_t flags, int32_t* dev_status, void* stream)
{
    return gemm_chunked_common(A, B, C, M, N, K, accum_mant_bits, acc_fmt, ps, k_chunk, alpha, operand_dtype, flags, nullptr, dev_status, stream);
}

int lof_gemm_chunked_rng_f32(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K,
                             int accum_mant_bits, const lof_format_t* acc_fmt, const lof_projspec_t* ps, int64_t k_chunk,
                             float alpha, int operand_dtype, uint32_t flags, const lof_rng_t* rng, int32_t* dev_status, void* stream)
{
    return gemm_chunked_common(A, B, C, M, N, K, accum_mant_bits, acc_fmt, ps, k_chunk, alpha, operand_dtype, flags, rng, dev_status, stream);
}

int lof_gemm_chunked_packed(const uint8_t* A, const uint8_t* Bt, float* C, int64_t M, int64_t N, int64_t K,
                            int accum_mant_bits, const lof_format_t* acc_fmt, const lof_projspec_t* ps, int64_t k_chunk,
                            float alpha, int operand_dtype, uint32_t flags, int32_t* dev_status, void* stream)
{
    if (operand_dtype != LOF_E4M3 && operand_dtype != LOF_E5M2 && operand_dtype != LOF_BF16)
        return fail(LOF_ERR_INVALID_ARG, "lof_gemm_chunked_packed: operand_dtype must be LOF_E4M3, LOF_E5M2 or LOF_BF16");
    if (flags & (LOF_GEMM_INTERNAL_PACKED_OPERANDS | LOF_GEMM_INTERNAL_FOLD_SPLIT2)) return fail(LOF_ERR_INVALID_ARG, "unknown flag");
    return lof_gemm_chunked_f32(reinterpret_cast<const float*>(A), reinterpret_cast<const float*>(Bt), C, M, N, K, accum_mant_bits,
                                acc_fmt, ps, k_chunk, alpha, operand_dtype, flags | LOF_GEMM_INTERNAL_PACKED_OPERANDS, dev_status, stream);
}

extern "C" LOF_API int lof_debug_mx_prof(unsigned long long* out16)     // not in the public header: profiling aid (scripts/mx_gemm_check.py)
{
    cudaDeviceSynchronize();
    const cudaError_t e = read_mx_prof(out16);
    return e == cudaSuccess ? LOF_OK : cuda_fail(e, "lof_debug_mx_prof");
}

int lof_p2p_post(int32_t* my_flags, int epoch, int expected_done, void* stream)
{
    if (!my_flags) return fail(LOF_ERR_INVALID_ARG, "null flag block");
    const cudaError_t e = launch_p2p_post(my_flags, epoch, expected_done, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "lof_p2p_post launch");
    g_launches.fetch_add(1);
    return LOF_OK;
}
int lof_p2p_wait(const int32_t* peer_flags, int32_t* my_flags, int epoch, void* stream)
{
    if (!peer_flags || !my_flags) return fail(LOF_ERR_INVALID_ARG, "null flag block");
    const cudaError_t e = launch_p2p_wait(peer_flags, my_flags, epoch, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "lof_p2p_wait launch");
    g_launches.fetch_add(1);
    return LOF_OK;
}
int lof_p2p_done(int32_t* peer_flags, void* stream)
{
    if (!peer_flags) return fail(LOF_ERR_INVALID_ARG, "null flag block");
    const cudaError_t e = launch_p2p_done(peer_flags, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "lof_p2p_done launch");
    g_launches.fetch_add(1);
    return LOF_OK;
}

int lof_p2p_enable_peer(int peer_device)
{
    int cur = -1;
    cudaGetDevice(&cur);
    if (peer_device == cur) return LOF_OK;
    int can = 0;
    cudaError_t e = cudaDeviceCanAccessPeer(&can, cur, peer_device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceCanAccessPeer");
    if (!can) return fail(LOF_ERR_UNSUPPORTED, "lof_p2p_enable_peer: the two devices have no peer access");
    e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return LOF_OK; }
    if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceEnablePeerAccess");
    return LOF_OK;
}
int lof_p2p_copy(void* dst, const void* src, int64_t bytes, void* stream)
{
    if (bytes < 0 || (bytes > 0 && (!dst || !src))) return fail(LOF_ERR_INVALID_ARG, "lof_p2p_copy: bad arguments");
    if (bytes == 0) return LOF_OK;
    const cudaError_t e = cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "lof_p2p_copy");
    return LOF_OK;
}

int lof_p2p_export(const void* ptr, uint8_t handle[64], uint64_t* offset)
{
    if (!ptr || !handle || !offset) return fail(LOF_ERR_INVALID_ARG, "lof_p2p_export: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    CUdeviceptr base = 0;
    size_t size = 0;
    // the driver entry point is looked up at run time: the library must keep loading on hosts without libcuda.so.1 (CPU test boxes)
    using RangeFn = CUresult (*)(CUdeviceptr*, size_t*, CUdeviceptr);
    static RangeFn range_fn = nullptr;
    if (!range_fn) {
        void* fp = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fp, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !fp)
            return fail(LOF_ERR_NO_DEVICE, "lof_p2p_export: cuMemGetAddressRange is not available");
        range_fn = reinterpret_cast<RangeFn>(fp);
    }
    if (range_fn(&base, &size, reinterpret_cast<CUdeviceptr>(ptr)) != CUDA_SUCCESS)
        return fail(LOF_ERR_INVALID_ARG, "lof_p2p_export: not a device allocation");
    cudaIpcMemHandle_t h;
    const cudaError_t e = cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(base));
    if (e != cudaSuccess) return cuda_fail(e, "cudaIpcGetMemHandle");
    std::memcpy(handle, &h, 64);
    *offset = (uint64_t)(reinterpret_cast<CUdeviceptr>(ptr) - base);
    return LOF_OK;
}
int lof_p2p_import(const uint8_t handle[64], uint64_t offset, void** ptr)
{
    if (!handle || !ptr) return fail(LOF_ERR_INVALID_ARG, "lof_p2p_import: null argument");
    static std::mutex mu;
    static std::vector<std::pair<std::string, void*>> opened;          // a handle can be opened once per process: cache by its bytes
    std::lock_guard<std::mutex> lock(mu);
    const std::string key(reinterpret_cast<const char*>(handle), 64);
    for (auto& kv : opened)
        if (kv.first == key) { *ptr = static_cast<uint8_t*>(kv.second) + offset; return LOF_OK; }
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, 64);
    void* base = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return cuda_fail(e, "cudaIpcOpenMemHandle");
    opened.emplace_back(key, base);
    *ptr = static_cast<uint8_t*>(base) + offset;
    return LOF_OK;
}

int lof_mx_pack_tc_f32(const float* in, int64_t n, uint8_t* elems, uint8_t* scales, int32_t* dev_status, void* stream)
{
    if (n < 0 || n % 32 != 0) return fail(LOF_ERR_INVALID_ARG, "lof_mx_pack_tc_f32: n must be a non-negative multiple of 32");
    if (n == 0) return LOF_OK;
    if (!in || !elems || !scales) return fail(LOF_ERR_INVALID_ARG, "null tensor pointer");
    if ((reinterpret_cast<uintptr_t>(in) & 15u) || (reinterpret_cast<uintptr_t>(elems) & 3u))
        return fail(LOF_ERR_INVALID_ARG, "lof_mx_pack_tc_f32: `in` must be 16-byte and `elems` 4-byte aligned");
    const int sms = num_sms();
    if (sms <= 0) return fail(LOF_ERR_NO_DEVICE, "no CUDA device (this library has no CPU fallback)");
    cudaError_t e = launch_mx_pack_tc(in, n, elems, scales, dev_status, sms, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "lof_mx_pack_tc launch");
    g_launches.fetch_add(1);
    return LOF_OK;
}

int lof_mx_gemm(const uint8_t* a_elems, const uint8_t* a_scales, const uint8_t* bt_elems, const uint8_t* bt_scales, float* C,
                int64_t M, int64_t N, int64_t K, int elem_dtype, int accum_mant_bits, const lof_format_t* acc_fmt,
                const lof_projspec_t* ps, int64_t k_chunk, float alpha, uint32_t flags, void* stream)
{
    if (M <= 0 || N <= 0 || K <= 0) return fail(LOF_ERR_INVALID_ARG, "GEMM dimensions must be positive");
    if (!a_elems || !a_scales || !bt_elems || !bt_scales || !C || !ps) return fail(LOF_ERR_INVALID_ARG, "null argument");
    if (flags & ~(LOF_GEMM_PARTIAL_ROUND | LOF_GEMM_NO_PACKED)) return fail(LOF_ERR_INVALID_ARG, "lof_mx_gemm: unknown flag");
    const int mode = ps->rounding_mode;
    if (mode < 0 || mode >= LOF_NUM_ROUNDING_MODES) return fail(LOF_ERR_INVALID_ARG, "unknown rounding mode");
    if (mode_uses_rng(mode)) return fail(LOF_ERR_UNSUPPORTED, "MX GEMM: stochastic accumulator rounding not implemented");
    RoundPlan plan;
    FastPlan fast;
    std::memset(&plan, 0, sizeof plan);
    std::memset(&fast, 0, sizeof fast);
    if (acc_fmt) {
        const int st = make_plan(acc_fmt, LOF_F32, mode, 0, &plan);
        if (st != LOF_OK) return st;
        if (g_opt.round_fast.load()) fast = make_fast_plan(plan, mode);
    } else if (accum_mant_bits < 0 || accum_mant_bits > 23) {
        return fail(LOF_ERR_INVALID_ARG, "accum_mant_bits must be in [0, 23]");
    }
    if (num_sms() <= 0) return fail(LOF_ERR_NO_DEVICE, "no CUDA device (this library has no CPU fallback)");
    int launches = 0;
    std::string err;
    cudaError_t e = launch_gemm_mx(a_elems, a_scales, bt_elems, bt_scales, C, M, N, K, elem_dtype, mode, accum_mant_bits, acc_fmt,
                                   acc_fmt ? &plan : nullptr, acc_fmt ? &fast : nullptr, k_chunk, alpha,
                                   flags | (g_opt.gemm_fold_split.load() == 2 ? LOF_GEMM_INTERNAL_FOLD_SPLIT2 : 0u) |
                                       (g_opt.gemm_mx_prof.load() == 1 ? LOF_GEMM_INTERNAL_PROF : 0u),
                                   static_cast<cudaStream_t>(stream), &launches, &err);
    g_launches.fetch_add((uint64_t)launches);
    if (!err.empty()) return fail(LOF_ERR_UNSUPPORTED, err);
    if (e != cudaSuccess) return cuda_fail(e, "lof_mx_gemm launch");
    return LOF_OK;
}

// ---------------------------------------------------------------------------------------
// fused callers: Hadamard rotation + rounding (LoFloat/layers.py:182-216, 363-378)
// ---------------------------------------------------------------------------------------
namespace {
int fwht_check(const float* in, float* out, int64_t n, int block, int max_block = 1024)
{
    if (n < 0) return fail(LOF_ERR_INVALID_ARG, "n < 0");
    if (block <= 0 || (block & (block - 1)) != 0) return fail(LOF_ERR_INVALID_ARG, "fwht: block_size must be a power of 2");   // layers.py:196-197
    if (block > max_block)
        return fail(LOF_ERR_UNSUPPORTED, max_block > 1024 ? "fwht: block sizes up to 32768 are handled" : "fwht: the fused kernels handle block sizes up to 1024 "
                                                            "(transform with lof_fwht_f32, then round)");
    if (n % block != 0) return fail(LOF_ERR_INVALID_ARG, "fwht: block_size must divide the tensor's last dimension");          // layers.py:198-199
    if (n % 128 != 0) return fail(LOF_ERR_UNSUPPORTED, "fwht: fused kernel needs the element count to be a multiple of 128");
    if (n > 0 && (!in || !out)) return fail(LOF_ERR_INVALID_ARG, "null tensor pointer");
    if (((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15u) != 0)
        return fail(LOF_ERR_INVALID_ARG, "fwht: pointers must be 16-byte aligned");
    return LOF_OK;
}
}  // namespace

int lof_fwht_f32(const float* in, float* out, int64_t n, int block_size, void* stream)
{
    const int st = fwht_check(in, out, n, block_size, 32768);
    if (st != LOF_OK) return st;
    if (n == 0) return LOF_OK;
    const int sms = num_sms();
    if (sms <= 0) return fail(LOF_ERR_NO_DEVICE, "no CUDA device (this library has no CPU fallback)");
    if (block_size > 1024) {                                   // two passes: inside the 1024-element chunks, then across them
        cudaError_t e2 = launch_fwht_large(in, out, n, block_size, sms * (int)g_opt.stream_ctas_per_sm.load(), static_cast<cudaStream_t>(stream));
        if (e2 != cudaSuccess) return cuda_fail(e2, "lof_fwht (large block) launch");
        g_launches.fetch_add(2);
        return LOF_OK;
    }
    const RngArgs none = make_rng(nullptr, kStreamElem);
    cudaError_t e = launch_fwht(0, 0, in, out, n, block_size, nullptr, 1.0f, nullptr, nullptr, nullptr, nullptr, none, none, 0, sms * (int)g_opt.stream_ctas_per_sm.load(), static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "lof_fwht launch");
    g_launches.fetch_add(1);
    return LOF_OK;
}

int lof_fwht_round_f32(const float* in, float* out, int64_t n, int block_size, const lof_format_t* fmt, const lof_projspec_t* ps,
                       double scale, const lof_rng_t* rng, uint64_t first_index, void* stream)
{
    int st = fwht_check(in, out, n, block_size);
    if (st != LOF_OK) return st;
    if (!ps) return fail(LOF_ERR_INVALID_ARG, "projspec is NULL");
    const int mode = ps->rounding_mode;
    if (mode < 0 || mode >= LOF_NUM_ROUNDING_MODES) return fail(LOF_ERR_INVALID_ARG, "unknown rounding mode");
    if (first_index % 4 != 0) return fail(LOF_ERR_INVALID_ARG, "first_index must be a multiple of 4");
    RoundPlan plan;
    std::memset(&plan, 0, sizeof plan);
    st = make_plan(fmt, LOF_F32, mode, ps->stoch_length, &plan);
    if (st != LOF_OK) return st;
    if (n == 0) return LOF_OK;
    const int sms = num_sms();
    if (sms <= 0) return fail(LOF_ERR_NO_DEVICE, "no CUDA device (this library has no CPU fallback)");
    const RngArgs ra = make_rng(rng, kStreamElem, mode, ps->stoch_length);
    MagicPlan rmagic;
    std::memset(&rmagic, 0, sizeof rmagic);
    if (g_opt.round_fast.load() >= 2) rmagic = make_magic_plan(plan, mode);
    cudaError_t e = launch_fwht(1, mode, in, out, n, block_size, &plan, (float)scale, nullptr, nullptr, nullptr, &rmagic, ra, ra, first_index, sms * (int)g_opt.stream_ctas_per_sm.load(),
                                static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "lof_fwht_round launch");
    g_launches.fetch_add(1);
    return LOF_OK;
}

int lof_fwht_mx_round_f32(const float* in, float* out, int64_t n, int block_size, int mx_block_size, const lof_format_t* scale_fmt,
                          const lof_format_t* elem_fmt, const lof_projspec_t* ps_scale, const lof_projspec_t* ps_elem,
                          const lof_rng_t* rng, uint64_t first_index, void* stream)
{
    int st = fwht_check(in, out, n, block_size);
    if (st != LOF_OK) return st;
    if (!ps_scale || !ps_elem) return fail(LOF_ERR_INVALID_ARG, "projspec is NULL");
    if (mx_block_size != 16 && mx_block_size != 32 && mx_block_size != 64 && mx_block_size != 128)
        return fail(LOF_ERR_UNSUPPORTED, "fwht + mx: mx block_size must be 16, 32, 64 or 128");
    if (first_index % (uint64_t)mx_block_size != 0) return fail(LOF_ERR_INVALID_ARG, "first_index must be a multiple of the mx block_size");
    MxArgs a;
    std::memset(&a, 0, sizeof a);
    st = make_plan(elem_fmt, LOF_F32, ps_elem->rounding_mode, ps_elem->stoch_length, &a.elem);
    if (st != LOF_OK) return st;
    st = make_plan(scale_fmt, LOF_F32, ps_scale->rounding_mode, 0, &a.scale);
    if (st != LOF_OK) return st;
    if (g_opt.round_fast.load()) a.fast = make_fast_plan(a.elem, ps_elem->rounding_mode);
    a.pmn = lof_mx_priv_max_normal(elem_fmt);
    a.scale_mode = ps_scale->rounding_mode;
    a.block_size = mx_block_size;
    if (n == 0) return LOF_OK;
    const int sms = num_sms();
    if (sms <= 0) return fail(LOF_ERR_NO_DEVICE, "no CUDA device (this library has no CPU fallback)");
    const RngArgs re = make_rng(rng, kStreamElem, ps_elem->rounding_mode, ps_elem->stoch_length), rs = make_rng(rng, kStreamScale);
    const MagicPlan smg = make_magic_plan(a.scale, ps_scale->rounding_mode);
    const FastPlan sfp = make_fast_plan(a.scale, ps_scale->rounding_mode);
    // magic-add element rounding for power-of-two scales: same eligibility as the second-generation MX kernel (lof_mx_fast.cu)
    const MagicPlan emg = make_magic_plan(a.elem, ps_elem->rounding_mode);
    const bool elem_fast = g_opt.round_fast.load() >= 2 && emg.valid && emg.limit == 0xFFFFFFFFu && !emg.check_denorm_in &&
                           scale_fmt->mantissa_bits == 0 && !mode_uses_rng(ps_elem->rounding_mode);
    cudaError_t e = launch_fwht(2, ps_elem->rounding_mode, in, out, n, block_size, nullptr, 1.0f, &a, &smg, &sfp, elem_fast ? &emg : nullptr, re, rs, first_index, sms * (int)g_opt.stream_ctas_per_sm.load(),
                                static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "lof_fwht_mx_round launch");
    g_launches.fetch_add(1);
    return LOF_OK;
}

// ---------------------------------------------------------------------------------------
// storage codec (lo_pack.hpp)
// ---------------------------------------------------------------------------------------
namespace {
int pack_gcd(int a, int b) { return b == 0 ? a : pack_gcd(b, a % b); }
int pack_group_bits(int n) { const int al = n <= 8 ? 8 : 16; return (n / pack_gcd(n, al)) * al; }     // lo_pack.hpp:58-67

int make_codec_plan(const lof_codec_format_t* cf, const lof_projspec_t* ps, int endian, CodecPlan* p)
{
    if (!cf) return fail(LOF_ERR_INVALID_ARG, "codec format is NULL");
    const lof_format_t& f = cf->fmt;
    if (f.bitwidth < 1 || f.bitwidth > 16) return fail(LOF_ERR_INVALID_ARG, "codec: bitwidth must be in [1, 16] (lo_pack.hpp)");
    if (f.is_signed != LOF_Signed && f.is_signed != LOF_Unsigned) return fail(LOF_ERR_INVALID_ARG, "codec: bad signedness");
    const int to_signed = f.is_signed == LOF_Signed;
    const int ebits = f.bitwidth - f.mantissa_bits - (to_signed ? 1 : 0);
    if (f.mantissa_bits < 0 || ebits < 0 || ebits > 30) return fail(LOF_ERR_INVALID_ARG, "codec: malformed format");
    if (endian != LOF_PACK_FORWARD && endian != LOF_PACK_REVERSE) return fail(LOF_ERR_INVALID_ARG, "codec: bad endian");
    std::memset(p, 0, sizeof *p);
    const int n = f.bitwidth;
    p->n = n; p->mant = f.mantissa_bits; p->bias = f.bias; p->to_signed = to_signed;
    p->exp_offset = f.bias - 127;
    p->digit_drop = 23 - f.mantissa_bits;
    p->to_min_exp = 1 - f.bias;
    const int to_max_exp = (1 << ebits) - 1 - f.bias, to_digits = f.mantissa_bits + 1;
    p->narrower = (to_max_exp < 128) || (to_max_exp == 128 && to_digits < 24);
    p->has_inf = f.ov_behavior != LOF_Saturating;
    p->signmask = to_signed ? (1u << (n - 1)) : 0u;
    p->rep_mask = (1u << n) - 1u;
    p->inf_mask = cf->inf_mask; p->inf_value = cf->inf_value;
    p->nan_mask = cf->nan_mask; p->nan_value = cf->nan_value; p->nan_mant_mask = cf->nan_mant_mask;
    p->sign_is_nan = (p->signmask & cf->nan_mask) == cf->nan_value && (cf->nan_mant_mask == 0u || (p->signmask & cf->nan_mant_mask) != 0u);
    p->max_finite_rep = p->has_inf ? (cf->min_pos_inf - 1u) : (to_signed ? (((1u << n) - 1u) >> 1) : ((1u << n) - 1u));
    const uint32_t inf_rep = p->has_inf ? cf->min_pos_inf : p->max_finite_rep;
    const int sat = ps ? ps->saturation_mode : LOF_OvfInf;
    p->inf_in_rep = (sat != LOF_SatFinite && p->has_inf) ? inf_rep : p->max_finite_rep;
    p->ovf_rep = (sat == LOF_OvfInf && p->has_inf) ? inf_rep : p->max_finite_rep;
    p->neg_unsigned_rep = (f.unsigned_behavior == LOF_NegtoZero) ? 0u : (f.na_behavior == LOF_NaN_3109 ? cf->qnan : cf->snan);
    p->qnan = cf->qnan;
    p->len = ps ? ps->stoch_length : 0;
    p->gc = pack_group_bits(n) / n;
    p->reverse = endian == LOF_PACK_REVERSE;
    p->rev_xor = p->reverse ? (uint32_t)(p->gc - 1) : 0u;                   // gc = 8 / gcd(n, 8) or 16 / gcd(n, 16): a power of two
    p->recip_n = n > 1 ? (uint32_t)(((1ull << 32) + (uint64_t)n - 1) / (uint64_t)n) : 0u;
    // magic-add fast path: IEEE-like modes, >= 1 mantissa bit (with m = 0 the reference's RNE / round-to-odd break ties
    // on the fp32 EXPONENT's LSB), target range inside fp32's normal range (bias <= 125 keeps lo_pack.hpp's branch 2 in
    // charge of fp32 subnormals), every magic constant a normal fp32 number
    const int mode = ps ? ps->rounding_mode : LOF_RoundToNearestEven;
    const int emin_b = 128 - f.bias;                                        // biased fp32 exponent of 2^(1 - bias)
    const int ds = p->digit_drop;
    const bool mode_ok = mode == LOF_RoundToNearestEven || mode == LOF_RoundTowardsZero || mode == LOF_RoundAwayFromZero ||
                         mode == LOF_RoundToOdd;
    p->fast_ok = g_opt.pack_fast.load() && mode_ok && f.mantissa_bits >= 1 && f.bias <= 125 && emin_b >= 1 && emin_b + ds <= 254 &&
                 to_max_exp + 127 + ds <= 254 && to_max_exp + 128 <= 254 && p->narrower && p->max_finite_rep < (1u << 16);
    p->f_cadd = (uint32_t)ds << 23;
    p->f_cfloor = (uint32_t)(emin_b + ds) << 23;
    p->f_big = (uint32_t)(to_max_exp + 128) << 23;
    p->lanes = g_opt.pack_fast.load() >= 2 ? (g_opt.pack_fast.load() >= 3 ? 1u : 2u) : 0u;     // 2: + block-per-thread MX encode (default); pack.fast = 3 keeps the warp-step MX encode
    // arithmetic decode (lof_pack.cu decode_lane_kernel): every finite code's value is a normal fp32 number or zero, and the
    // smallest magnitude among the infinity / NaN codes (the sign-bit-only NaN of the P3109 formats is tested separately)
    const int max_ef = (1 << ebits) - 1;
    p->dec_ok = f.bias >= 0 && f.bias <= 127 && f.bias + f.mantissa_bits <= 127 && max_ef - f.bias + 127 <= 254;
    p->d_special = 0xFFFFFFFFu;
    if (p->dec_ok) {
        const uint32_t magmask = to_signed ? p->signmask - 1u : p->rep_mask;
        for (uint32_t rep = 0; rep <= p->rep_mask; ++rep) {
            const bool is_inf = p->has_inf && (rep & p->inf_mask) == p->inf_value;
            const bool is_nan = (rep & p->nan_mask) == p->nan_value && (p->nan_mant_mask == 0u || (rep & p->nan_mant_mask) != 0u);
            if (!is_inf && !is_nan) continue;
            if (to_signed && rep == p->signmask && p->sign_is_nan) continue;
            if ((rep & magmask) < p->d_special) p->d_special = rep & magmask;
        }
        if (p->d_special == 0u) p->dec_ok = 0;               // a special code of magnitude 0 other than the sign-only NaN: table path
    }
    return LOF_OK;
}
}  // namespace

uint64_t lof_packed_size(int bitwidth, uint64_t count)
{
    if (bitwidth < 1 || bitwidth > 16) return 0;
    const uint64_t gc = (uint64_t)(pack_group_bits(bitwidth) / bitwidth), gb = (uint64_t)(pack_group_bits(bitwidth) / 8);
    return ((count + gc - 1) / gc) * gb;
}

int lof_encode_f32(const float* src, uint8_t* dst, int64_t count, const lof_codec_format_t* fmt, const lof_projspec_t* ps,
                   int endian, const lof_rng_t* rng, uint64_t first_index, void* stream)
{
    if (count < 0) return fail(LOF_ERR_INVALID_ARG, "count < 0");
    if (!ps) return fail(LOF_ERR_INVALID_ARG, "projspec is NULL");
    const int mode = ps->rounding_mode;
    if (mode < 0 || mode >= LOF_NUM_ROUNDING_MODES) return fail(LOF_ERR_INVALID_ARG, "unknown rounding mode");
    CodecPlan p;
    const int st = make_codec_plan(fmt, ps, endian, &p);
    if (st != LOF_OK) return st;
    if (mode == LOF_StochasticRoundingA || mode == LOF_StochasticRoundingB || mode == LOF_StochasticRoundingC ||
        mode == LOF_StochasticRoundingD) {
        // the reference shifts by (roundoff - len); roundoff >= 23 - mantissa_bits on every path
        if (ps->stoch_length < 0 || ps->stoch_length > p.digit_drop || ps->stoch_length > 31)
            return fail(LOF_ERR_INVALID_ARG, "stoch_length must be in [0, 23 - mantissa_bits]");
    }
    if (count == 0) return LOF_OK;
    if (!src || !dst) return fail(LOF_ERR_INVALID_ARG, "null tensor pointer");
    const int sms = num_sms();
    if (sms <= 0) return fail(LOF_ERR_NO_DEVICE, "no CUDA device (this library has no CPU fallback)");
    const RngArgs ra = make_rng(rng, kStreamElem, mode, ps->stoch_length);
    cudaError_t e = launch_encode(mode, src, dst, count, lof_packed_size(p.n, (uint64_t)count), p, ra, first_index, sms * (int)g_opt.stream_ctas_per_sm.load(),
                                  static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "lof_encode launch");
    g_launches.fetch_add(1);
    return LOF_OK;
}

int lof_round_strict_f32(const float* in, float* out, int64_t n, const lof_codec_format_t* fmt, const lof_projspec_t* ps,
                         const lof_rng_t* rng, uint64_t first_index, void* stream)
{
    if (n < 0) return fail(LOF_ERR_INVALID_ARG, "n < 0");
    if (!ps) return fail(LOF_ERR_INVALID_ARG, "projspec is NULL");
    const int mode = ps->rounding_mode;
    if (mode < 0 || mode >= LOF_NUM_ROUNDING_MODES) return fail(LOF_ERR_INVALID_ARG, "unknown rounding mode");
    CodecPlan p;
    const int st = make_codec_plan(fmt, ps, LOF_PACK_FORWARD, &p);
    if (st != LOF_OK) return st;
    if ((mode == LOF_StochasticRoundingA || mode == LOF_StochasticRoundingB || mode == LOF_StochasticRoundingC ||
         mode == LOF_StochasticRoundingD) && (ps->stoch_length < 0 || ps->stoch_length > p.digit_drop || ps->stoch_length > 31))
        return fail(LOF_ERR_INVALID_ARG, "stoch_length must be in [0, 23 - mantissa_bits]");
    if (n == 0) return LOF_OK;
    if (!in || !out) return fail(LOF_ERR_INVALID_ARG, "null tensor pointer");
    const int sms = num_sms();
    if (sms <= 0) return fail(LOF_ERR_NO_DEVICE, "no CUDA device (this library has no CPU fallback)");
    const RngArgs ra = make_rng(rng, kStreamElem, mode, ps->stoch_length);
    cudaError_t e = launch_round_strict(mode, in, out, n, p, ra, first_index, sms * (int)g_opt.stream_ctas_per_sm.load(), static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "lof_round_strict launch");
    g_launches.fetch_add(1);
    return LOF_OK;
}

namespace {
int mx_storage_check(int64_t n, int block_size, const lof_codec_format_t* scale_fmt, const lof_codec_format_t* elem_fmt)
{
    if (n < 0) return fail(LOF_ERR_INVALID_ARG, "n < 0");
    if (!scale_fmt || !elem_fmt) return fail(LOF_ERR_INVALID_ARG, "codec format is NULL");
    if (block_size != 16 && block_size != 32 && block_size != 64 && block_size != 128)
        return fail(LOF_ERR_UNSUPPORTED, "mx storage: block_size must be 16, 32, 64 or 128");
    if (n % block_size != 0) return fail(LOF_ERR_INVALID_ARG, "mx storage: n must be a multiple of block_size");
    if (elem_fmt->fmt.bitwidth != 4 && elem_fmt->fmt.bitwidth != 6 && elem_fmt->fmt.bitwidth != 8)
        return fail(LOF_ERR_UNSUPPORTED, "mx storage: element formats of 4, 6 or 8 bits (MXFP4 / MXFP6 / MXFP8)");
    if (scale_fmt->fmt.bitwidth < 1 || scale_fmt->fmt.bitwidth > 8)
        return fail(LOF_ERR_UNSUPPORTED, "mx storage: the scale format must fit one byte");
    return LOF_OK;
}
}  // namespace

int lof_mx_encode_f32(const float* in, int64_t n, int block_size, const lof_codec_format_t* scale_fmt,
                      const lof_codec_format_t* elem_fmt, const lof_projspec_t* ps_scale, const lof_projspec_t* ps_elem, int endian,
                      const lof_rng_t* rng, uint64_t first_index, uint8_t* scales, uint8_t* elems, void* stream)
{
    int st = mx_storage_check(n, block_size, scale_fmt, elem_fmt);
    if (st != LOF_OK) return st;
    if (!ps_scale || !ps_elem) return fail(LOF_ERR_INVALID_ARG, "projspec is NULL");
    if (first_index % (uint64_t)block_size != 0) return fail(LOF_ERR_INVALID_ARG, "first_index must be a multiple of block_size");
    MxArgs a;
    std::memset(&a, 0, sizeof a);
    st = make_plan(&elem_fmt->fmt, LOF_F32, ps_elem->rounding_mode, ps_elem->stoch_length, &a.elem);
    if (st != LOF_OK) return st;
    st = make_plan(&scale_fmt->fmt, LOF_F32, ps_scale->rounding_mode, 0, &a.scale);
    if (st != LOF_OK) return st;
    a.pmn = lof_mx_priv_max_normal(&elem_fmt->fmt);
    a.scale_mode = ps_scale->rounding_mode;
    a.block_size = block_size;
    lof_projspec_t exact = {LOF_RoundTowardsZero, LOF_OvfInf, 0};
    CodecPlan cs, ce;
    st = make_codec_plan(scale_fmt, &exact, LOF_PACK_FORWARD, &cs);
    if (st != LOF_OK) return st;
    st = make_codec_plan(elem_fmt, &exact, endian, &ce);
    if (st != LOF_OK) return st;
    // the element rounding + conversion collapse into the codec's magic-add path when that path is eligible for the
    // ELEMENT rounding mode and the two overflow policies agree (virtual_round: unchanged -> clamped by the conversion,
    // or +inf for Extended formats; codec under OvfInf: max finite for Saturating formats, inf for Extended ones)
    lof_projspec_t elem_ps = {ps_elem->rounding_mode, LOF_OvfInf, 0};
    CodecPlan probe;
    st = make_codec_plan(elem_fmt, &elem_ps, endian, &probe);
    if (st != LOF_OK) return st;
    const bool fast = probe.fast_ok != 0 && !probe.has_inf;     // Extended formats: virtual_round sends every overflow to +inf (sign dropped)
    if (fast) ce = probe;
    if (n == 0) return LOF_OK;
    if (!in || !scales || !elems) return fail(LOF_ERR_INVALID_ARG, "null tensor pointer");
    if ((reinterpret_cast<uintptr_t>(in) & 15u) || (reinterpret_cast<uintptr_t>(elems) & 3u))
        return fail(LOF_ERR_INVALID_ARG, "mx storage: `in` must be 16-byte and `elems` 4-byte aligned");
    const int sms = num_sms();
    if (sms <= 0) return fail(LOF_ERR_NO_DEVICE, "no CUDA device (this library has no CPU fallback)");
    const RngArgs re = make_rng(rng, kStreamElem, ps_elem->rounding_mode, ps_elem->stoch_length), rs = make_rng(rng, kStreamScale);
    cudaError_t e = launch_mx_encode(ps_elem->rounding_mode, in, n, a, cs, ce, re, rs, first_index, scales, elems, sms * (int)g_opt.stream_ctas_per_sm.load(), fast,
                                     make_magic_plan(a.scale, ps_scale->rounding_mode), make_fast_plan(a.scale, ps_scale->rounding_mode),
                                     static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "lof_mx_encode launch");
    g_launches.fetch_add(1);
    return LOF_OK;
}

int lof_mx_decode_f32(const uint8_t* scales, const uint8_t* elems, float* out, int64_t n, int block_size,
                      const lof_codec_format_t* scale_fmt, const lof_codec_format_t* elem_fmt, int endian, void* stream)
{
    int st = mx_storage_check(n, block_size, scale_fmt, elem_fmt);
    if (st != LOF_OK) return st;
    CodecPlan cs, ce;
    st = make_codec_plan(scale_fmt, nullptr, LOF_PACK_FORWARD, &cs);
    if (st != LOF_OK) return st;
    st = make_codec_plan(elem_fmt, nullptr, endian, &ce);
    if (st != LOF_OK) return st;
    if (n == 0) return LOF_OK;
    if (!out || !scales || !elems) return fail(LOF_ERR_INVALID_ARG, "null tensor pointer");
    if ((reinterpret_cast<uintptr_t>(out) & 15u) || (reinterpret_cast<uintptr_t>(elems) & 3u))
        return fail(LOF_ERR_INVALID_ARG, "mx storage: `out` must be 16-byte and `elems` 4-byte aligned");
    const int sms = num_sms();
    if (sms <= 0) return fail(LOF_ERR_NO_DEVICE, "no CUDA device (this library has no CPU fallback)");
    cudaError_t e = launch_mx_decode(scales, elems, out, n, block_size, cs, ce, sms * (int)g_opt.stream_ctas_per_sm.load(), static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "lof_mx_decode launch");
    g_launches.fetch_add(1);
    return LOF_OK;
}

int lof_decode_f32(const uint8_t* src, float* dst, int64_t count, const lof_codec_format_t* fmt, int endian, void* stream)
{
    if (count < 0) return fail(LOF_ERR_INVALID_ARG, "count < 0");
    CodecPlan p;
    const int st = make_codec_plan(fmt, nullptr, endian, &p);
    if (st != LOF_OK) return st;
    if (count == 0) return LOF_OK;
    if (!src || !dst) return fail(LOF_ERR_INVALID_ARG, "null tensor pointer");
    const int sms = num_sms();
    if (sms <= 0) return fail(LOF_ERR_NO_DEVICE, "no CUDA device (this library has no CPU fallback)");
    cudaError_t e = launch_decode(src, dst, count, lof_packed_size(p.n, (uint64_t)count), p, sms * (int)g_opt.stream_ctas_per_sm.load(), static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "lof_decode launch");
    g_launches.fetch_add(1);
    return LOF_OK;
}

int lof_round_params_host(const void* in_host, void* out_host, int64_t n, int dtype, const lof_format_t* fmt,
                          const lof_projspec_t* ps, double scale, const lof_rng_t* rng, uint64_t first_index, int device)
{
    SrcGeom g;
    if (!src_geom(dtype, &g)) return fail(LOF_ERR_INVALID_ARG, "unknown dtype");
    return run_host_pipeline(in_host, out_host, n, dtype_size(dtype), 4096, device,
        [&](void* d, int64_t ne, uint64_t e0, void* s) {
            return round_common(true, d, d, ne, dtype, fmt, 0, ps, scale, rng, first_index + e0, s);
        });
}

int lof_round_mantissa_host(const void* in_host, void* out_host, int64_t n, int dtype, int mantissa_bits,
                            const lof_projspec_t* ps, double scale, const lof_rng_t* rng, uint64_t first_index, int device)
{
    SrcGeom g;
    if (!src_geom(dtype, &g)) return fail(LOF_ERR_INVALID_ARG, "unknown dtype");
    return run_host_pipeline(in_host, out_host, n, dtype_size(dtype), 4096, device,
        [&](void* d, int64_t ne, uint64_t e0, void* s) {
            return round_common(false, d, d, ne, dtype, nullptr, mantissa_bits, ps, scale, rng, first_index + e0, s);
        });
}

int lof_mx_round_f32_host(const float* in_host, float* out_host, int64_t n, int block_size, const lof_format_t* scale_fmt,
                          const lof_format_t* elem_fmt, const lof_projspec_t* ps_scale, const lof_projspec_t* ps_elem,
                          const lof_rng_t* rng, uint64_t first_index, int device)
{
    if (block_size <= 0) return fail(LOF_ERR_INVALID_ARG, "virtual_mx_round: block_size must be positive");
    return run_host_pipeline(in_host, out_host, n, 4, 4096, device,      // 4096 is a multiple of every block size
        [&](void* d, int64_t ne, uint64_t e0, void* s) {
            return mx_common(static_cast<const float*>(d), static_cast<float*>(d), ne, block_size, scale_fmt, elem_fmt,
                             ps_scale, ps_elem, rng, first_index + e0, s);
        });
}

}  // extern "C"
