/*
 * lof_oracle.c -- CPU ORACLE (plain C) for the LoFloat virtual-rounding hot path.
 *
 * TEST INFRASTRUCTURE ONLY -- see lof_oracle.h for the usage rules and the pinning story.
 * Every function cites the reference lines it restates.
 */
#include "lof_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------
 * Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11).
 * Not a reference function: the reference has no reproducible RNG (see lof_oracle.h).
 * Known-answer vectors from the Random123 distribution are checked in
 * tests/test_oracle_golden.py.
 * ---------------------------------------------------------------------------------- */
void lofo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int round = 0; round < 10; ++round) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

uint32_t lofo_rand_word(const lofo_rng* rng, uint64_t index, uint32_t stream)
{
    if (!rng) return 0u;
    if (rng->fixed) return rng->fixed_word;
    const uint64_t blk = index >> 2;
    uint32_t ctr[4] = { (uint32_t)blk, (uint32_t)(blk >> 32),
                        (uint32_t)rng->offset, (uint32_t)(rng->offset >> 32) ^ stream };
    uint32_t key[2] = { (uint32_t)rng->seed, (uint32_t)(rng->seed >> 32) };
    uint32_t out[4];
    lofo_philox4x32_10(ctr, key, out);
    return out[index & 3u];
}

/* Random bits one rounding event consumes, and the event word of global element `index`:
 *   <= 8 bits : byte (index & 3) of the word owned by the 4-element group v = index >> 2, namely dense word
 *               4 * (((v >> 10) << 8) | (v & 255)) + ((v >> 8) & 3)   (4 groups 256 apart share one Philox block)
 *   <= 16 bits: half (index & 1) of dense word index >> 1
 *   else      : dense word index                                     (lofloat_b200/csrc/lof_core.cuh rand_class) */
int lofo_rand_bits_needed(int mode, int len)
{
    if (mode == LOFO_TRUE_SR) return 32;
    if (mode == LOFO_PROB) return 1;
    if (mode == LOFO_SR_D) return len + 1;
    return len;
}

uint32_t lofo_rand_event(const lofo_rng* rng, uint64_t index, uint32_t stream, int nbits)
{
    if (!rng) return 0u;
    if (rng->fixed) return rng->fixed_word;
    if (nbits > 16) return lofo_rand_word(rng, index, stream);
    if (nbits > 8) return (lofo_rand_word(rng, index >> 1, stream) >> (16u * (uint32_t)(index & 1u))) & 0xFFFFu;
    const uint64_t v = index >> 2;
    const uint64_t q = 4u * (((v >> 10) << 8) | (v & 255u)) + ((v >> 8) & 3u);
    return (lofo_rand_word(rng, q, stream) >> (8u * (uint32_t)(index & 3u))) & 0xFFu;
}


/* ------------------------------------------------------------------------------------
 * container conversions
 * ---------------------------------------------------------------------------------- */
static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint64_t d2u(double f) { uint64_t u; memcpy(&u, &f, 8); return u; }
static inline double u2d(uint64_t u) { double f; memcpy(&f, &u, 8); return f; }

static inline float half_bits_to_float(uint16_t h)
{
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1Fu;
    uint32_t man = h & 0x3FFu;
    if (exp == 0x1F) return u2f(sign | 0x7F800000u | (man << 13));
    if (exp == 0) {
        float v = (float)man * 0x1p-24f;    /* subnormal: exact */
        return sign ? -v : v;
    }
    return u2f(sign | ((exp + 112u) << 23) | (man << 13));
}

/* float -> IEEE half, round-to-nearest-even (what c10::Half(float) does) */
static inline uint16_t float_to_half_bits(float f)
{
    uint32_t x = f2u(f);
    uint32_t sign = (x >> 16) & 0x8000u;
    uint32_t a = x & 0x7FFFFFFFu;
    if (a > 0x7F800000u) return (uint16_t)(sign | 0x7E00u);            /* NaN */
    if (a >= 0x47800000u) return (uint16_t)(sign | 0x7C00u);           /* >= 65536 -> inf (incl. rounding) */
    if (a < 0x33000000u) return (uint16_t)sign;                        /* < 2^-25 -> 0 */
    int e = (int)(a >> 23) - 127;
    uint32_t man = (a & 0x7FFFFFu) | 0x800000u;
    int shift = (e < -14) ? (13 + (-14 - e)) : 13;
    uint32_t q = man >> shift;
    uint32_t rem = man & ((1u << shift) - 1u);
    uint32_t halfway = 1u << (shift - 1);
    if (rem > halfway || (rem == halfway && (q & 1u))) q += 1u;
    uint32_t h;
    if (e < -14) h = q;                                                /* subnormal (carry into normal is fine) */
    else h = ((uint32_t)(e + 15) << 10) + (q - 0x400u);                /* q in [0x400, 0x800] */
    if (h >= 0x7C00u) h = 0x7C00u;
    return (uint16_t)(sign | h);
}

static inline float bf16_bits_to_float(uint16_t h) { return u2f((uint32_t)h << 16); }

static inline uint16_t float_to_bf16_bits(float f)
{
    uint32_t x = f2u(f);
    if ((x & 0x7FFFFFFFu) > 0x7F800000u) return (uint16_t)((x >> 16) | 0x0040u);
    uint32_t lsb = (x >> 16) & 1u;
    x += 0x7FFFu + lsb;
    return (uint16_t)(x >> 16);
}

/* per-format constants of virtual_round(value, params): lo_float.h:2857-2861 */
typedef struct { int ToMax_exp, ToMin_exp; float To_min_val; } lofo_consts;

static lofo_consts lofo_consts_make(const lofo_format* f)
{
    lofo_consts k;
    k.ToMax_exp = (f->is_signed == LOFO_SIGNED)
                      ? (1 << (f->bitwidth - f->mantissa_bits - 1)) - 1 - f->bias
                      : (1 << (f->bitwidth - f->mantissa_bits)) - 1 - f->bias;
    k.ToMin_exp = 1 - f->bias;
    k.To_min_val = (float)(pow(2.0, (double)k.ToMin_exp) * pow(2.0, (double)(-f->mantissa_bits)));
    return k;
}

/* ------------------------------------------------------------------------------------
 * integer rounding core, one instantiation per container width
 * ---------------------------------------------------------------------------------- */
#define SFX b32
#define BITS_T uint32_t
#include "lof_oracle_core.inc"
#undef SFX
#undef BITS_T

#define SFX b16
#define BITS_T uint16_t
#include "lof_oracle_core.inc"
#undef SFX
#undef BITS_T

#define SFX b64
#define BITS_T uint64_t
#include "lof_oracle_core.inc"
#undef SFX
#undef BITS_T

/* ------------------------------------------------------------------------------------
 * virtual_round bodies, one instantiation per source type
 * ---------------------------------------------------------------------------------- */
/* float (Traits<float>, lo_float.h:2641-2653) */
#define TSFX f32
#define SFX b32
#define BITS_T uint32_t
#define KBITS 32
#define KMANT 23
#define KBIAS 127
#define QNAN_BITS 0x7FC00000u
#define INF_BITS 0x7F800000u
#define TO_FLOAT(b) u2f(b)
#define FROM_DOUBLE(d) f2u((float)(d))
#define IS_NAN(b) (((b) & 0x7FFFFFFFu) > 0x7F800000u)
#include "lof_oracle_type.inc"
#undef TSFX
#undef SFX
#undef BITS_T
#undef KBITS
#undef KMANT
#undef KBIAS
#undef QNAN_BITS
#undef INF_BITS
#undef TO_FLOAT
#undef FROM_DOUBLE
#undef IS_NAN

/* double (Traits<double>, lo_float.h:2656-2668); get_float() narrows with RN (:2689) */
#define TSFX f64
#define SFX b64
#define BITS_T uint64_t
#define KBITS 64
#define KMANT 52
#define KBIAS 1023
#define QNAN_BITS 0x7FF8000000000000ull
#define INF_BITS 0x7FF0000000000000ull
#define TO_FLOAT(b) ((float)u2d(b))
#define FROM_DOUBLE(d) d2u((double)(d))
#define IS_NAN(b) (((b) & 0x7FFFFFFFFFFFFFFFull) > 0x7FF0000000000000ull)
#include "lof_oracle_type.inc"
#undef TSFX
#undef SFX
#undef BITS_T
#undef KBITS
#undef KMANT
#undef KBIAS
#undef QNAN_BITS
#undef INF_BITS
#undef TO_FLOAT
#undef FROM_DOUBLE
#undef IS_NAN

/* c10::Half (Traits<c10::Half>, lo_float.h:2626-2638); from_double goes double->float->half */
#define TSFX f16
#define SFX b16
#define BITS_T uint16_t
#define KBITS 16
#define KMANT 10
#define KBIAS 15
#define QNAN_BITS 0x7E00u
#define INF_BITS 0x7C00u
#define TO_FLOAT(b) half_bits_to_float(b)
#define FROM_DOUBLE(d) float_to_half_bits((float)(d))
#define IS_NAN(b) (((b) & 0x7FFFu) > 0x7C00u)
#include "lof_oracle_type.inc"
#undef TSFX
#undef SFX
#undef BITS_T
#undef KBITS
#undef KMANT
#undef KBIAS
#undef QNAN_BITS
#undef INF_BITS
#undef TO_FLOAT
#undef FROM_DOUBLE
#undef IS_NAN

/* bfloat16: NOT a reference source type (LoPy_bind.cpp:83,133 dispatches float/double/half
 * only).  Defined by analogy: 16-bit container, 7 mantissa bits, bias 127. */
#define TSFX bf16
#define SFX b16
#define BITS_T uint16_t
#define KBITS 16
#define KMANT 7
#define KBIAS 127
#define QNAN_BITS 0x7FC0u
#define INF_BITS 0x7F80u
#define TO_FLOAT(b) bf16_bits_to_float(b)
#define FROM_DOUBLE(d) float_to_bf16_bits((float)(d))
#define IS_NAN(b) (((b) & 0x7FFFu) > 0x7F80u)
#include "lof_oracle_type.inc"
#undef TSFX
#undef SFX
#undef BITS_T
#undef KBITS
#undef KMANT
#undef KBIAS
#undef QNAN_BITS
#undef INF_BITS
#undef TO_FLOAT
#undef FROM_DOUBLE
#undef IS_NAN

/* ------------------------------------------------------------------------------------
 * scalar + array entry points
 * ---------------------------------------------------------------------------------- */
uint32_t lofo_round_params_bits_f32(uint32_t bits, const lofo_format* f, const lofo_projspec* ps, uint32_t w)
{
    const lofo_consts k = lofo_consts_make(f);
    return vround_params_f32(bits, f, &k, ps->rounding_mode, ps->stoch_length, w);
}

uint32_t lofo_round_mantissa_bits_f32(uint32_t bits, int mant_bits, const lofo_projspec* ps, uint32_t w)
{
    return vround_mant_f32(bits, mant_bits, ps->rounding_mode, ps->stoch_length, w);
}

static inline int mode_uses_rng(int mode)
{
    return mode == LOFO_SR_A || mode == LOFO_SR_B || mode == LOFO_SR_C || mode == LOFO_SR_D ||
           mode == LOFO_TRUE_SR || mode == LOFO_PROB;
}

/* fused pre-scale, Lof_kernel.cu:41,53: v = (scale == From(1)) ? in : From(scale * in),
 * evaluated in the container's own arithmetic (Half*Half and bf16*bf16 multiply in float
 * and round back to the container). */
static inline uint32_t scale_f32(uint32_t b, float s) { return (s == 1.0f) ? b : f2u(s * u2f(b)); }
static inline uint64_t scale_f64(uint64_t b, double s) { return (s == 1.0) ? b : d2u(s * u2d(b)); }
static inline uint16_t scale_f16(uint16_t b, uint16_t sb)
{
    float s = half_bits_to_float(sb);
    return (s == 1.0f) ? b : float_to_half_bits(s * half_bits_to_float(b));
}
static inline uint16_t scale_bf16(uint16_t b, uint16_t sb)
{
    float s = bf16_bits_to_float(sb);
    return (s == 1.0f) ? b : float_to_bf16_bits(s * bf16_bits_to_float(b));
}

#define ROUND_LOOP(BITS_T, SCALE_EXPR, CALL)                                              \
    do {                                                                                  \
        const BITS_T* src = (const BITS_T*)in;                                            \
        BITS_T* dst = (BITS_T*)out;                                                       \
        _Pragma("omp parallel for schedule(static)")                                      \
        for (int64_t i = 0; i < n; ++i) {                                                 \
            uint32_t w = need_rng ? lofo_rand_event(rng, first_index + (uint64_t)i, LOFO_STREAM_ELEM, nbits) : 0u; \
            BITS_T v = SCALE_EXPR;                                                        \
            dst[i] = CALL;                                                                \
        }                                                                                 \
    } while (0)

void lofo_round_params(const void* in, void* out, int64_t n, int dtype, const lofo_format* f,
                       const lofo_projspec* ps, double scale, const lofo_rng* rng, uint64_t first_index)
{
    const int mode = ps->rounding_mode, len = ps->stoch_length;
    const int need_rng = mode_uses_rng(mode) && rng != NULL;
    const int nbits = lofo_rand_bits_needed(mode, len);
    const lofo_consts kc = lofo_consts_make(f);
    switch (dtype) {
    case LOFO_F32: {
        const float s = (float)scale;
        ROUND_LOOP(uint32_t, scale_f32(src[i], s), vround_params_f32(v, f, &kc, mode, len, w));
        break;
    }
    case LOFO_F64: {
        ROUND_LOOP(uint64_t, scale_f64(src[i], scale), vround_params_f64(v, f, &kc, mode, len, w));
        break;
    }
    case LOFO_F16: {
        const uint16_t sb = float_to_half_bits((float)scale);
        ROUND_LOOP(uint16_t, scale_f16(src[i], sb), vround_params_f16(v, f, &kc, mode, len, w));
        break;
    }
    case LOFO_BF16: {
        const uint16_t sb = float_to_bf16_bits((float)scale);
        ROUND_LOOP(uint16_t, scale_bf16(src[i], sb), vround_params_bf16(v, f, &kc, mode, len, w));
        break;
    }
    default: break;
    }
}

void lofo_round_mantissa(const void* in, void* out, int64_t n, int dtype, int mant_bits,
                         const lofo_projspec* ps, double scale, const lofo_rng* rng, uint64_t first_index)
{
    const int mode = ps->rounding_mode, len = ps->stoch_length;
    const int need_rng = mode_uses_rng(mode) && rng != NULL;
    const int nbits = lofo_rand_bits_needed(mode, len);
    switch (dtype) {
    case LOFO_F32: {
        const float s = (float)scale;
        ROUND_LOOP(uint32_t, scale_f32(src[i], s), vround_mant_f32(v, mant_bits, mode, len, w));
        break;
    }
    case LOFO_F64: {
        ROUND_LOOP(uint64_t, scale_f64(src[i], scale), vround_mant_f64(v, mant_bits, mode, len, w));
        break;
    }
    case LOFO_F16: {
        const uint16_t sb = float_to_half_bits((float)scale);
        ROUND_LOOP(uint16_t, scale_f16(src[i], sb), vround_mant_f16(v, mant_bits, mode, len, w));
        break;
    }
    case LOFO_BF16: {
        const uint16_t sb = float_to_bf16_bits((float)scale);
        ROUND_LOOP(uint16_t, scale_bf16(src[i], sb), vround_mant_bf16(v, mant_bits, mode, len, w));
        break;
    }
    default: break;
    }
}

/* ------------------------------------------------------------------------------------
 * MX block rounding: LoPy_bind.cpp:197-215 (== lo_float.h:4311-4383 on NaN-free data)
 * ---------------------------------------------------------------------------------- */
float lofo_mx_priv_max_normal(const lofo_format* p)
{
    const int me = (p->is_signed == LOFO_SIGNED)
                       ? (1 << (p->bitwidth - p->mantissa_bits - 1)) - 1 - p->bias
                       : (1 << (p->bitwidth - p->mantissa_bits)) - 1 - p->bias;
    return ldexpf(1.0f, me) * (2.0f - ldexpf(1.0f, -p->mantissa_bits));   /* LoPy_bind.cpp:202 */
}

void lofo_mx_round_f32(const float* in, float* out, int64_t n, int block_size,
                       const lofo_format* pub, const lofo_format* priv,
                       const lofo_projspec* ps_pub, const lofo_projspec* ps_priv,
                       const lofo_rng* rng, uint64_t first_index)
{
    if (block_size <= 0 || n <= 0) return;
    const float pmn = lofo_mx_priv_max_normal(priv);
    const lofo_consts kpub = lofo_consts_make(pub), kpriv = lofo_consts_make(priv);
    const int need_rng_e = mode_uses_rng(ps_priv->rounding_mode) && rng != NULL;
    const int need_rng_s = mode_uses_rng(ps_pub->rounding_mode) && rng != NULL;
    const int nbits_e = lofo_rand_bits_needed(ps_priv->rounding_mode, ps_priv->stoch_length);
    const int64_t nblocks = (n + block_size - 1) / block_size;
#pragma omp parallel for schedule(static)
    for (int64_t blk = 0; blk < nblocks; ++blk) {
        const int64_t b = blk * block_size;
        const int len = (int)((n - b) < block_size ? (n - b) : block_size);
        float amax = 0.0f;
        for (int i = 0; i < len; ++i) {                     /* std::max(amax, fabs(x)): NaN is dropped */
            float v = fabsf(in[b + i]);
            amax = (amax < v) ? v : amax;
        }
        /* scale rounding: ProjSpec{scale_round_mode} -> stoch_len 0 (LoPy_bind.cpp:208) */
        const uint64_t gblk = (first_index + (uint64_t)b) / (uint64_t)block_size;
        uint32_t ws = need_rng_s ? lofo_rand_word(rng, gblk, LOFO_STREAM_SCALE) : 0u;
        const float scale = u2f(vround_params_f32(f2u(amax / pmn), pub, &kpub, ps_pub->rounding_mode, 0, ws));
        if (scale > 0.0f) {
            for (int i = 0; i < len; ++i) {
                uint32_t w = need_rng_e ? lofo_rand_event(rng, first_index + (uint64_t)(b + i), LOFO_STREAM_ELEM, nbits_e) : 0u;
                float q = u2f(vround_params_f32(f2u(in[b + i] / scale), priv, &kpriv, ps_priv->rounding_mode,
                                                ps_priv->stoch_length, w));
                out[b + i] = q * scale;
            }
        } else {
            for (int i = 0; i < len; ++i) out[b + i] = 0.0f;
        }
    }
}

/* ------------------------------------------------------------------------------------
 * round-to-odd FMA, blis_lof_gemm.hpp:63-99
 * ---------------------------------------------------------------------------------- */
float lofo_round_to_odd_fma(float a, float b, float c)
{
    const float rn = fmaf(a, b, c);
    if (!isfinite(rn)) return rn;
    const double p = (double)a * (double)b;
    const double cd = (double)c;
    const double s = p + cd;
    const double bb = s - p;
    const double e = (p - (s - bb)) + (cd - bb);
    const double res = (s - (double)rn) + e;
    if (res == 0.0) return rn;
    uint32_t brn = f2u(rn);
    if (brn & 1u) return rn;
    if (rn == 0.0f) {
        brn = (res > 0.0) ? 0x00000001u : 0x80000001u;
    } else {
        const int neg = (int)(brn >> 31);
        const int toward_pos = res > 0.0;
        if (toward_pos == neg) --brn; else ++brn;
    }
    return u2f(brn);
}

/* ------------------------------------------------------------------------------------
 * LoFloat-accumulation GEMM: lof_sgemm_ukr (blis_lof_gemm.hpp:147-262) under BLIS's
 * K-panelling.  Per output element the result depends only on the panel length kc
 * (MR/NR/MC/NC only change which elements share a micro-tile, never the k order).
 * ---------------------------------------------------------------------------------- */
void lofo_gemm_f32(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K,
                   int accum_mant_bits, const lofo_projspec* ps, int64_t kc,
                   float alpha, float beta, const lofo_rng* rng)
{
    const int mode = ps->rounding_mode, len = ps->stoch_length;
    const int need_rng = mode_uses_rng(mode) && rng != NULL;
    if (kc <= 0 || kc > K) kc = K > 0 ? K : 1;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < M; ++i) {
        for (int64_t j = 0; j < N; ++j) {
            float cij = C[i * N + j];
            int first = 1;
            if (K == 0) { C[i * N + j] = (beta == 0.0f) ? 0.0f : beta * cij; continue; }
            for (int64_t k0 = 0; k0 < K; k0 += kc) {
                const int64_t k1 = (k0 + kc < K) ? k0 + kc : K;
                float acc = 0.0f;
                if (accum_mant_bits >= 0) {
                    for (int64_t l = k0; l < k1; ++l) {                                  /* :217-228 */
                        uint32_t w = need_rng
                            ? lofo_rand_word(rng, ((uint64_t)i * (uint64_t)N + (uint64_t)j) * (uint64_t)K + (uint64_t)l, LOFO_STREAM_GEMM)
                            : 0u;
                        float t = lofo_round_to_odd_fma(A[i * K + l], B[l * N + j], acc);
                        acc = u2f(vround_mant_f32(f2u(t), accum_mant_bits, mode, len, w));
                    }
                } else {
                    for (int64_t l = k0; l < k1; ++l) {                                  /* :233-239 */
                        volatile float prod = A[i * K + l] * B[l * N + j];               /* no contraction: ab += a*b */
                        acc = acc + prod;
                    }
                }
                acc *= alpha;                                                            /* :244-245 */
                if (first) { cij = (beta == 0.0f) ? acc : beta * cij + acc; first = 0; } /* :248-261 */
                else       { cij = cij + acc; }                                          /* beta == 1 on later panels */
            }
            C[i * N + j] = cij;
        }
    }
}

/* Round-to-odd add used by the chunked mode: RZ(a + b) with the LSB OR-ed with "inexact".  Computed from
 * the RN sum and its exact TwoSum error term (Knuth), no rounding-mode switching. */
float lofo_rto_add(float a, float b)
{
    const float s = a + b;
    if (isnan(s)) return u2f(f2u(s) | 1u);                    /* NaN stays NaN (device: RZ NaN | sticky) */
    if (isinf(s)) {
        if (isinf(a) || isinf(b)) return s;                    /* genuine infinity: exact */
        return copysignf(FLT_MAX, s);                          /* finite sum beyond FLT_MAX: RZ = FLT_MAX, LSB already 1 */
    }
    const float bb = s - a;
    const float err = (a - (s - bb)) + (b - bb);               /* exact: a + b == s + err */
    if (err == 0.0f) return s;
    float rz = s;
    if (s == 0.0f) {
        /* cannot happen with err != 0 (a sum that rounds to zero is exact) */
        return s;
    }
    if ((err > 0.0f) != (s > 0.0f)) rz = nextafterf(s, 0.0f);  /* the exact value lies between s and zero */
    return u2f(f2u(rz) | 1u);
}

/* Oracle of OUR tensor-core (tcgen05) chunked mode -- not a reference function.
 * Each k-chunk's partial dot product P is taken as the correctly rounded fp32 value of the
 * exact sum (the tensor core's internal alignment may differ by a few fp32 ulps on wide-range
 * data; the GPU tests state that tolerance and also run exactly-summable data bit for bit), then
 *     partial_round == 0:  acc = round_to_format(acc + P)            -- single rounding (round-to-odd add, then format rounding)
 *     partial_round != 0:  acc = round_to_format(acc + round_to_format(P)) */
void lofo_gemm_chunked_rng_f32(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K,
                               int accum_mant_bits, const lofo_format* acc_fmt, const lofo_projspec* ps,
                               int64_t k_chunk, float alpha, int partial_round, const lofo_rng* rng);

void lofo_gemm_chunked_f32(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K,
                           int accum_mant_bits, const lofo_format* acc_fmt, const lofo_projspec* ps,
                           int64_t k_chunk, float alpha, int partial_round)
{
    lofo_gemm_chunked_rng_f32(A, B, C, M, N, K, accum_mant_bits, acc_fmt, ps, k_chunk, alpha, partial_round, NULL);
}

/* ... with stochastic accumulator rounding: one dense Philox word per rounding event (the stream of lofo_gemm_f32), event index
 * ((which * n_chunks + chunk) * M + i) * N + j, which = 0 for the partial's rounding, 1 for the accumulator's (rng == NULL: word 0). */
void lofo_gemm_chunked_rng_f32(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K,
                               int accum_mant_bits, const lofo_format* acc_fmt, const lofo_projspec* ps,
                               int64_t k_chunk, float alpha, int partial_round, const lofo_rng* rng)
{
    const int mode = ps->rounding_mode, len = ps->stoch_length;
    const int need_rng = mode_uses_rng(mode) && rng != NULL;
    if (k_chunk <= 0 || k_chunk > K) k_chunk = K > 0 ? K : 1;
    const int64_t n_chunks = (K + k_chunk - 1) / k_chunk;
    lofo_consts kacc = {0, 0, 0.0f};
    if (acc_fmt) kacc = lofo_consts_make(acc_fmt);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < M; ++i) {
        for (int64_t j = 0; j < N; ++j) {
            float acc = 0.0f;
            int64_t c = 0;
            for (int64_t k0 = 0; k0 < K; k0 += k_chunk, ++c) {
                const int64_t k1 = (k0 + k_chunk < K) ? k0 + k_chunk : K;
                double part = 0.0;
                for (int64_t l = k0; l < k1; ++l) part += (double)A[i * K + l] * (double)B[l * N + j];
                float partf = (float)part;
                const uint32_t wp = need_rng ? lofo_rand_word(rng, (((uint64_t)c) * (uint64_t)M + (uint64_t)i) * (uint64_t)N + (uint64_t)j, LOFO_STREAM_GEMM) : 0u;
                const uint32_t wa = need_rng ? lofo_rand_word(rng, (((uint64_t)(n_chunks + c)) * (uint64_t)M + (uint64_t)i) * (uint64_t)N + (uint64_t)j, LOFO_STREAM_GEMM) : 0u;
                if (partial_round) {
                    if (acc_fmt) partf = u2f(vround_params_f32(f2u(partf), acc_fmt, &kacc, mode, len, wp));
                    else         partf = u2f(vround_mant_f32(f2u(partf), accum_mant_bits, mode, len, wp));
                }
                const float t = lofo_rto_add(acc, partf);
                if (acc_fmt) acc = u2f(vround_params_f32(f2u(t), acc_fmt, &kacc, mode, len, wa));
                else         acc = u2f(vround_mant_f32(f2u(t), accum_mant_bits, mode, len, wa));
            }
            C[i * N + j] = acc * alpha;
        }
    }
}

/* ------------------------------------------------------------------------------------
 * Low-bit storage codec: lo_pack.hpp (runtime twin of ConvertImpl<float, Templated_Float<Fp>>::run()).
 * ---------------------------------------------------------------------------------- */
static inline int codec_is_inf(const lofo_codec_format* f, uint32_t rep) { return (rep & f->inf_mask) == f->inf_value; }
static inline int codec_is_nan(const lofo_codec_format* f, uint32_t rep)
{
    return (rep & f->nan_mask) == f->nan_value && (f->nan_mant_mask == 0u || (rep & f->nan_mant_mask) != 0u);
}

/* lo_pack.hpp:147-152 with_sign(): the negate operator's guard (no representable -0 when sign-bit-only is NaN) */
static inline uint32_t codec_with_sign(const lofo_codec_format* f, uint32_t mag, int neg, int to_signed, uint32_t signmask)
{
    if (!neg || !to_signed) return mag;
    if (mag == 0u && codec_is_nan(f, signmask)) return 0u;
    return mag | signmask;
}

/* lo_pack.hpp:109-267 convert_float_to_rep.  Every RoundMantissa call there omits `pos`, i.e. pos = false
 * (lo_float.h:1232): RoundUp / RoundDown treat every value as negative -- reproduced. */
uint32_t lofo_float_to_rep(float from, const lofo_codec_format* f, const lofo_projspec* ps, uint32_t w)
{
    const int mode = ps->rounding_mode, len = ps->stoch_length, sat_mode = ps->saturation_mode;
    const int n = f->fmt.bitwidth, toMant = f->fmt.mantissa_bits;
    const int to_signed = (f->fmt.is_signed == LOFO_SIGNED);
    const int toExpBits = to_signed ? n - toMant - 1 : n - toMant;
    const int toBias = f->fmt.bias;
    const int kExponentOffset = toBias - 127;
    const int kDigitShift = toMant - 23;
    const int toMinExp = 1 - toBias;
    const int toMaxExp = (1 << toExpBits) - 1 - toBias;
    const int toDigits = toMant + 1;
    const int has_inf = (f->fmt.ov_behavior != LOFO_SATURATING);
    const uint32_t signmask = to_signed ? (1u << (n - 1)) : 0u;
    const uint32_t maxFiniteRep = has_inf ? (f->min_pos_inf - 1u)
                                          : (to_signed ? (((1u << n) - 1u) >> 1) : ((1u << n) - 1u));
    const uint32_t infRep = has_inf ? f->min_pos_inf : maxFiniteRep;
    const uint32_t fb = f2u(from);
    const int neg = (fb >> 31) != 0;

    if (!to_signed && neg) {                                       /* :156-161 */
        if (f->fmt.unsigned_behavior == 0 /* NegtoZero */) return 0u;
        if (f->fmt.na_behavior == 0 /* _3109 */) return f->qnan;
        return f->snan;
    }
    const uint32_t from_bits = fb & 0x7FFFFFFFu;
    if (from_bits == 0x7F800000u) {                                /* :165-169 */
        if (sat_mode != LOFO_SAT_FINITE && has_inf) return codec_with_sign(f, infRep, neg, to_signed, signmask);
        return codec_with_sign(f, maxFiniteRep, neg, to_signed, signmask);
    }
    if (from_bits > 0x7F800000u) return f->qnan;                   /* :170 */
    if (from_bits == 0u) return codec_with_sign(f, 0u, neg, to_signed, signmask);
    const int biased_from_exponent = (int)(from_bits >> 23);

    if (toMinExp < -125) {                                         /* branch 1, :176-196 */
        if (biased_from_exponent == 0) {
            uint64_t bits = from_bits;
            const int normalization_factor = __builtin_clz(from_bits) - (32 - 23) + 1;
            const int biased_exponent = kExponentOffset - normalization_factor + 1;
            if (biased_exponent <= 0) {
                if (kExponentOffset < 64) bits <<= kExponentOffset;
            } else {
                bits <<= normalization_factor;
                bits &= ~((uint64_t)1 << 23);
                bits |= (uint64_t)biased_exponent << 23;
            }
            if (kDigitShift > 0) bits <<= kDigitShift;
            else { bits = round_mantissa_b64(bits, -kDigitShift, mode, len, 0, w); bits >>= -kDigitShift; }
            return codec_with_sign(f, (uint32_t)bits, neg, to_signed, signmask);
        }
    }
    if (toMinExp > -125) {                                         /* branch 2, :199-243 */
        const int unbiased_exponent = biased_from_exponent - 127;
        const int biased_to_exponent = unbiased_exponent + toBias;
        if (biased_to_exponent <= 0) {
            const uint32_t lead = (biased_from_exponent > 0) ? 1u : 0u;
            const int exponent_shift = -kDigitShift - biased_to_exponent + (int)lead;
            uint32_t rounded = (from_bits & 0x007FFFFFu) | (lead << 23);
            uint32_t bits = 0;
            if (exponent_shift <= 23 + 1) {
                rounded = round_mantissa_b32(rounded, exponent_shift, mode, len, 0, w);
                bits = (exponent_shift >= 32) ? 0u : (rounded >> exponent_shift);
            } else {
                uint64_t wide = rounded;
                switch (mode) {
                case LOFO_RTO: case LOFO_RAZ: bits = rounded > 0u ? 1u : 0u; break;
                case LOFO_RTZ: bits = 0u; break;
                case LOFO_RU: bits = (rounded && !neg) ? 1u : 0u; break;      /* `rounded && !sign > 0` */
                case LOFO_RD: bits = (rounded && neg) ? 1u : 0u; break;
                case LOFO_SR_A: case LOFO_SR_B: case LOFO_SR_C: case LOFO_SR_D: case LOFO_TRUE_SR: case LOFO_PROB:
                    /* the reference shifts a 64-bit word by exponent_shift, undefined for >= 64; defined here as 0 */
                    if (exponent_shift < 64) { wide = round_mantissa_b64(wide, exponent_shift, mode, len, 0, w); bits = (uint32_t)(wide >> exponent_shift); }
                    else bits = 0u;
                    break;
                default: bits = 0u; break;
                }
            }
            return codec_with_sign(f, bits, neg, to_signed, signmask);
        }
    }
    /* normal path, :246-266 */
    uint64_t rounded = from_bits;
    if (kDigitShift < 0) {
        rounded = round_mantissa_b64(rounded, -kDigitShift, mode, len, 0, w);
        rounded &= ~(((uint64_t)1 << (-kDigitShift)) - 1u);
    }
    rounded += (uint64_t)(int64_t)kExponentOffset << 23;          /* static_cast<WideBits>(int) sign-extends, then wraps */
    uint32_t bits;
    uint64_t aligned_highest = maxFiniteRep;
    if (kDigitShift < 0) {
        aligned_highest <<= -kDigitShift;
        bits = (uint32_t)(rounded >> -kDigitShift);
    } else {
        rounded <<= kDigitShift;
        bits = (uint32_t)rounded;
    }
    const int to_is_narrower = (toMaxExp < 128) || (toMaxExp == 128 && toDigits < 24);
    if (to_is_narrower && rounded > aligned_highest) {
        const int emit_inf = (sat_mode == LOFO_OVF_INF) && has_inf;
        return codec_with_sign(f, emit_inf ? infRep : maxFiniteRep, neg, to_signed, signmask);
    }
    return codec_with_sign(f, bits, neg, to_signed, signmask);
}

/* lo_pack.hpp:277-319 convert_rep_to_float */
float lofo_rep_to_float(uint32_t rep, const lofo_codec_format* f)
{
    const int n = f->fmt.bitwidth, mant = f->fmt.mantissa_bits, bias = f->fmt.bias;
    const int is_signed = (f->fmt.is_signed == LOFO_SIGNED);
    const int has_inf = (f->fmt.ov_behavior != LOFO_SATURATING);
    rep &= (n >= 32) ? 0xFFFFFFFFu : ((1u << n) - 1u);
    const int sign = is_signed ? (int)((rep >> (n - 1)) & 1u) : 0;
    const uint32_t mag = is_signed ? (rep & ((1u << (n - 1)) - 1u)) : rep;
    if (has_inf && codec_is_inf(f, rep)) return sign ? -INFINITY : INFINITY;
    if (codec_is_nan(f, rep)) return u2f(0x7FC00000u);
    const uint32_t exp_field = mag >> mant;
    const uint32_t mant_field = mag & ((1u << mant) - 1u);
    double val;
    if (exp_field == 0) val = ldexp((double)mant_field, 1 - bias - mant);
    else val = ldexp(1.0 + (double)mant_field / (double)(1u << mant), (int)exp_field - bias);
    return (float)(sign ? -val : val);
}

void lofo_float_to_rep_array(const float* in, uint32_t* out, int64_t n, const lofo_codec_format* f, const lofo_projspec* ps,
                             const lofo_rng* rng, uint64_t first_index)
{
    const int nbits = lofo_rand_bits_needed(ps->rounding_mode, ps->stoch_length);
    for (int64_t i = 0; i < n; ++i) {
        const uint32_t w = nbits ? lofo_rand_event(rng, first_index + (uint64_t)i, LOFO_STREAM_ELEM, nbits) : 0u;
        out[i] = lofo_float_to_rep(in[i], f, ps, w);
    }
}
void lofo_rep_to_float_array(const uint32_t* in, float* out, int64_t n, const lofo_codec_format* f)
{
    for (int64_t i = 0; i < n; ++i) out[i] = lofo_rep_to_float(in[i], f);
}

/* lo_pack.hpp:58-73 group geometry */
static int pk_gcd(int a, int b) { return b == 0 ? a : pk_gcd(b, a % b); }
static int pk_group_bits(int n) { const int al = n <= 8 ? 8 : 16; return (n / pk_gcd(n, al)) * al; }
int lofo_pack_group_count(int n) { return pk_group_bits(n) / n; }
int lofo_pack_group_bytes(int n) { return pk_group_bits(n) / 8; }
uint64_t lofo_packed_size(int n, uint64_t count)
{
    const uint64_t gc = (uint64_t)lofo_pack_group_count(n);
    return ((count + gc - 1) / gc) * (uint64_t)lofo_pack_group_bytes(n);
}

/* lo_pack.hpp:76-102 bit_offset / put_bits / get_bits (LSB-first within a byte) */
static void pk_put(uint8_t* dst, int off, uint32_t v, int n)
{
    for (int b = 0; b < n; ++b) { const int pos = off + b; dst[pos >> 3] |= (uint8_t)(((v >> b) & 1u) << (pos & 7)); }
}
static uint32_t pk_get(const uint8_t* src, int off, int n)
{
    uint32_t v = 0;
    for (int b = 0; b < n; ++b) { const int pos = off + b; v |= (uint32_t)((src[pos >> 3] >> (pos & 7)) & 1u) << b; }
    return v;
}
static int pk_off(int i, int n, int gc, int endian) { return ((endian == LOFO_PACK_FORWARD) ? i : (gc - 1 - i)) * n; }

uint64_t lofo_pack_reps(uint8_t* dst, const uint32_t* reps, uint64_t count, int n, int endian)
{
    const int gc = lofo_pack_group_count(n), gb = lofo_pack_group_bytes(n);
    const uint32_t mask = (n >= 32) ? 0xFFFFFFFFu : ((1u << n) - 1u);
    uint64_t written = 0;
    for (uint64_t i = 0; i < count; i += (uint64_t)gc) {
        uint8_t* d = dst + written;
        memset(d, 0, (size_t)gb);
        for (int j = 0; j < gc; ++j) {
            const uint32_t r = (i + (uint64_t)j < count) ? (reps[i + j] & mask) : 0u;      /* partial group: zero padded (:409-411) */
            pk_put(d, pk_off(j, n, gc, endian), r, n);
        }
        written += (uint64_t)gb;
    }
    return written;
}

uint64_t lofo_unpack_reps(uint32_t* reps, const uint8_t* src, uint64_t count, int n, int endian)
{
    const int gc = lofo_pack_group_count(n), gb = lofo_pack_group_bytes(n);
    uint64_t consumed = 0;
    for (uint64_t i = 0; i < count; i += (uint64_t)gc) {
        for (int j = 0; j < gc && i + (uint64_t)j < count; ++j) reps[i + j] = pk_get(src + consumed, pk_off(j, n, gc, endian), n);
        consumed += (uint64_t)gb;
    }
    return consumed;
}

uint64_t lofo_encode_f32(uint8_t* dst, const float* src, uint64_t count, const lofo_codec_format* f,
                         const lofo_projspec* ps, int endian, const lofo_rng* rng, uint64_t first_index)
{
    const int n = f->fmt.bitwidth;
    const int gc = lofo_pack_group_count(n), gb = lofo_pack_group_bytes(n);
    const int nbits = lofo_rand_bits_needed(ps->rounding_mode, ps->stoch_length);
    const int64_t groups = (int64_t)((count + (uint64_t)gc - 1) / (uint64_t)gc);
#pragma omp parallel for schedule(static)
    for (int64_t g = 0; g < groups; ++g) {
        uint8_t* d = dst + (uint64_t)g * (uint64_t)gb;
        memset(d, 0, (size_t)gb);
        for (int j = 0; j < gc; ++j) {
            const uint64_t i = (uint64_t)g * (uint64_t)gc + (uint64_t)j;
            const float x = (i < count) ? src[i] : 0.0f;                                    /* :425-429 */
            const uint32_t w = nbits ? lofo_rand_event(rng, first_index + i, LOFO_STREAM_ELEM, nbits) : 0u;
            pk_put(d, pk_off(j, n, gc, endian), lofo_float_to_rep(x, f, ps, w), n);
        }
    }
    return (uint64_t)groups * (uint64_t)gb;
}

uint64_t lofo_decode_f32(float* dst, const uint8_t* src, uint64_t count, const lofo_codec_format* f, int endian)
{
    const int n = f->fmt.bitwidth;
    const int gc = lofo_pack_group_count(n), gb = lofo_pack_group_bytes(n);
    const int64_t groups = (int64_t)((count + (uint64_t)gc - 1) / (uint64_t)gc);
#pragma omp parallel for schedule(static)
    for (int64_t g = 0; g < groups; ++g) {
        for (int j = 0; j < gc; ++j) {
            const uint64_t i = (uint64_t)g * (uint64_t)gc + (uint64_t)j;
            if (i < count) dst[i] = lofo_rep_to_float(pk_get(src + (uint64_t)g * (uint64_t)gb, pk_off(j, n, gc, endian), n), f);
        }
    }
    return (uint64_t)groups * (uint64_t)gb;
}

void lofo_mx_encode_f32(const float* in, int64_t n, int block_size, const lofo_codec_format* scale_fmt,
                        const lofo_codec_format* elem_fmt, const lofo_projspec* ps_scale, const lofo_projspec* ps_elem, int endian,
                        const lofo_rng* rng, uint64_t first_index, uint8_t* scales, uint8_t* elems)
{
    if (block_size <= 0 || n <= 0) return;
    const lofo_format* pub = &scale_fmt->fmt;
    const lofo_format* priv = &elem_fmt->fmt;
    const float pmn = lofo_mx_priv_max_normal(priv);
    const lofo_consts kpub = lofo_consts_make(pub), kpriv = lofo_consts_make(priv);
    const int need_rng_e = mode_uses_rng(ps_elem->rounding_mode) && rng != NULL;
    const int need_rng_s = mode_uses_rng(ps_scale->rounding_mode) && rng != NULL;
    const int nbits_e = lofo_rand_bits_needed(ps_elem->rounding_mode, ps_elem->stoch_length);
    const lofo_projspec exact = { LOFO_RTZ, LOFO_OVF_INF, 0 };          /* on-grid values convert exactly under any mode */
    const int64_t nblocks = n / block_size;
    uint32_t* reps = (uint32_t*)malloc((size_t)n * sizeof(uint32_t));
#pragma omp parallel for schedule(static)
    for (int64_t blk = 0; blk < nblocks; ++blk) {
        const int64_t b = blk * block_size;
        float amax = 0.0f;
        for (int i = 0; i < block_size; ++i) { const float v = fabsf(in[b + i]); amax = (amax < v) ? v : amax; }
        const uint64_t gblk = (first_index + (uint64_t)b) / (uint64_t)block_size;
        const uint32_t ws = need_rng_s ? lofo_rand_word(rng, gblk, LOFO_STREAM_SCALE) : 0u;
        const float scale = u2f(vround_params_f32(f2u(amax / pmn), pub, &kpub, ps_scale->rounding_mode, 0, ws));
        scales[blk] = (uint8_t)lofo_float_to_rep(scale, scale_fmt, &exact, 0u);
        for (int i = 0; i < block_size; ++i) {
            float q = 0.0f;
            if (scale > 0.0f) {
                const uint32_t w = need_rng_e ? lofo_rand_event(rng, first_index + (uint64_t)(b + i), LOFO_STREAM_ELEM, nbits_e) : 0u;
                q = u2f(vround_params_f32(f2u(in[b + i] / scale), priv, &kpriv, ps_elem->rounding_mode, ps_elem->stoch_length, w));
            }
            reps[b + i] = lofo_float_to_rep(q, elem_fmt, &exact, 0u);
        }
    }
    lofo_pack_reps(elems, reps, (uint64_t)n, elem_fmt->fmt.bitwidth, endian);
    free(reps);
}

void lofo_mx_decode_f32(const uint8_t* scales, const uint8_t* elems, float* out, int64_t n, int block_size,
                        const lofo_codec_format* scale_fmt, const lofo_codec_format* elem_fmt, int endian)
{
    if (block_size <= 0 || n <= 0) return;
    lofo_decode_f32(out, elems, (uint64_t)n, elem_fmt, endian);
    for (int64_t i = 0; i < n; ++i) out[i] = out[i] * lofo_rep_to_float(scales[i / block_size], scale_fmt);
}

int lofo_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
