/*
 * lof_oracle.h -- CPU ORACLE for the LoFloat virtual-rounding hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it, and only as the
 * checker or as the reported CPU baseline.  The product path (lofloat_b200/) never links,
 * imports or calls anything in oracle/.
 *
 * It is a plain-C restatement of the reference's scalar algorithms:
 *   - RoundMantissa + 15 helpers ........ /root/reference/src/lo_float.h:570-1272
 *   - virtual_round(value, mant_bits) ... /root/reference/src/lo_float.h:2715-2756
 *   - virtual_round(value, params) ...... /root/reference/src/lo_float.h:2843-2939
 *   - virtual_mx_round (CPU) ............ /root/reference/src/lo_float.h:4311-4383 and
 *                                         /root/reference/src/LoPy_bind.cpp:197-215
 *   - lof_round_to_odd_fma .............. /root/reference/src/blis_lof_gemm.hpp:63-99
 *   - lof_sgemm_ukr accumulation ........ /root/reference/src/blis_lof_gemm.hpp:147-262
 *     (+ BLIS generic KC=256 panelling, third_party/blis/ref_kernels/bli_cntx_ref.c:383)
 *   - storage codec ..................... /root/reference/src/lo_pack.hpp:58-319 (group geometry, put/get bits,
 *                                         convert_float_to_rep, convert_rep_to_float, LoF_encode / LoF_decode)
 * plus three statements that have NO reference function of their own ("parity unpinned" for these as a whole; each is a
 * composition of pinned pieces and is tied back to them by tests/test_oracle_golden.py):
 *   - lofo_gemm_chunked_f32 (k-chunk accumulation: the design intent of gemms.hpp:430-552, which is not runnable)
 *   - lofo_mx_encode_f32 / lofo_mx_decode_f32 (virtual_mx_round's rounding composed with the codec)
 *   - o_fwht in tests/oracle_lib.py (numpy restatement of LoFloat/layers.py:182-216)
 *
 * PARITY PINNING: the restatement is pinned against the reference itself, compiled from
 * /root/reference by oracle/Makefile into oracle/_ref/libref_shim.so (tests/test_oracle_vs_ref.py,
 * runs wherever /root/reference exists) and against golden vectors generated from that
 * shim and committed under tests/golden/ (tests/test_oracle_golden.py, runs everywhere).
 * float and double sources and the storage codec (17 formats x 9 deterministic modes x 3 saturation modes, decode tables,
 * packed bytes in both slot orders, the typed Templated_Float<Fp> ground truth) are pinned that way; the fp16 / bf16 source types go through
 * the same macro-expanded core but the reference's host build has no 16-bit source type
 * (c10::Half exists only under USE_CUDA), so for those two "parity unpinned" applies to
 * the container type handling (conversions), not to the rounding core.
 *
 * Randomness.  The reference's host helpers draw from a std::mt19937 re-seeded on every
 * call (lo_float.h:606,666,705,...) and its device helpers from a launch-geometry
 * dependent SplitMix64 hash (lo_float.h:83-98).  Neither is reproducible across
 * implementations, so this project defines ONE 32-bit random word W per rounding event:
 *     W = Philox4x32-10( key = seed, counter = (index>>2, offset ^ stream) )[index & 3]
 * and every stochastic recipe below consumes W exactly like the reference consumes its
 * draw (see lof_oracle.c).  rng.fixed != 0 replaces W by rng.fixed_word for every
 * element -- that is how the recipes are pinned against the reference's re-seeded
 * mt19937 (same draw for every element, tests/test_oracle_vs_ref.py).
 */
#ifndef LOF_ORACLE_H
#define LOF_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* enum values == reference (fp_tools.hpp:13-50), SURVEY.md Appendix D */
enum {
    LOFO_RNE = 0, LOFO_RTZ = 1, LOFO_RAZ = 2, LOFO_SR_A = 3, LOFO_RNO = 4, LOFO_RD = 5,
    LOFO_RU = 6, LOFO_RTA = 7, LOFO_SR_B = 8, LOFO_SR_C = 9, LOFO_TRUE_SR = 10,
    LOFO_PROB = 11, LOFO_RTO = 12, LOFO_RTTZ = 13, LOFO_SR_D = 14
};
enum { LOFO_EXTENDED = 0, LOFO_SATURATING = 1 };
enum { LOFO_SIGNED = 0, LOFO_UNSIGNED = 1 };
enum { LOFO_F32 = 0, LOFO_F16 = 1, LOFO_BF16 = 2, LOFO_F64 = 3 };
enum { LOFO_STREAM_ELEM = 0, LOFO_STREAM_SCALE = 0x5ca1e000u, LOFO_STREAM_GEMM = 0x6e3a0000u };

typedef struct {
    int32_t bitwidth, mantissa_bits, bias;
    int32_t ov_behavior, na_behavior, is_signed, unsigned_behavior;
} lofo_format;

typedef struct { int32_t rounding_mode, saturation_mode, stoch_length; } lofo_projspec;

typedef struct {
    uint64_t seed, offset;   /* Philox key / counter offset                         */
    int32_t  fixed;          /* != 0: every rounding event uses fixed_word          */
    uint32_t fixed_word;
} lofo_rng;

void     lofo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
uint32_t lofo_rand_word(const lofo_rng* rng, uint64_t index, uint32_t stream);      /* dense 32-bit word */
int      lofo_rand_bits_needed(int mode, int stoch_len);
uint32_t lofo_rand_event(const lofo_rng* rng, uint64_t index, uint32_t stream, int nbits); /* per-event word, see lof_oracle.c */

/* scalar entry points (bit patterns in / out, container = fp32) */
uint32_t lofo_round_params_bits_f32(uint32_t bits, const lofo_format* f, const lofo_projspec* ps, uint32_t w);
uint32_t lofo_round_mantissa_bits_f32(uint32_t bits, int mant_bits, const lofo_projspec* ps, uint32_t w);

/* array entry points. dtype = LOFO_F32/F16/BF16/F64; index of element i is first_index+i.
 * out may alias in.  `scale` follows Lof_kernel.cu:41,53 (skip the multiply when == 1). */
void lofo_round_params(const void* in, void* out, int64_t n, int dtype, const lofo_format* f,
                       const lofo_projspec* ps, double scale, const lofo_rng* rng, uint64_t first_index);
void lofo_round_mantissa(const void* in, void* out, int64_t n, int dtype, int mant_bits,
                         const lofo_projspec* ps, double scale, const lofo_rng* rng, uint64_t first_index);

/* MX block rounding, fp32 only, in-place semantics of LoPy_bind.cpp:197-215 (out may alias in).
 * ps_scale.stoch_length is forced to 0 like the reference. */
void lofo_mx_round_f32(const float* in, float* out, int64_t n, int block_size,
                       const lofo_format* scale_fmt, const lofo_format* elem_fmt,
                       const lofo_projspec* ps_scale, const lofo_projspec* ps_elem,
                       const lofo_rng* rng, uint64_t first_index);
float lofo_mx_priv_max_normal(const lofo_format* elem_fmt);

/* round-to-odd fused multiply-add, blis_lof_gemm.hpp:63-99 */
float lofo_round_to_odd_fma(float a, float b, float c);

/* LoFloat-accumulation GEMM, row-major A(MxK) B(KxN) C(MxN):
 *   per kc-panel: acc=0; for l: acc = round_mantissa(rto_fma(a_il, b_lj, acc), accum_mant_bits, ps)
 *   panel result *= alpha; C = (first panel ? (beta==0 ? ab : beta*C+ab) : C + ab)
 * kc <= 0 means one panel spanning all of K (the reference's CUDA leaf semantics,
 * cutlass_gemms.cuh:97-168); kc = 256 is the BLIS generic CPU path.
 * accum_mant_bits < 0 => plain fp32 accumulation (blis_lof_gemm.hpp:230-240). */
void lofo_gemm_f32(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K,
                   int accum_mant_bits, const lofo_projspec* ps, int64_t kc,
                   float alpha, float beta, const lofo_rng* rng);

/* round-to-odd add of the chunked mode: RZ(a + b) | inexact */
float lofo_rto_add(float a, float b);

/* chunked tensor-core semantics oracle (our tcgen05 mode, NOT a reference function; "parity unpinned" --
 * there is no runnable reference for chunked accumulation):
 *   per k-chunk: P = exact dot over the chunk, rounded once to fp32 (computed in double)
 *   acc = round(acc + P)  or, with partial_round, acc = round(acc + round(P)); round() into the
 *   accumulator format (mantissa-only if acc_fmt == NULL). */
void lofo_gemm_chunked_f32(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K,
                           int accum_mant_bits, const lofo_format* acc_fmt, const lofo_projspec* ps,
                           int64_t k_chunk, float alpha, int partial_round);
/* ... with stochastic accumulator rounding: one dense Philox word per event (lofo_gemm_f32's stream), event index
 * ((which * n_chunks + chunk) * M + i) * N + j, which = 0 partial / 1 accumulator; rng == NULL draws the word 0. */
void lofo_gemm_chunked_rng_f32(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K,
                               int accum_mant_bits, const lofo_format* acc_fmt, const lofo_projspec* ps,
                               int64_t k_chunk, float alpha, int partial_round, const lofo_rng* rng);

/* ------------------------------------------------------------------------------------
 * Low-bit storage codec ("next" row f-1): /root/reference/src/lo_pack.hpp.
 * The Inf / NaN functor objects of FloatingPointParams (fp_tools.hpp:187-252) travel as data:
 *     IsInf(rep) := (rep & inf_mask) == inf_value
 *     IsNaN(rep) := (rep & nan_mask) == nan_value  &&  (nan_mant_mask == 0 || (rep & nan_mant_mask) != 0)
 * ("never" = mask 0, value 1), which covers every checker the reference defines (lo_float.h:1411-1830,
 * fp_tools.hpp:257-440, LoFloat/formats.py).
 * ---------------------------------------------------------------------------------- */
typedef struct {
    lofo_format fmt;
    uint32_t inf_mask, inf_value;
    uint32_t nan_mask, nan_value, nan_mant_mask;
    uint32_t min_pos_inf;          /* IsInf.minPosInf()       */
    uint32_t qnan, snan;           /* IsNaN.qNanBitPattern() / sNanBitPattern() */
} lofo_codec_format;

enum { LOFO_PACK_FORWARD = 0, LOFO_PACK_REVERSE = 1 };      /* lo_pack.hpp:52 PackEndian */
enum { LOFO_OVF_INF = 0, LOFO_SAT_FINITE = 1, LOFO_SAT_PROPAGATE = 2 };   /* fp_tools.hpp:38-41 */

uint32_t lofo_float_to_rep(float from, const lofo_codec_format* f, const lofo_projspec* ps, uint32_t w);   /* lo_pack.hpp:109-267 */
float    lofo_rep_to_float(uint32_t rep, const lofo_codec_format* f);                                     /* lo_pack.hpp:277-319 */
void     lofo_float_to_rep_array(const float* in, uint32_t* out, int64_t n, const lofo_codec_format* f, const lofo_projspec* ps,
                                 const lofo_rng* rng, uint64_t first_index);
void     lofo_rep_to_float_array(const uint32_t* in, float* out, int64_t n, const lofo_codec_format* f);
int      lofo_pack_group_count(int n);                                                                    /* lo_pack.hpp:65-68 */
int      lofo_pack_group_bytes(int n);
uint64_t lofo_packed_size(int n, uint64_t count);                                                         /* lo_pack.hpp:70-73, 399 */
/* raw n-bit reps <-> packed bytes (the typed API 2 of lo_pack.hpp:329-341, 359-371); reps as uint32 */
uint64_t lofo_pack_reps(uint8_t* dst, const uint32_t* reps, uint64_t count, int n, int endian);
uint64_t lofo_unpack_reps(uint32_t* reps, const uint8_t* src, uint64_t count, int n, int endian);
/* float arrays <-> packed bytes (LoF_encode / LoF_decode, lo_pack.hpp:419-432, 452-465); stochastic modes draw the
 * event word of global element first_index + i like lofo_round_params.  Return bytes written / consumed. */
uint64_t lofo_encode_f32(uint8_t* dst, const float* src, uint64_t count, const lofo_codec_format* f,
                         const lofo_projspec* ps, int endian, const lofo_rng* rng, uint64_t first_index);
uint64_t lofo_decode_f32(float* dst, const uint8_t* src, uint64_t count, const lofo_codec_format* f, int endian);

/* MX quantisation with MATERIALISED scales (row f-1, second half): the blocks, amax, scale and element rounding of
 * virtual_mx_round (lo_float.h:4311-4383) composed with the storage codec (lo_pack.hpp):
 *     scales[b] = convert_float_to_rep(scale_b, scale codec format)                      one byte per block (bitwidth <= 8)
 *     elems     = packed convert_float_to_rep(virtual_round(x / scale_b, element format), element codec format)
 * (a zero / non-positive scale gives an all-zero block).  decode(elems) * decode(scale) reproduces virtual_mx_round
 * wherever the fake-quant value is representable (virtual_round returns an overflowing input unchanged for
 * Saturating formats; storage necessarily clamps it).  n % block_size == 0. */
void lofo_mx_encode_f32(const float* in, int64_t n, int block_size, const lofo_codec_format* scale_fmt,
                        const lofo_codec_format* elem_fmt, const lofo_projspec* ps_scale, const lofo_projspec* ps_elem, int endian,
                        const lofo_rng* rng, uint64_t first_index, uint8_t* scales, uint8_t* elems);
void lofo_mx_decode_f32(const uint8_t* scales, const uint8_t* elems, float* out, int64_t n, int block_size,
                        const lofo_codec_format* scale_fmt, const lofo_codec_format* elem_fmt, int endian);

int lofo_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
