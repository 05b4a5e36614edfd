/*
 * lofloat_b200.h -- C ABI of the B200-native LoFloat virtual-rounding hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  Every entry
 * point names the reference launcher it replaces (paths relative to /root/reference).
 * Unless stated otherwise pointers are DEVICE pointers on the current CUDA device, calls
 * are asynchronous on `stream` (a cudaStream_t passed as void*; NULL = legacy default
 * stream, which is what the reference uses, Lof_kernel.cu:47,59,80) and return a
 * lof_status_t.  There is no CPU fallback: without a usable CUDA device every compute
 * entry point returns LOF_ERR_NO_DEVICE / LOF_ERR_CUDA.
 *
 * Parameter structs keep the reference's field meaning and enum VALUES
 * (src/fp_tools.hpp:13-124, 78-88, 187-252), so a binding can copy them field by field.
 */
#ifndef LOFLOAT_B200_H
#define LOFLOAT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LOFLOAT_B200_ABI_VERSION 1

/* the library is built with -fvisibility=hidden; only the entry points below are exported */
#if defined(__GNUC__)
#define LOF_API __attribute__((visibility("default")))
#else
#define LOF_API
#endif

/* lo_float::Rounding_Mode, fp_tools.hpp:13-28 (same integer values) */
typedef enum {
    LOF_RoundToNearestEven = 0,
    LOF_RoundTowardsZero = 1,
    LOF_RoundAwayFromZero = 2,
    LOF_StochasticRoundingA = 3,
    LOF_RoundToNearestOdd = 4,
    LOF_RoundDown = 5,
    LOF_RoundUp = 6,
    LOF_RoundTiesToAway = 7,
    LOF_StochasticRoundingB = 8,
    LOF_StochasticRoundingC = 9,
    LOF_True_StochasticRounding = 10,
    LOF_ProbabilisticRounding = 11,
    LOF_RoundToOdd = 12,
    LOF_RoundTiesTowardsZero = 13,
    LOF_StochasticRoundingD = 14,
    LOF_NUM_ROUNDING_MODES = 15
} lof_rounding_mode_t;

typedef enum { LOF_OvfInf = 0, LOF_SatFinite = 1, LOF_SatPropagate = 2 } lof_saturation_mode_t; /* fp_tools.hpp:38-41 */
typedef enum { LOF_Signed = 0, LOF_Unsigned = 1 } lof_signedness_t;                             /* fp_tools.hpp:30-32 */
typedef enum { LOF_Extended = 0, LOF_Saturating = 1 } lof_inf_behavior_t;                       /* fp_tools.hpp:34-36 */
typedef enum { LOF_NaN_3109 = 0, LOF_NaN_754 = 1 } lof_nan_behavior_t;                          /* fp_tools.hpp:44-46 */
typedef enum { LOF_NegtoZero = 0, LOF_NegtoNaN = 1 } lof_unsigned_behavior_t;                   /* fp_tools.hpp:48-50 */

/* container (storage) type of the tensor being rounded.  The reference dispatches
 * float/double/c10::Half (LoPy_bind.cpp:83,133); bf16 is new. */
typedef enum { LOF_F32 = 0, LOF_F16 = 1, LOF_BF16 = 2, LOF_F64 = 3,
               /* tensor-core OPERAND encodings (lof_gemm_chunked_f32 only): OCP e4m3fn / e5m2 */
               LOF_E4M3 = 4, LOF_E5M2 = 5 } lof_dtype_t;

typedef enum {
    LOF_OK = 0,
    LOF_ERR_INVALID_ARG = 1,   /* null pointer, negative size, malformed format        */
    LOF_ERR_UNSUPPORTED = 2,   /* e.g. MX block size not in {16,32,64,128,256}         */
    LOF_ERR_CUDA = 3,          /* a CUDA call failed; see lof_last_error_string()      */
    LOF_ERR_NO_DEVICE = 4,     /* no CUDA device: there is NO CPU fallback             */
    LOF_ERR_ALLOC = 5
} lof_status_t;

/* lo_float::FloatingPointParams without the Inf/NaN functor objects (they are never read
 * on this path: lo_float.h:2843-2939), fp_tools.hpp:187-252 */
typedef struct lof_format_t {
    int32_t bitwidth;           /* total bits incl. sign                              */
    int32_t mantissa_bits;      /* explicit fraction bits                             */
    int32_t bias;               /* exponent bias                                      */
    int32_t ov_behavior;        /* lof_inf_behavior_t                                 */
    int32_t na_behavior;        /* lof_nan_behavior_t   (carried, not read)           */
    int32_t is_signed;          /* lof_signedness_t                                   */
    int32_t unsigned_behavior;  /* lof_unsigned_behavior_t (carried, not read)        */
} lof_format_t;

/* lo_float::ProjSpec, fp_tools.hpp:78-88 */
typedef struct lof_projspec_t {
    int32_t rounding_mode;      /* lof_rounding_mode_t                                */
    int32_t saturation_mode;    /* lof_saturation_mode_t (carried, not read)          */
    int32_t stoch_length;       /* random bits for StochasticRoundingA..D             */
} lof_projspec_t;

/* Counter-based RNG state for the stochastic modes: Philox4x32-10 keyed by `seed`; dense 32-bit word q of the
 * stream is word (q & 3) of block (q >> 2) at counter offset `offset`.  A rounding event (global element index i)
 * consumes only the random bits its mode needs (SR-A/B/C: stoch_length, SR-D: stoch_length + 1, Probabilistic: 1,
 * True_SR: 32), packed so that one Philox block serves up to 16 events:
 *     > 16 bits: dense word i;   9..16 bits: half (i & 1) of dense word i >> 1;
 *     <= 8 bits: byte (i & 3) of dense word 4 * (((v >> 10) << 8) | (v & 255)) + ((v >> 8) & 3),  v = i >> 2.
 * Results depend only on (seed, offset, global element index) -- never on launch geometry or sharding -- unlike
 * the reference's device hash (lo_float.h:83-98).  The GEMM and the MX block scale draw one dense word per event.
 * NULL = {0, 0}. */
typedef struct lof_rng_t {
    uint64_t seed;
    uint64_t offset;
} lof_rng_t;

/* ---------------------------------------------------------------------------------------
 * Elementwise virtual rounding.
 * ------------------------------------------------------------------------------------- */

/* out[i] = virtual_round(scale * in[i], *fmt, *ps)            (out may alias in)
 * Replaces  lo_float::round_fp_params<T>(T* inout, n, params, ps, scale)
 *           src/Lof_kernel.cu:49-60 (kernel :50, launcher :58), whose element function is
 *           lo_float::virtual_round(value, FloatingPointParams, ProjSpec) lo_float.h:2843-2939.
 * `first_index` is the global index of in[0] (row-sharded tensors pass their shard offset
 * so stochastic results do not depend on the partition). */
LOF_API int lof_round_params(const void* in, void* out, int64_t n, int dtype,
                     const lof_format_t* fmt, const lof_projspec_t* ps, double scale,
                     const lof_rng_t* rng, uint64_t first_index, void* stream);

/* out[i] = virtual_round(scale * in[i], mantissa_bits, *ps)   (out may alias in)
 * Replaces  lo_float::round_mantissa<T>(const T* in, T* out, n, mantissa_bits, ps, scale)
 *           src/Lof_kernel.cu:37-48; element function lo_float.h:2715-2756. */
LOF_API int lof_round_mantissa(const void* in, void* out, int64_t n, int dtype, int mantissa_bits,
                       const lof_projspec_t* ps, double scale,
                       const lof_rng_t* rng, uint64_t first_index, void* stream);

/* ---------------------------------------------------------------------------------------
 * MX (microscaling) block rounding, fp32.
 * ------------------------------------------------------------------------------------- */

/* Per contiguous block of `block_size` elements (16/32/64/128/256; trailing partial block
 * allowed): amax = max|x|; scale = virtual_round(amax / priv_max_normal, *scale_fmt,
 * {ps_scale->rounding_mode, -, 0}); out = scale > 0 ? virtual_round(in / scale, *elem_fmt,
 * *ps_elem) * scale : 0.                                     (out may alias in)
 * Replaces  lo_float::run_virtual_mx_round(float* inout, n, block_size, params_public,
 *           params_private, ps_public, ps_private)  src/Lof_kernel.cu:70-94 (kernel
 *           src/mx_round.cuh:27-214); results follow the reference's CPU definition
 *           LoPy_bind.cpp:197-215 == lo_float.h:4311-4383 (amax DIVIDED by priv_max_normal).
 * `first_index` must be a multiple of block_size. */
LOF_API int lof_mx_round_f32(const float* in, float* out, int64_t n, int block_size,
                     const lof_format_t* scale_fmt, const lof_format_t* elem_fmt,
                     const lof_projspec_t* ps_scale, const lof_projspec_t* ps_elem,
                     const lof_rng_t* rng, uint64_t first_index, void* stream);

/* ---------------------------------------------------------------------------------------
 * LoFloat-accumulation GEMM, fp32 row-major contiguous A(MxK), B(KxN), C(MxN).
 * ------------------------------------------------------------------------------------- */

/* Exact mode: per output element and per kc-panel of K,
 *     acc = 0; for l in panel: acc = virtual_round(round_to_odd_fma(a_il, b_lj, acc), accum_mant_bits, *ps)
 *     panel *= alpha; C = first panel ? (beta == 0 ? panel : beta*C + panel) : C + panel
 * kc = 256 reproduces the reference CPU path bit for bit (bli_sgemm + lof_sgemm_ukr,
 * src/blis_lof_gemm.hpp:147-262, BLIS generic KC = 256); kc <= 0 = one panel over all K,
 * the semantics of the reference CUDA leaf (src/cutlass_gemms.cuh:97-168) which, unlike
 * the CPU path, ignores *ps (always RNE) -- pass ps = RNE to reproduce it.
 * Replaces  torch::Tensor lo_float::LoF_gemm(A, B, accum_mant_bits, round_mode,
 *           stochastic_rounding_bits, scale_a, scale_b)  src/Lof_kernel.cu:96-154 with
 *           alpha = 1/(scale_a*scale_b), beta = 0 (cutlass_gemms.cuh:1777-1787). */
LOF_API int lof_gemm_f32(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K,
                 int accum_mant_bits, const lof_projspec_t* ps, int64_t kc,
                 float alpha, float beta, const lof_rng_t* rng, void* stream);

/* Batched form of lof_gemm_f32: `batch` independent GEMMs C_b = A_b * B_b in ONE launch (matrix b starts at
 * A + b * stride_a etc., row-major, contiguous rows), replacing the Python per-(batch, head) loop of `_lof_gemm_2d`
 * (LoFloat/layers.py:142-149).  Each C_b is bit-identical to the single call on (A_b, B_b); for the stochastic modes the
 * rounding-event index continues across the batch. */
LOF_API int lof_gemm_batched_f32(const float* A, const float* B, float* C, int64_t batch, int64_t M, int64_t N, int64_t K,
                                 int64_t stride_a, int64_t stride_b, int64_t stride_c, int accum_mant_bits,
                                 const lof_projspec_t* ps, int64_t kc, float alpha, float beta, const lof_rng_t* rng, void* stream);

/* x(M x K) times an nn.Linear weight stored (N x K) row-major, plus an optional (already quantised) bias of N values
 * added to alpha * acc in plain fp32: `_lof_linear` of LoFloat/layers.py:152-168 without the `weight.t().contiguous()`
 * copy and without the extra pass of `out + b_q` (layers.py:392-398).  Bit-identical to lof_gemm_f32 on the transposed
 * weight followed by the addition. */
LOF_API int lof_linear_f32(const float* x, const float* weight, const float* bias, float* out, int64_t M, int64_t N, int64_t K,
                           int accum_mant_bits, const lof_projspec_t* ps, int64_t kc, float alpha, const lof_rng_t* rng, void* stream);

/* Conv2d forward as an implicit GEMM with LoFloat accumulation (no im2col matrix in HBM): input NCHW, weight OIHW
 * (C_out, C_in / groups, R, S), out NCHW (batch, C_out, P, Q), all fp32 contiguous.  Every output element is the chain
 *     acc = virtual_round(round_to_odd_fma(w[co, k], x_col[k, (p, q)], acc), accum_mant_bits, *ps)   for k ascending
 * over the reduction axis in `k_order`:
 *   LOF_CONV_K_CRS  k = (c, r, s): F.unfold order == LoF_Conv2d(backend="im2col"), LoFloat/layers.py:710-745 (one
 *                   lof_gemm(w_col, x_col[b]) per image and group); kc = 256 matches that path on the reference CPU GEMM
 *   LOF_CONV_K_RSC  k = (r, s, c): the NHWC / KRSC order of lo_float::LoF_conv2d, src/Lof_kernel.cu:156-228
 *                   (CUTLASS implicit GEMM, one accumulator over all of K: kc <= 0), which this entry point replaces
 * Zero padding contributes fma(w, 0, acc) = acc, already on the accumulator grid.  out = alpha * acc (+ bias[co]),
 * alpha = 1 / (weight_scale * input_scale) (Lof_kernel.cu:209-212); `bias` may be NULL. */
enum { LOF_CONV_K_CRS = 0, LOF_CONV_K_RSC = 1 };
typedef struct lof_conv2d_t {
    int64_t batch, c_in, h, w;            /* input  (batch, c_in, h, w) */
    int64_t c_out, r, s;                  /* weight (c_out, c_in / groups, r, s) */
    int64_t pad_h, pad_w, stride_h, stride_w, dilation_h, dilation_w, groups;
    int32_t k_order;
} lof_conv2d_t;
LOF_API int lof_conv2d_f32(const float* input, const float* weight, const float* bias, float* out, const lof_conv2d_t* geometry,
                           int accum_mant_bits, const lof_projspec_t* ps, int64_t kc, float alpha, const lof_rng_t* rng, void* stream);
/* floor((in + 2 pad - dilation (kernel - 1) - 1) / stride) + 1, 0 if the window does not fit (Lof_kernel.cu:194-195) */
LOF_API int lof_conv2d_out_size(int64_t in, int64_t kernel, int64_t pad, int64_t stride, int64_t dilation, int64_t* out);

/* Device self-test of the exact GEMM's float-domain RNE step ((t + K) - K, lof_gemm_exact.cu): every fp32 pattern x every
 * accumulator width against the integer routine; adds the number of mismatches to *dev_mismatches (device memory, uint64). */
LOF_API int lof_selftest_gemm_rne_round(uint64_t* dev_mismatches, void* stream);

/* Chunked tensor-core mode (tcgen05 + TMEM): after every k_chunk columns of K the fp32 partial product P of
 * the chunk (tensor core, fp32 accumulate) is folded into the running accumulator, which lives in the
 * accumulator format:
 *     default                      acc = virtual_round(acc + P)                    (ONE rounding of the exact sum)
 *     LOF_GEMM_PARTIAL_ROUND       acc = virtual_round(acc + virtual_round(P))     (typed micro-accumulator)
 *     C = alpha * acc
 * virtual_round = mantissa-only rounding to accum_mant_bits like the reference GEMM leaf (acc_fmt == NULL) or
 * rounding into the full format *acc_fmt incl. its exponent range (e.g. E5M10), under any deterministic
 * rounding mode.  This is the design intent of src/gemms.hpp:430-552 (KR-chunk micro-accumulator); it has no
 * runnable reference -- its CPU statement is oracle/lof_oracle.c.
 * A and B are fp32 containers whose values must be exactly representable in `operand_dtype` (LOF_E4M3,
 * LOF_E5M2 or LOF_BF16); otherwise LOF_GEMM_STATUS_{A,B}_INEXACT is OR-ed into *dev_status (a DEVICE int32,
 * may be NULL) and C is meaningless.  k_chunk must be a multiple of 32 (8-bit operands) / 16 (bf16);
 * k_chunk <= 0 = one chunk over all K.  K * sizeof(operand) must be a multiple of 16.  Asynchronous on stream. */
#define LOF_GEMM_PARTIAL_ROUND       1u   /* flags */
#define LOF_GEMM_NO_PACKED           2u   /* never use the hardware-format shortcuts (packed half2/bfloat162 fold, fp16 TMEM accumulator) */
#define LOF_GEMM_NO_HW_ACCUM         4u   /* never let the tensor core keep the accumulator in fp16 (k_chunk == one MMA, E5M10, RNE) */
#define LOF_GEMM_NO_PAIR             8u   /* never use the persistent CTA-pair (cta_group::2, 256 x 256 tile) main loop */
#define LOF_GEMM_NO_PERSISTENT      16u   /* never use the persistent per-chunk fold kernel (keeps the first-generation one-tile-per-CTA kernel) */
#define LOF_GEMM_NO_FLOAT_FOLD      32u   /* never round the accumulator in the float domain (integer routine only) */
#define LOF_GEMM_STATUS_A_INEXACT    1    /* *dev_status bits */
#define LOF_GEMM_STATUS_B_INEXACT    2
#define LOF_GEMM_STATUS_PACKED_RANGE 4    /* informational: a hardware-format shortcut left the fp16 range, C was recomputed generically */
LOF_API int lof_gemm_chunked_f32(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K,
                         int accum_mant_bits, const lof_format_t* acc_fmt, const lof_projspec_t* ps,
                         int64_t k_chunk, float alpha, int operand_dtype, uint32_t flags, int32_t* dev_status,
                         void* stream);

/* The same with STOCHASTIC accumulator rounding (any of the six stochastic modes; mantissa-only accumulator, acc_fmt == NULL): the reference
 * GEMM leaf honours every rounding mode (src/blis_lof_gemm.hpp:217-228).  One dense Philox word per rounding event (stream of
 * lof_gemm_f32), event index = ((which * n_chunks + chunk) * M + row) * N + col with which = 0 for the partial's rounding
 * (LOF_GEMM_PARTIAL_ROUND) and 1 for the accumulator's; rng == NULL = seed 0.  Deterministic modes behave exactly like lof_gemm_chunked_f32. */
LOF_API int lof_gemm_chunked_rng_f32(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K,
                             int accum_mant_bits, const lof_format_t* acc_fmt, const lof_projspec_t* ps,
                             int64_t k_chunk, float alpha, int operand_dtype, uint32_t flags, const lof_rng_t* rng,
                             int32_t* dev_status, void* stream);

/* ---------------------------------------------------------------------------------------
 * Callers fused around the path: the orthonormal block-diagonal Walsh-Hadamard rotation
 * LoF_Linear applies to its activations (LoFloat/layers.py:182-216 `_fwht`, used at :363-365) fused with the rounding
 * that follows it -- STERound (:378, with its fused pre-scale) or STEMXRound (:366-369).  One pass over HBM; bit-identical
 * to `_fwht` followed by the separate op (same butterflies in the same stage order, then * fl32(1 / sqrt(block_size))).
 * block_size: a power of two up to 1024 that divides the last dimension; n % 128 == 0; 16-byte aligned; out may alias in.
 * lof_fwht_f32 (the plain transform) also takes block sizes 2048 ... 32768 -- `hadamard_block_size=None` on an LLM-sized
 * in_features -- as two passes (inside the 1024-element chunks, then across them), still bit-identical to `_fwht`.
 * ------------------------------------------------------------------------------------- */
LOF_API int lof_fwht_f32(const float* in, float* out, int64_t n, int block_size, void* stream);
LOF_API int lof_fwht_round_f32(const float* in, float* out, int64_t n, int block_size, const lof_format_t* fmt,
                               const lof_projspec_t* ps, double scale, const lof_rng_t* rng, uint64_t first_index, void* stream);
LOF_API int lof_fwht_mx_round_f32(const float* in, float* out, int64_t n, int block_size, int mx_block_size,
                                  const lof_format_t* scale_fmt, const lof_format_t* elem_fmt, const lof_projspec_t* ps_scale,
                                  const lof_projspec_t* ps_elem, const lof_rng_t* rng, uint64_t first_index, void* stream);

/* ---------------------------------------------------------------------------------------
 * Low-bit storage codec: fp32 -> packed n-bit representations and back.
 * Replaces LoF_encode_cuda / LoF_decode_cuda (src/lo_pack.hpp:499-529; CPU twins LoF_encode / LoF_decode
 * :419-465) with the conversion semantics of convert_float_to_rep / convert_rep_to_float (:109-319): the strict
 * typed conversion (Saturation_Mode, Inf / NaN checkers, signedness), NOT virtual_round's container semantics.
 * The Inf / NaN functor objects of FloatingPointParams (fp_tools.hpp:187-252) travel as data:
 *     IsInf(rep) := (rep & inf_mask) == inf_value
 *     IsNaN(rep) := (rep & nan_mask) == nan_value  &&  (nan_mant_mask == 0 || (rep & nan_mant_mask) != 0)
 * ("never" = mask 0, value 1); this form covers every checker the reference defines.
 * bitwidth must be in [1, 16] (lo_pack.hpp:4-9).  Packed layout: groups of lcm(n, 8 or 16) bits, LSB-first inside
 * a byte, element i of a group in slot i (LOF_PACK_FORWARD) or slot G-1-i (LOF_PACK_REVERSE); a trailing partial
 * group is zero-padded.  For n = 8 / 16 the packed buffer is simply the array of representations.
 * ------------------------------------------------------------------------------------- */
typedef struct lof_codec_format_t {
    lof_format_t fmt;
    uint32_t inf_mask, inf_value;
    uint32_t nan_mask, nan_value, nan_mant_mask;
    uint32_t min_pos_inf;           /* IsInf.minPosInf()                              */
    uint32_t qnan, snan;            /* IsNaN.qNanBitPattern() / sNanBitPattern()      */
} lof_codec_format_t;
typedef enum { LOF_PACK_FORWARD = 0, LOF_PACK_REVERSE = 1 } lof_pack_endian_t;            /* lo_pack.hpp:52 */

/* LoF_packed_size (lo_pack.hpp:399): bytes of the packed buffer for `count` elements of an n-bit format; 0 if n is
 * outside [1, 16].  Host-only helper. */
LOF_API uint64_t lof_packed_size(int bitwidth, uint64_t count);
/* replaces LoF_encode_cuda<E>(d_dst, d_src, params, count, ps, stream) (lo_pack.hpp:499-513).  Stochastic modes draw
 * the event word of global element first_index + i (same stream definition as lof_round_params). */
LOF_API int lof_encode_f32(const float* src, uint8_t* dst, int64_t count, const lof_codec_format_t* fmt,
                           const lof_projspec_t* ps, int endian, const lof_rng_t* rng, uint64_t first_index, void* stream);
/* replaces LoF_decode_cuda<E>(d_dst, d_src, params, count, stream) (lo_pack.hpp:516-529) */
LOF_API int lof_decode_f32(const uint8_t* src, float* dst, int64_t count, const lof_codec_format_t* fmt, int endian,
                           void* stream);

/* MX quantisation with MATERIALISED scales (real MXFP8 / MXFP6 / MXFP4 storage; containers src/Matrix.h:79-176
 * MX_Matrix = element array + shared_exps array): the blocks, amax, scale and element rounding of virtual_mx_round
 * (src/lo_float.h:4311-4383) composed with the storage codec (src/lo_pack.hpp):
 *     scales[b] = convert_float_to_rep(scale_b, scale format)                       one byte per block
 *     elems     = packed convert_float_to_rep(virtual_round(x / scale_b, element format))   4-, 6- or 8-bit elements
 * decode(elems) * decode(scales) == lof_mx_round_f32's output wherever that value is representable (virtual_round returns
 * an overflowing x / scale unchanged for Saturating formats; storage clamps it).  block_size in {16, 32, 64, 128},
 * n % block_size == 0, `in` / `out` 16-byte and `elems` 4-byte aligned.  elems holds lof_packed_size(bits, n) bytes. */
LOF_API int lof_mx_encode_f32(const float* in, int64_t n, int block_size, const lof_codec_format_t* scale_fmt,
                              const lof_codec_format_t* elem_fmt, const lof_projspec_t* ps_scale, const lof_projspec_t* ps_elem,
                              int endian, const lof_rng_t* rng, uint64_t first_index, uint8_t* scales, uint8_t* elems, void* stream);
LOF_API int lof_mx_decode_f32(const uint8_t* scales, const uint8_t* elems, float* out, int64_t n, int block_size,
                              const lof_codec_format_t* scale_fmt, const lof_codec_format_t* elem_fmt, int endian, void* stream);

/* Strict typed-semantics rounding as an elementwise op: out[i] = decode(encode(in[i]))
 * in one pass = (float)Templated_Float<Fp>(in[i]), the typed conversion ConvertImpl::run (src/lo_float.h:3057-3364,
 * runtime twin src/lo_pack.hpp:109-319; behaviour spec test/test_sat_mode.cpp:11-18, test/test_unsigned.cpp): overflow
 * goes to +-inf or +-max by Saturation_Mode and OV_behavior, NaN to the format's NaN (read back as +qNaN), negatives into
 * unsigned formats to zero or NaN.  virtual_round (lof_round_params) keeps the reference's container semantics instead.
 * fp32 only, bitwidth in [1, 16], out may alias in. */
LOF_API int lof_round_strict_f32(const float* in, float* out, int64_t n, const lof_codec_format_t* fmt,
                                 const lof_projspec_t* ps, const lof_rng_t* rng, uint64_t first_index, void* stream);

/* Chunked tensor-core GEMM on operands that are ALREADY stored as 8-bit representations (what lof_encode_f32
 * produces for e4m3fn / e5m2): A is M x K bytes row-major, Bt is N x K bytes row-major (B transposed, K-major --
 * the [out_features, in_features] weight layout of LoF_Linear, LoFloat/layers.py:160-161).  operand_dtype is
 * LOF_E4M3 or LOF_E5M2 -- or LOF_BF16, in which case A and Bt hold bf16 words (M x K and N x K, 2 bytes each): the
 * form in which MX-rounded operands (an e4m3 / e5m2 / e3m2 / e2m1 value times a power-of-two block scale is exact in
 * bf16) travel between GPUs at half the fp32 volume (sharding.gemm_2d_packed).  Same accumulator semantics, flags and
 * status word as lof_gemm_chunked_f32; no conversion pass and no workspace for the operands.  K * sizeof(operand) must be
 * a multiple of 16. */
LOF_API int lof_gemm_chunked_packed(const uint8_t* A, const uint8_t* Bt, float* C, int64_t M, int64_t N, int64_t K,
                                    int accum_mant_bits, const lof_format_t* acc_fmt, const lof_projspec_t* ps,
                                    int64_t k_chunk, float alpha, int operand_dtype, uint32_t flags, int32_t* dev_status,
                                    void* stream);

/* Block-scaled (MX) tensor-core GEMM on MX STORAGE (tcgen05.mma.kind::mxf8f6f4.block_scale): the operands are what
 * lof_mx_encode_f32 writes for block_size 32 with an E8M0 scale format and an 8-bit element format -- element codes
 * a_elems [M x K] / bt_elems [N x K] (B transposed, K-major: the [out_features, in_features] weight layout) and one scale byte per
 * 32 elements of K, a_scales [M x K/32] / bt_scales [N x K/32], row-major (container: src/Matrix.h:79-176 MX_Matrix).  Replaces
 * the composition LoF_Linear(scaling="mx") runs (LoFloat/layers.py:366-377: virtual_mx_round of both operands, then lof_gemm on
 * the fp32 containers) without ever materialising the de-quantised operands: the tensor core applies 2^(sa - 127) 2^(sb - 127)
 * per MX block.  Accumulator semantics, flags and rounding modes as lof_gemm_chunked_f32 (k_chunk a multiple of 32; <= 0 = one
 * chunk: plain MXFP8 GEMM with fp32 accumulation and one final rounding).  elem_dtype: LOF_E4M3 (OCP e4m3fn codes; every e3m2 /
 * e2m3 / e2m1 value is an e4m3 value, so MXFP6 / MXFP4 tensors re-coded as e4m3 bytes run here too) or LOF_E5M2.  K % 32 == 0,
 * element arrays 16-byte aligned.  A scale byte of 0xFF (NaN) makes its block NaN.  Asynchronous on stream. */
LOF_API int lof_mx_gemm(const uint8_t* a_elems, const uint8_t* a_scales, const uint8_t* bt_elems, const uint8_t* bt_scales, float* C,
                        int64_t M, int64_t N, int64_t K, int elem_dtype, int accum_mant_bits, const lof_format_t* acc_fmt,
                        const lof_projspec_t* ps, int64_t k_chunk, float alpha, uint32_t flags, void* stream);

/* fp32 containers holding MX-quantised values (lof_mx_round_f32's output, 32-element blocks along the last dimension) -> the
 * operand form of lof_mx_gemm: OCP e4m3fn element codes (n bytes) + one E8M0 scale byte per block (n / 32 bytes).  The split into
 * (power-of-two scale, element) is re-derived from each block's amax; every e3m2 / e2m3 / e2m1 (MXFP6 / MXFP4) block and every
 * e4m3 block converts exactly except where a value is not on the hardware's e4m3fn grid (the reference's saturating virtual_round
 * leaves a block's overflowing maximum unrounded, and its 8-bit e4m3 tops out at 480 where the hardware format ends at 448): those
 * are rounded to nearest-even / saturated and LOF_MX_PACK_STATUS_SNAPPED is OR-ed into *dev_status (device int32, may be NULL).
 * n % 32 == 0, `in` 16-byte and `elems` 4-byte aligned. */
#define LOF_MX_PACK_STATUS_SNAPPED 1
LOF_API int lof_mx_pack_tc_f32(const float* in, int64_t n, uint8_t* elems, uint8_t* scales, int32_t* dev_status, void* stream);

/* Flag hand-shake of the GEMM's peer-to-peer operand exchange (lofloat_b200/sharding.py PeerPanels): every rank maps its peers' operand
 * shards and a 16-byte flag block per rank (int32[4], zero-initialised) through CUDA IPC and pulls the panels it needs over NVLink with plain
 * copies; these one-thread kernels order the pulls against the owners' writes.  flags[0] = published epoch, flags[1] = completed remote
 * pulls (all epochs), flags[3] = sticky protocol-error flag (a bounded spin expired).
 *   lof_p2p_post(my_flags, epoch, expected_done)   owner's stream, after the shards were written: waits for flags[1] >= expected_done (all pulls
 *                                                 of the previous epoch), then publishes `epoch`
 *   lof_p2p_wait(peer_flags, my_flags, epoch)      consumer's stream, before copying from that peer
 *   lof_p2p_done(peer_flags)                       consumer's stream, after its copies from that peer */
LOF_API int lof_p2p_post(int32_t* my_flags, int epoch, int expected_done, void* stream);
LOF_API int lof_p2p_wait(const int32_t* peer_flags, int32_t* my_flags, int epoch, void* stream);
LOF_API int lof_p2p_done(int32_t* peer_flags, void* stream);
/* cudaDeviceEnablePeerAccess(peer_device) from the current device (kernels of this device may then dereference IPC-mapped memory of that peer;
 * already-enabled is not an error), and a plain asynchronous copy between two UVA pointers (peer or local) on `stream`. */
LOF_API int lof_p2p_enable_peer(int peer_device);
LOF_API int lof_p2p_copy(void* dst, const void* src, int64_t bytes, void* stream);
/* CUDA IPC by hand, so that the mapping is opened from the CONSUMER's device (cudaIpcMemLazyEnablePeerAccess): export = the 64-byte
 * cudaIpcMemHandle_t of the allocation that contains `ptr` (cudaMalloc memory; not stream-ordered pools) + the offset of `ptr` inside it;
 * import opens the handle once per process (cached by its bytes) and returns base + offset, valid for kernels and copies of the current device. */
LOF_API int lof_p2p_export(const void* ptr, uint8_t handle[64], uint64_t* offset);
LOF_API int lof_p2p_import(const uint8_t handle[64], uint64_t offset, void** ptr);

/* ---------------------------------------------------------------------------------------
 * Host-buffer entry points: what a binding calls when the tensor lives in host memory.
 * The library stages through device memory in chunks and overlaps H2D, kernel and D2H on
 * three streams; it blocks until `out_host` is complete.  Pinned host buffers give full
 * PCIe bandwidth, pageable ones work but are slower.  `device` < 0 = current device.
 * ------------------------------------------------------------------------------------- */
LOF_API int lof_round_params_host(const void* in_host, void* out_host, int64_t n, int dtype,
                          const lof_format_t* fmt, const lof_projspec_t* ps, double scale,
                          const lof_rng_t* rng, uint64_t first_index, int device);
LOF_API int lof_round_mantissa_host(const void* in_host, void* out_host, int64_t n, int dtype, int mantissa_bits,
                            const lof_projspec_t* ps, double scale,
                            const lof_rng_t* rng, uint64_t first_index, int device);
LOF_API int lof_mx_round_f32_host(const float* in_host, float* out_host, int64_t n, int block_size,
                          const lof_format_t* scale_fmt, const lof_format_t* elem_fmt,
                          const lof_projspec_t* ps_scale, const lof_projspec_t* ps_elem,
                          const lof_rng_t* rng, uint64_t first_index, int device);

/* ---------------------------------------------------------------------------------------
 * Introspection / diagnostics.
 * ------------------------------------------------------------------------------------- */
LOF_API int         lof_abi_version(void);
LOF_API const char* lof_status_string(int status);
LOF_API const char* lof_last_error_string(void);          /* thread-local detail of the last failure  */
LOF_API uint64_t    lof_kernel_launch_count(void);        /* kernels launched by this library so far  */
LOF_API float       lof_mx_priv_max_normal(const lof_format_t* elem_fmt);  /* (2 - 2^-m) * 2^max_exp  */

/* Device self-test of the MX fast kernel's division shortcut: counts the non-negative finite fp32 values a for
 * which its 3-FMA quotient a / priv_max_normal(elem_fmt) differs from IEEE division (must be 0).  Synchronous. */
LOF_API int lof_selftest_mx_division(const lof_format_t* elem_fmt, uint64_t* mismatches, void* stream);

/* Tuning knobs used by bench.py / the profiling scripts to compare kernel variants on the
 * same inputs.  key: "round.unroll", "round.ctas_per_sm", "round.fast" (0 = always the
 * exact integer kernel, 1 = branch-free fast path, 2 = + magic-add path [default]), "mx.variant" (0 = fast kernel
 * + generic remainder [default], 1 = TMA-staged, 2 = first-generation shuffle kernel), "mx.ctas_per_sm",
 * "host.chunk_bytes".  Returns LOF_ERR_INVALID_ARG for
 * an unknown key. */
LOF_API int lof_set_option(const char* key, int64_t value);
LOF_API int lof_get_option(const char* key, int64_t* value);

#ifdef __cplusplus
}
#endif
#endif /* LOFLOAT_B200_H */
