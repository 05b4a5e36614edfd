// lof_gemm_mx.cu -- block-scaled (MX) chunked tensor-core GEMM for sm_100a:
//     tcgen05.mma.kind::mxf8f6f4.block_scale, operands = MX storage as lof_mx_encode_f32 writes it
//     (8-bit element codes + one E8M0 scale byte per 32 elements along K; container: /root/reference/src/Matrix.h:79-176 MX_Matrix).
//
// What it computes -- the composition LoF_Linear(scaling="mx") runs (/root/reference/LoFloat/layers.py:366-377: mx_round of both operands,
// then the LoFloat-accumulation GEMM), with the MX operands consumed in their storage form instead of as fp32 / bf16 containers:
//     acc = 0
//     for every k-chunk c (k_chunk = a multiple of the MX block, 32):
//         P_c = sum_{blocks b in c} 2^(sa_ib - 127) 2^(sb_jb - 127) sum_{k in b} a_ik b_jk      tensor core, fp32 in TMEM; the scales are
//                                                                                              applied by the MMA itself
//         acc = round(acc + P_c)    or    acc = round(acc + round(P_c))   (LOF_GEMM_PARTIAL_ROUND)
//     C = alpha * acc
// i.e. the accumulator semantics of lof_gemm_tc.cu (CPU statement: oracle/lof_oracle.c lofo_gemm_chunked_f32 on the decoded operands).
// Element codes: OCP e4m3 (also carries every e3m2 / e2m3 / e2m1 value exactly, so MXFP6 / MXFP4 tensors re-coded as e4m3 bytes run
// here at the FP8 rate, which is also the tensor core's FP6 rate) or e5m2.
//
// Kernel: the persistent per-chunk fold pipeline of lof_gemm_tc3.cuh (one CTA per SM, warp 0 TMA producer, warp 1 one-thread MMA
// issuer, warps 4..11 own the accumulators in registers) with three changes:
//   * the CTA tile is 128 x 224: TMEM holds two accumulator stages of 224 fp32 columns (at columns 0 and 256) and, in the 2 x 32
//     columns they leave free, the scale factors of the four shared-memory stages (SFA 4 columns + SFB 8 columns each);
//   * every shared-memory stage (128 bytes of K = four MX blocks) also carries the stage's scale bytes, pre-arranged by a tiny
//     pre-pass in the layout tcgen05.cp.32x128b.warpx4 copies into TMEM -- 512-byte chunks of [32 rows][4 row groups][4 blocks]
//     (row r of a 128-row group -> lane r % 32, column r / 32, byte = block; the layout cuBLAS / CUTLASS call "128x4 tiled") -- and
//     fetched by plain bulk copies on the stage's mbarrier;
//   * the issuing thread copies the stage's scales into TMEM (three tcgen05.cp) and then issues the four K = 32 MMAs with
//     a_sf_id = b_sf_id = block-in-stage; tcgen05.cp and tcgen05.mma execute in issue order, and the stage's commit covers both.
#include "lof_gemm_mx.cuh"

namespace lof {

namespace {

// Scale bytes [rows, kb] (one E8M0 byte per 32 elements of K, row-major -- what lof_mx_encode_f32 writes for a [rows, K] tensor)
// -> [tiles][kb4][CH][32 lanes][4 row groups][4 blocks] with tiles of TR rows (CH = ceil(TR / 128) chunks of 512 bytes).
// Rows / blocks outside the matrix get 0x7F (scale 1; their elements are TMA zero fill).  One thread per 32-bit word.
__global__ void __launch_bounds__(256)
mx_scale_arrange_kernel(const uint8_t* __restrict__ s, uint8_t* __restrict__ out, int rows, int kb, int kb4, int TR, int CH, int64_t n_words)
{
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    const int j = (int)(w & 3);                  // row group inside the chunk
    const int l = (int)((w >> 2) & 31);          // lane
    int64_t rest = w >> 7;
    const int ch = (int)(rest % CH); rest /= CH;
    const int g = (int)(rest % kb4);
    const int tile = (int)(rest / kb4);
    const int rt = ch * 128 + j * 32 + l;        // row inside the tile
    const int64_t row = (int64_t)tile * TR + rt;
    uint32_t word = 0x7F7F7F7Fu;
    if (rt < TR && row < rows) {
        const uint8_t* p = s + row * kb + (int64_t)g * 4;
        uint32_t b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) b[i] = (g * 4 + i < kb) ? (uint32_t)p[i] : 0x7Fu;
        word = b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24);
    }
    reinterpret_cast<uint32_t*>(out)[w] = word;
}

// fp32 containers holding MX-quantised values (the output of lof_mx_round_f32 with 32-element blocks) -> the tensor core's MX
// storage: OCP e4m3fn element codes + one E8M0 scale byte per block.  Any (power-of-two scale, element) split of a block gives the
// same products, so the block exponent is re-derived from the block's amax: the largest element lands in [256, 448] (one binade
// up when its mantissa exceeds 1.75, e.g. the 1.875 maxima of e2m3 and of the reference's 480-max e4m3).  Every e3m2 / e2m3 / e2m1
// block converts exactly; values that are not on the e4m3fn grid after scaling (the reference leaves a block's overflowing maximum
// unrounded; sub-2^-5 odd steps below a 480 maximum) are rounded to nearest-even / saturated at 448 and
// LOF_MX_PACK_STATUS_SNAPPED is raised.  A block holding an Inf / NaN gets the NaN scale 0xFF.  8 lanes x float4 per block.
__global__ void __launch_bounds__(256)
mx_pack_tc_kernel(const float4* __restrict__ y, uint32_t* __restrict__ elems, uint8_t* __restrict__ scales, int64_t n_vec, int* status)
{
    bool snapped = false;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n_round = (n_vec + 31) & ~int64_t(31);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
        const bool live = i < n_vec;
        const float4 v = live ? __ldcs(y + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        uint32_t m = max(max(__float_as_uint(v.x) & 0x7FFFFFFFu, __float_as_uint(v.y) & 0x7FFFFFFFu),
                         max(__float_as_uint(v.z) & 0x7FFFFFFFu, __float_as_uint(v.w) & 0x7FFFFFFFu));
        m = max(m, __shfl_xor_sync(0xFFFFFFFFu, m, 1));
        m = max(m, __shfl_xor_sync(0xFFFFFFFFu, m, 2));
        m = max(m, __shfl_xor_sync(0xFFFFFFFFu, m, 4));
        const int ex = (int)(m >> 23);
        int es = ex - 8 + ((m & 0x007FFFFFu) > 0x00600000u ? 1 : 0);      // biased exponent of the block scale
        es = min(max(es, 1), 253);
        const bool special = ex == 255;
        const float inv = __uint_as_float((uint32_t)(254 - es) << 23);   // 2^(127 - es)
        const float q[4] = {v.x * inv, v.y * inv, v.z * inv, v.w * inv};
        uint32_t word = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const __nv_fp8_storage_t b = __nv_cvt_float_to_fp8(q[e], __NV_SATFINITE, __NV_E4M3);
            const float back = __half2float(__half(__nv_cvt_fp8_to_halfraw(b, __NV_E4M3)));
            snapped |= !special && !(back == q[e]);
            word |= (uint32_t)b << (8 * e);
        }
        if (live) {
            elems[i] = word;
            if ((i & 7) == 0) scales[i >> 3] = special ? (uint8_t)0xFF : (uint8_t)es;
        }
    }
    if (__any_sync(0xFFFFFFFFu, snapped) && (threadIdx.x & 31) == 0 && status) atomicOr(status, 1);
}

using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeFn x_get_encode()
{
    static EncodeFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeFn>(p);
    }
    return fn;
}
bool x_operand_map(CUtensorMap* map, const void* base, int64_t rows, int64_t k_bytes, int box_rows)
{
    EncodeFn enc = x_get_encode();
    if (!enc) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)k_bytes, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)k_bytes};
    const cuuint32_t box[2] = {(cuuint32_t)X_ROW_BYTES, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

// elem_dtype: LOF_E4M3 or LOF_E5M2 (the element CODES; scales are E8M0 bytes).  *err is set for unsupported requests (no launch).
cudaError_t launch_gemm_mx(const uint8_t* A, const uint8_t* SA, const uint8_t* Bt, const uint8_t* SB, float* C, int64_t M, int64_t N, int64_t K,
                           int elem_dtype, int mode, int accum_mant_bits, const lof_format_t* acc_fmt, const RoundPlan* acc_plan,
                           const FastPlan* acc_fast, int64_t k_chunk, float alpha, uint32_t flags, cudaStream_t stream, int* launches,
                           std::string* err)
{
    if (elem_dtype != LOF_E4M3 && elem_dtype != LOF_E5M2) { *err = "MX GEMM: elem_dtype must be LOF_E4M3 or LOF_E5M2"; return cudaSuccess; }
    if (K % 32 != 0) { *err = "MX GEMM: K must be a multiple of the MX block (32)"; return cudaSuccess; }
    if (k_chunk <= 0 || k_chunk > K) k_chunk = K;
    if (k_chunk % 32 != 0) { *err = "MX GEMM: k_chunk must be a multiple of the MX block (32)"; return cudaSuccess; }
    if (M > (1 << 24) || N > (1 << 24) || K > (1ll << 30)) { *err = "MX GEMM: dimension too large"; return cudaSuccess; }
    if (mode_uses_rng(mode)) { *err = "MX GEMM: stochastic accumulator rounding is not implemented"; return cudaSuccess; }
    if (((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(Bt)) & 15u) != 0) { *err = "MX GEMM: element arrays must be 16-byte aligned (TMA)"; return cudaSuccess; }

    MxGemmArgs a;
    memset(&a, 0, sizeof a);
    a.M = (int)M; a.N = (int)N; a.K = (int)K;
    a.k_steps = (int)(K / 32);
    a.steps_per_chunk = (int)(k_chunk / 32);
    a.kb4 = (int)((K + 127) / 128);
    a.tiles_m = (int)((M + X_BM - 1) / X_BM);
    a.tiles_n = (int)((N + X_BN - 1) / X_BN);
    a.n_tiles = a.tiles_m * a.tiles_n;
    a.alpha = alpha;
    a.r = acc_fmt ? 0 : 23 - accum_mant_bits;
    if (acc_plan) a.plan = *acc_plan;
    if (acc_fast) a.fast = *acc_fast;
    // instruction descriptor (block-scaled form): element formats at bits 7-9 / 10-12 (0 = e4m3, 1 = e5m2), K-major A and B, N >> 3 at
    // 17-22, scale format E8M0 (bit 23), M >> 4 at 24-28; b_sf_id (bits 4-5) and a_sf_id (29-30) are filled in per MMA
    const uint32_t ef = elem_dtype == LOF_E4M3 ? 0u : 1u;
    a.idesc = (ef << 7) | (ef << 10) | ((uint32_t)(X_BN >> 3) << 17) | (1u << 23) | ((uint32_t)(X_BM >> 4) << 24);

    const size_t bytes_sfa = (size_t)a.tiles_m * a.kb4 * X_SFA_BYTES, bytes_sfb = (size_t)a.tiles_n * a.kb4 * X_SFB_BYTES;
    const size_t off_b = (bytes_sfa + 1023) & ~size_t(1023), off_s = off_b + ((bytes_sfb + 1023) & ~size_t(1023));
    uint8_t* ws = nullptr;
    cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&ws), off_s + 256, stream);
    if (e != cudaSuccess) return e;
    auto fail = [&](cudaError_t ce) { cudaFreeAsync(ws, stream); return ce; };
    a.sfa = ws;
    a.sfb = ws + off_b;
    {
        const int kb = (int)(K / 32);
        const int64_t wa = (int64_t)bytes_sfa / 4, wb = (int64_t)bytes_sfb / 4;
        mx_scale_arrange_kernel<<<(unsigned)((wa + 255) / 256), 256, 0, stream>>>(SA, ws, (int)M, kb, a.kb4, X_BM, 1, wa);
        mx_scale_arrange_kernel<<<(unsigned)((wb + 255) / 256), 256, 0, stream>>>(SB, ws + off_b, (int)N, kb, a.kb4, X_BN, 2, wb);
        *launches += 2;
        e = cudaGetLastError();
        if (e != cudaSuccess) return fail(e);
    }
    CUtensorMap ma, mb;
    if (!x_operand_map(&ma, A, M, K, X_BM) || !x_operand_map(&mb, Bt, N, K, X_BN)) {
        cudaFreeAsync(ws, stream);
        *err = "MX GEMM: cuTensorMapEncodeTiled failed";
        return cudaSuccess;
    }
    int sms = 0, dev_id = 0;
    cudaGetDevice(&dev_id);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev_id);
    const int grid = std::max(1, std::min(a.n_tiles, sms));
    const bool partial = (flags & LOF_GEMM_PARTIAL_ROUND) != 0;
    const bool one_chunk = a.steps_per_chunk >= a.k_steps;
    // bfloat162 fold: bf16-like accumulator (7 mantissa bits, mantissa-only), RNE, partial rounding (with ONE chunk the two forms coincide:
    // acc = round(0 + P)) -- cvt.rn.bf16x2.f32 + HADD2.BF16, bit-identical to the generic routine for every finite value
    const bool bf162 = (flags & LOF_GEMM_NO_PACKED) == 0 && !acc_fmt && accum_mant_bits == 7 && mode == LOF_RoundToNearestEven && (partial || one_chunk);
    const bool fw8 = (flags & LOF_GEMM_INTERNAL_FOLD_SPLIT2) == 0;     // 8 fold warps x 112 columns by default; option gemm.fold_split = 2: 16 x 56
    const bool prof = (flags & LOF_GEMM_INTERNAL_PROF) != 0;
    if (bf162 && prof) e = fw8 ? x_launch_one<LOF_RoundToNearestEven, kXBf162, true, 8, true>(ma, mb, C, a, grid, stream)
                               : x_launch_one<LOF_RoundToNearestEven, kXBf162, true, 16, true>(ma, mb, C, a, grid, stream);
    else if (bf162) e = fw8 ? x_launch_one<LOF_RoundToNearestEven, kXBf162, true, 8>(ma, mb, C, a, grid, stream)
                            : x_launch_one<LOF_RoundToNearestEven, kXBf162, true, 16>(ma, mb, C, a, grid, stream);
    else e = launch_gemm_mx_generic(mode, acc_fmt != nullptr, partial, ma, mb, C, &a, grid, stream);
    ++*launches;
    if (e != cudaSuccess) return fail(e);
    return cudaFreeAsync(ws, stream);
}

cudaError_t read_mx_prof(unsigned long long* out16)
{
    return cudaMemcpyFromSymbol(out16, g_mx_prof, sizeof(unsigned long long) * 16);
}

cudaError_t launch_mx_pack_tc(const float* y, int64_t n, uint8_t* elems, uint8_t* scales, int* dev_status, int num_sms, cudaStream_t stream)
{
    const int64_t n_vec = n / 4;
    const int64_t want = (n_vec + 255) / 256;
    const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)num_sms * 8));
    mx_pack_tc_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const float4*>(y), reinterpret_cast<uint32_t*>(elems), scales, n_vec, dev_status);
    return cudaGetLastError();
}

}  // namespace lof
