"""ctypes bindings for the CPU oracle (oracle/liblof_oracle.so) and, where it was built, the
compiled reference shim (oracle/_ref/libref_shim.so).

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
ORACLE_SO = ORACLE_DIR / "liblof_oracle.so"
REF_SO = ORACLE_DIR / "_ref" / "libref_shim.so"

F32, F16, BF16, F64 = 0, 1, 2, 3
STREAM_ELEM, STREAM_SCALE, STREAM_GEMM = 0, 0x5CA1E000, 0x6E3A0000

# rounding modes (fp_tools.hpp:13-28)
RNE, RTZ, RAZ, SR_A, RNO, RD, RU, RTA, SR_B, SR_C, TRUE_SR, PROB, RTO, RTTZ, SR_D = range(15)
DETERMINISTIC_MODES = (RNE, RTZ, RAZ, RNO, RD, RU, RTA, RTO, RTTZ)
STOCHASTIC_MODES = (SR_A, SR_B, SR_C, SR_D, TRUE_SR, PROB)
MODE_NAMES = {
    RNE: "RNE", RTZ: "RTZ", RAZ: "RAZ", SR_A: "SR_A", RNO: "RNO", RD: "RD", RU: "RU", RTA: "RTA",
    SR_B: "SR_B", SR_C: "SR_C", TRUE_SR: "TRUE_SR", PROB: "PROB", RTO: "RTO", RTTZ: "RTTZ", SR_D: "SR_D",
}
EXTENDED, SATURATING = 0, 1
SIGNED, UNSIGNED = 0, 1


class Format(C.Structure):
    _fields_ = [("bitwidth", C.c_int32), ("mantissa_bits", C.c_int32), ("bias", C.c_int32),
                ("ov_behavior", C.c_int32), ("na_behavior", C.c_int32), ("is_signed", C.c_int32),
                ("unsigned_behavior", C.c_int32)]

    def key(self):
        return (self.bitwidth, self.mantissa_bits, self.bias, self.ov_behavior, self.is_signed)


class ProjSpec(C.Structure):
    _fields_ = [("rounding_mode", C.c_int32), ("saturation_mode", C.c_int32), ("stoch_length", C.c_int32)]


class Rng(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("offset", C.c_uint64), ("fixed", C.c_int32), ("fixed_word", C.c_uint32)]


def fmt(bitwidth, mantissa_bits, bias, ov=SATURATING, signed=SIGNED, na=0) -> Format:
    return Format(bitwidth, mantissa_bits, bias, ov, na, signed, 0)


# the named formats of BASELINE.json / SURVEY.md section 8(d)
FORMATS = {
    "e4m3_b7": fmt(8, 3, 7),
    "p3109_8p4": fmt(8, 3, 8),            # create_p3109_params(8, 4): bias 2^(k-p-1)
    "e5m2": fmt(8, 2, 15),
    "e3m2": fmt(6, 2, 3),
    "e2m3": fmt(6, 3, 1),
    "e2m1": fmt(4, 1, 1),
    "bf16_like": fmt(16, 7, 127, ov=EXTENDED),
    "half": fmt(16, 10, 15, ov=EXTENDED),
    "single": fmt(32, 23, 127, ov=EXTENDED),
    "e8m0": fmt(8, 0, 127, ov=EXTENDED, signed=UNSIGNED),
    "e4m3_b7_ext": fmt(8, 3, 7, ov=EXTENDED),
}



# ------------------------------------------------------------------------------ storage codec formats
class CodecFormat(C.Structure):
    """lofo_codec_format / lof_codec_format_t: FloatingPointParams with the Inf / NaN functors as data."""
    _fields_ = [("fmt", Format), ("inf_mask", C.c_uint32), ("inf_value", C.c_uint32), ("nan_mask", C.c_uint32),
                ("nan_value", C.c_uint32), ("nan_mant_mask", C.c_uint32), ("min_pos_inf", C.c_uint32),
                ("qnan", C.c_uint32), ("snan", C.c_uint32)]


NEVER = (0, 1)      # (mask, value): (rep & 0) == 1 is never true
ALL = 0xFFFFFFFF


def codec_p3109(k, p, signed=True, saturating=True, bias=None, unsigned_behavior=0) -> CodecFormat:
    """param_float_p_3109<k, p, signedness, has_inf> (lo_float.h:1679-1768) == LoFloat/formats.py create_p3109_params."""
    bias = (1 << (k - p - 1)) if bias is None else bias
    f = Format(k, p - 1, bias, SATURATING if saturating else EXTENDED, 0, SIGNED if signed else UNSIGNED, unsigned_behavior)
    nan = (1 << (k - 1)) if signed else (1 << k) - 1
    if saturating:
        inf_m, inf_v = NEVER
    else:   # (bits | 1 << (k-1)) == (signed ? 2^k - 1 : 2^k - 2)
        inf_m, inf_v = (1 << (k - 1)) - 1, ((1 << k) - 1 if signed else (1 << k) - 2) & ((1 << (k - 1)) - 1)
    min_pos = ((1 << (k - 1)) - 1) if signed else (1 << k) - 2
    return CodecFormat(f, inf_m, inf_v, ALL, nan, 0, min_pos, nan, nan)


def codec_ocp(k, m, bias, signed=True) -> CodecFormat:
    """The OCP-style presets (e4m3fn, e5m2, e3m2, e2m3, e2m1: lo_float.h:1448-1660): no infinities, NaN == 0xFF
    (OCP_F8E4M3_{Inf,NaN}Checker -- the reference uses these two functors for ALL of them)."""
    f = Format(k, m, bias, SATURATING, 0, SIGNED if signed else UNSIGNED, 0)
    return CodecFormat(f, NEVER[0], NEVER[1], ALL, 0xFF, 0, 0, 0xFF, 0xFF)


def codec_half() -> CodecFormat:
    """halfPrecisionParams (fp_tools.hpp:334-370): IEEE binary16."""
    f = Format(16, 10, 15, EXTENDED, 1, SIGNED, 0)
    return CodecFormat(f, 0x7FFF, 0x7C00, 0x7C00, 0x7C00, 0x03FF, 0x7C00, 0x7E00, 0x7D00)


def codec_e8m0() -> CodecFormat:
    """param_ocp_e8m0 (lo_float.h:1777-1795): unsigned, NaN == 0xFF, saturating."""
    f = Format(8, 0, 127, SATURATING, 0, UNSIGNED, 0)
    return CodecFormat(f, NEVER[0], NEVER[1], ALL, 0xFF, 0, 0, 0xFF, 0xFF)


# the formats test/test_pack.cpp:230-244 runs plus the MX element / scale formats
CODEC_FORMATS = {
    "p3109_3p2": codec_p3109(3, 2), "e2m1": codec_ocp(4, 1, 1), "p3109_5p2": codec_p3109(5, 2), "e3m2": codec_ocp(6, 2, 3),
    "e2m3": codec_ocp(6, 3, 1), "p3109_7p3": codec_p3109(7, 3), "e5m2": codec_ocp(8, 2, 15), "e4m3fn": codec_ocp(8, 3, 7),
    "p3109_8p4": codec_p3109(8, 4), "p3109_8p4_ext": codec_p3109(8, 4, saturating=False),
    "p3109_8p3_unsigned": codec_p3109(8, 3, signed=False), "p3109_8p3_unsigned_ext_nan": codec_p3109(8, 3, signed=False, saturating=False, unsigned_behavior=1),
    "p3109_9p3": codec_p3109(9, 3), "p3109_11p5": codec_p3109(11, 5), "p3109_12p8": codec_p3109(12, 8), "half": codec_half(),
    "e8m0": codec_e8m0(),
}

_NP_OF = {F32: np.float32, F16: np.float16, F64: np.float64, BF16: np.uint16}


def build_oracle(force: bool = False) -> Path:
    srcs = [ORACLE_DIR / n for n in ("lof_oracle.c", "lof_oracle.h", "lof_oracle_core.inc", "lof_oracle_type.inc")]
    stale = (not ORACLE_SO.exists()) or any(s.stat().st_mtime > ORACLE_SO.stat().st_mtime for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", str(ORACLE_DIR), "liblof_oracle.so"], check=True,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return ORACLE_SO


def build_ref(reference_dir: str = "/root/reference") -> Path | None:
    """Compile the reference shim where the reference tree exists (this container)."""
    if not Path(reference_dir, "src", "lo_float.h").exists():
        return REF_SO if REF_SO.exists() else None
    subprocess.run(["make", "-C", str(ORACLE_DIR), "ref", f"REF={reference_dir}"], check=True,
                   stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return REF_SO


_oracle = None
_ref = None


def oracle():
    global _oracle
    if _oracle is None:
        build_oracle()
        lib = C.CDLL(str(ORACLE_SO))
        vp, i64, i32, u64, u32, f64 = C.c_void_p, C.c_int64, C.c_int, C.c_uint64, C.c_uint32, C.c_double
        PF, PP, PR = C.POINTER(Format), C.POINTER(ProjSpec), C.POINTER(Rng)
        lib.lofo_philox4x32_10.argtypes = [C.POINTER(u32), C.POINTER(u32), C.POINTER(u32)]
        lib.lofo_rand_word.argtypes = [PR, u64, u32]
        lib.lofo_rand_word.restype = u32
        lib.lofo_rand_event.argtypes = [PR, u64, u32, i32]
        lib.lofo_rand_event.restype = u32
        lib.lofo_rand_bits_needed.argtypes = [i32, i32]
        lib.lofo_rand_bits_needed.restype = i32
        lib.lofo_round_params_bits_f32.argtypes = [u32, PF, PP, u32]
        lib.lofo_round_params_bits_f32.restype = u32
        lib.lofo_round_mantissa_bits_f32.argtypes = [u32, i32, PP, u32]
        lib.lofo_round_mantissa_bits_f32.restype = u32
        lib.lofo_round_params.argtypes = [vp, vp, i64, i32, PF, PP, f64, PR, u64]
        lib.lofo_round_mantissa.argtypes = [vp, vp, i64, i32, i32, PP, f64, PR, u64]
        lib.lofo_mx_round_f32.argtypes = [vp, vp, i64, i32, PF, PF, PP, PP, PR, u64]
        lib.lofo_mx_priv_max_normal.argtypes = [PF]
        lib.lofo_mx_priv_max_normal.restype = C.c_float
        lib.lofo_round_to_odd_fma.argtypes = [C.c_float] * 3
        lib.lofo_round_to_odd_fma.restype = C.c_float
        lib.lofo_gemm_f32.argtypes = [vp, vp, vp, i64, i64, i64, i32, PP, i64, C.c_float, C.c_float, PR]
        lib.lofo_gemm_chunked_f32.argtypes = [vp, vp, vp, i64, i64, i64, i32, PF, PP, i64, C.c_float, i32]
        lib.lofo_gemm_chunked_rng_f32.argtypes = [vp, vp, vp, i64, i64, i64, i32, PF, PP, i64, C.c_float, i32, PR]
        lib.lofo_num_threads.restype = i32
        PC, u8p = C.POINTER(CodecFormat), C.c_void_p
        lib.lofo_float_to_rep.argtypes = [C.c_float, PC, PP, u32]
        lib.lofo_float_to_rep.restype = u32
        lib.lofo_rep_to_float.argtypes = [u32, PC]
        lib.lofo_rep_to_float.restype = C.c_float
        lib.lofo_float_to_rep_array.argtypes = [vp, vp, i64, PC, PP, PR, u64]
        lib.lofo_rep_to_float_array.argtypes = [vp, vp, i64, PC]
        lib.lofo_pack_group_count.argtypes = [i32]
        lib.lofo_pack_group_bytes.argtypes = [i32]
        lib.lofo_packed_size.argtypes = [i32, u64]
        lib.lofo_packed_size.restype = u64
        lib.lofo_pack_reps.argtypes = [u8p, vp, u64, i32, i32]
        lib.lofo_pack_reps.restype = u64
        lib.lofo_unpack_reps.argtypes = [vp, u8p, u64, i32, i32]
        lib.lofo_unpack_reps.restype = u64
        lib.lofo_encode_f32.argtypes = [u8p, vp, u64, PC, PP, i32, PR, u64]
        lib.lofo_encode_f32.restype = u64
        lib.lofo_decode_f32.argtypes = [vp, u8p, u64, PC, i32]
        lib.lofo_decode_f32.restype = u64
        lib.lofo_mx_encode_f32.argtypes = [vp, i64, i32, PC, PC, PP, PP, i32, PR, u64, vp, vp]
        lib.lofo_mx_decode_f32.argtypes = [vp, vp, vp, i64, i32, PC, PC, i32]
        _oracle = lib
    return _oracle


def ref():
    """The compiled reference shim, or None when it is not available (GPU box without a prebuilt
    copy, or a checkout that never had /root/reference)."""
    global _ref
    if _ref is None:
        if not REF_SO.exists():
            return None
        lib = C.CDLL(str(REF_SO))
        vp, i64, i32, u32 = C.c_void_p, C.c_int64, C.c_int, C.c_uint
        lib.ref_round_params_f32.argtypes = [vp, vp, i64] + [i32] * 8 + [u32, i32]
        lib.ref_round_params_f64.argtypes = [vp, vp, i64] + [i32] * 8 + [u32, i32]
        lib.ref_round_mantissa_f32.argtypes = [vp, vp, i64, i32, i32, i32, u32, i32]
        lib.ref_round_mantissa_f64.argtypes = [vp, vp, i64, i32, i32, i32, u32, i32]
        lib.ref_round_params_array_f32.argtypes = [vp, vp, i32] + [i32] * 8
        lib.ref_mx_round_f32.argtypes = [vp, i32, i32] + [i32] * 13 + [u32]
        lib.ref_draw_uint.argtypes = [u32, u32]
        lib.ref_draw_uint.restype = u32
        lib.ref_draw_uint_then_coin.argtypes = [u32, u32]
        lib.ref_draw_uint_then_coin.restype = u32
        lib.ref_draw_01.argtypes = [u32]
        lib.ref_draw_01.restype = i32
        lib.ref_draw_real.argtypes = [u32]
        lib.ref_draw_real.restype = C.c_double
        lib.ref_shim_threads.restype = i32
        lib.ref_has_blis.restype = i32
        PC = C.POINTER(CodecFormat)
        lib.ref_float_to_rep.argtypes = [vp, vp, i64, PC, i32, i32, i32, u32]
        lib.ref_rep_to_float.argtypes = [vp, vp, i64, PC]
        lib.ref_packed_size.argtypes = [i32, C.c_uint64]
        lib.ref_packed_size.restype = C.c_uint64
        lib.ref_encode_f32.argtypes = [vp, vp, C.c_uint64, PC, i32, i32, i32, i32]
        lib.ref_encode_f32.restype = i64
        lib.ref_decode_f32.argtypes = [vp, vp, C.c_uint64, PC, i32]
        lib.ref_decode_f32.restype = i64
        lib.ref_typed_e4m3fn.argtypes = [vp, vp, vp, i64]
        lib.ref_typed_p3109_8_4.argtypes = [vp, vp, vp, i64]
        if lib.ref_has_blis():
            lib.ref_round_to_odd_fma.argtypes = [C.c_float] * 3
            lib.ref_round_to_odd_fma.restype = C.c_float
            lib.ref_gemm_f32.argtypes = [vp, vp, vp, i64, i64, i64, i32, i32, i32, C.c_float, C.c_float, u32, i32]
        _ref = lib
    return _ref


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _rng(seed=0, offset=0, fixed_word=None) -> Rng:
    if fixed_word is None:
        return Rng(seed, offset, 0, 0)
    return Rng(seed, offset, 1, int(fixed_word) & 0xFFFFFFFF)


# ------------------------------------------------------------------------------ oracle calls
def o_round_params(x: np.ndarray, f: Format, mode=RNE, stoch_len=0, scale=1.0, dtype=F32,
                   seed=0, offset=0, fixed_word=None, first_index=0) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=_NP_OF[dtype])
    out = np.empty_like(x)
    ps = ProjSpec(mode, 0, stoch_len)
    rng = _rng(seed, offset, fixed_word)
    oracle().lofo_round_params(_ptr(x), _ptr(out), x.size, dtype, C.byref(f), C.byref(ps), float(scale),
                               C.byref(rng), first_index)
    return out


def o_round_mantissa(x: np.ndarray, mant_bits: int, mode=RNE, stoch_len=0, scale=1.0, dtype=F32,
                     seed=0, offset=0, fixed_word=None, first_index=0) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=_NP_OF[dtype])
    out = np.empty_like(x)
    ps = ProjSpec(mode, 0, stoch_len)
    rng = _rng(seed, offset, fixed_word)
    oracle().lofo_round_mantissa(_ptr(x), _ptr(out), x.size, dtype, mant_bits, C.byref(ps), float(scale),
                                 C.byref(rng), first_index)
    return out


def o_mx_round(x: np.ndarray, block: int, scale_fmt: Format, elem_fmt: Format, mode=RNE, scale_mode=RNE,
               stoch_len=0, seed=0, offset=0, fixed_word=None, first_index=0) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(x)
    ps_s, ps_e = ProjSpec(scale_mode, 0, 0), ProjSpec(mode, 0, stoch_len)
    rng = _rng(seed, offset, fixed_word)
    oracle().lofo_mx_round_f32(_ptr(x), _ptr(out), x.size, block, C.byref(scale_fmt), C.byref(elem_fmt),
                               C.byref(ps_s), C.byref(ps_e), C.byref(rng), first_index)
    return out


def o_gemm(A: np.ndarray, B: np.ndarray, accum_mant_bits: int, mode=RNE, stoch_len=0, kc=256, alpha=1.0,
           beta=0.0, C0: np.ndarray | None = None, seed=0, offset=0, fixed_word=None) -> np.ndarray:
    A = np.ascontiguousarray(A, dtype=np.float32)
    B = np.ascontiguousarray(B, dtype=np.float32)
    M, K = A.shape
    K2, N = B.shape
    assert K == K2
    out = np.zeros((M, N), np.float32) if C0 is None else np.ascontiguousarray(C0, np.float32).copy()
    ps = ProjSpec(mode, 0, stoch_len)
    rng = _rng(seed, offset, fixed_word)
    oracle().lofo_gemm_f32(_ptr(A), _ptr(B), _ptr(out), M, N, K, accum_mant_bits, C.byref(ps), kc,
                           alpha, beta, C.byref(rng))
    return out


def o_gemm_chunked(A, B, accum_mant_bits, k_chunk, mode=RNE, acc_fmt: Format | None = None, alpha=1.0,
                   partial_round=False, stoch_len=0, seed=None, offset=0):
    A = np.ascontiguousarray(A, dtype=np.float32)
    B = np.ascontiguousarray(B, dtype=np.float32)
    M, K = A.shape
    _, N = B.shape
    out = np.zeros((M, N), np.float32)
    ps = ProjSpec(mode, 0, stoch_len)
    if seed is None:
        oracle().lofo_gemm_chunked_f32(_ptr(A), _ptr(B), _ptr(out), M, N, K, accum_mant_bits,
                                       C.byref(acc_fmt) if acc_fmt is not None else None, C.byref(ps), k_chunk, alpha,
                                       1 if partial_round else 0)
    else:
        rng = _rng(seed, offset)
        oracle().lofo_gemm_chunked_rng_f32(_ptr(A), _ptr(B), _ptr(out), M, N, K, accum_mant_bits,
                                           C.byref(acc_fmt) if acc_fmt is not None else None, C.byref(ps), k_chunk, alpha,
                                           1 if partial_round else 0, C.byref(rng))
    return out


def o_philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    oracle().lofo_philox4x32_10(c, k, o)
    return [int(v) for v in o]


def o_rand_event(seed, offset, index, nbits, stream=STREAM_ELEM) -> int:
    r = _rng(seed, offset)
    return int(oracle().lofo_rand_event(C.byref(r), index, stream, nbits))


def o_rand_word(seed, offset, index, stream=STREAM_ELEM) -> int:
    r = Rng(seed, offset, 0, 0)
    return int(oracle().lofo_rand_word(C.byref(r), index, stream))


# --------------------------------------------------------------------------- reference calls
def r_round_params(x, f: Format, mode=RNE, stoch_len=0, seed=12345, threads=None, dtype=F32):
    lib = ref()
    x = np.ascontiguousarray(x, dtype=_NP_OF[dtype])
    out = np.empty_like(x)
    fn = lib.ref_round_params_f32 if dtype == F32 else lib.ref_round_params_f64
    fn(_ptr(x), _ptr(out), x.size, f.bitwidth, f.mantissa_bits, f.bias, f.ov_behavior, f.na_behavior,
       f.is_signed, mode, stoch_len, seed, threads or os.cpu_count())
    return out


def r_round_mantissa(x, mant_bits, mode=RNE, stoch_len=0, seed=12345, threads=None, dtype=F32):
    lib = ref()
    x = np.ascontiguousarray(x, dtype=_NP_OF[dtype])
    out = np.empty_like(x)
    fn = lib.ref_round_mantissa_f32 if dtype == F32 else lib.ref_round_mantissa_f64
    fn(_ptr(x), _ptr(out), x.size, mant_bits, mode, stoch_len, seed, threads or os.cpu_count())
    return out


def r_mx_round(x, block, scale_fmt: Format, elem_fmt: Format, mode=RNE, scale_mode=RNE, stoch_len=0, seed=12345):
    lib = ref()
    buf = np.ascontiguousarray(x, dtype=np.float32).copy()
    p, e = scale_fmt, elem_fmt
    lib.ref_mx_round_f32(_ptr(buf), buf.size, block, p.bitwidth, p.mantissa_bits, p.bias, p.ov_behavior,
                         p.is_signed, scale_mode, e.bitwidth, e.mantissa_bits, e.bias, e.ov_behavior,
                         e.is_signed, mode, stoch_len, seed)
    return buf


def r_gemm(A, B, accum_mant_bits, mode=RNE, stoch_len=0, alpha=1.0, beta=0.0, C0=None, seed=12345, threads=None):
    lib = ref()
    A = np.ascontiguousarray(A, dtype=np.float32)
    B = np.ascontiguousarray(B, dtype=np.float32)
    M, K = A.shape
    _, N = B.shape
    out = np.zeros((M, N), np.float32) if C0 is None else np.ascontiguousarray(C0, np.float32).copy()
    lib.ref_gemm_f32(_ptr(A), _ptr(B), _ptr(out), M, N, K, accum_mant_bits, mode, stoch_len, alpha, beta, seed,
                     threads or os.cpu_count())
    return out


def ref_fixed_word(mode: int, stoch_len: int, seed: int) -> int:
    """The event word W that makes the oracle consume exactly the draw the reference's host
    helper makes from std::mt19937(seed) (lo_float.h:606,666,705,764,835,872)."""
    lib = ref()
    maxv = (1 << stoch_len) - 1 if stoch_len > 0 else 0
    if mode in (SR_A, SR_B, SR_C):
        return int(lib.ref_draw_uint(seed, maxv))
    if mode == SR_D:
        v = int(lib.ref_draw_uint_then_coin(seed, maxv))
        return (v & 0x7FFFFFFF) | ((v >> 31) << stoch_len)
    if mode == PROB:
        return int(lib.ref_draw_01(seed))
    if mode == TRUE_SR:
        return int(np.floor(lib.ref_draw_real(seed) * 4294967296.0))
    return 0


# ------------------------------------------------------------------------------- input makers
def special_values_f32() -> np.ndarray:
    """Bit patterns that exercise every regime boundary for the named formats."""
    vals = [0.0, -0.0, np.inf, -np.inf, np.nan, 1.0, -1.0, 1.0625, 1.1875, 447, 464, 480, 496, 512, -1e6, 1e6,
            0.001, 0.0009765625, 0.00146484375, 1e-30, -1e-30, 1e-38, 1e-39, 3e38, -3e38, 1.4, 1.5, 2.9, 3.0,
            3.1, 0.75, 0.1, 0.25, 0.3, 1.25, -1.25, 2.5, -2.5, 5, 6, 6.9, 7, 7.5, 8, -0.2, 57344, 61440, 65504,
            65520, 114688, 122880, 2.0 ** -126, 2.0 ** -127, 2.0 ** -130, 2.0 ** -133, 2.0 ** -134, 2.0 ** -149,
            3.0 * 2.0 ** -149, 1.5 * 2.0 ** -9, 2.5 * 2.0 ** -9, 2.0 ** -9, 2.0 ** -10, 2.0 ** -6, 0.0154, 0.0157]
    a = np.array(vals, dtype=np.float32)
    bits = [0x7F800001, 0x7FFFFFFF, 0xFFC00000, 0x7F7FFFFF, 0xFF7FFFFF, 0x00000001, 0x80000001, 0x007FFFFF,
            0x00800000, 0x00400000, 0x3F7FFFFF, 0x3F800001, 0x7F000000, 0x7EFFFFFF]
    b = np.array(bits, dtype=np.uint32).view(np.float32)
    return np.concatenate([a, -a, b])


def wide_random_f32(n: int, seed: int) -> np.ndarray:
    """randn * 2^U{-12..12} + raw random bit patterns."""
    rng = np.random.default_rng(seed)
    a = (rng.standard_normal(n // 2) * np.exp2(rng.integers(-12, 13, n // 2))).astype(np.float32)
    b = rng.integers(0, 2 ** 32, n - n // 2, dtype=np.uint64).astype(np.uint32).view(np.float32)
    return np.concatenate([a, b])


def bits_equal(a: np.ndarray, b: np.ndarray) -> bool:
    ua = a.view({2: np.uint16, 4: np.uint32, 8: np.uint64}[a.dtype.itemsize])
    ub = b.view(ua.dtype)
    return bool(np.array_equal(ua, ub))


def first_mismatch(x, a, b):
    ux = x.view(np.uint32) if x.dtype == np.float32 else x
    ua, ub = a.view(np.uint32), b.view(np.uint32)
    idx = np.nonzero(ua != ub)[0]
    if idx.size == 0:
        return None
    i = int(idx[0])
    return f"{idx.size} mismatches; first at {i}: in=0x{int(ux[i]):08x} got=0x{int(ua[i]):08x} want=0x{int(ub[i]):08x}"


# ------------------------------------------------------------------------------ storage codec (lo_pack.hpp)
def packed_size(n: int, count: int) -> int:
    return int(oracle().lofo_packed_size(n, count))


def o_float_to_rep(x: np.ndarray, f: CodecFormat, mode=RNE, sat=0, stoch_len=0, seed=0, offset=0, first_index=0, fixed_word=None) -> np.ndarray:
    """convert_float_to_rep element by element (no packing)."""
    x = np.ascontiguousarray(x, np.float32).ravel()
    ps, r = ProjSpec(mode, sat, stoch_len), _rng(seed, offset, fixed_word)
    out = np.empty(x.size, np.uint32)
    oracle().lofo_float_to_rep_array(_ptr(x), _ptr(out), x.size, C.byref(f), C.byref(ps), C.byref(r), first_index)
    return out


def o_rep_to_float(reps: np.ndarray, f: CodecFormat) -> np.ndarray:
    reps = np.ascontiguousarray(reps, np.uint32).ravel()
    out = np.empty(reps.size, np.float32)
    oracle().lofo_rep_to_float_array(_ptr(reps), _ptr(out), reps.size, C.byref(f))
    return out


def o_encode(x: np.ndarray, f: CodecFormat, mode=RNE, sat=0, stoch_len=0, endian=0, seed=0, offset=0, first_index=0) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32)
    out = np.full(packed_size(f.fmt.bitwidth, x.size), 0xAB, np.uint8)
    ps, r = ProjSpec(mode, sat, stoch_len), _rng(seed, offset)
    n = oracle().lofo_encode_f32(_ptr(out), _ptr(x), x.size, C.byref(f), C.byref(ps), endian, C.byref(r), first_index)
    assert n == out.size
    return out


def o_decode(buf: np.ndarray, count: int, f: CodecFormat, endian=0) -> np.ndarray:
    buf = np.ascontiguousarray(buf, np.uint8)
    out = np.empty(count, np.float32)
    oracle().lofo_decode_f32(_ptr(out), _ptr(buf), count, C.byref(f), endian)
    return out


def o_pack_reps(reps: np.ndarray, n: int, endian=0) -> np.ndarray:
    reps = np.ascontiguousarray(reps, np.uint32)
    out = np.full(packed_size(n, reps.size), 0xAB, np.uint8)
    oracle().lofo_pack_reps(_ptr(out), _ptr(reps), reps.size, n, endian)
    return out


def o_unpack_reps(buf: np.ndarray, count: int, n: int, endian=0) -> np.ndarray:
    buf = np.ascontiguousarray(buf, np.uint8)
    out = np.empty(count, np.uint32)
    oracle().lofo_unpack_reps(_ptr(out), _ptr(buf), count, n, endian)
    return out


def r_float_to_rep(x: np.ndarray, f: CodecFormat, mode=RNE, sat=0, stoch_len=0, seed=12345) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty(x.size, np.uint32)
    ref().ref_float_to_rep(_ptr(x), _ptr(out), x.size, C.byref(f), mode, sat, stoch_len, seed)
    return out


def r_rep_to_float(reps: np.ndarray, f: CodecFormat) -> np.ndarray:
    reps = np.ascontiguousarray(reps, np.uint32)
    out = np.empty(reps.size, np.float32)
    ref().ref_rep_to_float(_ptr(reps), _ptr(out), reps.size, C.byref(f))
    return out


def r_encode(x: np.ndarray, f: CodecFormat, mode=RNE, sat=0, stoch_len=0, endian=0) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32)
    out = np.full(int(ref().ref_packed_size(f.fmt.bitwidth, x.size)), 0xAB, np.uint8)
    n = ref().ref_encode_f32(_ptr(out), _ptr(x), x.size, C.byref(f), mode, sat, stoch_len, endian)
    assert n == out.size
    return out


def r_decode(buf: np.ndarray, count: int, f: CodecFormat, endian=0) -> np.ndarray:
    buf = np.ascontiguousarray(buf, np.uint8)
    out = np.empty(count, np.float32)
    ref().ref_decode_f32(_ptr(out), _ptr(buf), count, C.byref(f), endian)
    return out


def codec_probe_floats(f: CodecFormat) -> np.ndarray:
    """The probe set of test/test_pack.cpp:77-115: every representable value and its two fp32 neighbours, the exact
    midpoints between adjacent values, zeros, overflow, infinities, fp32 subnormals -- plus NaN and wide random data."""
    n = f.fmt.bitwidth
    grid = o_rep_to_float(np.arange(1 << n, dtype=np.uint32), f) if n <= 12 else \
        o_rep_to_float(np.arange(0, 1 << n, 37, dtype=np.uint32), f)
    fin = np.unique(grid[np.isfinite(grid)])
    mids = (0.5 * (fin[1:].astype(np.float64) + fin[:-1])).astype(np.float32)
    extras = np.array([0.0, -0.0, 1e30, -1e30, np.inf, -np.inf, 1e-40, -1e-40, 1.4e-45, np.nan, 3.4e38, -3.4e38], np.float32)
    with np.errstate(over="ignore", invalid="ignore"):
        return np.concatenate([fin, np.nextafter(fin, np.float32(np.inf)), np.nextafter(fin, np.float32(-np.inf)), mids,
                               np.nextafter(mids, np.float32(np.inf)), np.nextafter(mids, np.float32(-np.inf)), extras,
                               wide_random_f32(4096, 77 + n)]).astype(np.float32)


def o_mx_encode(x: np.ndarray, block: int, scale_f: CodecFormat, elem_f: CodecFormat, mode=RNE, scale_mode=RNE, stoch_len=0, endian=0,
                seed=0, offset=0, first_index=0):
    """-> (scales uint8 [n / block], packed elements uint8)"""
    x = np.ascontiguousarray(x, np.float32).ravel()
    scales = np.full(x.size // block, 0xAB, np.uint8)
    elems = np.full(packed_size(elem_f.fmt.bitwidth, x.size), 0xAB, np.uint8)
    ps_s, ps_e, r = ProjSpec(scale_mode, 0, 0), ProjSpec(mode, 0, stoch_len), _rng(seed, offset)
    oracle().lofo_mx_encode_f32(_ptr(x), x.size, block, C.byref(scale_f), C.byref(elem_f), C.byref(ps_s), C.byref(ps_e), endian,
                                C.byref(r), first_index, _ptr(scales), _ptr(elems))
    return scales, elems


def o_mx_decode(scales: np.ndarray, elems: np.ndarray, n: int, block: int, scale_f: CodecFormat, elem_f: CodecFormat, endian=0):
    out = np.empty(n, np.float32)
    oracle().lofo_mx_decode_f32(_ptr(np.ascontiguousarray(scales, np.uint8)), _ptr(np.ascontiguousarray(elems, np.uint8)), _ptr(out), n,
                                block, C.byref(scale_f), C.byref(elem_f), endian)
    return out


def o_fwht(x: np.ndarray, block: int) -> np.ndarray:
    """Orthonormal block Walsh-Hadamard transform along the last axis, restating LoFloat/layers.py:182-216 (`_fwht`):
    stages h = 1, 2, 4, ...: (a, b) = (x[i], x[i + h]) -> (a + b, a - b), all in float32, then * float32(1 / sqrt(block))."""
    x = np.ascontiguousarray(x, np.float32)
    flat = x.reshape(-1, block)
    M = flat.shape[0]
    h = 1
    while h < block:
        v = flat.reshape(M, block // (2 * h), 2, h)
        a, b = v[:, :, 0, :], v[:, :, 1, :]
        flat = np.stack([a + b, a - b], axis=2).reshape(M, block)
        h *= 2
    return (flat * np.float32(1.0 / np.sqrt(float(block)))).reshape(x.shape)
