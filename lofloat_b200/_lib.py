"""ctypes loader for liblofloat_b200.so (the C ABI declared in include/lofloat_b200.h).

The library is the product: there is no Python or CPU fallback.  If the shared object is
missing, or no CUDA device is usable, every op raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "liblofloat_b200.so"

# status codes (include/lofloat_b200.h)
LOF_OK, LOF_ERR_INVALID_ARG, LOF_ERR_UNSUPPORTED, LOF_ERR_CUDA, LOF_ERR_NO_DEVICE, LOF_ERR_ALLOC = range(6)
# dtypes
LOF_F32, LOF_F16, LOF_BF16, LOF_F64, LOF_E4M3, LOF_E5M2 = range(6)
# lof_gemm_chunked_f32 flags / status bits
LOF_GEMM_PARTIAL_ROUND, LOF_GEMM_NO_PACKED, LOF_GEMM_NO_HW_ACCUM, LOF_GEMM_NO_PAIR, LOF_GEMM_NO_PERSISTENT, LOF_GEMM_NO_FLOAT_FOLD = 1, 2, 4, 8, 16, 32
LOF_GEMM_STATUS_A_INEXACT, LOF_GEMM_STATUS_B_INEXACT, LOF_GEMM_STATUS_PACKED_RANGE = 1, 2, 4

# every symbol include/lofloat_b200.h declares (tests check the library exports all of them)
ABI_SYMBOLS = (
    "lof_round_params", "lof_round_mantissa", "lof_mx_round_f32", "lof_gemm_f32", "lof_gemm_chunked_f32",
    "lof_round_params_host", "lof_round_mantissa_host", "lof_mx_round_f32_host",
    "lof_abi_version", "lof_status_string", "lof_last_error_string", "lof_kernel_launch_count",
    "lof_mx_priv_max_normal", "lof_set_option", "lof_get_option", "lof_selftest_mx_division",
    "lof_packed_size", "lof_encode_f32", "lof_decode_f32", "lof_gemm_chunked_packed", "lof_round_strict_f32", "lof_mx_encode_f32", "lof_mx_decode_f32", "lof_gemm_batched_f32", "lof_fwht_f32", "lof_fwht_round_f32", "lof_fwht_mx_round_f32",
    "lof_linear_f32", "lof_conv2d_f32", "lof_conv2d_out_size", "lof_selftest_gemm_rne_round", "lof_mx_gemm", "lof_mx_pack_tc_f32", "lof_gemm_chunked_rng_f32", "lof_p2p_post", "lof_p2p_wait", "lof_p2p_done", "lof_p2p_enable_peer", "lof_p2p_copy", "lof_p2p_export", "lof_p2p_import",
)
LOF_PACK_FORWARD, LOF_PACK_REVERSE = 0, 1
LOF_CONV_K_CRS, LOF_CONV_K_RSC = 0, 1


class CFormat(C.Structure):
    """lof_format_t"""
    _fields_ = [("bitwidth", C.c_int32), ("mantissa_bits", C.c_int32), ("bias", C.c_int32),
                ("ov_behavior", C.c_int32), ("na_behavior", C.c_int32), ("is_signed", C.c_int32),
                ("unsigned_behavior", C.c_int32)]


class CCodecFormat(C.Structure):
    """lof_codec_format_t"""
    _fields_ = [("fmt", CFormat), ("inf_mask", C.c_uint32), ("inf_value", C.c_uint32), ("nan_mask", C.c_uint32),
                ("nan_value", C.c_uint32), ("nan_mant_mask", C.c_uint32), ("min_pos_inf", C.c_uint32),
                ("qnan", C.c_uint32), ("snan", C.c_uint32)]


class CProjSpec(C.Structure):
    """lof_projspec_t"""
    _fields_ = [("rounding_mode", C.c_int32), ("saturation_mode", C.c_int32), ("stoch_length", C.c_int32)]


class CConv2d(C.Structure):
    """lof_conv2d_t"""
    _fields_ = [(n, C.c_int64) for n in ("batch", "c_in", "h", "w", "c_out", "r", "s", "pad_h", "pad_w", "stride_h", "stride_w",
                                         "dilation_h", "dilation_w", "groups")] + [("k_order", C.c_int32)]


class CRng(C.Structure):
    """lof_rng_t"""
    _fields_ = [("seed", C.c_uint64), ("offset", C.c_uint64)]


class LoFloatError(RuntimeError):
    """A C-ABI call failed (mirrors the reference's TORCH_CHECK -> RuntimeError)."""

    def __init__(self, status: int, message: str):
        super().__init__(message)
        self.status = status


_lib = None


def _declare(lib):
    vp, i64, i32, u64, f64, f32 = C.c_void_p, C.c_int64, C.c_int, C.c_uint64, C.c_double, C.c_float
    PF, PP, PR = C.POINTER(CFormat), C.POINTER(CProjSpec), C.POINTER(CRng)
    lib.lof_round_params.argtypes = [vp, vp, i64, i32, PF, PP, f64, PR, u64, vp]
    lib.lof_round_mantissa.argtypes = [vp, vp, i64, i32, i32, PP, f64, PR, u64, vp]
    lib.lof_mx_round_f32.argtypes = [vp, vp, i64, i32, PF, PF, PP, PP, PR, u64, vp]
    lib.lof_gemm_f32.argtypes = [vp, vp, vp, i64, i64, i64, i32, PP, i64, f32, f32, PR, vp]
    lib.lof_gemm_chunked_f32.argtypes = [vp, vp, vp, i64, i64, i64, i32, PF, PP, i64, f32, i32, C.c_uint32, vp, vp]
    lib.lof_round_params_host.argtypes = [vp, vp, i64, i32, PF, PP, f64, PR, u64, i32]
    lib.lof_round_mantissa_host.argtypes = [vp, vp, i64, i32, i32, PP, f64, PR, u64, i32]
    lib.lof_mx_round_f32_host.argtypes = [vp, vp, i64, i32, PF, PF, PP, PP, PR, u64, i32]
    for name in ("lof_round_params", "lof_round_mantissa", "lof_mx_round_f32", "lof_gemm_f32", "lof_gemm_chunked_f32",
                 "lof_round_params_host", "lof_round_mantissa_host", "lof_mx_round_f32_host", "lof_abi_version"):
        getattr(lib, name).restype = i32
    lib.lof_status_string.argtypes = [i32]
    lib.lof_status_string.restype = C.c_char_p
    lib.lof_last_error_string.restype = C.c_char_p
    lib.lof_kernel_launch_count.restype = u64
    lib.lof_mx_priv_max_normal.argtypes = [PF]
    lib.lof_mx_priv_max_normal.restype = f32
    lib.lof_selftest_mx_division.argtypes = [PF, C.POINTER(u64), vp]
    lib.lof_selftest_mx_division.restype = i32
    PC = C.POINTER(CCodecFormat)
    lib.lof_packed_size.argtypes = [i32, u64]
    lib.lof_packed_size.restype = u64
    lib.lof_encode_f32.argtypes = [vp, vp, i64, PC, PP, i32, PR, u64, vp]
    lib.lof_encode_f32.restype = i32
    lib.lof_decode_f32.argtypes = [vp, vp, i64, PC, i32, vp]
    lib.lof_decode_f32.restype = i32
    lib.lof_gemm_batched_f32.argtypes = [vp, vp, vp, i64, i64, i64, i64, i64, i64, i64, i32, PP, i64, f32, f32, PR, vp]
    lib.lof_gemm_batched_f32.restype = i32
    lib.lof_fwht_f32.argtypes = [vp, vp, i64, i32, vp]
    lib.lof_fwht_round_f32.argtypes = [vp, vp, i64, i32, PF, PP, f64, PR, u64, vp]
    lib.lof_fwht_mx_round_f32.argtypes = [vp, vp, i64, i32, i32, PF, PF, PP, PP, PR, u64, vp]
    for _n in ("lof_gemm_batched_f32", "lof_fwht_f32", "lof_fwht_round_f32", "lof_fwht_mx_round_f32"):
        getattr(lib, _n).restype = i32
    lib.lof_mx_encode_f32.argtypes = [vp, i64, i32, PC, PC, PP, PP, i32, PR, u64, vp, vp, vp]
    lib.lof_mx_encode_f32.restype = i32
    lib.lof_mx_decode_f32.argtypes = [vp, vp, vp, i64, i32, PC, PC, i32, vp]
    lib.lof_mx_decode_f32.restype = i32
    lib.lof_round_strict_f32.argtypes = [vp, vp, i64, PC, PP, PR, u64, vp]
    lib.lof_round_strict_f32.restype = i32
    lib.lof_gemm_chunked_packed.argtypes = [vp, vp, vp, i64, i64, i64, i32, PF, PP, i64, f32, i32, C.c_uint32, vp, vp]
    lib.lof_gemm_chunked_packed.restype = i32
    lib.lof_gemm_chunked_rng_f32.argtypes = [vp, vp, vp, i64, i64, i64, i32, PF, PP, i64, f32, i32, C.c_uint32, PR, vp, vp]
    lib.lof_gemm_chunked_rng_f32.restype = i32
    lib.lof_p2p_post.argtypes = [vp, i32, i32, vp]
    lib.lof_p2p_wait.argtypes = [vp, vp, i32, vp]
    lib.lof_p2p_done.argtypes = [vp, vp]
    lib.lof_p2p_enable_peer.argtypes = [i32]
    lib.lof_p2p_enable_peer.restype = i32
    lib.lof_p2p_copy.argtypes = [vp, vp, i64, vp]
    lib.lof_p2p_copy.restype = i32
    lib.lof_p2p_export.argtypes = [vp, vp, C.POINTER(u64)]
    lib.lof_p2p_export.restype = i32
    lib.lof_p2p_import.argtypes = [vp, u64, C.POINTER(vp)]
    lib.lof_p2p_import.restype = i32
    for _n in ("lof_p2p_post", "lof_p2p_wait", "lof_p2p_done"):
        getattr(lib, _n).restype = i32
    lib.lof_mx_gemm.argtypes = [vp, vp, vp, vp, vp, i64, i64, i64, i32, i32, PF, PP, i64, f32, C.c_uint32, vp]
    lib.lof_mx_gemm.restype = i32
    lib.lof_mx_pack_tc_f32.argtypes = [vp, i64, vp, vp, vp, vp]
    lib.lof_mx_pack_tc_f32.restype = i32
    lib.lof_linear_f32.argtypes = [vp, vp, vp, vp, i64, i64, i64, i32, PP, i64, f32, PR, vp]
    lib.lof_linear_f32.restype = i32
    lib.lof_conv2d_f32.argtypes = [vp, vp, vp, vp, C.POINTER(CConv2d), i32, PP, i64, f32, PR, vp]
    lib.lof_conv2d_f32.restype = i32
    lib.lof_conv2d_out_size.argtypes = [i64, i64, i64, i64, i64, C.POINTER(i64)]
    lib.lof_conv2d_out_size.restype = i32
    lib.lof_selftest_gemm_rne_round.argtypes = [vp, vp]
    lib.lof_selftest_gemm_rne_round.restype = i32
    lib.lof_set_option.argtypes = [C.c_char_p, i64]
    lib.lof_set_option.restype = i32
    lib.lof_get_option.argtypes = [C.c_char_p, C.POINTER(i64)]
    lib.lof_get_option.restype = i32


def lib():
    """Load (once) and return the C-ABI library.  Raises if it has not been built."""
    global _lib
    if _lib is None:
        path = Path(os.environ.get("LOFLOAT_B200_LIB", LIB_PATH))
        if not path.exists():
            raise ImportError(
                f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C lofloat_b200/csrc`). lofloat_b200 has no CPU/Python fallback.")
        handle = C.CDLL(str(path))
        _declare(handle)
        _lib = handle
    return _lib


def check(status: int, what: str) -> None:
    if status != LOF_OK:
        L = lib()
        detail = L.lof_last_error_string().decode()
        kind = L.lof_status_string(status).decode()
        raise LoFloatError(status, f"{what}: {detail or kind}")


def set_option(key: str, value: int) -> None:
    check(lib().lof_set_option(key.encode(), int(value)), f"lof_set_option({key})")


def get_option(key: str) -> int:
    v = C.c_int64(0)
    check(lib().lof_get_option(key.encode(), C.byref(v)), f"lof_get_option({key})")
    return int(v.value)


def kernel_launch_count() -> int:
    return int(lib().lof_kernel_launch_count())
