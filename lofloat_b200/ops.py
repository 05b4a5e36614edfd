"""Host-side mirror of the reference's Python extension module `LoFloat.LoFloat`
(/root/reference/src/LoPy_bind.cpp:281-393): same names, argument meaning and error
behaviour, implemented as thin calls into the C ABI (include/lofloat_b200.h).

PyTorch is plumbing only: it owns device memory and the current stream.
"""
from __future__ import annotations

import ctypes as C
import enum
import threading

import torch

from . import _lib
from ._lib import CFormat, CProjSpec, CRng, LoFloatError


# --------------------------------------------------------------------------------------
# enums: same names and integer values as the pybind enums (LoPy_bind.cpp:284-313), which
# are the reference's C++ enums (fp_tools.hpp:13-50).  RoundToNearestOdd and
# RoundTiesToAway exist in C++ but are missing from the reference binding (:294-308); they
# are exposed here too.
# --------------------------------------------------------------------------------------
class RoundingMode(enum.IntEnum):
    RoundToNearestEven = 0
    RoundTowardsZero = 1
    RoundAwayFromZero = 2
    StochasticRoundingA = 3
    RoundToNearestOdd = 4
    RoundDown = 5
    RoundUp = 6
    RoundTiesToAway = 7
    StochasticRoundingB = 8
    StochasticRoundingC = 9
    True_StochasticRounding = 10
    ProbabilisticRounding = 11
    RoundToOdd = 12
    RoundTiesTowardsZero = 13
    StochasticRoundingD = 14


class Signedness(enum.IntEnum):
    Signed = 0
    Unsigned = 1


class InfBehavior(enum.IntEnum):
    Extended = 0
    Saturating = 1


class NaNBehavior(enum.IntEnum):
    _3109 = 0
    _754 = 1


class SaturationMode(enum.IntEnum):      # lo_float::Saturation_Mode (fp_tools.hpp:38-41); read by the storage codec only
    OvfInf = 0
    SatFinite = 1
    SatPropagate = 2


class PackEndian(enum.IntEnum):          # lo_float::PackEndian (lo_pack.hpp:52)
    Forward = 0
    Reverse = 1


# pybind's .export_values(): the enumerators are also module attributes
globals().update({m.name: m for e in (RoundingMode, Signedness, InfBehavior) for m in e})

_STOCHASTIC = {RoundingMode.StochasticRoundingA, RoundingMode.StochasticRoundingB, RoundingMode.StochasticRoundingC,
               RoundingMode.StochasticRoundingD, RoundingMode.True_StochasticRounding,
               RoundingMode.ProbabilisticRounding}


class FloatFormatDescriptor:
    """Python twin of FloatingPointParamsPy (LoPy_bind.cpp:45-68, 315-327): read-only
    attributes, checker objects validated the same way (std::invalid_argument -> ValueError)."""

    __slots__ = ("_total_bits", "_mantissa_bits", "_bias", "_inf_behavior", "_nan_behavior", "_signedness",
                 "_is_inf_checker", "_is_nan_checker")

    def __init__(self, total_bits, mantissa_bits, bias, inf_behavior, nan_behavior, signedness,
                 is_inf_checker, is_nan_checker):
        if not callable(is_inf_checker):
            raise ValueError("inf_checker must be callable")
        if not hasattr(is_inf_checker, "minNegInf"):
            raise ValueError("inf_checker must have minNegInf()")
        if not hasattr(is_inf_checker, "minPosInf"):
            raise ValueError("inf_checker must have minPosInf()")
        if not callable(is_nan_checker):
            raise ValueError("nan_checker must be callable")
        if not hasattr(is_nan_checker, "qNanBitPattern") and not hasattr(is_nan_checker, "sNanBitPattern"):
            raise ValueError("nan_checker must have qNanBitPattern() or sNanBitPattern()")
        self._total_bits = int(total_bits)
        self._mantissa_bits = int(mantissa_bits)
        self._bias = int(bias)
        self._inf_behavior = InfBehavior(int(inf_behavior))
        self._nan_behavior = NaNBehavior(int(nan_behavior))
        self._signedness = Signedness(int(signedness))
        self._is_inf_checker = is_inf_checker
        self._is_nan_checker = is_nan_checker

    total_bits = property(lambda self: self._total_bits)
    mantissa_bits = property(lambda self: self._mantissa_bits)
    bias = property(lambda self: self._bias)
    inf_behavior = property(lambda self: self._inf_behavior)
    nan_behavior = property(lambda self: self._nan_behavior)
    signedness = property(lambda self: self._signedness)
    is_inf_checker = property(lambda self: self._is_inf_checker)
    is_nan_checker = property(lambda self: self._is_nan_checker)

    def _c(self) -> CFormat:
        return CFormat(self._total_bits, self._mantissa_bits, self._bias, int(self._inf_behavior),
                       int(self._nan_behavior), int(self._signedness), 0)

    def __repr__(self):
        return (f"FloatFormatDescriptor(total_bits={self._total_bits}, mantissa_bits={self._mantissa_bits}, "
                f"bias={self._bias}, {self._inf_behavior.name}, {self._signedness.name})")


# --------------------------------------------------------------------------------------
# RNG state for the stochastic modes.  The reference has a thread_local `lof_seed`
# (lo_float.h:58, not bound to Python); here the seed keys a Philox stream and every
# stochastic call advances `offset`, so repeated calls draw fresh, reproducible randomness.
# --------------------------------------------------------------------------------------
_rng_lock = threading.Lock()
_rng_seed = 12345          # the reference's default lof_seed
_rng_offset = 0


def manual_seed(seed: int) -> None:
    """Seed the stochastic-rounding stream and rewind its offset (cf. lo_float::set_seed)."""
    global _rng_seed, _rng_offset
    with _rng_lock:
        _rng_seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        _rng_offset = 0


set_seed = manual_seed


def _next_rng(stochastic: bool, rng) -> CRng:
    """rng: None -> global stream; (seed, offset) tuple -> explicit, not advanced."""
    global _rng_offset
    if rng is not None:
        return CRng(int(rng[0]) & 0xFFFFFFFFFFFFFFFF, int(rng[1]) & 0xFFFFFFFFFFFFFFFF)
    with _rng_lock:
        r = CRng(_rng_seed, _rng_offset)
        if stochastic:
            _rng_offset += 1
    return r


def _chunk_rng(r: CRng, chunk: int) -> CRng:
    """Philox stream of the `chunk`-th launch of one logical call that had to be split (grid limits): the event index restarts at 0 in
    every launch, so the launches differ in the high word of the offset (the global stream only ever advances the low word)."""
    return r if chunk == 0 else CRng(r.seed, (r.offset + (int(chunk) << 32)) & 0xFFFFFFFFFFFFFFFF)


_DTYPES = {torch.float32: _lib.LOF_F32, torch.float16: _lib.LOF_F16, torch.bfloat16: _lib.LOF_BF16,
           torch.float64: _lib.LOF_F64}


def _dtype_code(t: torch.Tensor, what: str) -> int:
    code = _DTYPES.get(t.dtype)
    if code is None:
        # AT_DISPATCH_FLOATING_TYPES_AND_HALF error text (LoPy_bind.cpp:83,133)
        raise RuntimeError(f'"{what}" not implemented for \'{t.dtype}\'')
    return code


def _stream_ptr(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _require_gpu():
    if not torch.cuda.is_available():
        raise LoFloatError(_lib.LOF_ERR_NO_DEVICE,
                           "lofloat_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


def _round(input: torch.Tensor, params, to_mantissa_bits, round_mode, stoch_len, scale, rng, first_index):
    L = _lib.lib()
    what = "virtual_round_params" if params is not None else "virtual_round_mantissa"
    code = _dtype_code(input, what)
    mode = RoundingMode(int(round_mode))
    ps = CProjSpec(int(mode), 0, int(stoch_len))
    r = _next_rng(mode in _STOCHASTIC, rng)
    x = input if input.is_contiguous() else input.contiguous()
    out = torch.empty_like(x)
    n = x.numel()
    _require_gpu()
    if x.is_cuda:
        with torch.cuda.device(x.device):
            if params is not None:
                fmt = params._c()
                st = L.lof_round_params(x.data_ptr(), out.data_ptr(), n, code, C.byref(fmt), C.byref(ps), float(scale),
                                        C.byref(r), int(first_index), _stream_ptr(x))
            else:
                st = L.lof_round_mantissa(x.data_ptr(), out.data_ptr(), n, code, int(to_mantissa_bits), C.byref(ps),
                                          float(scale), C.byref(r), int(first_index), _stream_ptr(x))
    else:
        # host tensor: staged through the GPU by the library's host-buffer pipeline (still no CPU compute)
        if params is not None:
            fmt = params._c()
            st = L.lof_round_params_host(x.data_ptr(), out.data_ptr(), n, code, C.byref(fmt), C.byref(ps), float(scale),
                                         C.byref(r), int(first_index), -1)
        else:
            st = L.lof_round_mantissa_host(x.data_ptr(), out.data_ptr(), n, code, int(to_mantissa_bits), C.byref(ps),
                                           float(scale), C.byref(r), int(first_index), -1)
    _lib.check(st, what)
    return out


def virtual_round(input: torch.Tensor, params=None, round_mode=RoundingMode.RoundToNearestEven, stoch_len: int = 0,
                  scale: float = 1.0, *, to_mantissa_bits=None, rng=None, first_index: int = 0) -> torch.Tensor:
    """Both reference overloads (LoPy_bind.cpp:329-339):

        virtual_round(input, to_mantissa_bits: int, round_mode, stoch_len, scale)
        virtual_round(input, params: FloatFormatDescriptor, round_mode, stoch_len, scale)

    Returns a new tensor; if scale != 1 computes round(scale * input) in one pass.
    Extensions (keyword-only): `rng=(seed, offset)` pins the Philox stream, `first_index`
    is the global index of element 0 (row-sharded tensors).
    """
    if to_mantissa_bits is not None and params is None:
        return _round(input, None, to_mantissa_bits, round_mode, stoch_len, scale, rng, first_index)
    if isinstance(params, FloatFormatDescriptor):
        return _round(input, params, None, round_mode, stoch_len, scale, rng, first_index)
    if isinstance(params, int) and not isinstance(params, bool):
        return _round(input, None, params, round_mode, stoch_len, scale, rng, first_index)
    raise TypeError("virtual_round(): incompatible function arguments. The following argument types are supported:\n"
                    "    1. (input: Tensor, to_mantissa_bits: int, round_mode=RoundToNearestEven, stoch_len=0, scale=1.0)\n"
                    "    2. (input: Tensor, params: FloatFormatDescriptor, round_mode=RoundToNearestEven, stoch_len=0, scale=1.0)")


def virtual_mx_round(input: torch.Tensor, block_size: int, element_format: FloatFormatDescriptor,
                     scale_format: FloatFormatDescriptor, round_mode=RoundingMode.RoundToNearestEven,
                     scale_round_mode=RoundingMode.RoundToNearestEven, stoch_len: int = 0, *, rng=None,
                     first_index: int = 0) -> torch.Tensor:
    """MX block fake-quantisation, float32 only (LoPy_bind.cpp:165-217, 341-351)."""
    if input.dtype != torch.float32:
        raise RuntimeError("virtual_mx_round: float32 tensors only")
    if block_size <= 0:
        raise RuntimeError("virtual_mx_round: block_size must be positive")
    L = _lib.lib()
    mode, smode = RoundingMode(int(round_mode)), RoundingMode(int(scale_round_mode))
    ps_e = CProjSpec(int(mode), 0, int(stoch_len))
    ps_s = CProjSpec(int(smode), 0, 0)
    r = _next_rng(mode in _STOCHASTIC or smode in _STOCHASTIC, rng)
    x = input.contiguous()
    out = torch.empty_like(x)
    efmt, sfmt = element_format._c(), scale_format._c()
    _require_gpu()
    if x.is_cuda:
        with torch.cuda.device(x.device):
            st = L.lof_mx_round_f32(x.data_ptr(), out.data_ptr(), x.numel(), int(block_size), C.byref(sfmt),
                                    C.byref(efmt), C.byref(ps_s), C.byref(ps_e), C.byref(r), int(first_index),
                                    _stream_ptr(x))
    else:
        st = L.lof_mx_round_f32_host(x.data_ptr(), out.data_ptr(), x.numel(), int(block_size), C.byref(sfmt),
                                     C.byref(efmt), C.byref(ps_s), C.byref(ps_e), C.byref(r), int(first_index), -1)
    _lib.check(st, "virtual_mx_round")
    return out.view_as(input)


# GEMM semantics switch.  kc = 0 (default) is the reference CUDA leaf -- the only path the reference's Python
# lof_gemm can reach (LoPy_bind.cpp:360-372 is inside #ifdef USE_CUDA): one accumulator over all of K
# (cutlass_gemms.cuh:97-168).  set_gemm_panel(256) reproduces the reference CPU path (bli_sgemm + lof_sgemm_ukr,
# blis_lof_gemm.hpp:147-262, BLIS generic KC = 256) bit for bit.
_gemm_kc = 0


# Opt-in for callers that cannot pass keywords -- the reference's LoF_Linear / torch.ops.lofloat.lof_gemm call
# lof_gemm(A, B, bits, mode, sr_bits, scale_a, scale_b) positionally (LoFloat/_custom_ops.py:92-99, layers.py:152-168): with
# set_gemm_tensor_core(32) every such call runs the tensor-core chunked mode (accumulator re-rounded every k_chunk) instead of the
# per-FMA SIMT kernel.  None (default) = the reference semantics, bit-exact vs BLIS / the CUDA leaf.
_gemm_tc = None


def set_gemm_tensor_core(k_chunk: int | None, operand_dtype: str = "e4m3", partial_round: bool = False, check: bool = False) -> None:
    global _gemm_tc
    if k_chunk is None:
        _gemm_tc = None
        return
    if operand_dtype not in _OPERAND_DTYPES:
        raise RuntimeError("set_gemm_tensor_core: operand_dtype must be 'e4m3', 'e5m2' or 'bf16'")
    _gemm_tc = (int(k_chunk), operand_dtype, bool(partial_round), bool(check))


def set_gemm_panel(kc: int) -> None:
    global _gemm_kc
    _gemm_kc = int(kc)


_OPERAND_DTYPES = {"e4m3": _lib.LOF_E4M3, "e5m2": _lib.LOF_E5M2, "bf16": _lib.LOF_BF16}


def lof_gemm(A: torch.Tensor, B: torch.Tensor, accum_mant_bits: int, round_mode=RoundingMode.RoundToNearestEven,
             stochastic_rounding_bits: int = 0, scale_a: float = 1.0, scale_b: float = 1.0, *, kc=None,
             k_chunk=None, accum_format: FloatFormatDescriptor | None = None, operand_dtype: str = "e4m3",
             partial_round: bool = False, allow_packed: bool = True, allow_hw_accum: bool = True, allow_pair: bool = True,
             allow_persistent: bool = True, allow_float_fold: bool = True, check: bool = True, rng=None) -> torch.Tensor:
    """Low-precision-accumulation GEMM, float32 row-major (Lof_kernel.cu:96-154, LoPy_bind.cpp:361-372).
    The output is divided by scale_a*scale_b (a zero product counts as 1, cutlass_gemms.cuh:1784-1786).

    Extensions (keyword-only): `kc` panel length of the exact per-FMA mode.  `k_chunk` selects the
    tensor-core chunked mode (lof_gemm_chunked_f32): the accumulator is re-rounded once per k_chunk,
    mantissa-only (accum_mant_bits) or into the full `accum_format` incl. its exponent range; A and B must
    be exactly representable in `operand_dtype` ("e4m3", "e5m2" or "bf16") -- with check=True (default) the
    call synchronises like the reference (Lof_kernel.cu:149) and raises if they are not.
    """
    if k_chunk is None and kc is None and _gemm_tc is not None:
        k_chunk, operand_dtype, partial_round, check = _gemm_tc          # global opt-in (set_gemm_tensor_core)
    if A.dim() != 2 or B.dim() != 2:
        raise RuntimeError("LoF_gemm: A and B must be 2-D")
    if A.size(1) != B.size(0):
        raise RuntimeError(f"LoF_gemm: A.cols ({A.size(1)}) != B.rows ({B.size(0)})")
    if A.dtype != torch.float32 or B.dtype != torch.float32:
        raise RuntimeError("LoF_gemm: only float32 inputs are supported")
    if not (A.is_cuda and B.is_cuda):
        raise RuntimeError(f"LoF_gemm requires CUDA tensors (A on {A.device}, B on {B.device})")
    L = _lib.lib()
    mode = RoundingMode(int(round_mode))
    ps = CProjSpec(int(mode), 0, int(stochastic_rounding_bits))
    denom = float(scale_a) * float(scale_b)
    alpha = 1.0 if denom == 0.0 else 1.0 / denom
    A_c, B_c = A.contiguous(), B.contiguous()
    M, K = A_c.shape
    N = B_c.shape[1]
    out = torch.empty((M, N), dtype=torch.float32, device=A.device)      # every element is written (beta = 0); the reference zero-fills first (Lof_kernel.cu:119)
    status = None
    with torch.cuda.device(A.device):
        if k_chunk is not None:
            fmt = accum_format._c() if accum_format is not None else None
            flags = ((_lib.LOF_GEMM_PARTIAL_ROUND if partial_round else 0) | (0 if allow_packed else _lib.LOF_GEMM_NO_PACKED) |
                     (0 if allow_hw_accum else _lib.LOF_GEMM_NO_HW_ACCUM) | (0 if allow_pair else _lib.LOF_GEMM_NO_PAIR) |
                     (0 if allow_persistent else _lib.LOF_GEMM_NO_PERSISTENT) | (0 if allow_float_fold else _lib.LOF_GEMM_NO_FLOAT_FOLD))
            status = torch.zeros(1, dtype=torch.int32, device=A.device)
            r = _next_rng(mode in _STOCHASTIC, rng)          # stochastic accumulator rounding: one Philox word per rounding event
            st = L.lof_gemm_chunked_rng_f32(A_c.data_ptr(), B_c.data_ptr(), out.data_ptr(), M, N, K, int(accum_mant_bits),
                                            C.byref(fmt) if fmt is not None else None, C.byref(ps), int(k_chunk), alpha,
                                            _OPERAND_DTYPES[operand_dtype], flags, C.byref(r), status.data_ptr(), _stream_ptr(A_c))
        else:
            r = _next_rng(mode in _STOCHASTIC, rng)
            st = L.lof_gemm_f32(A_c.data_ptr(), B_c.data_ptr(), out.data_ptr(), M, N, K, int(accum_mant_bits),
                                C.byref(ps), int(_gemm_kc if kc is None else kc), alpha, 0.0, C.byref(r),
                                _stream_ptr(A_c))
    _lib.check(st, "LoF_gemm")
    if status is not None and check:
        bits = int(status.item())
        if bits & (_lib.LOF_GEMM_STATUS_A_INEXACT | _lib.LOF_GEMM_STATUS_B_INEXACT):
            which = [n for b, n in ((1, "A"), (2, "B")) if bits & b]
            raise LoFloatError(_lib.LOF_ERR_INVALID_ARG,
                               f"LoF_gemm (chunked): {' and '.join(which)} not exactly representable in {operand_dtype}")
        lof_gemm.last_status = bits
    return out


lof_gemm.last_status = 0


# --------------------------------------------------------------------------------------
# Low-bit storage codec: host mirror of LoF_encode_cuda / LoF_decode_cuda (lo_pack.hpp:499-529; the reference
# has no Python binding for them yet -- "trivial to bind to Python later", lo_pack.hpp:17 -- so the names
# follow the C++ API).
# --------------------------------------------------------------------------------------
def _fit_checker(pred, n: int, what: str):
    """Express a checker callable over n-bit representations as (mask, value, mant_mask):
    pred(rep) == ((rep & mask) == value and (mant_mask == 0 or (rep & mant_mask) != 0))."""
    full = (1 << n) - 1
    S = [r for r in range(1 << n) if pred(r)]
    if not S:
        return 0, 1, 0
    and_all, or_all = full, 0
    for r in S:
        and_all &= r
        or_all |= r
    mask = ~(and_all ^ or_all) & full
    value = and_all & mask
    free = full & ~mask
    if len(S) == 1 << bin(free).count("1"):
        return mask, value, 0
    Sset = set(S)
    excluded_or = 0
    for r in range(1 << n):
        if (r & mask) == value and r not in Sset:
            excluded_or |= r
    mant_mask = free & ~excluded_or
    ok = mant_mask != 0 and all(((r & mask) == value and (r & mant_mask) != 0) == (r in Sset) for r in range(1 << n))
    if not ok:
        raise ValueError(f"{what} is not expressible as (rep & mask) == value [and (rep & mant_mask) != 0]")
    return mask, value, mant_mask


_codec_cache: dict = {}


def codec_format(params: FloatFormatDescriptor, unsigned_behavior: int = 0) -> "_lib.CCodecFormat":
    """lof_codec_format_t of a descriptor: the Inf / NaN checker callables are evaluated over every n-bit
    representation once (n <= 16) and folded into masks."""
    key = (id(params), int(unsigned_behavior))
    hit = _codec_cache.get(key)
    if hit is not None and hit[0] is params:
        return hit[1]
    n = params.total_bits
    if not 1 <= n <= 16:
        raise ValueError("storage codec: total_bits must be in [1, 16]")
    inf_m, inf_v, inf_mm = _fit_checker(params.is_inf_checker, n, "is_inf_checker")
    if inf_mm:
        raise ValueError("is_inf_checker is not expressible as (rep & mask) == value")
    nan_m, nan_v, nan_mm = _fit_checker(params.is_nan_checker, n, "is_nan_checker")
    q = params.is_nan_checker.qNanBitPattern() if hasattr(params.is_nan_checker, "qNanBitPattern") else params.is_nan_checker.sNanBitPattern()
    s_ = params.is_nan_checker.sNanBitPattern() if hasattr(params.is_nan_checker, "sNanBitPattern") else q
    f = params._c()
    f.unsigned_behavior = int(unsigned_behavior)
    cf = _lib.CCodecFormat(f, inf_m, inf_v, nan_m, nan_v, nan_mm, int(params.is_inf_checker.minPosInf()) & 0xFFFFFFFF,
                           int(q) & 0xFFFFFFFF, int(s_) & 0xFFFFFFFF)
    _codec_cache[key] = (params, cf)
    return cf


def packed_size(total_bits: int, count: int) -> int:
    """LoF_packed_size (lo_pack.hpp:399)."""
    return int(_lib.lib().lof_packed_size(int(total_bits), int(count)))


def lof_encode(input: torch.Tensor, params: FloatFormatDescriptor, round_mode=RoundingMode.RoundToNearestEven,
               stoch_len: int = 0, saturation_mode=SaturationMode.OvfInf, endian=PackEndian.Forward, *,
               unsigned_behavior: int = 0, rng=None, first_index: int = 0) -> torch.Tensor:
    """float32 CUDA tensor -> packed uint8 tensor of packed_size(bits, numel) bytes (LoF_encode_cuda, lo_pack.hpp:499)."""
    if input.dtype != torch.float32:
        raise RuntimeError("lof_encode: float32 tensors only")
    if not input.is_cuda:
        raise RuntimeError("lof_encode requires a CUDA tensor")
    L = _lib.lib()
    mode = RoundingMode(int(round_mode))
    ps = CProjSpec(int(mode), int(saturation_mode), int(stoch_len))
    r = _next_rng(mode in _STOCHASTIC, rng)
    cf = codec_format(params, unsigned_behavior)
    x = input.contiguous()
    out = torch.empty(packed_size(params.total_bits, x.numel()), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        st = L.lof_encode_f32(x.data_ptr(), out.data_ptr(), x.numel(), C.byref(cf), C.byref(ps), int(endian), C.byref(r),
                              int(first_index), _stream_ptr(x))
    _lib.check(st, "lof_encode")
    return out


def lof_decode(packed: torch.Tensor, count: int, params: FloatFormatDescriptor, endian=PackEndian.Forward, *,
               unsigned_behavior: int = 0) -> torch.Tensor:
    """packed uint8 CUDA tensor -> `count` float32 values (LoF_decode_cuda, lo_pack.hpp:516)."""
    if packed.dtype != torch.uint8:
        raise RuntimeError("lof_decode: uint8 tensors only")
    if not packed.is_cuda:
        raise RuntimeError("lof_decode requires a CUDA tensor")
    need = packed_size(params.total_bits, count)
    if packed.numel() < need:
        raise RuntimeError(f"lof_decode: packed buffer has {packed.numel()} bytes, {need} needed for {count} elements")
    L = _lib.lib()
    cf = codec_format(params, unsigned_behavior)
    src = packed.contiguous()
    out = torch.empty(int(count), dtype=torch.float32, device=packed.device)
    with torch.cuda.device(packed.device):
        st = L.lof_decode_f32(src.data_ptr(), out.data_ptr(), int(count), C.byref(cf), int(endian), _stream_ptr(src))
    _lib.check(st, "lof_decode")
    return out


def virtual_round_strict(input: torch.Tensor, params: FloatFormatDescriptor, round_mode=RoundingMode.RoundToNearestEven,
                         stoch_len: int = 0, saturation_mode=SaturationMode.OvfInf, *, unsigned_behavior: int = 0, rng=None,
                         first_index: int = 0) -> torch.Tensor:
    """Strict typed-semantics rounding: (float)Templated_Float<Fp>(x) elementwise (ConvertImpl::run, lo_float.h:3057-3364)
    -- overflow to +-inf / +-max by `saturation_mode`, NaN propagation, NegtoZero / NegtoNaN -- as one fused pass."""
    if input.dtype != torch.float32:
        raise RuntimeError("virtual_round_strict: float32 tensors only")
    if not input.is_cuda:
        raise RuntimeError("virtual_round_strict requires a CUDA tensor")
    L = _lib.lib()
    mode = RoundingMode(int(round_mode))
    ps = CProjSpec(int(mode), int(saturation_mode), int(stoch_len))
    r = _next_rng(mode in _STOCHASTIC, rng)
    cf = codec_format(params, unsigned_behavior)
    x = input.contiguous()
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        st = L.lof_round_strict_f32(x.data_ptr(), out.data_ptr(), x.numel(), C.byref(cf), C.byref(ps), C.byref(r), int(first_index),
                                    _stream_ptr(x))
    _lib.check(st, "virtual_round_strict")
    return out


def lof_gemm_batched(A: torch.Tensor, B: torch.Tensor, accum_mant_bits: int, round_mode=RoundingMode.RoundToNearestEven,
                     stochastic_rounding_bits: int = 0, scale_a: float = 1.0, scale_b: float = 1.0, *, kc=None, rng=None,
                     k_chunk=None, **tensor_core_kwargs) -> torch.Tensor:
    """(..., M, K) x (..., K, N) -> (..., M, N) with LoFloat accumulation, the `_lof_gemm_2d` of LoFloat/layers.py:121-150,
    as ONE launch over the flattened batch instead of a Python loop; every slice equals lof_gemm on that slice.
    With `k_chunk` (or after set_gemm_tensor_core()) every slice runs the tensor-core chunked mode of lof_gemm (same keyword
    arguments: accum_format, operand_dtype, partial_round, check, ...): one launch per slice, written into one result tensor."""
    if A.dim() < 2 or B.dim() < 2 or A.shape[:-2] != B.shape[:-2] or A.shape[-1] != B.shape[-2]:
        raise RuntimeError(f"lof_gemm_batched: incompatible shapes {tuple(A.shape)} x {tuple(B.shape)}")
    if A.dtype != torch.float32 or B.dtype != torch.float32:
        raise RuntimeError("LoF_gemm: only float32 inputs are supported")
    if not (A.is_cuda and B.is_cuda):
        raise RuntimeError("LoF_gemm requires CUDA tensors")
    if k_chunk is not None or (kc is None and _gemm_tc is not None):
        M, K, N = A.shape[-2], A.shape[-1], B.shape[-1]
        A_f, B_f = A.reshape(-1, M, K), B.reshape(-1, K, N)
        out = torch.empty((A_f.shape[0], M, N), dtype=torch.float32, device=A.device)
        for b in range(A_f.shape[0]):
            out[b] = lof_gemm(A_f[b], B_f[b], accum_mant_bits, round_mode, stochastic_rounding_bits, scale_a, scale_b, k_chunk=k_chunk,
                              rng=rng, **tensor_core_kwargs)
        return out.reshape(*A.shape[:-2], M, N)
    if tensor_core_kwargs:
        raise TypeError(f"lof_gemm_batched: {sorted(tensor_core_kwargs)} need k_chunk (tensor-core chunked mode)")
    L = _lib.lib()
    mode = RoundingMode(int(round_mode))
    ps = CProjSpec(int(mode), 0, int(stochastic_rounding_bits))
    denom = float(scale_a) * float(scale_b)
    alpha = 1.0 if denom == 0.0 else 1.0 / denom
    batch_shape = A.shape[:-2]
    M, K, N = A.shape[-2], A.shape[-1], B.shape[-1]
    A_f, B_f = A.reshape(-1, M, K).contiguous(), B.reshape(-1, K, N).contiguous()
    nb = A_f.shape[0]
    out = torch.empty((nb, M, N), dtype=torch.float32, device=A.device)
    r = _next_rng(mode in _STOCHASTIC, rng)
    with torch.cuda.device(A.device):
        for b0 in range(0, nb, 65535):                       # grid.z limit
            cnt = min(65535, nb - b0)
            rc = _chunk_rng(r, b0 // 65535)                  # a later chunk must not replay the event stream of the first
            st = L.lof_gemm_batched_f32(A_f[b0].data_ptr(), B_f[b0].data_ptr(), out[b0].data_ptr(), cnt, M, N, K, M * K, K * N, M * N,
                                        int(accum_mant_bits), C.byref(ps), int(_gemm_kc if kc is None else kc), alpha, 0.0, C.byref(rc),
                                        _stream_ptr(A_f))
            _lib.check(st, "LoF_gemm (batched)")
    return out.reshape(*batch_shape, M, N)


def lof_linear(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None, accum_mant_bits: int,
               round_mode=RoundingMode.RoundToNearestEven, stochastic_rounding_bits: int = 0, scale_a: float = 1.0, scale_b: float = 1.0,
               *, kc=None, rng=None, k_chunk=None, **tensor_core_kwargs) -> torch.Tensor:
    """F.linear with LoFloat accumulation: x (..., K), weight (N, K) -> (..., N) -- `_lof_linear` of LoFloat/layers.py:152-168
    as ONE kernel: the weight is read as stored (no `weight.t().contiguous()` copy) and the already quantised `bias` is added in
    the epilogue (plain fp32 add, `out + bias`).  Bit-identical to lof_gemm(x_2d, weight.t().contiguous(), ...) + bias.
    With `k_chunk` (or after set_gemm_tensor_core()) the product runs the tensor-core chunked mode of lof_gemm (same keyword
    arguments) and the bias is added afterwards."""
    if weight.dim() != 2 or x.dim() < 1 or x.shape[-1] != weight.shape[1]:
        raise RuntimeError(f"lof_linear: incompatible shapes {tuple(x.shape)} x {tuple(weight.shape)}")
    if x.dtype != torch.float32 or weight.dtype != torch.float32 or (bias is not None and bias.dtype != torch.float32):
        raise RuntimeError("LoF_gemm: only float32 inputs are supported")
    if not (x.is_cuda and weight.is_cuda and (bias is None or bias.is_cuda)):
        raise RuntimeError("LoF_gemm requires CUDA tensors")
    N, K = weight.shape
    if bias is not None and bias.numel() != N:
        raise RuntimeError(f"lof_linear: bias has {bias.numel()} values, expected {N}")
    if k_chunk is not None or (kc is None and _gemm_tc is not None):
        out = lof_gemm(x.reshape(-1, K), weight.t().contiguous(), accum_mant_bits, round_mode, stochastic_rounding_bits, scale_a, scale_b,
                       k_chunk=k_chunk, rng=rng, **tensor_core_kwargs)
        if bias is not None:
            out += bias
        return out.reshape(*x.shape[:-1], N)
    if tensor_core_kwargs:
        raise TypeError(f"lof_linear: {sorted(tensor_core_kwargs)} need k_chunk (tensor-core chunked mode)")
    mode = RoundingMode(int(round_mode))
    ps = CProjSpec(int(mode), 0, int(stochastic_rounding_bits))
    denom = float(scale_a) * float(scale_b)
    alpha = 1.0 if denom == 0.0 else 1.0 / denom
    x2, w_c = x.reshape(-1, K).contiguous(), weight.contiguous()
    b_c = bias.contiguous() if bias is not None else None
    out = torch.empty((x2.shape[0], N), dtype=torch.float32, device=x.device)
    r = _next_rng(mode in _STOCHASTIC, rng)
    with torch.cuda.device(x.device):
        st = _lib.lib().lof_linear_f32(x2.data_ptr(), w_c.data_ptr(), b_c.data_ptr() if b_c is not None else None, out.data_ptr(),
                                       x2.shape[0], N, K, int(accum_mant_bits), C.byref(ps), int(_gemm_kc if kc is None else kc), alpha,
                                       C.byref(r), _stream_ptr(x2))
    _lib.check(st, "lof_linear")
    return out.reshape(*x.shape[:-1], N)


def lof_conv2d(input: torch.Tensor, weight: torch.Tensor, pad_h: int, pad_w: int, stride_h: int, stride_w: int, dilation_h: int,
               dilation_w: int, accum_mant_bits: int, round_mode=RoundingMode.RoundToNearestEven, stochastic_rounding_bits: int = 0,
               weight_scale: float = 1.0, input_scale: float = 1.0, *, groups: int = 1, bias: torch.Tensor | None = None,
               k_order: str = "rsc", kc=None, rng=None) -> torch.Tensor:
    """Low-precision Conv2d forward, float32 NCHW input x OIHW weight -> NCHW output, divided by weight_scale * input_scale
    (LoPy_bind.cpp:374-392, Lof_kernel.cu:156-228).  Implicit GEMM: no im2col matrix is written.

    Defaults reproduce the reference entry point: reduction order (r, s, c) and one accumulator over all of K.  Keyword
    extensions: `k_order="crs"` (+ `kc=256` for the CPU GEMM's panels) is LoF_Conv2d's im2col backend -- F.unfold followed
    by one lof_gemm per image and group (LoFloat/layers.py:710-745) -- as one launch, `groups`, and a fused `bias` epilogue
    (`out + b_q.view(1, -1, 1, 1)`, layers.py:747-751)."""
    if input.dim() != 4 or weight.dim() != 4:
        raise RuntimeError("LoF_conv2d: input and weight must be 4-D (NCHW / OIHW)")
    if input.dtype != torch.float32 or weight.dtype != torch.float32 or (bias is not None and bias.dtype != torch.float32):
        raise RuntimeError("LoF_conv2d: only float32 inputs are supported")
    if not (input.is_cuda and weight.is_cuda and (bias is None or bias.is_cuda)):
        raise RuntimeError(f"LoF_conv2d requires CUDA tensors (input on {input.device}, weight on {weight.device})")
    if groups < 1 or input.size(1) % groups or weight.size(0) % groups:
        raise RuntimeError("LoF_conv2d: groups must divide C_in and C_out")
    if input.size(1) != weight.size(1) * groups:
        raise RuntimeError(f"LoF_conv2d: input C_in ({input.size(1)}) != weight C_in ({weight.size(1) * groups})")
    if k_order not in ("rsc", "crs"):
        raise ValueError("k_order must be 'rsc' or 'crs'")
    L = _lib.lib()
    B, C_in, H, W = input.shape
    C_out, _, R, S = weight.shape
    if bias is not None and bias.numel() != C_out:
        raise RuntimeError(f"LoF_conv2d: bias has {bias.numel()} values, expected {C_out}")
    P, Q = C.c_int64(0), C.c_int64(0)
    _lib.check(L.lof_conv2d_out_size(H, R, pad_h, stride_h, dilation_h, C.byref(P)), "LoF_conv2d")
    _lib.check(L.lof_conv2d_out_size(W, S, pad_w, stride_w, dilation_w, C.byref(Q)), "LoF_conv2d")
    if P.value <= 0 or Q.value <= 0:
        raise RuntimeError(f"LoF_conv2d: non-positive output spatial size (P={P.value}, Q={Q.value})")
    mode = RoundingMode(int(round_mode))
    ps = CProjSpec(int(mode), 0, int(stochastic_rounding_bits))
    denom = float(weight_scale) * float(input_scale)
    alpha = 1.0 if denom == 0.0 else 1.0 / denom
    x_c, w_c = input.contiguous(), weight.contiguous()
    b_c = bias.contiguous() if bias is not None else None
    out = torch.empty((B, C_out, P.value, Q.value), dtype=torch.float32, device=input.device)
    r = _next_rng(mode in _STOCHASTIC, rng)
    per_call = max(1, 65535 // groups)
    with torch.cuda.device(input.device):
        for b0 in range(0, B, per_call):
            cnt = min(per_call, B - b0)
            geo = _lib.CConv2d(cnt, C_in, H, W, C_out, R, S, pad_h, pad_w, stride_h, stride_w, dilation_h, dilation_w, groups,
                               _lib.LOF_CONV_K_RSC if k_order == "rsc" else _lib.LOF_CONV_K_CRS)
            st = L.lof_conv2d_f32(x_c[b0].data_ptr(), w_c.data_ptr(), b_c.data_ptr() if b_c is not None else None, out[b0].data_ptr(),
                                  C.byref(geo), int(accum_mant_bits), C.byref(ps), int(0 if kc is None else kc), alpha,
                                  C.byref(_chunk_rng(r, b0 // per_call)), _stream_ptr(x_c))
            _lib.check(st, "LoF_conv2d")
    return out


def fwht(input: torch.Tensor, block_size=None, *, params: FloatFormatDescriptor | None = None, round_mode=RoundingMode.RoundToNearestEven,
         stoch_len: int = 0, scale: float = 1.0, mx_block_size: int | None = None, scale_format: FloatFormatDescriptor | None = None,
         scale_round_mode=RoundingMode.RoundToNearestEven, rng=None, first_index: int = 0) -> torch.Tensor:
    """Orthonormal Walsh-Hadamard transform along the last dim in blocks of `block_size` (None = the whole last dim), the
    `_fwht` of LoFloat/layers.py:182-216, optionally FUSED with the rounding LoF_Linear applies next: `params` alone ->
    virtual_round(scale * y, params); `params` + `mx_block_size` (+ `scale_format`) -> virtual_mx_round(y, ...)."""
    if input.dtype != torch.float32 or not input.is_cuda:
        raise RuntimeError("fwht: float32 CUDA tensors only")
    n_last = input.shape[-1]
    if block_size is None:
        block_size = n_last
    if block_size <= 0 or (block_size & (block_size - 1)) != 0:
        raise ValueError(f"block_size must be a power of 2, got {block_size}")
    if n_last % block_size != 0:
        raise ValueError(f"block_size {block_size} must divide last dim {n_last}")
    L = _lib.lib()
    x = input.contiguous()
    out = torch.empty_like(x)
    n = x.numel()
    mode = RoundingMode(int(round_mode))
    if block_size > 1024 and params is not None:
        # blocks of 2048 .. 32768 (block_size=None on an LLM-sized in_features): two-pass transform, then the rounding op (the
        # single-pass fused kernels hold a whole Hadamard block in one warp's registers, up to 1024 elements)
        y = fwht(x, block_size)
        if mx_block_size is None:
            return virtual_round(y, params, mode, stoch_len, scale, rng=rng, first_index=first_index)
        sf = scale_format if scale_format is not None else _default_e8m0()
        return virtual_mx_round(y, int(mx_block_size), params, sf, mode, RoundingMode(int(scale_round_mode)), stoch_len, rng=rng,
                                first_index=first_index)
    with torch.cuda.device(x.device):
        if params is None:
            st = L.lof_fwht_f32(x.data_ptr(), out.data_ptr(), n, int(block_size), _stream_ptr(x))
        elif mx_block_size is None:
            ps = CProjSpec(int(mode), 0, int(stoch_len))
            r = _next_rng(mode in _STOCHASTIC, rng)
            f = params._c()
            st = L.lof_fwht_round_f32(x.data_ptr(), out.data_ptr(), n, int(block_size), C.byref(f), C.byref(ps), float(scale), C.byref(r),
                                      int(first_index), _stream_ptr(x))
        else:
            if n_last % int(mx_block_size) != 0:
                raise ValueError(f"mx scaling needs the last dim ({n_last}) divisible by mx_block_size ({mx_block_size})")
            sf = scale_format if scale_format is not None else _default_e8m0()
            smode = RoundingMode(int(scale_round_mode))
            ps_e, ps_s = CProjSpec(int(mode), 0, int(stoch_len)), CProjSpec(int(smode), 0, 0)
            r = _next_rng(mode in _STOCHASTIC or smode in _STOCHASTIC, rng)
            fe, fs = params._c(), sf._c()
            st = L.lof_fwht_mx_round_f32(x.data_ptr(), out.data_ptr(), n, int(block_size), int(mx_block_size), C.byref(fs), C.byref(fe),
                                         C.byref(ps_s), C.byref(ps_e), C.byref(r), int(first_index), _stream_ptr(x))
    _lib.check(st, "fwht")
    return out


def _default_e8m0():
    from .formats import create_e8m0_params
    return create_e8m0_params()


def mx_encode(input: torch.Tensor, block_size: int, element_format: FloatFormatDescriptor, scale_format: FloatFormatDescriptor,
              round_mode=RoundingMode.RoundToNearestEven, scale_round_mode=RoundingMode.RoundToNearestEven, stoch_len: int = 0,
              endian=PackEndian.Forward, *, rng=None, first_index: int = 0):
    """MX quantisation to real storage: returns (scales uint8 [numel / block_size], packed elements uint8).  Blocks run
    along the flattened tensor (keep K last), same scale / element rounding as virtual_mx_round."""
    if input.dtype != torch.float32 or not input.is_cuda:
        raise RuntimeError("mx_encode: float32 CUDA tensors only")
    L = _lib.lib()
    mode, smode = RoundingMode(int(round_mode)), RoundingMode(int(scale_round_mode))
    ps_e, ps_s = CProjSpec(int(mode), 0, int(stoch_len)), CProjSpec(int(smode), 0, 0)
    r = _next_rng(mode in _STOCHASTIC or smode in _STOCHASTIC, rng)
    x = input.contiguous()
    n = x.numel()
    if int(block_size) <= 0 or n % int(block_size) != 0:
        raise LoFloatError(_lib.LOF_ERR_INVALID_ARG, f"mx_encode: numel ({n}) must be a multiple of block_size ({block_size})")
    cfe, cfs = codec_format(element_format), codec_format(scale_format)
    scales = torch.empty(n // int(block_size), dtype=torch.uint8, device=x.device)
    elems = torch.empty(packed_size(element_format.total_bits, n), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        st = L.lof_mx_encode_f32(x.data_ptr(), n, int(block_size), C.byref(cfs), C.byref(cfe), C.byref(ps_s), C.byref(ps_e), int(endian),
                                 C.byref(r), int(first_index), scales.data_ptr(), elems.data_ptr(), _stream_ptr(x))
    _lib.check(st, "mx_encode")
    return scales, elems


def mx_decode(scales: torch.Tensor, elems: torch.Tensor, count: int, block_size: int, element_format: FloatFormatDescriptor,
              scale_format: FloatFormatDescriptor, endian=PackEndian.Forward) -> torch.Tensor:
    """Inverse of mx_encode: float32 values element * scale."""
    if scales.dtype != torch.uint8 or elems.dtype != torch.uint8 or not elems.is_cuda or not scales.is_cuda:
        raise RuntimeError("mx_decode: uint8 CUDA tensors only")
    if scales.device != elems.device:
        raise RuntimeError(f"mx_decode: scales ({scales.device}) and elems ({elems.device}) must be on the same device")
    count, block_size = int(count), int(block_size)
    if block_size <= 0 or count < 0 or count % block_size != 0:
        raise LoFloatError(_lib.LOF_ERR_INVALID_ARG, f"mx_decode: count ({count}) must be a non-negative multiple of block_size ({block_size})")
    if scales.numel() < count // block_size:
        raise LoFloatError(_lib.LOF_ERR_INVALID_ARG, f"mx_decode: {scales.numel()} scale bytes for {count // block_size} blocks")
    need = packed_size(element_format.total_bits, count)
    if elems.numel() < need:
        raise LoFloatError(_lib.LOF_ERR_INVALID_ARG, f"mx_decode: {elems.numel()} packed bytes, {need} needed for {count} {element_format.total_bits}-bit elements")
    L = _lib.lib()
    cfe, cfs = codec_format(element_format), codec_format(scale_format)
    out = torch.empty(int(count), dtype=torch.float32, device=elems.device)
    with torch.cuda.device(elems.device):
        st = L.lof_mx_decode_f32(scales.contiguous().data_ptr(), elems.contiguous().data_ptr(), out.data_ptr(), int(count), int(block_size),
                                 C.byref(cfs), C.byref(cfe), int(endian), _stream_ptr(elems))
    _lib.check(st, "mx_decode")
    return out


def _gemm_out(out, M: int, N: int, device, what: str) -> torch.Tensor:
    """The caller's output block (a contiguous float32 [M, N] CUDA tensor, e.g. a row slice of a larger result) or a fresh one."""
    if out is None:
        return torch.empty((M, N), dtype=torch.float32, device=device)
    if out.dtype != torch.float32 or tuple(out.shape) != (M, N) or not out.is_contiguous() or out.device != device:
        raise RuntimeError(f"{what}: out must be a contiguous float32 [{M}, {N}] tensor on {device}")
    return out


def lof_gemm_packed(A: torch.Tensor, Bt: torch.Tensor, accum_mant_bits: int, round_mode=RoundingMode.RoundToNearestEven,
                    scale_a: float = 1.0, scale_b: float = 1.0, *, k_chunk: int = 32,
                    accum_format: FloatFormatDescriptor | None = None, operand_dtype: str = "e4m3",
                    partial_round: bool = False, allow_packed: bool = True, allow_hw_accum: bool = True,
                    allow_pair: bool = True, allow_persistent: bool = True, allow_float_fold: bool = True,
                    out: torch.Tensor | None = None) -> torch.Tensor:
    """Chunked tensor-core GEMM on operands already stored in the tensor core's operand type: OCP 8-bit representations
    (uint8, operand_dtype "e4m3" / "e5m2") or bfloat16 tensors (operand_dtype is then "bf16").  A is [M, K], Bt is
    [N, K] (the [out_features, in_features] weight layout).  C = A @ Bt^T / (scale_a * scale_b), float32 [M, N]."""
    if A.dtype == torch.bfloat16 and Bt.dtype == torch.bfloat16:
        operand_dtype = "bf16"
    elif A.dtype != torch.uint8 or Bt.dtype != torch.uint8:
        raise RuntimeError("lof_gemm_packed: A and Bt must both be uint8 (8-bit representations) or both bfloat16")
    elif operand_dtype == "bf16":
        raise RuntimeError("lof_gemm_packed: operand_dtype 'bf16' needs bfloat16 tensors")
    if A.dim() != 2 or Bt.dim() != 2:
        raise RuntimeError("lof_gemm_packed: A and Bt must be 2-D")
    if A.size(1) != Bt.size(1):
        raise RuntimeError(f"lof_gemm_packed: A.cols ({A.size(1)}) != Bt.cols ({Bt.size(1)})")
    if not (A.is_cuda and Bt.is_cuda):
        raise RuntimeError("lof_gemm_packed requires CUDA tensors")
    L = _lib.lib()
    ps = CProjSpec(int(RoundingMode(int(round_mode))), 0, 0)
    denom = float(scale_a) * float(scale_b)
    alpha = 1.0 if denom == 0.0 else 1.0 / denom
    A_c, B_c = A.contiguous(), Bt.contiguous()
    M, K = A_c.shape
    N = B_c.shape[0]
    out = _gemm_out(out, M, N, A.device, "lof_gemm_packed")
    status = torch.zeros(1, dtype=torch.int32, device=A.device)
    fmt = accum_format._c() if accum_format is not None else None
    flags = ((_lib.LOF_GEMM_PARTIAL_ROUND if partial_round else 0) | (0 if allow_packed else _lib.LOF_GEMM_NO_PACKED) |
             (0 if allow_hw_accum else _lib.LOF_GEMM_NO_HW_ACCUM) | (0 if allow_pair else _lib.LOF_GEMM_NO_PAIR) |
             (0 if allow_persistent else _lib.LOF_GEMM_NO_PERSISTENT) | (0 if allow_float_fold else _lib.LOF_GEMM_NO_FLOAT_FOLD))
    with torch.cuda.device(A.device):
        st = L.lof_gemm_chunked_packed(A_c.data_ptr(), B_c.data_ptr(), out.data_ptr(), M, N, K, int(accum_mant_bits),
                                       C.byref(fmt) if fmt is not None else None, C.byref(ps), int(k_chunk), alpha,
                                       _OPERAND_DTYPES[operand_dtype], flags, status.data_ptr(), _stream_ptr(A_c))
    _lib.check(st, "lof_gemm_packed")
    lof_gemm_packed.status = status
    return out


def mx_pack_operand(y: torch.Tensor):
    """[rows, K] float32 holding MX-quantised values (virtual_mx_round's output, blocks of 32 along K) -> the operand form of
    mx_gemm: (elems uint8 [rows, K] OCP e4m3fn codes, scales uint8 [rows, K / 32] E8M0 bytes, status int32[1] on the device:
    bit 0 = some value was not on the hardware's e4m3fn grid and was rounded / saturated)."""
    if y.dtype != torch.float32 or not y.is_cuda or y.dim() != 2 or y.size(1) % 32 != 0:
        raise RuntimeError("mx_pack_operand: a 2-D float32 CUDA tensor whose last dimension is a multiple of 32")
    L = _lib.lib()
    yc = y.contiguous()
    rows, K = yc.shape
    elems = torch.empty((rows, K), dtype=torch.uint8, device=yc.device)
    scales = torch.empty((rows, K // 32), dtype=torch.uint8, device=yc.device)
    status = torch.zeros(1, dtype=torch.int32, device=yc.device)
    with torch.cuda.device(yc.device):
        st = L.lof_mx_pack_tc_f32(yc.data_ptr(), yc.numel(), elems.data_ptr(), scales.data_ptr(), status.data_ptr(), _stream_ptr(yc))
    _lib.check(st, "mx_pack_operand")
    return elems, scales, status


def mx_quantize_operand(x: torch.Tensor, element_format: FloatFormatDescriptor, scale_format: FloatFormatDescriptor | None = None,
                        round_mode=RoundingMode.RoundToNearestEven, scale_round_mode=RoundingMode.RoundToNearestEven, stoch_len: int = 0,
                        *, rng=None, first_index: int = 0):
    """virtual_mx_round(x, 32, element_format, scale_format) -- the reference's MX quantisation, any element format that is a subset of
    e4m3 x 2^k (MXFP8 e4m3, MXFP6 e3m2 / e2m3, MXFP4 e2m1) -- followed by mx_pack_operand: (elems, scales, status)."""
    if scale_format is None:
        scale_format = _default_e8m0()
    y = virtual_mx_round(x, 32, element_format, scale_format, round_mode, scale_round_mode, stoch_len, rng=rng, first_index=first_index)
    return mx_pack_operand(y)


def mx_gemm(a_elems: torch.Tensor, a_scales: torch.Tensor, bt_elems: torch.Tensor, bt_scales: torch.Tensor, accum_mant_bits: int,
            round_mode=RoundingMode.RoundToNearestEven, *, k_chunk: int = 32, accum_format: FloatFormatDescriptor | None = None,
            element: str = "e4m3", partial_round: bool = False, allow_packed: bool = True, alpha: float = 1.0,
            out: torch.Tensor | None = None) -> torch.Tensor:
    """Block-scaled tensor-core GEMM on MX storage (tcgen05.mma.kind::mxf8f6f4.block_scale): C[M, N] = alpha * A_mx @ Bt_mx^T with the
    accumulator re-rounded every k_chunk (a multiple of the MX block, 32; 0 = fp32 accumulation over all of K, one final rounding).
    a_elems [M, K] / bt_elems [N, K] uint8 element codes, a_scales [M, K/32] / bt_scales [N, K/32] uint8 E8M0 bytes -- what
    mx_quantize_operand returns.  The fused form of LoF_Linear(scaling="mx") (LoFloat/layers.py:366-377)."""
    for t, what in ((a_elems, "a_elems"), (a_scales, "a_scales"), (bt_elems, "bt_elems"), (bt_scales, "bt_scales")):
        if t.dtype != torch.uint8 or not t.is_cuda or t.dim() != 2:
            raise RuntimeError(f"mx_gemm: {what} must be a 2-D uint8 CUDA tensor")
        if t.device != a_elems.device:
            raise RuntimeError("mx_gemm: all operands must live on the same device")
    M, K = a_elems.shape
    N = bt_elems.shape[0]
    if bt_elems.shape[1] != K:
        raise RuntimeError(f"mx_gemm: a_elems.cols ({K}) != bt_elems.cols ({bt_elems.shape[1]})")
    if K % 32 != 0 or tuple(a_scales.shape) != (M, K // 32) or tuple(bt_scales.shape) != (N, K // 32):
        raise RuntimeError("mx_gemm: K must be a multiple of 32 and the scale tensors [rows, K / 32]")
    L = _lib.lib()
    ps = CProjSpec(int(RoundingMode(int(round_mode))), 0, 0)
    ae, asc, be, bsc = a_elems.contiguous(), a_scales.contiguous(), bt_elems.contiguous(), bt_scales.contiguous()
    out = _gemm_out(out, M, N, ae.device, "mx_gemm")
    fmt = accum_format._c() if accum_format is not None else None
    flags = (_lib.LOF_GEMM_PARTIAL_ROUND if partial_round else 0) | (0 if allow_packed else _lib.LOF_GEMM_NO_PACKED)
    with torch.cuda.device(ae.device):
        st = L.lof_mx_gemm(ae.data_ptr(), asc.data_ptr(), be.data_ptr(), bsc.data_ptr(), out.data_ptr(), M, N, K,
                           _OPERAND_DTYPES[element], int(accum_mant_bits), C.byref(fmt) if fmt is not None else None, C.byref(ps),
                           int(k_chunk), float(alpha), flags, _stream_ptr(ae))
    _lib.check(st, "mx_gemm")
    return out
