"""Format factories with the reference's names and argument meaning
(/root/reference/LoFloat/formats.py:55,74,117,155).  The Inf/NaN checker objects are carried
for API compatibility; the rounding path never reads them (lo_float.h:2843-2939).
"""
from __future__ import annotations

from .ops import FloatFormatDescriptor, InfBehavior, NaNBehavior, Signedness


class _BitPatternNaN:
    """NaN checker described by one bit pattern (IEEE P3109 style)."""

    def __init__(self, pattern: int):
        self._pattern = pattern

    def __call__(self, bits):
        return bits == self._pattern

    def qNanBitPattern(self):
        return self._pattern

    def sNanBitPattern(self):
        return self._pattern


class _P3109Inf:
    def __init__(self, k: int, is_signed: bool, saturating: bool):
        self.k, self.is_signed, self.saturating = k, is_signed, saturating

    def __call__(self, bits):
        if self.saturating:
            return False
        target = (1 << self.k) - 1 if self.is_signed else (1 << self.k) - 2
        return (bits | (1 << (self.k - 1))) == target

    def minNegInf(self):
        return (1 << self.k) - 1 if self.is_signed else 0

    def minPosInf(self):
        return (1 << (self.k - 1)) - 1 if self.is_signed else (1 << self.k) - 2


class _IeeeInf:
    def __init__(self, exp_mask: int, mant_mask: int, sign_mask: int):
        self.exp_mask, self.mant_mask, self.sign_mask = exp_mask, mant_mask, sign_mask

    def __call__(self, bits):
        return (bits & self.exp_mask) == self.exp_mask and (bits & self.mant_mask) == 0

    def minNegInf(self):
        return self.sign_mask | self.exp_mask

    def minPosInf(self):
        return self.exp_mask


class _IeeeNaN:
    def __init__(self, exp_mask: int, mant_mask: int, qnan: int, snan: int):
        self.exp_mask, self.mant_mask, self._q, self._s = exp_mask, mant_mask, qnan, snan

    def __call__(self, bits):
        return (bits & self.exp_mask) == self.exp_mask and (bits & self.mant_mask) != 0

    def qNanBitPattern(self):
        return self._q

    def sNanBitPattern(self):
        return self._s


def create_p3109_params(k, p, is_signed=True, saturating=True, bias=None):
    """binary{k}p{p}: k total bits, p precision bits (p - 1 stored); default bias 2^(k-p-1)
    -- the P3109 bias, NOT the OCP one (reference formats.py:55-71): pass bias=7 for OCP E4M3."""
    if bias is None:
        bias = 1 << (k - p - 1)
    nan_pattern = (1 << (k - 1)) if is_signed else (1 << k) - 1
    return FloatFormatDescriptor(
        k, p - 1, bias,
        inf_behavior=InfBehavior.Saturating if saturating else InfBehavior.Extended,
        nan_behavior=NaNBehavior._3109,
        signedness=Signedness.Signed if is_signed else Signedness.Unsigned,
        is_inf_checker=_P3109Inf(k, is_signed, saturating),
        is_nan_checker=_BitPatternNaN(nan_pattern))


def create_e8m0_params():
    """E8M0 shared-scale format for MX: 8-bit unsigned, no mantissa, bias 127 (reference formats.py:74-91)."""
    return FloatFormatDescriptor(
        8, 0, 127, inf_behavior=InfBehavior.Extended, nan_behavior=NaNBehavior._3109, signedness=Signedness.Unsigned,
        is_inf_checker=_P3109Inf(8, False, False), is_nan_checker=_BitPatternNaN(0xFF))


def create_half_params():
    """IEEE binary16 (reference formats.py:117-128)."""
    return FloatFormatDescriptor(
        16, 10, 15, inf_behavior=InfBehavior.Extended, nan_behavior=NaNBehavior._3109, signedness=Signedness.Signed,
        is_inf_checker=_IeeeInf(0x7C00, 0x03FF, 0x8000), is_nan_checker=_IeeeNaN(0x7C00, 0x03FF, 0x7E00, 0x7D00))


def create_single_params():
    """IEEE binary32 (reference formats.py:155-166)."""
    return FloatFormatDescriptor(
        32, 23, 127, inf_behavior=InfBehavior.Extended, nan_behavior=NaNBehavior._3109, signedness=Signedness.Signed,
        is_inf_checker=_IeeeInf(0x7F800000, 0x007FFFFF, 0x80000000),
        is_nan_checker=_IeeeNaN(0x7F800000, 0x007FFFFF, 0x7FC00000, 0x7FA00000))


def create_bfloat16_params():
    """bfloat16-like target (8 exponent bits, 7 mantissa bits); not in the reference's factory list."""
    return FloatFormatDescriptor(
        16, 7, 127, inf_behavior=InfBehavior.Extended, nan_behavior=NaNBehavior._3109, signedness=Signedness.Signed,
        is_inf_checker=_IeeeInf(0x7F80, 0x007F, 0x8000), is_nan_checker=_IeeeNaN(0x7F80, 0x007F, 0x7FC0, 0x7FA0))


# MX element formats named in BASELINE.json (OCP biases)
def create_mx_element_params(name: str):
    table = {"e4m3": (8, 4, 7), "e5m2": (8, 3, 15), "e3m2": (6, 3, 3), "e2m3": (6, 4, 1), "e2m1": (4, 2, 1)}
    k, p, bias = table[name.lower()]
    return create_p3109_params(k, p, True, True, bias=bias)
