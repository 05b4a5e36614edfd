"""lofloat_b200 -- B200-native (sm_100a) implementation of LoFloat's virtual-rounding hot path.

Drop-in for the names the reference package exports for this path
(/root/reference/LoFloat/__init__.py:1-3): `virtual_round`, `virtual_mx_round`, `lof_gemm`,
`RoundingMode`, `InfBehavior`, `NaNBehavior`, `Signedness`, `FloatFormatDescriptor`, the
`create_*_params` factories and `torch.ops.lofloat.*`.  Everything computes on the GPU through
liblofloat_b200.so (C ABI: include/lofloat_b200.h); there is no CPU fallback.
"""
from ._lib import LoFloatError, get_option, kernel_launch_count, set_option  # noqa: F401
from .ops import (FloatFormatDescriptor, InfBehavior, NaNBehavior, PackEndian, RoundingMode, SaturationMode, Signedness,  # noqa: F401
                  codec_format, fwht, lof_decode, lof_encode, lof_gemm, lof_conv2d, lof_gemm_batched, lof_gemm_packed, lof_linear, mx_decode, mx_encode, mx_gemm, mx_pack_operand, mx_quantize_operand, packed_size, manual_seed, set_gemm_panel, set_gemm_tensor_core, set_seed, virtual_mx_round, virtual_round, virtual_round_strict)
from .formats import (create_bfloat16_params, create_e8m0_params, create_half_params,  # noqa: F401
                      create_mx_element_params, create_p3109_params, create_single_params)
from . import _custom_ops  # noqa: F401  (registers torch.ops.lofloat.*)
from ._custom_ops import get_format, register_format  # noqa: F401
from . import sharding  # noqa: F401
from .ste import STEMXRound, STERound  # noqa: F401

__all__ = [
    "virtual_round", "virtual_mx_round", "lof_gemm", "RoundingMode", "InfBehavior", "NaNBehavior", "Signedness",
    "FloatFormatDescriptor", "create_p3109_params", "create_half_params", "create_single_params",
    "create_e8m0_params", "create_bfloat16_params", "create_mx_element_params", "register_format", "get_format",
    "manual_seed", "set_seed", "set_gemm_panel", "set_gemm_tensor_core", "set_option", "get_option", "kernel_launch_count", "LoFloatError",
    "sharding", "lof_encode", "lof_decode", "lof_gemm_packed", "packed_size", "codec_format", "SaturationMode", "PackEndian", "virtual_round_strict", "mx_encode", "mx_decode", "mx_gemm", "mx_pack_operand", "mx_quantize_operand", "fwht", "lof_gemm_batched", "lof_linear", "lof_conv2d", "STERound", "STEMXRound",
]
