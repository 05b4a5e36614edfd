"""Straight-through-estimator wrappers of the rounding ops: the training-path callers of the hot path
(/root/reference/LoFloat/layers.py:8-57).  Same class names, argument order and gradient
rules as the reference; forward runs the CUDA ops of this package (eager objects or the int format ids of
torch.ops.lofloat.* for traced graphs).  Every forward is a single asynchronous launch on the current stream with no
allocation inside the library and no host synchronisation, so the ops can be captured in CUDA graphs."""
from __future__ import annotations

import torch

from .ops import RoundingMode, virtual_mx_round, virtual_round


class STERound(torch.autograd.Function):
    """round(scale * x) in forward (fused scale-then-round); d/dx = scale in backward (layers.py:8-31)."""

    @staticmethod
    def forward(ctx, tensor, params, rounding_mode=RoundingMode.RoundToNearestEven, scale=1.0):
        ctx.scale = float(scale)
        if isinstance(params, int) and not isinstance(params, bool):        # pre-resolved format id (traceable path)
            return torch.ops.lofloat.virtual_round(tensor, params, int(rounding_mode), 0, float(scale))
        return virtual_round(tensor, params, round_mode=rounding_mode, stoch_len=0, scale=float(scale))

    @staticmethod
    def backward(ctx, grad_output):
        if ctx.scale == 1.0:
            return grad_output, None, None, None
        return grad_output * ctx.scale, None, None, None


class STEMXRound(torch.autograd.Function):
    """MX fake-quantise along the last axis in forward; identity in backward (layers.py:34-57)."""

    @staticmethod
    def forward(ctx, tensor, element_params, scale_params, block_size, rounding_mode=RoundingMode.RoundToNearestEven):
        if isinstance(element_params, int) and not isinstance(element_params, bool):
            return torch.ops.lofloat.virtual_mx_round(tensor, int(block_size), element_params, scale_params, int(rounding_mode),
                                                      int(RoundingMode.RoundToNearestEven), 0)
        return virtual_mx_round(tensor, int(block_size), element_params, scale_params, round_mode=rounding_mode)

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output, None, None, None, None
