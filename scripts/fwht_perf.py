"""Fused Hadamard + MX rounding vs the unfused sequence (torch butterflies as in layers.py:182-216, then virtual_mx_round)."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import lofloat_b200 as lof  # noqa: E402


def torch_fwht(x, block):
    flat = x.reshape(-1, block)
    M, h = flat.shape[0], 1
    while h < block:
        v = flat.view(M, block // (2 * h), 2, h)
        flat = torch.stack([v[:, :, 0, :] + v[:, :, 1, :], v[:, :, 0, :] - v[:, :, 1, :]], dim=2).reshape(M, block)
        h *= 2
    return (flat * (1.0 / float(np.sqrt(block)))).view(x.shape)


dev = torch.device("cuda:0")
x = torch.randn(16384, 16384, device=dev)
e4m3, e8m0 = lof.create_p3109_params(8, 4, bias=7), lof.create_e8m0_params()
n = x.numel()
for block in (32, 128, 1024):
    for name, fn in ((f"fused fwht{block} + mx_round", lambda: lof.fwht(x, block, params=e4m3, mx_block_size=32, scale_format=e8m0)),
                     (f"fused fwht{block} + virtual_round e4m3 RNE", lambda: lof.fwht(x, block, params=e4m3)),
                     (f"fused fwht{block} + virtual_round e4m3 SR-A8", lambda: lof.fwht(x, block, params=e4m3, round_mode=3, stoch_len=8)),
                     (f"fused fwht{block} + virtual_round, magic add off", lambda: (lof.set_option("round.fast", 1), lof.fwht(x, block, params=e4m3), lof.set_option("round.fast", 2))),
                     (f"fused fwht{block} only", lambda: lof.fwht(x, block)),
                     (f"torch fwht{block} then lof.virtual_mx_round", lambda: lof.virtual_mx_round(torch_fwht(x, block), 32, e4m3, e8m0))):
        fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            fn()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 3
        print(f"{name:48s} {ms:8.3f} ms   {8.0 * n / (ms * 1e-3) / 1e9:7.0f} GB/s (8 B/element)", flush=True)
