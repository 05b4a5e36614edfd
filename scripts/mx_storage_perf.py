"""Throughput of MX quantisation to real storage (scale bytes + packed elements) on the C3 shape (GPU box)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import lofloat_b200 as lof  # noqa: E402
from lofloat_b200 import formats  # noqa: E402

import os  # noqa: E402

lof.set_option("pack.fast", int(os.environ.get("LOF_PACK_FAST", "2")))     # 2 = register-lane kernels (default), 1 = second-generation kernels
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(7)
x = torch.randn(16384, 16384, device=dev, generator=g) * torch.exp2(torch.randint(-8, 9, (16384, 1), device=dev, generator=g).float())
d8 = lof.FloatFormatDescriptor(8, 0, 127, lof.InfBehavior.Saturating, lof.NaNBehavior._3109, lof.Signedness.Unsigned,
                               formats._P3109Inf(8, False, True), formats._BitPatternNaN(0xFF))
n = x.numel()
for name, d in (("mxfp8_e4m3", lof.create_p3109_params(8, 4, bias=7)), ("mxfp6_e3m2", lof.create_p3109_params(6, 3, bias=3)),
                ("mxfp4_e2m1", lof.create_p3109_params(4, 2, bias=1))):
    sc, el = lof.mx_encode(x, 32, d, d8)
    by = (4 + d.total_bits / 8 + 1 / 32) * n
    for what, fn in (("encode RNE", lambda: lof.mx_encode(x, 32, d, d8)), ("encode RoundUp (exact routines)", lambda: lof.mx_encode(x, 32, d, d8, 6)),
                     ("decode", lambda: lof.mx_decode(sc, el, n, 32, d, d8))):
        fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            fn()
        b.record()
        torch.cuda.synchronize()
        print(f"{name:12s} {what:32s} {by * 5 / (a.elapsed_time(b) * 1e-3) / 1e9:7.0f} GB/s", flush=True)
