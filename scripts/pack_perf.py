"""Throughput of the storage codec on 2^28 fp32 elements (GPU box): algorithmic bytes = 4 + n/8 per element."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import lofloat_b200 as lof  # noqa: E402

import os  # noqa: E402

lof.set_option("pack.fast", int(os.environ.get("LOF_PACK_FAST", "2")))     # 2 = register-lane kernels (default), 1 = second-generation kernels
dev = torch.device("cuda:0")
n = 1 << 28
g = torch.Generator(device=dev).manual_seed(1)
x = torch.randn(n, device=dev, generator=g) * torch.exp2(torch.randint(-6, 7, (n,), device=dev, generator=g).float())
for name, d in (("e4m3/8", lof.create_p3109_params(8, 4, bias=7)), ("e3m2/6", lof.create_p3109_params(6, 3, bias=3)),
                ("e2m1/4", lof.create_p3109_params(4, 2, bias=1)), ("half/16", lof.create_half_params()),
                ("p3109<5,2>/5", lof.create_p3109_params(5, 2))):
    p = lof.lof_encode(x, d)
    for what, fn in (("encode RNE", lambda: lof.lof_encode(x, d)), ("encode RoundUp (exact routine)", lambda: lof.lof_encode(x, d, 6)),
                     ("decode", lambda: lof.lof_decode(p, n, d))):
        fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            fn()
        b.record()
        torch.cuda.synchronize()
        by = (4 + d.total_bits / 8) * n * 5
        print(f"{name:14s} {what:32s} {by / (a.elapsed_time(b) * 1e-3) / 1e9:7.0f} GB/s", flush=True)
