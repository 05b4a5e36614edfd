"""stream.ctas_per_sm sweep for the codec / strict / MX storage / fused Hadamard kernels (2^28 elements)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import lofloat_b200 as lof  # noqa: E402
from lofloat_b200 import formats  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(7)
x = torch.randn(16384, 16384, device=dev, generator=g) * torch.exp2(torch.randint(-8, 9, (16384, 1), device=dev, generator=g).float())
n = x.numel()
e4m3, e8m0 = lof.create_p3109_params(8, 4, bias=7), lof.create_e8m0_params()
d8 = lof.FloatFormatDescriptor(8, 0, 127, lof.InfBehavior.Saturating, lof.NaNBehavior._3109, lof.Signedness.Unsigned,
                               formats._P3109Inf(8, False, True), formats._BitPatternNaN(0xFF))
p = lof.lof_encode(x, e4m3)
sc, el = lof.mx_encode(x, 32, e4m3, d8)
ops = (("strict", lambda: lof.virtual_round_strict(x, e4m3), 8.0), ("encode8", lambda: lof.lof_encode(x, e4m3), 5.0),
       ("decode8", lambda: lof.lof_decode(p, n, e4m3), 5.0), ("fwht128", lambda: lof.fwht(x, 128), 8.0),
       ("fwht128+mx", lambda: lof.fwht(x, 128, params=e4m3, mx_block_size=32, scale_format=e8m0), 8.0),
       ("mx_encode", lambda: lof.mx_encode(x, 32, e4m3, d8), 5.03), ("mx_decode", lambda: lof.mx_decode(sc, el, n, 32, e4m3, d8), 5.03))
for ctas in (2, 3, 4, 6, 8):
    lof.set_option("stream.ctas_per_sm", ctas)
    row = []
    for name, fn, bpe in ops:
        fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            fn()
        b.record()
        torch.cuda.synchronize()
        row.append(f"{name} {bpe * n * 5 / (a.elapsed_time(b) * 1e-3) / 1e9:5.0f}")
    print(f"ctas {ctas}: " + " | ".join(row), flush=True)
