"""Launch-geometry sweep of the fp32 rounding kernel (unroll x CTAs per SM) on the C2 tensor; prints GB/s."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import lofloat_b200 as lof  # noqa: E402

dev = torch.device("cuda:0")
n = 1 << 28
g = torch.Generator(device=dev).manual_seed(2024)
x = torch.randn(n, device=dev, generator=g) * torch.exp2(torch.randint(-12, 13, (n,), device=dev, generator=g).float())
e4m3 = lof.create_p3109_params(8, 4, bias=7)
y = torch.empty_like(x)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    y.copy_(x)
b.record()
torch.cuda.synchronize()
print(f"torch copy_: {8.0 * n * 10 / (a.elapsed_time(b) * 1e-3) / 1e9:.0f} GB/s")
for mode, sl, name in ((0, 0, "RNE"), (3, 8, "SR-A8")):
    for unroll in (2, 4, 8):
        for ctas in (4, 8, 12, 16, 32):
            lof.set_option("round.unroll", unroll)
            lof.set_option("round.ctas_per_sm", ctas)
            lof.virtual_round(x, e4m3, mode, sl)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10):
                lof.virtual_round(x, e4m3, mode, sl)
            b.record()
            torch.cuda.synchronize()
            print(f"{name:6s} unroll {unroll} ctas/SM {ctas:2d}: {8.0 * n * 10 / (a.elapsed_time(b) * 1e-3) / 1e9:6.0f} GB/s", flush=True)
