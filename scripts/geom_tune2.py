"""Finer residency sweep (CTAs per SM) for the rounding and MX kernels; each point measured 3 times."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import lofloat_b200 as lof  # noqa: E402

dev = torch.device("cuda:0")
n = 1 << 28
g = torch.Generator(device=dev).manual_seed(2024)
x = torch.randn(n, device=dev, generator=g) * torch.exp2(torch.randint(-12, 13, (n,), device=dev, generator=g).float())
x2 = x.view(16384, 16384)
e4m3, e8m0 = lof.create_p3109_params(8, 4, bias=7), lof.create_e8m0_params()


def timeit(fn):
    fn()
    res = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            fn()
        b.record()
        torch.cuda.synchronize()
        res.append(8.0 * n * 10 / (a.elapsed_time(b) * 1e-3) / 1e9)
    return " ".join(f"{r:6.0f}" for r in res)


for ctas in (2, 3, 4, 5, 6, 8):
    lof.set_option("mx.ctas_per_sm", ctas)
    print(f"mx ctas {ctas}: {timeit(lambda: lof.virtual_mx_round(x2, 32, e4m3, e8m0))}", flush=True)
for mode, sl, name in ((0, 0, "RNE"), (3, 8, "SR-A8"), (12, 0, "RTO")):
    for unroll in (4, 8):
        for ctas in (2, 3, 4, 5, 6, 8):
            lof.set_option("round.unroll", unroll)
            lof.set_option("round.ctas_per_sm", ctas)
            print(f"round {name} unroll {unroll} ctas {ctas}: {timeit(lambda: lof.virtual_round(x, e4m3, mode, sl))}", flush=True)
