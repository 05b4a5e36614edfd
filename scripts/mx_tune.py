"""mx.ctas_per_sm sweep of the second-generation MX kernel on the C3 shape."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import lofloat_b200 as lof  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(7)
x = torch.randn(16384, 16384, device=dev, generator=g) * torch.exp2(torch.randint(-8, 9, (16384, 1), device=dev, generator=g).float())
e4m3, e8m0 = lof.create_p3109_params(8, 4, bias=7), lof.create_e8m0_params()
for ctas in (2, 3, 4, 6, 8, 12, 16, 32):
    lof.set_option("mx.ctas_per_sm", ctas)
    lof.virtual_mx_round(x, 32, e4m3, e8m0)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        lof.virtual_mx_round(x, 32, e4m3, e8m0)
    b.record()
    torch.cuda.synchronize()
    print(f"mx.ctas_per_sm {ctas:2d}: {8.0 * x.numel() * 10 / (a.elapsed_time(b) * 1e-3) / 1e9:6.0f} GB/s", flush=True)
