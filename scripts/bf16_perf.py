"""bf16-container virtual_round throughput (2^29 elements = 1 GiB, 4 B / element) and fast == exact on every bf16 pattern."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import lofloat_b200 as lof  # noqa: E402

dev = torch.device("cuda:0")
allbits = torch.arange(0, 65536, device=dev, dtype=torch.int32).to(torch.int16).view(torch.bfloat16).repeat(64)
bad = 0
for name, d in (("e4m3", lof.create_p3109_params(8, 4, bias=7)), ("e5m2", lof.create_p3109_params(8, 3, bias=15)), ("e2m1", lof.create_p3109_params(4, 2, bias=1)),
                ("e3m2", lof.create_p3109_params(6, 3, bias=3)), ("e8m0", lof.create_e8m0_params()), ("bf16like", lof.create_bfloat16_params()),
                ("p3109_8p4_ext", lof.create_p3109_params(8, 4, saturating=False))):
    for mode in (0, 1, 2, 4, 5, 6, 7, 12, 13):
        lof.set_option("round.fast", 0)
        a = lof.virtual_round(allbits, d, mode)
        lof.set_option("round.fast", 2)
        b = lof.virtual_round(allbits, d, mode)
        lof.set_option("round.fast", 1)
        c = lof.virtual_round(allbits, d, mode)
        lof.set_option("round.fast", 2)
        if not (torch.equal(a.view(torch.int16), b.view(torch.int16)) and torch.equal(a.view(torch.int16), c.view(torch.int16))):
            bad += 1
            print("MISMATCH", name, mode)
print("bf16 fast == exact on all 65536 patterns:", "OK" if bad == 0 else f"{bad} FAILED")
n = 1 << 29
g = torch.Generator(device=dev).manual_seed(2024)
x = (torch.randn(n // 2, device=dev, generator=g) * torch.exp2(torch.randint(-12, 13, (n // 2,), device=dev, generator=g).float())).to(torch.bfloat16).repeat(2)
for name, d, mode, sl in (("e4m3 RNE", lof.create_p3109_params(8, 4, bias=7), 0, 0), ("e2m1 RNE", lof.create_p3109_params(4, 2, bias=1), 0, 0),
                          ("e4m3 RoundToOdd", lof.create_p3109_params(8, 4, bias=7), 12, 0), ("e4m3 SR-A4 (exact kernel)", lof.create_p3109_params(8, 4, bias=7), 3, 4),
                          ("mantissa 3 RNE", 3, 0, 0)):
    fn = lambda: lof.virtual_round(x, d, mode, sl)
    fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        fn()
    b.record()
    torch.cuda.synchronize()
    print(f"bf16 container {name:28s} {4.0 * n * 5 / (a.elapsed_time(b) * 1e-3) / 1e9:7.0f} GB/s")
xh = (torch.randn(n // 2, device=dev, generator=g) * torch.exp2(torch.randint(-10, 6, (n // 2,), device=dev, generator=g).float())).to(torch.float16).repeat(2)
for name, d, mode in (("e4m3 RNE", lof.create_p3109_params(8, 4, bias=7), 0), ("e2m1 RNE", lof.create_p3109_params(4, 2, bias=1), 0),
                      ("e5m2 RNE (exact kernel: denorm_min < 2^-14)", lof.create_p3109_params(8, 3, bias=15), 0)):
    for fast in (0, 2):
        lof.set_option("round.fast", fast)
        fn = lambda: lof.virtual_round(xh, d, mode)
        fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            fn()
        b.record()
        torch.cuda.synchronize()
        print(f"fp16 container {name:44s} round.fast={fast} {4.0 * n * 5 / (a.elapsed_time(b) * 1e-3) / 1e9:7.0f} GB/s")
lof.set_option("round.fast", 2)
