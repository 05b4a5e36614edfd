"""Row f-3 measurement: implicit-GEMM lof_conv2d vs the reference's composition (F.unfold + one lof_gemm per image,
LoFloat/layers.py:710-745) on a ResNet-style layer; lof_linear vs lof_gemm on a transposed copy + bias add."""
import json
import sys
from pathlib import Path

import torch
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import lofloat_b200 as lof  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(5)
e4m3 = lof.create_p3109_params(8, 4, bias=7)


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


res = {}
B, Ci, H, W, Co, R = 32, 64, 56, 56, 64, 3
x = lof.virtual_round(torch.randn(B, Ci, H, W, device=dev, generator=g), e4m3)
w = lof.virtual_round(torch.randn(Co, Ci, R, R, device=dev, generator=g) * 0.1, e4m3)
bias = torch.randn(Co, device=dev, generator=g)
flops = 2.0 * B * Co * H * W * Ci * R * R


def ref_style():
    cols = F.unfold(x, (R, R), padding=1)
    wc = w.view(Co, -1)
    out = torch.stack([lof.lof_gemm(wc, cols[b].contiguous(), 10, kc=0) for b in range(B)]).view(B, Co, H, W)
    return out + bias.view(1, -1, 1, 1)


def batched_style():
    cols = F.unfold(x, (R, R), padding=1)
    out = lof.lof_gemm_batched(w.view(1, Co, -1).expand(B, -1, -1), cols, 10, kc=0).view(B, Co, H, W)
    return out + bias.view(1, -1, 1, 1)


def fused():
    return lof.lof_conv2d(x, w, 1, 1, 1, 1, 1, 1, 10, k_order="crs", bias=bias)


assert torch.equal(ref_style(), fused()) and torch.equal(batched_style(), fused())
for name, fn in (("unfold_plus_per_image_lof_gemm_loop", ref_style), ("unfold_plus_lof_gemm_batched", batched_style), ("lof_conv2d_implicit", fused)):
    ms = timeit(fn)
    res[f"conv3x3_B32_C64_56x56_{name}"] = {"ms": ms, "TFLOPs": flops / ms / 1e9}

M, K, N = 4096, 4096, 4096
xa = torch.rand(M, K, device=dev, generator=g) * 2 - 1
wt = torch.rand(N, K, device=dev, generator=g) * 2 - 1
bb = torch.randn(N, device=dev, generator=g)
assert torch.equal(lof.lof_linear(xa, wt, bb, 10), lof.lof_gemm(xa, wt.t().contiguous(), 10) + bb)
res["linear_4096^3_lof_gemm_transposed_copy_plus_bias"] = {"ms": timeit(lambda: lof.lof_gemm(xa, wt.t().contiguous(), 10) + bb, 3)}
res["linear_4096^3_lof_linear_fused"] = {"ms": timeit(lambda: lof.lof_linear(xa, wt, bb, 10), 3)}
A2 = torch.rand(2048, 2048, device=dev, generator=g)
ms = timeit(lambda: lof.lof_gemm(A2, A2, 10), 3)
res["lof_gemm_exact_2048^3"] = {"ms": ms, "TFLOPs": 2.0 * 2048 ** 3 / ms / 1e9}
for fast in (0, 2, 1):
    lof.set_option("gemm.fast", fast)
    for mode, name in ((0, "RNE"), (1, "RTZ"), (3, "SR_A")):
        ms = timeit(lambda: lof.lof_gemm(A2, A2, 10, mode, 4 if mode == 3 else 0), 3)
        res[f"lof_gemm_exact_2048^3_{name}_gemm.fast={fast}"] = {"ms": ms, "TFLOPs": 2.0 * 2048 ** 3 / ms / 1e9}
print(json.dumps(res, indent=1))
