"""Tiny driver for ncu captures: runs one op a few times on BASELINE-sized inputs.

    python scripts/prof_target.py round [mode] [format]     # 2^28 fp32, randn * 2^U{-12..12}
    python scripts/prof_target.py mx [variant] [format]     # 16384 x 16384 fp32, block 32
    python scripts/prof_target.py gemm [M]                  # exact-mode GEMM
    python scripts/prof_target.py gemm_tc [M] [packed|generic]   # tcgen05 chunked GEMM, k_chunk 32, accum mant 10
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import lofloat_b200 as lof  # noqa: E402

FMT = {"e4m3": (8, 4, 7), "e5m2": (8, 3, 15), "e3m2": (6, 3, 3), "e2m1": (4, 2, 1)}


def fmt(name):
    if name == "bf16like":
        return lof.create_bfloat16_params()
    k, p, b = FMT[name]
    return lof.create_p3109_params(k, p, bias=b)


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "round"
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(2024)
    if what == "round":
        mode = int(sys.argv[2]) if len(sys.argv) > 2 else 0
        f = fmt(sys.argv[3] if len(sys.argv) > 3 else "e4m3")
        n = 1 << 28
        x = torch.randn(n, device=dev, generator=g)
        x *= torch.exp2(torch.randint(-12, 13, (n,), device=dev, generator=g).float())
        for _ in range(4):
            y = lof.virtual_round(x, f, mode, 8 if mode in (3, 8, 9, 14) else 0)
    elif what == "round16":                          # 16-bit containers: python scripts/prof_target.py round16 [bf16|f16]
        dt = torch.bfloat16 if (len(sys.argv) < 3 or sys.argv[2] == "bf16") else torch.float16
        n = 1 << 29
        x = (torch.randn(n, device=dev, generator=g) * torch.exp2(torch.randint(-8, 6, (n,), device=dev, generator=g).float())).to(dt)
        for _ in range(4):
            y = lof.virtual_round(x, fmt("e4m3"), 0)
        y = y.float()
    elif what == "mx":
        variant = int(sys.argv[2]) if len(sys.argv) > 2 else 0
        f = fmt(sys.argv[3] if len(sys.argv) > 3 else "e4m3")
        lof.set_option("mx.variant", variant)
        x = torch.randn(16384, 16384, device=dev, generator=g)
        x *= torch.exp2(torch.randint(-8, 9, (16384, 1), device=dev, generator=g).float())
        for _ in range(4):
            y = lof.virtual_mx_round(x, 32, f, lof.create_e8m0_params())
    elif what == "newops":
        from lofloat_b200 import formats
        e4m3, e8m0 = fmt("e4m3"), lof.create_e8m0_params()
        d8 = lof.FloatFormatDescriptor(8, 0, 127, lof.InfBehavior.Saturating, lof.NaNBehavior._3109, lof.Signedness.Unsigned,
                                       formats._P3109Inf(8, False, True), formats._BitPatternNaN(0xFF))
        x = torch.randn(16384, 16384, device=dev, generator=g)
        x *= torch.exp2(torch.randint(-8, 9, (16384, 1), device=dev, generator=g).float())
        xb = x.to(torch.bfloat16)
        for _ in range(2):
            y = lof.virtual_round_strict(x, e4m3)                                   # round_strict_kernel
            p = lof.lof_encode(x, e4m3)                                             # encode_kernel
            y = lof.lof_decode(p, x.numel(), e4m3)                                  # decode_kernel
            y = lof.fwht(x, 128)                                                    # fwht_kernel (plain)
            y = lof.fwht(x, 128, params=e4m3, mx_block_size=32, scale_format=e8m0)  # fwht_kernel (fused MX)
            sc, el = lof.mx_encode(x, 32, e4m3, d8)                                 # mx_encode_kernel
            y = lof.mx_decode(sc, el, x.numel(), 32, e4m3, d8)                      # mx_decode_kernel
            yb = lof.virtual_round(xb, e4m3)                                        # round_vec_kernel<bf16>
            y = yb.float()
    elif what == "mxenc":
        from lofloat_b200 import formats
        d8 = lof.FloatFormatDescriptor(8, 0, 127, lof.InfBehavior.Saturating, lof.NaNBehavior._3109, lof.Signedness.Unsigned,
                                       formats._P3109Inf(8, False, True), formats._BitPatternNaN(0xFF))
        x = torch.randn(16384, 16384, device=dev, generator=g)
        x *= torch.exp2(torch.randint(-8, 9, (16384, 1), device=dev, generator=g).float())
        for _ in range(3):
            sc, el = lof.mx_encode(x, 32, fmt("e4m3"), d8)
            y = lof.mx_decode(sc, el, x.numel(), 32, fmt("e4m3"), d8)
    elif what == "pack":
        f = fmt(sys.argv[2] if len(sys.argv) > 2 else "e4m3")
        n = 1 << 28
        x = torch.randn(n, device=dev, generator=g)
        x *= torch.exp2(torch.randint(-6, 7, (n,), device=dev, generator=g).float())
        for _ in range(3):
            p = lof.lof_encode(x, f)
            y = lof.lof_decode(p, n, f)
    elif what == "gemm_tc":
        import numpy as np
        M = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
        variant = sys.argv[3] if len(sys.argv) > 3 else "packed"
        kw = {"packed": dict(k_chunk=32, partial_round=True), "hw": dict(k_chunk=32),
              "generic": dict(k_chunk=32, allow_hw_accum=False)}[variant]
        e4m3 = fmt("e4m3")
        A = lof.virtual_round(torch.rand(M, M, device=dev, generator=g) * 2 - 1, e4m3)
        B = lof.virtual_round(torch.rand(M, M, device=dev, generator=g) * 2 - 1, e4m3)
        for _ in range(3):
            y = lof.lof_gemm(A, B, 10, check=False, **kw)
    elif what == "gemm_fold":
        # the persistent per-chunk fold kernel (lof_gemm_tc3.cuh) on operands already stored as e4m3 bytes / bf16 words
        M = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
        variant = sys.argv[3] if len(sys.argv) > 3 else "bf162"
        e4m3 = fmt("e4m3")
        A = lof.virtual_round(torch.rand(M, M, device=dev, generator=g) * 2 - 1, e4m3)
        B = lof.virtual_round(torch.rand(M, M, device=dev, generator=g) * 2 - 1, e4m3)
        if variant == "bf16ops":
            a, bt = A.bfloat16(), B.t().contiguous().bfloat16()
        else:
            a, bt = lof.lof_encode(A, e4m3).view(M, M), lof.lof_encode(B.t().contiguous(), e4m3).view(M, M)
        mant, kw = {"bf162": (7, dict(k_chunk=32, partial_round=True)), "half2": (10, dict(k_chunk=32, partial_round=True)),
                    "bf16ops": (7, dict(k_chunk=32, partial_round=True)), "generic": (7, dict(k_chunk=32))}[variant]
        for _ in range(3):
            y = lof.lof_gemm_packed(a, bt, mant, **kw)
    elif what == "mx_gemm":
        # block-scaled MX GEMM (lof_gemm_mx.cu): python scripts/prof_target.py mx_gemm [M] [fold|onechunk|k128]
        M = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
        variant = sys.argv[3] if len(sys.argv) > 3 else "fold"
        e4m3, e8m0 = fmt("e4m3"), lof.create_e8m0_params()
        ae, asc, _ = lof.mx_quantize_operand(torch.randn(M, M, device=dev, generator=g), e4m3, e8m0)
        be, bsc, _ = lof.mx_quantize_operand(torch.randn(M, M, device=dev, generator=g), e4m3, e8m0)
        kw = {"fold": dict(k_chunk=32, partial_round=True), "k128": dict(k_chunk=128, partial_round=True), "onechunk": dict(k_chunk=0)}[variant]
        for _ in range(3):
            y = lof.mx_gemm(ae, asc, be, bsc, 7, **kw)
    else:
        M = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
        A = torch.rand(M, M, device=dev, generator=g) * 2 - 1
        B = torch.rand(M, M, device=dev, generator=g) * 2 - 1
        for _ in range(2):
            y = lof.lof_gemm(A, B, 10)
    torch.cuda.synchronize()
    print("done", float(y.flatten()[0]))


if __name__ == "__main__":
    main()
