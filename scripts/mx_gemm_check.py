#!/usr/bin/env python
"""Block-scaled MX GEMM (lof_mx_gemm) on the GPU: operand packing checked against a torch decode of the bytes, the GEMM against
the chunked-accumulation oracle on the decoded operands (small sizes) and against the bf16-word tensor-core path (large sizes),
then timings.  TEST INFRASTRUCTURE (executes oracle/)."""
import json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np, torch
import lofloat_b200 as lof
import oracle_lib as L

dev = torch.device("cuda", 0)
out = {}

def e4m3fn_table():
    t = np.zeros(256, dtype=np.float32)
    for c in range(256):
        s = -1.0 if c & 0x80 else 1.0
        e, m = (c >> 3) & 15, c & 7
        if e == 15 and m == 7: v = np.nan
        elif e == 0: v = m * 2.0 ** -9
        else: v = (1 + m / 8.0) * 2.0 ** (e - 7)
        t[c] = s * v
    return torch.from_numpy(t).to(dev)
TAB = e4m3fn_table()

def decode(elems, scales):
    v = TAB[elems.long()]
    sc = torch.exp2(scales.float() - 127.0)
    return (v.view(elems.size(0), -1, 32) * sc.unsqueeze(-1)).view(elems.shape)

e8 = lof.create_e8m0_params()
fmts = {"e4m3": lof.create_p3109_params(8, 4, bias=7), "e3m2": lof.create_p3109_params(6, 3, bias=3),
        "e2m3": lof.create_p3109_params(6, 4, bias=1), "e2m1": lof.create_p3109_params(4, 2, bias=1)}
g = torch.Generator(device=dev).manual_seed(5)

# 1. packing: decode(pack(mx_round(x))) == mx_round(x) wherever nothing was snapped
for name, f in fmts.items():
    x = torch.randn(512, 1024, device=dev, generator=g) * torch.exp2(torch.randint(-6, 7, (512, 1), device=dev, generator=g).float())
    y = lof.virtual_mx_round(x, 32, f, e8)
    el, sc, st = lof.mx_pack_operand(y)
    d = decode(el, sc)
    neq = int((d != y).sum().item())
    out[f"pack_{name}"] = {"snapped_flag": int(st.item()), "elements_differing_from_mx_round": neq, "of": y.numel()}
    print("pack", name, out[f"pack_{name}"], flush=True)

# 2. GEMM vs the oracle on the decoded operands
def check(M, N, K, name, k_chunk, mant, partial, mode=0):
    f = fmts[name]
    A = torch.randn(M, K, device=dev, generator=g); Bt = torch.randn(N, K, device=dev, generator=g)
    ae, asc, _ = lof.mx_quantize_operand(A, f); be, bsc, _ = lof.mx_quantize_operand(Bt, f)
    Ad, Bd = decode(ae, asc), decode(be, bsc)
    Cg = lof.mx_gemm(ae, asc, be, bsc, mant, mode, k_chunk=k_chunk, partial_round=partial)
    want = L.o_gemm_chunked(Ad.cpu().numpy(), Bd.t().contiguous().cpu().numpy(), mant, k_chunk if k_chunk > 0 else K, mode=mode, partial_round=partial)
    got = Cg.cpu().numpy()
    same = int((got.view(np.uint32) == want.view(np.uint32)).sum())
    ulp = np.abs(got - want) / np.maximum(np.abs(want) * 2.0 ** -mant, 1e-30)
    r = {"bit_equal": same, "of": got.size, "max_err_in_ulp_of_accumulator": float(ulp.max())}
    # against the bf16-word tensor-core path on the same decoded operands (same tensor-core arithmetic?)
    Cb = lof.lof_gemm_packed(Ad.bfloat16(), Bd.bfloat16(), mant, mode, k_chunk=(k_chunk if k_chunk > 0 else K), partial_round=partial)
    r["bit_equal_to_bf16_word_path"] = int((Cb == Cg).sum().item())
    out[f"gemm_{name}_{M}x{N}x{K}_k{k_chunk}_m{mant}_{'S2' if partial else 'S1'}_mode{mode}"] = r
    print("gemm", name, M, N, K, k_chunk, mant, partial, mode, r, flush=True)

for name in ("e4m3", "e2m1", "e3m2"):
    check(256, 480, 256, name, 32, 7, True)
check(200, 300, 160, "e4m3", 32, 7, True)       # ragged M / N, K not a multiple of 128
check(256, 320, 512, "e4m3", 64, 10, False)     # generic fold, single rounding
check(256, 320, 512, "e4m3", 128, 4, True, mode=1)
check(128, 224, 256, "e4m3", 0, 7, False)       # one chunk: plain MXFP8 GEMM + one final rounding
check(128, 224, 256, "e4m3", 0, 23, False)      # ... no rounding at all: the tensor core's fp32 result vs the correctly rounded sum

# 3. timing
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

from lofloat_b200 import _lib
for n in (8192,):
    A = torch.randn(n, n, device=dev, generator=g); Bt = torch.randn(n, n, device=dev, generator=g)
    ae, asc, _ = lof.mx_quantize_operand(A, fmts["e4m3"]); be, bsc, _ = lof.mx_quantize_operand(Bt, fmts["e4m3"])
    Ad16, Bd16 = decode(ae, asc).bfloat16(), decode(be, bsc).bfloat16()
    def fw16():                                  # option gemm.fold_split = 2: 16 fold warps x 56 columns instead of 8 x 112
        _lib.set_option("gemm.fold_split", 2)
        r = lof.mx_gemm(ae, asc, be, bsc, 7, k_chunk=32, partial_round=True)
        _lib.set_option("gemm.fold_split", 1)
        return r
    for label, fn in (("mx_k32_bf16acc_partial", lambda: lof.mx_gemm(ae, asc, be, bsc, 7, k_chunk=32, partial_round=True)),
                      ("mx_k32_bf16acc_partial_16foldwarps", fw16),
                      ("mx_k64_bf16acc_partial", lambda: lof.mx_gemm(ae, asc, be, bsc, 7, k_chunk=64, partial_round=True)),
                      ("mx_k128_bf16acc_partial", lambda: lof.mx_gemm(ae, asc, be, bsc, 7, k_chunk=128, partial_round=True)),
                      ("mx_onechunk_fp32acc", lambda: lof.mx_gemm(ae, asc, be, bsc, 7, k_chunk=0)),
                      ("mx_k32_mant10_generic", lambda: lof.mx_gemm(ae, asc, be, bsc, 10, k_chunk=32)),
                      ("bf16words_k32_bf16acc_partial", lambda: lof.lof_gemm_packed(Ad16, Bd16, 7, k_chunk=32, partial_round=True)),
                      ("quantize_A (mx_round + pack)", lambda: lof.mx_quantize_operand(A, fmts["e4m3"]))):
        ms = timed(fn)
        out[f"time_{n}_{label}"] = {"ms": ms, "TFLOPs": 2.0 * n ** 3 / ms / 1e9}
        print(n, label, round(ms, 4), "ms", round(2.0 * n ** 3 / ms / 1e9, 1), "TFLOP/s", flush=True)
    import ctypes as C
    for fs in (1, 2):
        _lib.set_option("gemm.fold_split", fs); _lib.set_option("gemm.mx_prof", 1)
        for kc in (32, 128, 0):
            lof.mx_gemm(ae, asc, be, bsc, 7, k_chunk=kc, partial_round=True)
            arr = (C.c_ulonglong * 16)()
            _lib.lib().lof_debug_mx_prof(arr)
            v = list(arr)
            ch = max(v[3], 1); fch = max(v[13], 1)
            print(f"prof fold_warps={8 if fs == 1 else 16} k_chunk={kc}: MMA thread total {v[0]} clk, {v[3]} chunks -> {v[0]/ch:.0f} clk/chunk; wait smem full {v[1]/ch:.0f}, "
                  f"wait tmem empty {v[2]/ch:.0f} | fold warp0: total {v[8]/fch:.0f} clk/chunk: wait tmem full {v[9]/fch:.0f}, read-back {v[10]/fch:.0f}, "
                  f"fold {v[11]/fch:.0f}, C store {v[12]/fch:.0f}", flush=True)
            out[f"prof_fw{8 if fs == 1 else 16}_k{kc}"] = v
        _lib.set_option("gemm.fold_split", 1); _lib.set_option("gemm.mx_prof", 0)
    same = int((lof.mx_gemm(ae, asc, be, bsc, 7, k_chunk=32, partial_round=True) == lof.lof_gemm_packed(Ad16, Bd16, 7, k_chunk=32, partial_round=True)).sum().item())
    out[f"big_{n}_bit_equal_to_bf16_word_path"] = {"equal": same, "of": n * n}
    print("8192 equal to bf16-word path:", same, "of", n * n)
print(json.dumps(out))
