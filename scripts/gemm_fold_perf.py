"""Times the chunked tensor-core GEMM variants at n^3 (default 8192): the persistent per-chunk fold kernel (lof_gemm_tc3.cuh)
against the first-generation kernel, for the packed half2 / bfloat162 folds, the generic fold and bf16 operands (C5's kernel).
Usage: python scripts/gemm_fold_perf.py [n] [reps]"""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import lofloat_b200 as lof  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(42)
e4m3 = lof.create_p3109_params(8, 4, bias=7)
A = lof.virtual_round(torch.rand(n, n, device=dev, generator=g) * 2 - 1, e4m3)
B = lof.virtual_round(torch.rand(n, n, device=dev, generator=g) * 2 - 1, e4m3)
A8, Bt8 = lof.lof_encode(A, e4m3).view(n, n), lof.lof_encode(B.t().contiguous(), e4m3).view(n, n)
A16, Bt16 = A.bfloat16(), B.t().contiguous().bfloat16()
flops = 2.0 * n ** 3


def timeit(fn):
    for _ in range(2):
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
cases = {
    "e4m3_k32_mant10_partial_half2": (A8, Bt8, 10, dict(k_chunk=32, partial_round=True)),
    "e4m3_k32_mant7_partial_bf162": (A8, Bt8, 7, dict(k_chunk=32, partial_round=True)),
    "bf16ops_k32_mant7_partial_bf162": (A16, Bt16, 7, dict(k_chunk=32, partial_round=True)),
    "e4m3_k32_mant10_single_generic": (A8, Bt8, 10, dict(k_chunk=32, allow_hw_accum=False)),
    "e4m3_k32_mant7_single_generic": (A8, Bt8, 7, dict(k_chunk=32)),
    "e4m3_k32_mant7_RTZ_partial_generic": (A8, Bt8, 7, dict(k_chunk=32, partial_round=True, round_mode=1)),
    "e4m3_k128_mant7_partial_bf162": (A8, Bt8, 7, dict(k_chunk=128, partial_round=True)),
    "e4m3_k512_mant7_single_generic": (A8, Bt8, 7, dict(k_chunk=512)),
}
split_cases = ("e4m3_k32_mant7_partial_bf162", "bf16ops_k32_mant7_partial_bf162", "e4m3_k128_mant7_partial_bf162")
for name in split_cases:                                # packed fold, 128 x 256 tile: one N = 256 MMA per k-step vs two N = 128 MMAs
    a, bt, mant, kw = cases[name]
    for split in (1, 2):
        lof.set_option("gemm.fold_split", split)
        ms = timeit(lambda: lof.lof_gemm_packed(a, bt, mant, **kw))
        res[f"{name}_split{split}"] = {"ms": round(ms, 4), "TFLOPs": round(flops / (ms * 1e-3) / 1e12, 1)}
lof.set_option("gemm.fold_split", 1)
for name, (a, bt, mant, kw) in cases.items():
    for pers in (True, False):
        if not pers and "generic" in name and n > 4096 and "k512" not in name:
            continue                                   # the first-generation generic fold takes ~14 ms at 8192^3: measured in round 1
        ms = timeit(lambda: lof.lof_gemm_packed(a, bt, mant, allow_persistent=pers, **kw))
        res[f"{name}{'' if pers else '_gen1'}"] = {"ms": round(ms, 4), "TFLOPs": round(flops / (ms * 1e-3) / 1e12, 1)}
    c1 = lof.lof_gemm_packed(a, bt, mant, **kw)
    c0 = lof.lof_gemm_packed(a[:512], bt, mant, allow_persistent=False, **kw)
    res[name]["equals_gen1_first_512_rows"] = bool(torch.equal(c1[:512].view(torch.int32), c0.view(torch.int32)))
print(json.dumps({"n": n, "reps": reps, "results": res}, indent=1))
