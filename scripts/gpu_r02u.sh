#!/usr/bin/env bash
# stochastic accumulator rounding in the chunked GEMM, Hadamard blocks up to 32768, full GPU suite
set -u
OUT=gpurun_out/r02u; mkdir -p $OUT
(time timeout 1500 python -m pytest tests -m gpu -x -q) > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -6 $OUT/pytest_gpu.log | cut -c1-400
python - <<'PY'
import torch, sys
sys.path.insert(0, '.')
import lofloat_b200 as lof
dev = torch.device("cuda", 0)
n = 8192
e4m3 = lof.create_p3109_params(8, 4, bias=7)
A = lof.virtual_round(torch.rand(n, n, device=dev) * 2 - 1, e4m3); B = lof.virtual_round(torch.rand(n, n, device=dev) * 2 - 1, e4m3)
for label, kw in (("SR-A6 k32 single", dict(round_mode=3, stochastic_rounding_bits=6, k_chunk=32)), ("TRUE_SR k128 single", dict(round_mode=10, k_chunk=128)),
                  ("SR-A6 k32 partial", dict(round_mode=3, stochastic_rounding_bits=6, k_chunk=32, partial_round=True))):
    f = lambda: lof.lof_gemm(A, B, 7, check=False, rng=(1, 2), **kw)
    f(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); f(); f(); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 2
    print(f"chunked GEMM 8192^3 stochastic accumulator {label}: {ms:.2f} ms = {2.0 * n ** 3 / ms / 1e9:.1f} TFLOP/s")
x = torch.randn(16384, 16384, device=dev)
for blk in (8192, 16384):
    f = lambda: lof.fwht(x, blk)
    f(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); f(); f(); f(); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    print(f"fwht block {blk} on 16384^2: {ms:.3f} ms = {8.0 * x.numel() / ms / 1e6:.0f} GB/s algorithmic")
PY
