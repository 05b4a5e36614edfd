#!/usr/bin/env bash
# after hoisting the shared-window addresses out of the pipelines' loops (pair / fold / MX kernels): GEMM parity + timings
set -u
OUT=gpurun_out/r02q; mkdir -p $OUT
(time timeout 1200 python -m pytest tests -m gpu -x -q -k "gemm or reference or mx_") > $OUT/pytest_gemm.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/pytest_gemm.log
timeout 300 python scripts/gemm_fold_perf.py 8192 5 > $OUT/gemm_fold_perf.json 2> $OUT/gemm_fold_perf.err; echo "fold perf exit $?"; python - <<'PY'
import json
d=json.load(open("gpurun_out/r02q/gemm_fold_perf.json"))
for k,v in d["results"].items(): print(k, v)
PY
timeout 500 python scripts/mx_gemm_check.py > $OUT/mx_gemm_check.out 2> $OUT/mx_gemm_check.err; echo "check exit $?"; grep -E "^8192" $OUT/mx_gemm_check.out | cut -c1-400
python - <<'PY'
import torch, sys
sys.path.insert(0, '.')
import lofloat_b200 as lof
from lofloat_b200 import formats as _formats
dev = torch.device("cuda", 0)
n = 8192
g = torch.Generator(device=dev).manual_seed(42)
e4m3 = lof.create_p3109_params(8, 4, bias=7)
e5m10 = lof.FloatFormatDescriptor(16, 10, 15, lof.InfBehavior.Saturating, lof.NaNBehavior._3109, lof.Signedness.Signed, lof.create_half_params().is_inf_checker, lof.create_half_params().is_nan_checker)
ocp = lof.FloatFormatDescriptor(8, 3, 7, lof.InfBehavior.Saturating, lof.NaNBehavior._3109, lof.Signedness.Signed, _formats._P3109Inf(8, True, True), _formats._BitPatternNaN(0xFF))
A = lof.virtual_round(torch.rand(n, n, device=dev, generator=g) * 2 - 1, e4m3)
B = lof.virtual_round(torch.rand(n, n, device=dev, generator=g) * 2 - 1, e4m3)
Ap = lof.lof_encode(A, ocp).view(n, n); Btp = lof.lof_encode(B.t().contiguous(), ocp).view(n, n)
for _ in range(3): lof.lof_gemm_packed(Ap, Btp, 0, k_chunk=32, accum_format=e5m10)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20): lof.lof_gemm_packed(Ap, Btp, 0, k_chunk=32, accum_format=e5m10)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 20
print(f"C4 pair kernel through lof_gemm_packed: {ms:.4f} ms = {2.0 * n ** 3 / ms / 1e9:.1f} TFLOP/s")
PY
