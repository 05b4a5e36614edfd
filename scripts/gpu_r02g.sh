#!/usr/bin/env bash
set -u
OUT=gpurun_out/r02g; mkdir -p $OUT
(time timeout 1500 python -m pytest tests -m gpu -x -q) > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/pytest_gpu.log
timeout 300 python scripts/gemm_fold_perf.py 8192 5 > $OUT/gemm_fold_perf.json 2> $OUT/gemm_fold_perf.err; echo "fold perf exit $?"; python - <<'PY'
import json
d=json.load(open("gpurun_out/r02g/gemm_fold_perf.json"))
for k,v in d["results"].items(): print(k, v)
PY
tail -3 $OUT/gemm_fold_perf.err
(time timeout 900 python bench.py --steps 10 --warmup 3 --no-side) > $OUT/bench_n1_noside.json 2> $OUT/bench_n1_noside.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02g/bench_n1_noside.json").read().strip().splitlines()[-1])
print({k:round(v) for k,v in d["sweep_GBps"].items()}); print(d["value"], d["e2e"])
PY
timeout 300 python scripts/bf16_perf.py 2>&1 | grep -E "bf16 container|OK|FAILED"
