#!/usr/bin/env bash
set -u
OUT=gpurun_out/r02f; mkdir -p $OUT
timeout 300 python scripts/gemm_fold_perf.py 8192 5 > $OUT/gemm_fold_perf.json 2> $OUT/gemm_fold_perf.err; echo "fold perf exit $?"; python - <<'PY'
import json
d=json.load(open("gpurun_out/r02f/gemm_fold_perf.json"))
for k,v in d["results"].items(): print(k, v)
PY
tail -3 $OUT/gemm_fold_perf.err
(time timeout 1500 python -m pytest tests -m gpu -x -q) > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/pytest_gpu.log
(time timeout 900 python bench.py --steps 5 --warmup 3) > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "bench exit $?"; tail -3 $OUT/bench_n1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02f/bench_n1.json").read().strip().splitlines()[-1])
print(json.dumps({k:v for k,v in d.items() if k not in ("other","workloads","sweep_GBps","kernels")})[:1500])
print(json.dumps(d.get("workloads"),indent=1)[:6000])
PY
