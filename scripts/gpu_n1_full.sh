#!/usr/bin/env bash
# N = 1: full GPU test suite, bench with all workloads (C1 own entry, C5 on the MX path)
set -u
OUT=gpurun_out/${OUTDIR:-r02x}; mkdir -p $OUT
(time timeout 1500 python -m pytest tests -m gpu -x -q) > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/pytest_gpu.log
(time timeout 900 python bench.py --impl reference --steps 2 --warmup 1) > $OUT/bench_reference_n1.json 2> $OUT/bench_reference_n1.err; echo "reference arm exit $?"
(time timeout 900 python bench.py --steps 10 --warmup 3) > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "bench exit $?"; tail -3 $OUT/bench_n1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/"+__import__("os").environ.get("OUTDIR","r02x")+"/bench_n1.json").read().strip().splitlines()[-1])
print(json.dumps({k:v for k,v in d.items() if k not in ("other","workloads","sweep_GBps","kernels")})[:1200])
for k,v in d.get("workloads",{}).items():
    print(k, json.dumps(v)[:1500])
PY
