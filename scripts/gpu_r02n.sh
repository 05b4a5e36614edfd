#!/usr/bin/env bash
# bf16-container rounding after the single out-of-line rare path: parity tests of the 16-bit containers + throughput + N = 1 bench
set -u
OUT=gpurun_out/r02n; mkdir -p $OUT
(time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "container or stochastic or round_params_all") > $OUT/pytest_round.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/pytest_round.log
timeout 300 python scripts/bf16_perf.py 2>&1 | tee $OUT/bf16_perf.txt | grep -E "bf16 container|OK|FAILED"
(time timeout 900 python bench.py --steps 10 --warmup 3) > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "bench exit $?"; tail -3 $OUT/bench_n1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02n/bench_n1.json").read().strip().splitlines()[-1])
for k,v in d["other"].items():
    if k.startswith("reference_gpu") or "bf16" in k: print(k, json.dumps(v)[:420])
print(d["cpu_baseline"])
PY
