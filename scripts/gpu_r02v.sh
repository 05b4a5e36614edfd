#!/usr/bin/env bash
# fp32 checked magic kinds on running extrema: exhaustive fast == exact, rounding parity tests, sweep numbers
set -u
OUT=gpurun_out/r02v; mkdir -p $OUT
(time timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "round or fast_kernel or stochastic or container") > $OUT/pytest_round.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/pytest_round.log | cut -c1-300
(time timeout 900 python bench.py --steps 10 --warmup 3 --no-side) > $OUT/bench_n1_noside.json 2> $OUT/bench_n1_noside.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02v/bench_n1_noside.json").read().strip().splitlines()[-1])
print({k:round(v) for k,v in d["sweep_GBps"].items()}); print(d["value"], d["roofline"]["frac"], d["e2e"]["value"])
PY
