#!/usr/bin/env bash
# smoke + full GPU tests + bench + profiles of the rewritten kernels
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?"; tail -2 $OUT/smoke.log
(time timeout 2400 python -m pytest tests -m gpu -x -q) > $OUT/pytest_gpu_full.log 2>&1; echo "pytest exit $?"; tail -15 $OUT/pytest_gpu_full.log
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; head -c 6000 $OUT/bench.json; tail -3 $OUT/bench.err
bash scripts/gpu_profile.sh ${1:-r01b} round_rne mx0
