#!/usr/bin/env bash
# Full single-GPU validation of the current tree on the GPU box: smoke, GPU parity tests, both bench arms, ncu launch lists.
# Usage: scripts/gpu_validate.sh <tag>
set -u
TAG=${1:-r01h}
OUT=gpurun_out/val_$TAG; mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?"; tail -1 $OUT/smoke.log
(time timeout 2400 python -m pytest tests -m gpu -x -q) > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/pytest_gpu.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "bench reference exit $?"
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; tail -2 $OUT/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file $OUT/launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --e2e-steps 0 --no-side > $OUT/ncu_bench.log 2>&1; echo "launches exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_gemm.csv \
    python scripts/prof_target.py gemm_tc 8192 hw > $OUT/ncu_gemm.log 2>&1; echo "gemm launches exit $?"
du -sh gpurun_out
