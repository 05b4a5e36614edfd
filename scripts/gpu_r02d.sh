#!/usr/bin/env bash
set -u
OUT=gpurun_out/r02d; mkdir -p $OUT
timeout 120 scripts/build/probe_fold_pipeline.bin > $OUT/probe_fold_pipeline.txt 2>&1; echo "probe exit $?"; cat $OUT/probe_fold_pipeline.txt
(time timeout 1500 python -m pytest tests -m gpu -x -q) > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/pytest_gpu.log
