#!/usr/bin/env bash
set -u
OUT=gpurun_out/${OUTDIR:-r02bb}; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mx_storage" > $OUT/pytest_mx.log 2>&1; grep -E "passed|failed|Error" $OUT/pytest_mx.log | head -8
LOF_PACK_FAST=2 python scripts/mx_storage_perf.py 2>&1 | grep "encode RNE"
