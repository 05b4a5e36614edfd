#!/usr/bin/env bash
set -u
OUT=gpurun_out/${OUTDIR:-r02ao}; mkdir -p $OUT
(time timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --durations=6 -k "gemm_exact or exact_gemm or batched or conv2d or via_torch_op") > $OUT/pytest_sel.log 2>&1; echo "pytest exit $?"; tail -10 $OUT/pytest_sel.log | cut -c1-300
