#!/usr/bin/env bash
# MX GEMM: new parity tests + check script
set -u
OUT=gpurun_out/r02k; mkdir -p $OUT
(time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mx_pack or mx_gemm") > $OUT/pytest_mx.log 2>&1; echo "pytest exit $?"; tail -25 $OUT/pytest_mx.log
timeout 500 python scripts/mx_gemm_check.py > $OUT/mx_gemm_check.out 2> $OUT/mx_gemm_check.err; echo "check exit $?"; grep -E "^8192|^prof" $OUT/mx_gemm_check.out | cut -c1-400; tail -3 $OUT/mx_gemm_check.err
