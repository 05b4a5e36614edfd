#!/usr/bin/env bash
# compute-sanitizer memcheck over the tests that drive the round-2 kernels on small / ragged inputs
set -u
OUT=gpurun_out/${OUTDIR:-r02an}; mkdir -p $OUT
(time timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pack_codec_bit_exact or fused_hadamard or batched or mx_storage") > $OUT/memcheck.log 2>&1; echo "memcheck exit $?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" $OUT/memcheck.log | head -20
