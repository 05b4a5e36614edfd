#!/usr/bin/env bash
# round 2, second GPU call: persistent fold kernel (parity + perf), C1 / bf16 stochastic perf, reference layers on the extension
set -u
OUT=gpurun_out/r02b; mkdir -p $OUT
(time timeout 900 python -m pytest tests -m gpu -x -q -k "gemm or reference or mx_modes or round_other or bf16 or stochastic or edge_shapes or unroll") > $OUT/pytest_subset.log 2>&1; echo "pytest subset exit $?"; tail -5 $OUT/pytest_subset.log
timeout 600 python scripts/gemm_fold_perf.py 8192 5 > $OUT/gemm_fold_perf.json 2> $OUT/gemm_fold_perf.err; echo "fold perf exit $?"; cat $OUT/gemm_fold_perf.json; tail -3 $OUT/gemm_fold_perf.err
timeout 300 python scripts/bf16_perf.py > $OUT/bf16_perf.txt 2>&1; echo "bf16 perf exit $?"; grep -E "SR|RNE|OK|FAILED" $OUT/bf16_perf.txt
timeout 300 python scripts/c1_tune.py > $OUT/c1_tune.txt 2>&1; echo "c1 exit $?"; cat $OUT/c1_tune.txt
