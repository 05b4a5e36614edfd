#!/usr/bin/env bash
# strict rounding on the register lanes, fused Hadamard + rounding (FFMA butterflies, cold paths out of line): parity tests + throughput
set -u
OUT=gpurun_out/r02af; mkdir -p $OUT
(time timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --durations=6 -k "pack or mx_storage or strict or hadamard or mx_bit_exact or mx_modes or ste_backward") > $OUT/pytest_sel.log 2>&1; echo "pytest exit $?"; tail -12 $OUT/pytest_sel.log | cut -c1-300
timeout 300 python scripts/strict_perf.py 2>&1 | grep -v Warning | tee $OUT/strict_perf.txt
timeout 300 python scripts/fwht_perf.py 2>&1 | grep -v Warning | tee $OUT/fwht_perf.txt
timeout 300 python scripts/mx_storage_perf.py 2>&1 | grep -v Warning | tee $OUT/mx_storage_perf.txt
