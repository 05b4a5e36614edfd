#!/usr/bin/env bash
set -u
OUT=gpurun_out/${OUTDIR:-r02aq}; mkdir -p $OUT
(time timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_reference_layers.py -m gpu -x -q --durations=4 -k "hadamard or ste_backward or reference or layers") > $OUT/pytest_sel.log 2>&1; echo "pytest exit $?"; tail -8 $OUT/pytest_sel.log | cut -c1-300
timeout 300 python scripts/fwht_perf.py 2>&1 | grep -v Warning | tee $OUT/fwht_perf.txt
