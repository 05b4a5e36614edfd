#!/usr/bin/env bash
set -u
OUT=gpurun_out/${OUTDIR:-r02au}; mkdir -p $OUT
(time timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --durations=4 -k "mx_storage or ste_backward or mx_gemm_block or mx_pack") > $OUT/pytest_sel.log 2>&1; echo "pytest exit $?"; tail -6 $OUT/pytest_sel.log | cut -c1-300
for gen in 2 3; do echo "== pack.fast = $gen"; LOF_PACK_FAST=$gen timeout 300 python scripts/mx_storage_perf.py 2>&1 | grep -v Warning | grep -v exact | tee $OUT/mx_storage_perf_gen$gen.txt; done
