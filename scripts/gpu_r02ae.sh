#!/usr/bin/env bash
# register-lane codec kernels, overflow handled in line (only infinity / NaN inputs take the rare path): parity tests + throughput
set -u
OUT=gpurun_out/${OUTDIR:-r02ae}; mkdir -p $OUT
(time timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --durations=5 -k "pack or mx_storage or strict") > $OUT/pytest_pack.log 2>&1; echo "pytest exit $?"; tail -12 $OUT/pytest_pack.log | cut -c1-300
for gen in 2; do
  echo "== pack.fast = $gen"
  LOF_PACK_FAST=$gen timeout 300 python scripts/pack_perf.py 2>&1 | grep -v Warning | tee $OUT/pack_perf_gen$gen.txt
  LOF_PACK_FAST=$gen timeout 300 python scripts/mx_storage_perf.py 2>&1 | grep -v Warning | tee $OUT/mx_storage_perf_gen$gen.txt
done
