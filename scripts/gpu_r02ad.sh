#!/usr/bin/env bash
# register-lane codec kernels: parity tests (codec, exhaustive fast == exact, MX storage), throughput with pack.fast = 2 vs 1, bf16-container timing
set -u
OUT=gpurun_out/r02ad; mkdir -p $OUT
(time timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --durations=8 -k "pack or mx_storage or strict") > $OUT/pytest_pack.log 2>&1; echo "pytest exit $?"; tail -14 $OUT/pytest_pack.log | cut -c1-300
for gen in 2 1; do
  echo "== pack.fast = $gen"
  LOF_PACK_FAST=$gen timeout 300 python scripts/pack_perf.py 2>&1 | grep -v Warning | tee $OUT/pack_perf_gen$gen.txt
  LOF_PACK_FAST=$gen timeout 300 python scripts/mx_storage_perf.py 2>&1 | grep -v Warning | tee $OUT/mx_storage_perf_gen$gen.txt
done
timeout 300 python scripts/bf16_perf.py 2>&1 | grep -v Warning | tee $OUT/bf16_perf.txt | head -20
