#!/usr/bin/env bash
# round 2, first GPU call: parity suite at HEAD, TMEM read-back probe, host<->device copy ceiling
set -u
OUT=gpurun_out/r02a; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
lscpu | head -25 > $OUT/lscpu.txt; nvidia-smi topo -m > $OUT/topo.txt 2>&1
timeout 120 scripts/build/probe_tmem_ld.bin > $OUT/probe_tmem_ld.txt 2>&1; echo "tmem probe exit $?"
timeout 200 python scripts/probe_pcie.py > $OUT/probe_pcie.json 2> $OUT/probe_pcie.err; echo "pcie exit $?"; cat $OUT/probe_pcie.json
(time timeout 1500 python -m pytest tests -m gpu -x -q) > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/pytest_gpu.log
cat $OUT/probe_tmem_ld.txt
