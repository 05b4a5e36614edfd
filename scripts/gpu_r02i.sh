#!/usr/bin/env bash
# exhaustive <= 8-bit sweep: every fp32 pattern x 7 formats x 9 deterministic modes, CUDA path vs the compiled reference on the host cores
set -u
OUT=gpurun_out/r02i; mkdir -p $OUT
nproc > $OUT/nproc.txt
(time timeout 2400 python scripts/exhaustive_sweep.py --log $OUT/exhaustive_sweep.log) > $OUT/exhaustive_sweep.out 2>&1; echo "sweep exit $?"
tail -70 $OUT/exhaustive_sweep.log
