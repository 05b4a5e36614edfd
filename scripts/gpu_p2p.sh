#!/usr/bin/env bash
set -u
N=${1:-2}
OUT=gpurun_out/${OUTDIR:-r02z}; mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 scripts/multi_gpu_check.py > $OUT/multi_gpu_check_n$N.txt 2> $OUT/multi_gpu_check_n$N.err; echo "check exit $?"; grep -E "PASS|FAIL" $OUT/multi_gpu_check_n$N.txt; tail -5 $OUT/multi_gpu_check_n$N.err | cut -c1-300
