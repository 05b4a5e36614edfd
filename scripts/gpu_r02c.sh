#!/usr/bin/env bash
# round 2, third GPU call: fold kernel after the N = 256 descriptor fix (parity + perf), conversion-pipe probe, ncu of the fold kernel
set -u
OUT=gpurun_out/r02c; mkdir -p $OUT
timeout 60 scripts/build/probe_cvt.bin > $OUT/probe_cvt.txt 2>&1; echo "cvt probe exit $?"; cat $OUT/probe_cvt.txt
(time timeout 900 python -m pytest tests -m gpu -x -q -k "gemm or reference") > $OUT/pytest_subset.log 2>&1; echo "pytest subset exit $?"; tail -5 $OUT/pytest_subset.log
timeout 600 python scripts/gemm_fold_perf.py 8192 5 > $OUT/gemm_fold_perf.json 2> $OUT/gemm_fold_perf.err; echo "fold perf exit $?"; python - <<'PY'
import json
d=json.load(open("gpurun_out/r02c/gemm_fold_perf.json"))
for k,v in d["results"].items(): print(k, v)
PY
tail -3 $OUT/gemm_fold_perf.err
timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:gemm_tc_fold_kernel -s 2 -c 1 -o $OUT/fold_bf162 python scripts/prof_target.py gemm_fold 8192 bf162 > $OUT/ncu_fold.log 2>&1; echo "ncu exit $?"
ncu -i $OUT/fold_bf162.ncu-rep --page raw --csv > $OUT/fold_bf162.raw.csv 2>/dev/null
ncu -i $OUT/fold_bf162.ncu-rep --page details --csv > $OUT/fold_bf162.details.csv 2>/dev/null
ncu -i $OUT/fold_bf162.ncu-rep --page source --csv > $OUT/fold_bf162.source.csv 2>/dev/null
ls -la $OUT; sz=$(stat -c %s $OUT/fold_bf162.ncu-rep 2>/dev/null || echo 0); if [ "$sz" -gt 20000000 ]; then rm -f $OUT/fold_bf162.ncu-rep; fi
