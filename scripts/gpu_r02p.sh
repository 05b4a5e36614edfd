#!/usr/bin/env bash
# ncu evidence for round 2: launch list of the bench command, --set full of the MX GEMM kernel (per-block fold and fp32 accumulation) and of the C1-sized rounding launch
set -u
OUT=gpurun_out/r02p; mkdir -p $OUT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_bench.csv python bench.py --steps 2 --warmup 1 --e2e-steps 0 > $OUT/bench_under_ncu.log 2>&1; echo "launch list exit $?"
for v in fold onechunk; do
  timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:gemm_mx_kernel -s 1 -c 1 -o $OUT/mx_gemm_$v python scripts/prof_target.py mx_gemm 8192 $v > $OUT/ncu_mx_$v.log 2>&1; echo "ncu $v exit $?"
  ncu -i $OUT/mx_gemm_$v.ncu-rep --page raw --csv > $OUT/mx_gemm_$v.raw.csv 2>/dev/null
  ncu -i $OUT/mx_gemm_$v.ncu-rep --page details --csv > $OUT/mx_gemm_$v.details.csv 2>/dev/null
  ncu -i $OUT/mx_gemm_$v.ncu-rep --page source --csv > $OUT/mx_gemm_$v.source.csv 2>/dev/null
  sz=$(stat -c %s $OUT/mx_gemm_$v.ncu-rep 2>/dev/null || echo 0); if [ "$sz" -gt 15000000 ]; then rm -f $OUT/mx_gemm_$v.ncu-rep; fi
done
timeout 300 ncu --set full --clock-control none -f -k regex:mx_pack_tc_kernel -s 1 -c 1 -o $OUT/mx_pack python scripts/prof_target.py mx_gemm 8192 fold > $OUT/ncu_pack.log 2>&1; echo "ncu pack exit $?"
ncu -i $OUT/mx_pack.ncu-rep --page raw --csv > $OUT/mx_pack.raw.csv 2>/dev/null; ncu -i $OUT/mx_pack.ncu-rep --page details --csv > $OUT/mx_pack.details.csv 2>/dev/null; rm -f $OUT/mx_pack.ncu-rep
python scripts/ncu_summary.py $OUT/mx_gemm_fold.raw.csv $OUT/mx_gemm_onechunk.raw.csv $OUT/mx_pack.raw.csv 2>&1 | tee $OUT/ncu_summary.txt | head -90
ls -la $OUT
