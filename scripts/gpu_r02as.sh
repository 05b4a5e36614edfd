#!/usr/bin/env bash
# round-2 final ncu evidence: launch list of the bench command, --set full of the register-lane codec kernels and the specialised fused Hadamard kernels
set -u
OUT=gpurun_out/${OUTDIR:-r02as}; mkdir -p $OUT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file $OUT/launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --e2e-steps 0 --no-side > $OUT/ncu_bench.log 2>&1; echo "launches exit $?"
NCU_FULL="ncu --set full --clock-control none --import-source on -f"
for pair in "encode_lane:encode_lane_kernel" "decode_lane:decode_lane_kernel" "strict_lane:round_strict_lane_kernel" "mx_decode_lane:mx_decode_lane_kernel" "fwht_mx_full:fwht_kernel"; do
  n=${pair%%:*}; k=${pair##*:}
  skip=0; if [ $n = fwht_mx_full ]; then skip=1; fi
  timeout 600 $NCU_FULL -k regex:$k -s $skip -c 1 -o $OUT/$n python scripts/prof_target.py newops > $OUT/$n.log 2>&1; echo "ncu $n exit $?"
  ncu -i $OUT/$n.ncu-rep --page raw --csv > $OUT/$n.raw.csv 2>/dev/null
  ncu -i $OUT/$n.ncu-rep --page details --csv > $OUT/$n.details.csv 2>/dev/null
  python scripts/ncu_summary.py $OUT/$n.raw.csv | grep -E "kernel|time_duration|dram__bytes|issue_active|pipe_alu|pipe_fma|inst_executed.sum|registers" 
  rm -f $OUT/$n.ncu-rep
done
