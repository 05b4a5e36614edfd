#!/usr/bin/env bash
# strict lanes after the NaN-code fix: parity; ncu --set full of the fused Hadamard + MX kernel and of the MX encode lane kernel
set -u
OUT=gpurun_out/r02ag; mkdir -p $OUT
(time timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --durations=6 -k "strict or ste_backward or hadamard") > $OUT/pytest_sel.log 2>&1; echo "pytest exit $?"; tail -8 $OUT/pytest_sel.log | cut -c1-300
NCU_FULL="ncu --set full --clock-control none --import-source on -f"
timeout 600 $NCU_FULL -k regex:fwht_kernel -s 1 -c 1 -o $OUT/fwht_mx python scripts/prof_target.py newops > $OUT/fwht_mx.log 2>&1; echo "ncu fwht exit $?"
timeout 600 $NCU_FULL -k regex:mx_encode_lane_kernel -s 1 -c 1 -o $OUT/mx_encode_lane python scripts/prof_target.py mxenc > $OUT/mx_encode_lane.log 2>&1; echo "ncu mxenc exit $?"
for n in fwht_mx mx_encode_lane; do
  ncu -i $OUT/$n.ncu-rep --page raw --csv > $OUT/$n.raw.csv 2>/dev/null
  ncu -i $OUT/$n.ncu-rep --page details --csv > $OUT/$n.details.csv 2>/dev/null
  ncu -i $OUT/$n.ncu-rep --page source --csv > $OUT/$n.source.csv 2>/dev/null
  python scripts/ncu_summary.py $OUT/$n.raw.csv | head -40
  python scripts/ncu_source_top.py $OUT/$n.source.csv > $OUT/$n.source_top.txt 2>&1; head -45 $OUT/$n.source_top.txt
  sz=$(stat -c %s $OUT/$n.ncu-rep); if [ $sz -gt 6000000 ]; then rm -f $OUT/$n.ncu-rep; fi
  rm -f $OUT/$n.source.csv
done
