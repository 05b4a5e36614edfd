#!/usr/bin/env bash
set -u
OUT=gpurun_out/${OUTDIR:-r02ay}; mkdir -p $OUT
NCU_FULL="ncu --set full --clock-control none --import-source on -f"
n=mx_encode_block
timeout 600 $NCU_FULL -k regex:mx_encode_block_kernel -s 1 -c 1 -o $OUT/$n python scripts/prof_target.py mxenc > $OUT/$n.log 2>&1; echo "ncu $n exit $?"
ncu -i $OUT/$n.ncu-rep --page raw --csv > $OUT/$n.raw.csv 2>/dev/null
ncu -i $OUT/$n.ncu-rep --page details --csv > $OUT/$n.details.csv 2>/dev/null
ncu -i $OUT/$n.ncu-rep --page source --csv > $OUT/$n.source.csv 2>/dev/null
python scripts/ncu_summary.py $OUT/$n.raw.csv
python scripts/ncu_source_top.py $OUT/$n.source.csv 25 > $OUT/$n.source_top.txt 2>&1; head -30 $OUT/$n.source_top.txt
grep -E "l1tex__t_sector|l1tex__data_pipe|lsu|l1tex__throughput|L1/TEX Hit|Mem Busy|Max Bandwidth|Mem Pipes" $OUT/$n.details.csv | cut -c1-300 | head -20
rm -f $OUT/$n.ncu-rep $OUT/$n.source.csv
