#!/usr/bin/env bash
# Runs on the GPU box under gpurun: ncu launch list of the bench command + one `--set full` capture per hot
# kernel.  The .ncu-rep files are exported to CSV ON THE BOX (raw page + source page) and only kept when small,
# because gpurun_out/ is capped at 64 MiB.
# Usage: scripts/gpu_profile.sh <tag> [targets...]     targets: launches round_rne round_sra mx0 mx1 gemm_exact gemm_tc
set -u
TAG=${1:-r01}; shift || true
TARGETS=${*:-launches round_rne round_sra mx0 mx1}
OUT=gpurun_out/prof_$TAG
mkdir -p $OUT
NCU_FULL="ncu --set full --clock-control none --import-source on -f"
export_rep () {   # $1 = basename (without extension)
  local rep=$OUT/$1.ncu-rep
  [ -f $rep ] || { echo "no report $rep"; return; }
  ncu -i $rep --page raw --csv > $OUT/$1.raw.csv 2>/dev/null
  ncu -i $rep --page details --csv > $OUT/$1.details.csv 2>/dev/null
  ncu -i $rep --page source --csv > $OUT/$1.source.csv 2>/dev/null
  local sz=$(stat -c %s $rep)
  echo "$1.ncu-rep size $sz"
  if [ $sz -gt 6000000 ]; then rm -f $rep; fi
}
for t in $TARGETS; do
  case $t in
    launches)
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file $OUT/launches.csv \
          python bench.py --steps 2 --warmup 1 --e2e-steps 0 --no-side > $OUT/ncu_bench.log 2>&1; echo "launches exit $?" ;;
    round_rne)
      timeout 600 $NCU_FULL -k regex:round_vec_kernel -s 2 -c 1 -o $OUT/round_rne python scripts/prof_target.py round 0 e4m3 > $OUT/round_rne.log 2>&1
      echo "round_rne exit $?"; export_rep round_rne ;;
    round_bf16)
      timeout 600 $NCU_FULL -k regex:round_vec_kernel -s 2 -c 1 -o $OUT/round_bf16 python scripts/prof_target.py round16 bf16 > $OUT/round_bf16.log 2>&1
      echo "round_bf16 exit $?"; export_rep round_bf16 ;;
    round_sra)
      timeout 600 $NCU_FULL -k regex:round_vec_kernel -s 2 -c 1 -o $OUT/round_sra python scripts/prof_target.py round 3 e4m3 > $OUT/round_sra.log 2>&1
      echo "round_sra exit $?"; export_rep round_sra ;;
    mx0)
      timeout 600 $NCU_FULL -k regex:"mx_fast_kernel|mx_shuffle_kernel" -s 2 -c 1 -o $OUT/mx_shuffle python scripts/prof_target.py mx 0 e4m3 > $OUT/mx0.log 2>&1
      echo "mx0 exit $?"; export_rep mx_shuffle ;;
    mx1)
      timeout 600 $NCU_FULL -k regex:mx_tma_kernel -s 2 -c 1 -o $OUT/mx_tma python scripts/prof_target.py mx 1 e4m3 > $OUT/mx1.log 2>&1
      echo "mx1 exit $?"; export_rep mx_tma ;;
    gemm_exact)
      timeout 600 $NCU_FULL -k regex:gemm_exact_kernel -s 1 -c 1 -o $OUT/gemm_exact python scripts/prof_target.py gemm 2048 > $OUT/gemm_exact.log 2>&1
      echo "gemm_exact exit $?"; export_rep gemm_exact ;;
    gemm_tc_hw)
      timeout 600 $NCU_FULL -k regex:gemm_tc_kernel -s 2 -c 1 -o $OUT/gemm_tc_hw python scripts/prof_target.py gemm_tc 8192 hw > $OUT/gemm_tc_hw.log 2>&1
      echo "gemm_tc_hw exit $?"; export_rep gemm_tc_hw ;;
    newops)
      timeout 900 $NCU_FULL -k regex:"round_strict_kernel|encode_kernel|decode_kernel|fwht_kernel|round_vec_kernel" -s 8 -c 8 -o $OUT/newops python scripts/prof_target.py newops > $OUT/newops.log 2>&1
      echo "newops exit $?"; export_rep newops ;;
    mxenc)
      timeout 600 $NCU_FULL -k regex:"mx_encode_kernel" -s 1 -c 1 -o $OUT/mxenc python scripts/prof_target.py mxenc > $OUT/mxenc.log 2>&1
      echo "mxenc exit $?"; export_rep mxenc ;;
    pack8)
      timeout 600 $NCU_FULL -k regex:"encode_kernel|decode_kernel" -s 2 -c 2 -o $OUT/pack8 python scripts/prof_target.py pack e4m3 > $OUT/pack8.log 2>&1
      echo "pack8 exit $?"; export_rep pack8 ;;
    pack4)
      timeout 600 $NCU_FULL -k regex:"encode_kernel|decode_kernel" -s 2 -c 2 -o $OUT/pack4 python scripts/prof_target.py pack e2m1 > $OUT/pack4.log 2>&1
      echo "pack4 exit $?"; export_rep pack4 ;;
    gemm_pair)
      timeout 600 $NCU_FULL -k regex:gemm_tc_pair_kernel -s 1 -c 1 -o $OUT/gemm_pair python scripts/prof_target.py gemm_tc 8192 hw > $OUT/gemm_pair.log 2>&1
      echo "gemm_pair exit $?"; export_rep gemm_pair ;;
    gemm_launches)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/gemm_launches.csv \
          python scripts/prof_target.py gemm_tc 8192 hw > $OUT/gemm_launches.log 2>&1; echo "gemm_launches exit $?" ;;
    gemm_tc)
      timeout 600 $NCU_FULL -k regex:gemm_tc_kernel -s 1 -c 1 -o $OUT/gemm_tc python scripts/prof_target.py gemm_tc 8192 packed > $OUT/gemm_tc.log 2>&1
      echo "gemm_tc exit $?"; export_rep gemm_tc ;;
  esac
done
du -sh gpurun_out
