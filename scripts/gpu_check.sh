#!/usr/bin/env bash
# Runs on the GPU box under gpurun: smoke, GPU parity tests, bench, ncu launch list.
# Usage: scripts/gpu_check.sh [quick|full]
set -u
MODE=${1:-quick}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt
lscpu | head -20 > $OUT/lscpu.txt
echo "== smoke" | tee $OUT/smoke.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" >> $OUT/smoke.log 2>&1; echo "smoke exit $?" | tee -a $OUT/smoke.log
echo "== pytest gpu"
if [ "$MODE" = quick ]; then
  timeout 1500 python -m pytest tests -m "gpu and not slow" -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log
else
  timeout 2400 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log
fi
tail -5 $OUT/pytest_gpu.log
echo "== bench"
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
cat $OUT/bench.json | head -c 3000
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "bench ref exit $?"
cat $OUT/bench_ref.json | head -c 1500
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --e2e-steps 0 > $OUT/ncu_bench.log 2>&1; echo "ncu exit $?"
tail -3 $OUT/launches.csv
if [ "${PROFILE:-0}" = 1 ]; then
  echo "== ncu full captures"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:round_vec_kernel -s 2 -c 1 -f -o $OUT/prof_round_rne \
      python scripts/prof_target.py round 0 e4m3 > $OUT/ncu_round.log 2>&1; echo "ncu round exit $?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:round_vec_kernel -s 2 -c 1 -f -o $OUT/prof_round_sra \
      python scripts/prof_target.py round 3 e4m3 > $OUT/ncu_round_sra.log 2>&1; echo "ncu round sra exit $?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:mx_shuffle_kernel -s 2 -c 1 -f -o $OUT/prof_mx_shuffle \
      python scripts/prof_target.py mx 0 e4m3 > $OUT/ncu_mx0.log 2>&1; echo "ncu mx0 exit $?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:mx_tma_kernel -s 2 -c 1 -f -o $OUT/prof_mx_tma \
      python scripts/prof_target.py mx 1 e4m3 > $OUT/ncu_mx1.log 2>&1; echo "ncu mx1 exit $?"
fi
