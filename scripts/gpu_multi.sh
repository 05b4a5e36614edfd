#!/usr/bin/env bash
# 8 GPUs: sharded ops == single GPU (incl. the MX-storage GEMM), bench.py under torchrun (sweep + C1 / C3 / C4 / C5 on the 2x4 grid), host copy ceiling
set -u
N=${1:-8}
OUT=gpurun_out/${OUTDIR:-r02o}; mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 scripts/multi_gpu_check.py > $OUT/multi_gpu_check_n$N.txt 2> $OUT/multi_gpu_check_n$N.err; echo "check exit $?"; grep -E "PASS|FAIL" $OUT/multi_gpu_check_n$N.txt
timeout 200 python scripts/probe_pcie.py $N > $OUT/probe_pcie_n$N.json 2> $OUT/probe_pcie_n$N.err; cat $OUT/probe_pcie_n$N.json
(time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29537 bench.py --gpus $N --steps 10 --warmup 3) > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "bench exit $?"; tail -5 $OUT/bench_n$N.err
python - $N ${OUTDIR:-r02o} <<'PY'
import json, sys
n, o = sys.argv[1], sys.argv[2]
d=json.loads(open(f"gpurun_out/{o}/bench_n{n}.json").read().strip().splitlines()[-1])
print(json.dumps({k:v for k,v in d.items() if k not in ("other","workloads","sweep_GBps","kernels")})[:1500])
for k,v in d.get("workloads",{}).items():
    if isinstance(v,dict): print(k, {kk:vv for kk,vv in v.items() if kk in ("value","unit","ms","grid","comm","error","e2e","mx_passes","fp32_accumulation_one_final_rounding")})
PY
