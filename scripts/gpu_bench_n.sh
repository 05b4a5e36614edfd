#!/usr/bin/env bash
# bench.py under torchrun on N GPUs (the driver's launch line), result summary
set -u
N=${1:-4}
OUT=gpurun_out/${OUTDIR:-r02am}; mkdir -p $OUT
(time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29537 bench.py --gpus $N --steps 10 --warmup 3) > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "bench exit $?"; tail -4 $OUT/bench_n$N.err | cut -c1-300
python - $N $OUT <<'PY'
import json, sys
n, o = sys.argv[1], sys.argv[2]
d=json.loads(open(f"{o}/bench_n{n}.json").read().strip().splitlines()[-1])
print(json.dumps({k:v for k,v in d.items() if k not in ("other","workloads","sweep_GBps","kernels")})[:1500])
for k,v in d.get("workloads",{}).items():
    if isinstance(v,dict): print(k, {kk:vv for kk,vv in v.items() if kk in ("value","unit","ms","grid","comm","error","e2e","fp32_accumulation_one_final_rounding")})
PY
