#!/usr/bin/env bash
# 2 GPUs: bench.py under torchrun (headline sweep + C3 / C4 / C5 on the 1x2 grid), host<->device copy ceiling on both GPUs
set -u
OUT=gpurun_out/r02z; mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo_n2.txt 2>&1
timeout 200 python scripts/probe_pcie.py 2 > $OUT/probe_pcie_n2.json 2> $OUT/probe_pcie_n2.err; cat $OUT/probe_pcie_n2.json
(time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29529 bench.py --gpus 2 --steps 10 --warmup 3) > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "bench n2 exit $?"; tail -5 $OUT/bench_n2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02z/bench_n2.json").read().strip().splitlines()[-1])
print(json.dumps({k:v for k,v in d.items() if k not in ("other","workloads","sweep_GBps","kernels")})[:1200])
w=d.get("workloads",{})
for k,v in w.items():
    if isinstance(v,dict): print(k, {kk:vv for kk,vv in v.items() if kk in ("value","unit","ms","grid","comm","error","e2e","mx_passes")})
PY
