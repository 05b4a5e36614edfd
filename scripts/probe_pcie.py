"""Host <-> device copy ceiling of the box (what bounds bench.py's `e2e`): pinned cudaMemcpyAsync H2D only, D2H only and both
directions at once, per GPU and on all visible GPUs concurrently (one thread of this process per GPU is not needed: the copies
are asynchronous).  Usage: python scripts/probe_pcie.py [n_gpus]"""
import json
import sys
import time

import torch

n_gpus = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
SZ = 1 << 30
res = {"n_gpus": n_gpus, "bytes": SZ}
bufs = []
for d in range(n_gpus):
    torch.cuda.set_device(d)
    h_in = torch.empty(SZ, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(SZ, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(SZ, dtype=torch.uint8, device=f"cuda:{d}")
    d_out = torch.empty(SZ, dtype=torch.uint8, device=f"cuda:{d}")
    bufs.append((h_in, h_out, d_in, d_out, torch.cuda.Stream(device=d), torch.cuda.Stream(device=d)))


def run(kind, devs, reps=4):
    best = 0.0
    for _ in range(reps):
        for d in devs:
            torch.cuda.synchronize(d)
        t0 = time.perf_counter()
        for d in devs:
            h_in, h_out, d_in, d_out, s1, s2 = bufs[d]
            if kind in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if kind in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        for d in devs:
            torch.cuda.synchronize(d)
        dt = time.perf_counter() - t0
        nbytes = SZ * len(devs) * (2 if kind == "both" else 1)
        best = max(best, nbytes / dt / 1e9)
    return best


for kind in ("h2d", "d2h", "both"):
    res[f"{kind}_gpu0_GBps"] = round(run(kind, [0]), 2)
    if n_gpus > 1:
        res[f"{kind}_all{n_gpus}_aggregate_GBps"] = round(run(kind, list(range(n_gpus))), 2)
        res[f"{kind}_per_gpu_alone_GBps"] = [round(run(kind, [d], 2), 2) for d in range(n_gpus)]
print(json.dumps(res))
