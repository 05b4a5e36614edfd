"""Where does the time of the 2-D partitioned C4 GEMM (8192^3, packed e4m3 operands) go at N GPUs?  Run under torchrun.
Times (CUDA events, max over ranks): the two NCCL collectives alone, the GEMM on already assembled panels, the whole gemm_2d_packed call with
and without the own-shard-first overlap, and a P2P pull of the same panels through CUDA-IPC mapped peer tensors."""
import os, sys, json
from pathlib import Path
import torch, torch.distributed as dist
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import lofloat_b200 as lof
from lofloat_b200 import sharding as S
from lofloat_b200 import formats as _formats

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
dist.init_process_group("nccl", device_id=dev)
grid = S.GemmGrid.create(world, rank)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
M = N = K = n
e4m3 = lof.create_p3109_params(8, 4, bias=7)
e5m10 = lof.FloatFormatDescriptor(16, 10, 15, lof.InfBehavior.Saturating, lof.NaNBehavior._3109, lof.Signedness.Signed, lof.create_half_params().is_inf_checker, lof.create_half_params().is_nan_checker)
ocp = lof.FloatFormatDescriptor(8, 3, 7, lof.InfBehavior.Saturating, lof.NaNBehavior._3109, lof.Signedness.Signed, _formats._P3109Inf(8, True, True), _formats._BitPatternNaN(0xFF))
g = torch.Generator(device=dev).manual_seed(42 + rank)
A8 = lof.lof_encode(lof.virtual_round(torch.rand(M // world, K, device=dev, generator=g) * 2 - 1, e4m3), ocp).view(M // world, K)
W8 = lof.lof_encode(lof.virtual_round(torch.rand(N // world, K, device=dev, generator=g) * 2 - 1, e4m3), ocp).view(N // world, K)

def timed(fn, reps=20, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / reps], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

R, Cn, W = grid.R, grid.C, world
out = {"world": world, "grid": f"{R}x{Cn}", "n": n}
A_panel = torch.empty((M // R, K), dtype=torch.uint8, device=dev)
Bt_panel = torch.empty((N // Cn, K), dtype=torch.uint8, device=dev)
if world > 1:
    out["all_gather_A_ms"] = timed(lambda: dist.all_gather_into_tensor(A_panel, A8, group=grid.row_group))
    rows = N // W; my_col = rank // R
    n_dst = sum(1 for d in range(W) if d % Cn == my_col)
    in_split = [rows if d % Cn == my_col else 0 for d in range(W)]
    out_split = [rows if s // R == grid.c else 0 for s in range(W)]
    send = W8.unsqueeze(0).expand(n_dst, rows, K).reshape(n_dst * rows, K).contiguous()
    out["all_to_all_Bt_ms"] = timed(lambda: dist.all_to_all_single(Bt_panel, send, out_split, in_split))
    out["barrier_ms"] = timed(lambda: dist.barrier())
else:
    A_panel.copy_(A8); Bt_panel.copy_(W8)
out["gemm_on_panels_ms"] = timed(lambda: lof.lof_gemm_packed(A_panel, Bt_panel, 0, k_chunk=32, accum_format=e5m10))
out["gemm_2d_packed_overlap_ms"] = timed(lambda: S.gemm_2d_packed(A8, W8, grid, M, N, K, 0, k_chunk=32, accum_format=e5m10))
out["gemm_2d_packed_no_overlap_ms"] = timed(lambda: S.gemm_2d_packed(A8, W8, grid, M, N, K, 0, k_chunk=32, accum_format=e5m10, overlap=False))
# P2P pull through CUDA IPC: every rank publishes its shards, peers map them and copy what they need
if world > 1:
    try:
        from torch.multiprocessing.reductions import reduce_tensor
        handles = [None] * world
        dist.all_gather_object(handles, (reduce_tensor(A8), reduce_tensor(W8)))
        peers_A, peers_W = {}, {}
        for r_, (ha, hw) in enumerate(handles):
            if r_ == rank:
                peers_A[r_], peers_W[r_] = A8, W8
            else:
                peers_A[r_], peers_W[r_] = ha[0](*ha[1]), hw[0](*hw[1])
        mw, nw = M // W, N // W
        def pull():
            for j in range(Cn):
                src = grid.r * Cn + j
                A_panel[j * mw:(j + 1) * mw].copy_(peers_A[src], non_blocking=True)
            for i in range(R):
                src = grid.c * R + i
                Bt_panel[i * nw:(i + 1) * nw].copy_(peers_W[src], non_blocking=True)
        pull(); torch.cuda.synchronize()
        ok = True
        chk_A = torch.empty_like(A_panel); dist.all_gather_into_tensor(chk_A, A8, group=grid.row_group)
        ok = bool(torch.equal(chk_A, A_panel))
        out["p2p_pull_panels_ms"] = timed(pull)
        out["p2p_pull_equals_nccl"] = ok
        out["p2p_pull_plus_gemm_ms"] = timed(lambda: (pull(), lof.lof_gemm_packed(A_panel, Bt_panel, 0, k_chunk=32, accum_format=e5m10)))
    except Exception as exc:
        out["p2p_error"] = repr(exc)[:300]
dist.barrier(); dist.destroy_process_group()
if rank == 0: print(json.dumps(out))
