"""C5 (BASELINE.json configs[4]): MXFP8 GEMM with bf16-rounded accumulation, 2-D output-partitioned across the GPUs of
one box with NCCL operand exchange.  Composition as LoF_Linear(scaling="mx") defines it (LoFloat/layers.py:254-257,
366-377): both operands are MX fake-quantised (E4M3 elements, E8M0 scale, block 32 along K), then multiplied with the
accumulator rounded to 7 mantissa bits -- here once per k-chunk of 32 (= one MX block) on the tensor cores
(bf16 operands: an e4m3 value times a power-of-two block scale is exactly representable in bf16).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29541 \\
        scripts/c5_mx_gemm.py [M=N=K, default 4096 * N]

Prints one JSON line from rank 0: device time (CUDA events, max over ranks) of (a) the two MX passes, (b) the operand
exchange + GEMM, and the aggregate TFLOP/s.
"""
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import lofloat_b200 as lof  # noqa: E402
from lofloat_b200 import sharding as S  # noqa: E402


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096 * world
    M = N = K = n
    grid = S.GemmGrid.create(world, rank)
    g = torch.Generator(device=dev).manual_seed(99 + rank)
    A_loc = torch.randn(M // world, K, device=dev, generator=g)
    Wt_loc = torch.randn(N // world, K, device=dev, generator=g)          # weight layout [out, in]: MX blocks run along K
    e4m3, e8m0 = lof.create_p3109_params(8, 4, bias=7), lof.create_e8m0_params()

    def sync_max(ms):
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, reps):
        for _ in range(3):                      # warm-up: kernels, and the caching allocator's blocks for the outputs
            keep = fn()
        del keep
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            out = fn()
        b.record()
        torch.cuda.synchronize()
        return sync_max(a.elapsed_time(b) / reps), out

    def mx_pass():
        return lof.virtual_mx_round(A_loc, 32, e4m3, e8m0), lof.virtual_mx_round(Wt_loc, 32, e4m3, e8m0)
    ms_mx, (A_q, Wt_q) = timed(mx_pass, 3)
    # the reference's saturating virtual_round leaves a block's overflowing maxima unrounded: snap the operands
    # to bf16 once so that the tensor-core operands are exact and both exchange paths below multiply the same numbers
    A_q, Wt_q = A_q.bfloat16().float(), Wt_q.bfloat16().float()
    Wt_keep = Wt_q
    # B = Wt^T, row-sharded the way gemm_2d expects ([K / W, N]): one all-to-all of the quantised weight (setup, not timed)
    if world > 1:
        send = torch.stack([Wt_q[:, r * (K // world):(r + 1) * (K // world)].t().contiguous() for r in range(world)])  # [W, K/W, N/W]
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send)
        B_loc = torch.cat([recv[r] for r in range(world)], dim=1).contiguous()                                       # [K/W, N]
        del send, recv
    else:
        B_loc = Wt_q.t().contiguous()
    del Wt_q

    def gemm():
        return S.gemm_2d(A_q, B_loc, grid, M, N, K, 7, k_chunk=32, operand_dtype="bf16", partial_round=True, check=False)
    ms_gemm_fp32, C_blk32 = timed(gemm, 2)
    del B_loc
    # the same product with the operands exchanged as bf16 (exact for MX-rounded values) in the weight layout: half the NVLink
    # volume, no transposes, no operand conversion pass inside the GEMM call
    Wt_q = Wt_keep

    def gemm_packed():
        return S.gemm_2d_packed(A_q.bfloat16(), Wt_q.bfloat16(), grid, M, N, K, 7, k_chunk=32, partial_round=True)
    ms_gemm, C_blk = timed(gemm_packed, 2)
    same = bool(torch.equal(C_blk, C_blk32))
    del C_blk32

    def gemm_packed_seq():
        return S.gemm_2d_packed(A_q.bfloat16(), Wt_q.bfloat16(), grid, M, N, K, 7, overlap=False, k_chunk=32, partial_round=True)
    ms_gemm_seq, C_seq = timed(gemm_packed_seq, 2)
    same = same and bool(torch.equal(C_seq, C_blk))
    del C_seq
    zero_rows_ok = True
    finite = bool(torch.isfinite(C_blk).all().item())
    flops = 2.0 * M * N * K
    if rank == 0:
        print(json.dumps({"config": "C5 MXFP8 (E4M3 / E8M0, block 32) GEMM, accumulator rounded to 7 mantissa bits every k = 32",
                          "M": M, "N": N, "K": K, "n_gpus": world, "grid": f"{grid.R}x{grid.C}",
                          "ms_mx_round_both_operands": ms_mx,
                          "mx_GBps_aggregate": 8.0 * (M * K + N * K) / (ms_mx * 1e-3) / 1e9,
                          "ms_exchange_plus_gemm": ms_gemm, "TFLOPs_aggregate": flops / (ms_gemm * 1e-3) / 1e12,
                          "exchange": "bf16 operands (cast inside the timed region), all-to-all of B^T row shards, all-gather of A row shards overlapped with the GEMM on the own shard",
                          "ms_exchange_plus_gemm_bf16_not_overlapped": ms_gemm_seq,
                          "ms_exchange_plus_gemm_fp32_operands": ms_gemm_fp32, "TFLOPs_aggregate_fp32_operands": flops / (ms_gemm_fp32 * 1e-3) / 1e12,
                          "bf16_paths_bit_identical_to_fp32_path": same,
                          "finite": finite, "timing": "CUDA events, max over ranks"}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
