"""Multi-GPU check, run under torchrun on the GPU box (N = 2, 4 or 8):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 scripts/multi_gpu_check.py

* row-sharded virtual_round (deterministic + stochastic) and virtual_mx_round == the single-GPU result on the full tensor
  (stochastic rounding is keyed on the GLOBAL element index, so the partition must not matter);
* 2-D output-partitioned lof_gemm (sharding.gemm_2d, NCCL all-gather / all-to-all of the operand panels) == the single-GPU GEMM.
Prints one PASS / FAIL line per check from rank 0.
"""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import lofloat_b200 as lof  # noqa: E402
from lofloat_b200 import sharding as S  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok_all = True

    def report(name, ok):
        nonlocal ok_all
        t = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        ok_all &= bool(t.item())
        if rank == 0:
            print(("PASS " if t.item() else "FAIL ") + name, flush=True)

    g = torch.Generator(device=dev).manual_seed(7)                  # same seed on every rank: identical full tensors
    rows, cols = 4096 + 8 * 3, 2048
    full = torch.randn(rows, cols, device=dev, generator=g) * torch.exp2(torch.randint(-10, 11, (rows, 1), device=dev, generator=g).float())
    e4m3 = lof.create_p3109_params(8, 4, bias=7)
    local_x, sh = S.shard_rows(full, world, rank)
    for name, mode, sl in (("RNE", 0, 0), ("RoundToOdd", 12, 0), ("SR-A8", 3, 8), ("True-SR", 10, 0)):
        want = lof.virtual_round(full, e4m3, mode, sl, rng=(11, 5))[sh.row_start:sh.row_end]
        got = S.sharded_virtual_round(local_x.contiguous(), sh, e4m3, mode, sl, rng=(11, 5))
        report(f"row-sharded virtual_round {name} == single GPU (world {world})", torch.equal(got.view(torch.int32), want.view(torch.int32)))
    want = lof.virtual_mx_round(full, 32, e4m3, lof.create_e8m0_params())[sh.row_start:sh.row_end]
    got = S.sharded_virtual_mx_round(local_x.contiguous(), sh, 32, e4m3, lof.create_e8m0_params())
    report(f"row-sharded virtual_mx_round == single GPU (world {world})", torch.equal(got.view(torch.int32), want.view(torch.int32)))

    # 2-D partitioned GEMM: C4 semantics (E4M3 operands, E5M10 accumulator re-rounded every k = 32) and the exact per-FMA mode
    M = N = K = 1024 * world if world <= 2 else 2048
    A = lof.virtual_round(torch.rand(M, K, device=dev, generator=g) * 2 - 1, e4m3)
    B = lof.virtual_round(torch.rand(K, N, device=dev, generator=g) * 2 - 1, e4m3)
    grid = S.GemmGrid.create(world, rank)
    A_loc = A[rank * (M // world):(rank + 1) * (M // world)].contiguous()
    B_loc = B[rank * (K // world):(rank + 1) * (K // world)].contiguous()
    e5m10 = lof.FloatFormatDescriptor(16, 10, 15, lof.InfBehavior.Saturating, lof.NaNBehavior._3109, lof.Signedness.Signed,
                                      lof.create_half_params().is_inf_checker, lof.create_half_params().is_nan_checker)
    Cfull = lof.lof_gemm(A, B, 0, k_chunk=32, accum_format=e5m10)
    Cblk = S.gemm_2d(A_loc, B_loc, grid, M, N, K, 0, k_chunk=32, accum_format=e5m10)
    r0, c0 = grid.r * (M // grid.R), grid.c * (N // grid.C)
    report(f"gemm_2d {grid.R}x{grid.C} chunked tcgen05 == single GPU", torch.equal(Cblk, Cfull[r0:r0 + M // grid.R, c0:c0 + N // grid.C]))
    m2 = 256 * world
    Cfull = lof.lof_gemm(A[:m2, :512].contiguous(), B[:512, :m2].contiguous(), 10)
    Cblk = S.gemm_2d(A[:m2, :512][rank * (m2 // world):(rank + 1) * (m2 // world)].contiguous(),
                     B[:512, :m2][rank * (512 // world):(rank + 1) * (512 // world)].contiguous(), grid, m2, m2, 512, 10)
    r0, c0 = grid.r * (m2 // grid.R), grid.c * (m2 // grid.C)
    report(f"gemm_2d {grid.R}x{grid.C} exact per-FMA mode == single GPU", torch.equal(Cblk, Cfull[r0:r0 + m2 // grid.R, c0:c0 + m2 // grid.C]))
    # packed-operand exchange: MX-rounded operands travel as bf16 (exact), second operand in the weight layout [N, K]
    e4m3, e8m0 = lof.create_p3109_params(8, 4, bias=7), lof.create_e8m0_params()
    # (.bfloat16().float(): the reference's saturating virtual_round leaves a block's overflowing maxima UNROUNDED, SURVEY
    #  Appendix B, so a few MX-rounded elements are arbitrary fp32 values; both paths get the same bf16-exact operands)
    Aq = lof.virtual_mx_round(torch.randn(M, K, device=dev, generator=g), 32, e4m3, e8m0).bfloat16().float()
    Wq = lof.virtual_mx_round(torch.randn(N, K, device=dev, generator=g), 32, e4m3, e8m0).bfloat16().float()
    exact_bf16 = torch.equal(Aq.bfloat16().float(), Aq) and torch.equal(Wq.bfloat16().float(), Wq)
    Cfull = lof.lof_gemm(Aq, Wq.t().contiguous(), 7, k_chunk=32, operand_dtype="bf16", partial_round=True)
    Cblk = S.gemm_2d_packed(Aq[rank * (M // world):(rank + 1) * (M // world)].bfloat16().contiguous(),
                            Wq[rank * (N // world):(rank + 1) * (N // world)].bfloat16().contiguous(), grid, M, N, K, 7,
                            k_chunk=32, partial_round=True)
    r0, c0 = grid.r * (M // grid.R), grid.c * (N // grid.C)
    report(f"gemm_2d_packed {grid.R}x{grid.C} (bf16 exchange of MX-rounded operands, exact: {exact_bf16}) == single GPU fp32-container GEMM",
           exact_bf16 and torch.equal(Cblk, Cfull[r0:r0 + M // grid.R, c0:c0 + N // grid.C]))
    # MX-storage exchange + block-scaled tensor-core GEMM (C5's path): rows quantised where they live, bytes + scale bytes on the wire
    Ax, Wx = torch.randn(M, K, device=dev, generator=g), torch.randn(N, K, device=dev, generator=g)
    ae, asc, _ = lof.mx_quantize_operand(Ax, e4m3, e8m0)
    we, wsc, _ = lof.mx_quantize_operand(Wx, e4m3, e8m0)
    Cfull = lof.mx_gemm(ae, asc, we, wsc, 7, k_chunk=32, partial_round=True)
    sa, sw = slice(rank * (M // world), (rank + 1) * (M // world)), slice(rank * (N // world), (rank + 1) * (N // world))
    ael, ascl, _ = lof.mx_quantize_operand(Ax[sa].contiguous(), e4m3, e8m0)            # quantising the shard == the shard of the quantised tensor
    wel, wscl, _ = lof.mx_quantize_operand(Wx[sw].contiguous(), e4m3, e8m0)
    same_q = torch.equal(ael, ae[sa]) and torch.equal(ascl, asc[sa]) and torch.equal(wel, we[sw]) and torch.equal(wscl, wsc[sw])
    Cblk = S.gemm_2d_mx(ael, ascl, wel, wscl, grid, M, N, K, 7, k_chunk=32, partial_round=True)
    report(f"gemm_2d_mx {grid.R}x{grid.C} (MX storage exchange, block-scaled MMA; shard quantisation == global: {same_q}) == single GPU",
           same_q and torch.equal(Cblk, Cfull[r0:r0 + M // grid.R, c0:c0 + N // grid.C]))
    # the same two products with the operand panels PULLED over NVLink peer memory (CUDA IPC mappings, device-side flag hand-shake) instead of NCCL
    try:
        pp = S.PeerPanels(grid)
        pp.register([ael, ascl], [wel, wscl])
        fn = lambda a_e, a_s, b_e, b_s, out=None: lof.mx_gemm(a_e, a_s, b_e, b_s, 7, k_chunk=32, partial_round=True, out=out)
        ok = True
        for it in range(3):
            for split in (False, True):
                ok = ok and torch.equal(pp.gemm(M, N, K, fn, split=split), Cblk)
        # shards rewritten between calls (activations): the hand-shake orders the pulls against the owner's writes without any host sync
        Ax2 = torch.randn(M, K, device=dev, generator=g)
        ae2, asc2, _ = lof.mx_quantize_operand(Ax2, e4m3, e8m0)
        want2 = lof.mx_gemm(ae2, asc2, we, wsc, 7, k_chunk=32, partial_round=True)[r0:r0 + M // grid.R, c0:c0 + N // grid.C]
        pp.begin_update()                                   # every peer has finished pulling the old contents
        ael.copy_(ae2[sa]); ascl.copy_(asc2[sa])
        ok2 = torch.equal(pp.gemm(M, N, K, fn), want2)
        pp.begin_update()
        ael.copy_(ae[sa]); ascl.copy_(asc[sa])
        ok3 = torch.equal(pp.gemm(M, N, K, fn), Cblk)
        report(f"PeerPanels {grid.R}x{grid.C} (P2P pulls of MX storage over NVLink) == gemm_2d_mx (NCCL), also after the shards were rewritten in place "
               f"[static {ok}, rewritten {ok2}, restored {ok3}, error flag {pp.error_flag()}]", ok and ok2 and ok3 and pp.error_flag() == 0)
    except Exception as exc:
        report(f"PeerPanels raised {exc!r}", False)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("ALL PASS" if ok_all else "SOME FAILED", flush=True)


if __name__ == "__main__":
    main()
