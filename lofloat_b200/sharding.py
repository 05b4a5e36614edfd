"""Multi-GPU layer (new: the reference has no distributed code at all).

One process per GPU (`torch.distributed`, NCCL over NVLink on the GPU box, gloo in the CPU
tests).  The three ops shard like this:

* virtual_round / virtual_mx_round: independent elements / independent 32-element blocks ->
  contiguous ROW ranges per rank, NO collective on the data path.  Stochastic modes key
  Philox on the GLOBAL element index, so the result is independent of the partition.
* lof_gemm: 2-D output partition on an R x C process grid.  Rank (r, c) owns
  C[r-th row panel, c-th column panel] and needs A's row panel r and B's column panel c; K is
  never split (that would change the accumulation order).  Operands arrive row-sharded
  (the layout the rounding ops leave them in) and are exchanged with two collectives:
  an all-gather of A inside the row group, and an all-gather of B's column slice over
  the world (the "operand broadcast").
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional

import torch
import torch.distributed as dist


# --------------------------------------------------------------------------------------
# row sharding (rounding / MX)
# --------------------------------------------------------------------------------------
def row_partition(n_rows: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced row range [start, end) of `rank` (the first n_rows % world_size
    ranks get one extra row)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size: {rank}/{world_size}")
    if n_rows < 0:
        raise ValueError("n_rows < 0")
    base, rem = divmod(n_rows, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


@dataclass(frozen=True)
class RowShard:
    """A rank's view of a row-sharded 2-D (or flattened N-D) tensor."""
    n_rows: int          # global
    row_len: int         # elements per row (product of the trailing dims)
    row_start: int
    row_end: int

    @property
    def first_index(self) -> int:
        return self.row_start * self.row_len

    @property
    def local_rows(self) -> int:
        return self.row_end - self.row_start

    @staticmethod
    def make(n_rows: int, row_len: int, world_size: int, rank: int) -> "RowShard":
        s, e = row_partition(n_rows, world_size, rank)
        return RowShard(n_rows, row_len, s, e)


def shard_rows(full: torch.Tensor, world_size: int, rank: int) -> tuple[torch.Tensor, RowShard]:
    """Slice a full tensor (dim 0 = rows) into this rank's shard."""
    row_len = 1
    for d in full.shape[1:]:
        row_len *= d
    sh = RowShard.make(full.shape[0], row_len, world_size, rank)
    return full[sh.row_start:sh.row_end], sh


def sharded_virtual_round(local: torch.Tensor, shard: RowShard, params, round_mode=0, stoch_len: int = 0,
                          scale: float = 1.0, rng=None, round_fn: Optional[Callable] = None) -> torch.Tensor:
    """Round this rank's rows.  No communication.  `round_fn` defaults to the CUDA op."""
    if round_fn is None:
        from .ops import virtual_round as round_fn
    return round_fn(local, params, round_mode, stoch_len, scale, rng=rng, first_index=shard.first_index)


def sharded_virtual_mx_round(local: torch.Tensor, shard: RowShard, block_size: int, element_format, scale_format,
                             round_mode=0, scale_round_mode=0, stoch_len: int = 0, rng=None,
                             mx_fn: Optional[Callable] = None) -> torch.Tensor:
    """MX-round this rank's rows.  Blocks run along the last axis, so a row shard never
    splits a block as long as block_size divides the row length (layers.py:92,261 enforce
    the same)."""
    if shard.row_len % block_size != 0:
        raise ValueError(f"mx sharding needs the row length ({shard.row_len}) divisible by block_size ({block_size})")
    if mx_fn is None:
        from .ops import virtual_mx_round as mx_fn
    return mx_fn(local, block_size, element_format, scale_format, round_mode, scale_round_mode, stoch_len,
                 rng=rng, first_index=shard.first_index)


def gather_rows(local: torch.Tensor, shard: RowShard, world_size: int, group=None) -> torch.Tensor:
    """Reassemble the full tensor on every rank (verification / hand-off helper, not on the
    data path of the rounding ops)."""
    sizes = [row_partition(shard.n_rows, world_size, r) for r in range(world_size)]
    max_rows = max(e - s for s, e in sizes)
    pad = torch.zeros((max_rows,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world_size)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[: e - s] for b, (s, e) in zip(bufs, sizes)], dim=0)


# --------------------------------------------------------------------------------------
# 2-D output-partitioned GEMM
# --------------------------------------------------------------------------------------
def process_grid(world_size: int) -> tuple[int, int]:
    """R x C grid, as square as possible with R <= C (8 -> 2x4, 4 -> 2x2, 2 -> 1x2, 1 -> 1x1)."""
    r = int(world_size ** 0.5)
    while world_size % r != 0:
        r -= 1
    return r, world_size // r


@dataclass
class GemmGrid:
    world_size: int
    rank: int
    R: int
    C: int
    row_group: object = None     # ranks with the same r
    world_group: object = None

    @property
    def r(self) -> int:
        return self.rank // self.C

    @property
    def c(self) -> int:
        return self.rank % self.C

    @staticmethod
    def create(world_size: Optional[int] = None, rank: Optional[int] = None) -> "GemmGrid":
        """Collective: every rank must call it (new_group is collective)."""
        if world_size is None:
            world_size = dist.get_world_size() if dist.is_initialized() else 1
        if rank is None:
            rank = dist.get_rank() if dist.is_initialized() else 0
        R, C = process_grid(world_size)
        grid = GemmGrid(world_size, rank, R, C)
        if world_size > 1:
            for rr in range(R):
                g = dist.new_group(ranks=[rr * C + cc for cc in range(C)])
                if rr == grid.r:
                    grid.row_group = g
        return grid


def gemm_2d(A_local: torch.Tensor, B_local: torch.Tensor, grid: GemmGrid, M: int, N: int, K: int,
            accum_mant_bits: int, round_mode=0, stochastic_rounding_bits: int = 0, scale_a: float = 1.0,
            scale_b: float = 1.0, gemm_fn: Optional[Callable] = None, **gemm_kwargs) -> torch.Tensor:
    """C = A @ B with LoFloat accumulation on an R x C grid.

    A_local: this rank's rows of A under a 1-D row partition over the world, in rank order
             ([M/W, K]); B_local: this rank's rows of B ([K/W, N]).  M must be divisible by
             the world size, K by the world size and N by C.
    Returns this rank's block C[r*M/R:(r+1)*M/R, c*N/C:(c+1)*N/C].
    """
    if gemm_fn is None:
        from .ops import lof_gemm as gemm_fn
    W, R, Cn = grid.world_size, grid.R, grid.C
    if M % W or K % W or N % Cn:
        raise ValueError(f"gemm_2d needs M % world == 0, K % world == 0, N % C == 0 (M={M}, N={N}, K={K}, grid={R}x{Cn})")
    if tuple(A_local.shape) != (M // W, K) or tuple(B_local.shape) != (K // W, N):
        raise ValueError("gemm_2d: local operand shapes do not match the 1-D row partition")
    if W == 1:
        return gemm_fn(A_local, B_local, accum_mant_bits, round_mode, stochastic_rounding_bits, scale_a, scale_b,
                       **gemm_kwargs)
    # A row panel r: the C ranks of row group r hold consecutive row shards -> all-gather in the group
    A_panel = torch.empty((M // R, K), dtype=A_local.dtype, device=A_local.device)
    dist.all_gather_into_tensor(A_panel, A_local.contiguous(), group=grid.row_group)
    # B column panel c: every rank contributes its K/W rows; exchange only the N/C columns each peer needs
    nc = N // Cn
    send = torch.stack([B_local[:, (dst % Cn) * nc:(dst % Cn + 1) * nc] for dst in range(W)]).contiguous()
    recv = torch.empty_like(send)                      # [W, K/W, nc]: slice c of every rank's rows
    dist.all_to_all_single(recv, send, group=grid.world_group)
    B_panel = recv.reshape(K, nc)
    return gemm_fn(A_panel, B_panel, accum_mant_bits, round_mode, stochastic_rounding_bits, scale_a, scale_b,
                   **gemm_kwargs)


def gemm_2d_packed(A_local: torch.Tensor, Bt_local: torch.Tensor, grid: GemmGrid, M: int, N: int, K: int,
                   accum_mant_bits: int, round_mode=0, scale_a: float = 1.0, scale_b: float = 1.0,
                   gemm_fn: Optional[Callable] = None, overlap: bool = True, **gemm_kwargs) -> torch.Tensor:
    """gemm_2d for operands that already are in the tensor core's operand type (bf16, or uint8 e4m3 / e5m2
    representations) with the second operand in the weight layout: A_local = this rank's rows of A ([M/W, K]), Bt_local =
    this rank's rows of B^T ([N/W, K]) under 1-D row partitions over the world.  Half (bf16) or a quarter (8-bit) of the
    fp32 exchange volume, no transposes, and both panels are unions of whole row shards:
      A row panel r    = shards of ranks r*C .. r*C + C - 1      -> all-gather inside the row group
      Bt column panel c = shards of ranks c*R .. c*R + R - 1      -> one all-to-all in which shard s goes to column s // R
    With overlap=True (default) the all-gather of A runs while the rows of C that need only this rank's own A shard are
    being computed.  Returns this rank's block C[r*M/R:(r+1)*M/R, c*N/C:(c+1)*N/C] (fp32)."""
    if gemm_fn is None:
        from .ops import lof_gemm_packed as gemm_fn
    W, R, Cn = grid.world_size, grid.R, grid.C
    if M % W or N % W:
        raise ValueError(f"gemm_2d_packed needs M % world == 0 and N % world == 0 (M={M}, N={N}, grid={R}x{Cn})")
    if tuple(A_local.shape) != (M // W, K) or tuple(Bt_local.shape) != (N // W, K):
        raise ValueError("gemm_2d_packed: local operand shapes do not match the 1-D row partitions")
    if W == 1:
        return gemm_fn(A_local, Bt_local, accum_mant_bits, round_mode, scale_a, scale_b, **gemm_kwargs)
    rows = N // W
    my_col = grid.rank // R                                    # the column panel this rank's Bt shard belongs to
    dsts = [d for d in range(W) if d % Cn == my_col]            # the R ranks that own a block in that column
    send = Bt_local.contiguous().unsqueeze(0).expand(len(dsts), rows, K).reshape(len(dsts) * rows, K)
    in_split = [rows if d % Cn == my_col else 0 for d in range(W)]
    out_split = [rows if s // R == grid.c else 0 for s in range(W)]
    Bt_panel = torch.empty((N // Cn, K), dtype=Bt_local.dtype, device=Bt_local.device)
    A_panel = torch.empty((M // R, K), dtype=A_local.dtype, device=A_local.device)
    A_own = A_local.contiguous()
    if not overlap:
        dist.all_gather_into_tensor(A_panel, A_own, group=grid.row_group)
        dist.all_to_all_single(Bt_panel, send.contiguous(), out_split, in_split, group=grid.world_group)
        return gemm_fn(A_panel, Bt_panel, accum_mant_bits, round_mode, scale_a, scale_b, **gemm_kwargs)
    # overlapped: B^T first (every output needs all of it), then the rows of C that come from this rank's OWN A shard are
    # computed while the other A shards of the row panel are still in flight (output rows are independent)
    w_b = dist.all_to_all_single(Bt_panel, send.contiguous(), out_split, in_split, group=grid.world_group, async_op=True)
    w_a = dist.all_gather_into_tensor(A_panel, A_own, group=grid.row_group, async_op=True)
    mw = M // W
    parts = [None] * Cn
    w_b.wait()
    parts[grid.c] = gemm_fn(A_own, Bt_panel, accum_mant_bits, round_mode, scale_a, scale_b, **gemm_kwargs)
    w_a.wait()
    for j in range(Cn):
        if j != grid.c:
            parts[j] = gemm_fn(A_panel[j * mw:(j + 1) * mw], Bt_panel, accum_mant_bits, round_mode, scale_a, scale_b, **gemm_kwargs)
    return torch.cat(parts, dim=0)


def gemm_2d_mx(A_elems: torch.Tensor, A_scales: torch.Tensor, Bt_elems: torch.Tensor, Bt_scales: torch.Tensor, grid: GemmGrid,
               M: int, N: int, K: int, accum_mant_bits: int, round_mode=0, gemm_fn: Optional[Callable] = None, overlap: bool = True,
               **gemm_kwargs) -> torch.Tensor:
    """gemm_2d_packed for MX STORAGE (ops.mx_quantize_operand: uint8 element codes [rows, K] + uint8 E8M0 scale bytes [rows, K / 32]),
    consumed by the block-scaled tensor-core GEMM (ops.mx_gemm): 1.03 bytes per element on the wire instead of 2 (bf16 words) or 4.
    Same panels and collectives as gemm_2d_packed, issued once for the element arrays and once for the scale arrays:
      A row panel r     = shards of ranks r*C .. r*C + C - 1       -> all_gather_into_tensor inside the row group
      Bt column panel c = shards of ranks c*R .. c*R + R - 1       -> all_to_all_single over the world (shard s goes to column s // R)
    Returns this rank's block C[r*M/R:(r+1)*M/R, c*N/C:(c+1)*N/C] (fp32)."""
    if gemm_fn is None:
        from .ops import mx_gemm as gemm_fn
    W, R, Cn = grid.world_size, grid.R, grid.C
    if M % W or N % W or K % 32:
        raise ValueError(f"gemm_2d_mx needs M % world == 0, N % world == 0 and K % 32 == 0 (M={M}, N={N}, K={K}, grid={R}x{Cn})")
    kb = K // 32
    if (tuple(A_elems.shape) != (M // W, K) or tuple(A_scales.shape) != (M // W, kb) or tuple(Bt_elems.shape) != (N // W, K)
            or tuple(Bt_scales.shape) != (N // W, kb)):
        raise ValueError("gemm_2d_mx: local operand shapes do not match the 1-D row partitions")
    if W == 1:
        return gemm_fn(A_elems, A_scales, Bt_elems, Bt_scales, accum_mant_bits, round_mode, **gemm_kwargs)
    rows = N // W
    my_col = grid.rank // R
    n_dst = sum(1 for d in range(W) if d % Cn == my_col)
    in_split = [rows if d % Cn == my_col else 0 for d in range(W)]
    out_split = [rows if s // R == grid.c else 0 for s in range(W)]
    works, panels = [], {}
    for name, t in (("bt_e", Bt_elems), ("bt_s", Bt_scales)):           # B^T first: every output block needs all of it
        send = t.contiguous().unsqueeze(0).expand(n_dst, rows, t.size(1)).reshape(n_dst * rows, t.size(1)).contiguous()
        panels[name] = torch.empty((N // Cn, t.size(1)), dtype=t.dtype, device=t.device)
        works.append(dist.all_to_all_single(panels[name], send, out_split, in_split, group=grid.world_group, async_op=True))
    own = {"a_e": A_elems.contiguous(), "a_s": A_scales.contiguous()}
    a_works = []
    for name in ("a_e", "a_s"):
        t = own[name]
        panels[name] = torch.empty((M // R, t.size(1)), dtype=t.dtype, device=t.device)
        a_works.append(dist.all_gather_into_tensor(panels[name], t, group=grid.row_group, async_op=True))
    for w in works:
        w.wait()
    if not overlap:
        for w in a_works:
            w.wait()
        return gemm_fn(panels["a_e"], panels["a_s"], panels["bt_e"], panels["bt_s"], accum_mant_bits, round_mode, **gemm_kwargs)
    mw = M // W
    parts = [None] * Cn
    parts[grid.c] = gemm_fn(own["a_e"], own["a_s"], panels["bt_e"], panels["bt_s"], accum_mant_bits, round_mode, **gemm_kwargs)
    for w in a_works:
        w.wait()
    for j in range(Cn):
        if j != grid.c:
            parts[j] = gemm_fn(panels["a_e"][j * mw:(j + 1) * mw], panels["a_s"][j * mw:(j + 1) * mw], panels["bt_e"], panels["bt_s"],
                               accum_mant_bits, round_mode, **gemm_kwargs)
    return torch.cat(parts, dim=0)


# --------------------------------------------------------------------------------------
# Peer-to-peer operand exchange over NVLink (CUDA IPC mappings + plain copies) instead of NCCL collectives
# --------------------------------------------------------------------------------------
class PeerPanels:
    """The operand broadcast of the 2-D partitioned GEMM as PULLS over NVLink peer memory.

    Every rank registers its operand shards once (`register`): their CUDA-IPC handles and a 16-byte flag block travel through one
    `all_gather_object`, every rank maps its peers' shards, and from then on a call assembles
        A row panel r     = shards of ranks r*C .. r*C + C - 1      (this rank's row group)
        Bt column panel c = shards of ranks c*R .. c*R + R - 1
    with `cudaMemcpyAsync` peer copies on side streams -- no collective, no rendez-vous: ~0.09 ms for a 32 MB shard where the NCCL
    all-gather + all-to-all pair costs ~0.15 ms of latency alone -- while the rows of C that need only the rank's own A shard are computed.
    Ordering is a flag hand-shake on the device (csrc/lof_p2p.cu: post / wait / done), so the shards may be rewritten between calls
    (activations) without any host synchronisation: `begin_update()` makes the owner's stream wait until every pull of the latest epoch has
    completed (call it before overwriting the shards), `post` publishes epoch e after the owner's writes, a consumer's side stream waits for the
    peer's epoch before copying and reports `done` after.  Every rank must make the same sequence of calls.  One list of tensors per operand: [elements] for packed operands, [elements, scales] for MX storage."""

    def __init__(self, grid: GemmGrid):
        import ctypes as _C
        from . import _lib
        self._C, self._lib, self.grid = _C, _lib, grid
        self.epoch = 0
        self.dev = None
        self.mine = None
        self.peers = None

    # -- registration (collective) ----------------------------------------------------
    def _export(self, t: torch.Tensor):
        C_ = self._C
        h = (C_.c_uint8 * 64)()
        off = C_.c_uint64(0)
        self._lib.check(self._lib.lib().lof_p2p_export(t.data_ptr(), h, C_.byref(off)), "lof_p2p_export")
        return (bytes(h), int(off.value), tuple(t.shape), t.numel() * t.element_size())

    def _import(self, desc):
        C_ = self._C
        h = (C_.c_uint8 * 64).from_buffer_copy(desc[0])
        p = C_.c_void_p(0)
        self._lib.check(self._lib.lib().lof_p2p_import(h, desc[1], C_.byref(p)), "lof_p2p_import")
        return int(p.value), desc[3]                                  # (pointer valid on THIS device, bytes)

    def register(self, a_tensors, bt_tensors) -> None:
        g = self.grid
        self.a_mine = [t.contiguous() for t in a_tensors]
        self.bt_mine = [t.contiguous() for t in bt_tensors]
        self.dev = self.a_mine[0].device
        self.flags = torch.zeros(4, dtype=torch.int32, device=self.dev)
        self.epoch = 0
        if g.world_size == 1:
            return
        torch.cuda.synchronize(self.dev)
        with torch.cuda.device(self.dev):
            payload = ([self._export(t) for t in self.a_mine], [self._export(t) for t in self.bt_mine], self._export(self.flags))
        allp = [None] * g.world_size
        dist.all_gather_object(allp, payload)
        self.peer_a, self.peer_bt, self.peer_flags = {}, {}, {}
        need = {g.r * g.C + j for j in range(g.C)} | {g.c * g.R + i for i in range(g.R)}
        my_col = g.rank // g.R
        consumers = {g.r * g.C + j for j in range(g.C)} | {d for d in range(g.world_size) if d % g.C == my_col}
        self.n_consumers = len(consumers - {g.rank})                  # ranks that pull at least one of my shards
        with torch.cuda.device(self.dev):                            # the mappings are opened FROM this device (lazy peer access)
            for r_, (ha, hb, hf) in enumerate(allp):
                if r_ == g.rank or r_ not in need:
                    continue
                self.peer_a[r_] = [self._import(d) for d in ha]
                self.peer_bt[r_] = [self._import(d) for d in hb]
                self.peer_flags[r_] = self._import(hf)[0]
        self.side = {r_: torch.cuda.Stream(device=self.dev) for r_ in self.peer_flags}
        torch.cuda.synchronize(self.dev)
        dist.barrier()

    # -- one exchange + GEMM ------------------------------------------------------------
    def _pull(self, M: int, N: int):
        """Start the pulls of this call's epoch; returns (a_panels, bt_panels, events_bt, events_a)."""
        g, L, C_ = self.grid, self._lib.lib(), self._C
        W, R, Cn = g.world_size, g.R, g.C
        mw, nw = M // W, N // W
        a_pan = [torch.empty((M // R, t.size(1)), dtype=t.dtype, device=self.dev) for t in self.a_mine]
        bt_pan = [torch.empty((N // Cn, t.size(1)), dtype=t.dtype, device=self.dev) for t in self.bt_mine]
        self.epoch += 1
        cur = torch.cuda.current_stream(self.dev)
        st = L.lof_p2p_post(self.flags.data_ptr(), self.epoch, (self.epoch - 1) * self.n_consumers, cur.cuda_stream)
        self._lib.check(st, "lof_p2p_post")
        ev_bt, ev_a = [], []
        srcs_a = [g.r * Cn + j for j in range(Cn)]
        srcs_b = [g.c * R + i for i in range(R)]
        for src in sorted(set(srcs_a) | set(srcs_b)):
            if src == g.rank:
                continue
            s = self.side[src]
            s.wait_stream(cur)                                       # the panels were allocated on the current stream
            with torch.cuda.stream(s):
                self._lib.check(L.lof_p2p_wait(self.peer_flags[src], self.flags.data_ptr(), self.epoch, s.cuda_stream), "lof_p2p_wait")
                if src in srcs_b:
                    i = srcs_b.index(src)
                    for pan, (ptr, nbytes) in zip(bt_pan, self.peer_bt[src]):  # cudaMemcpyAsync peer -> local on THIS device's side stream
                        dst = pan[i * nw:(i + 1) * nw]
                        self._lib.check(L.lof_p2p_copy(dst.data_ptr(), ptr, nbytes, s.cuda_stream), "lof_p2p_copy")
                    e = torch.cuda.Event(); e.record(s); ev_bt.append(e)
                if src in srcs_a:
                    j = srcs_a.index(src)
                    for pan, (ptr, nbytes) in zip(a_pan, self.peer_a[src]):
                        dst = pan[j * mw:(j + 1) * mw]
                        self._lib.check(L.lof_p2p_copy(dst.data_ptr(), ptr, nbytes, s.cuda_stream), "lof_p2p_copy")
                    e = torch.cuda.Event(); e.record(s); ev_a.append(e)
                self._lib.check(L.lof_p2p_done(self.peer_flags[src], s.cuda_stream), "lof_p2p_done")
        if g.rank in srcs_b:
            i = srcs_b.index(g.rank)
            for pan, t in zip(bt_pan, self.bt_mine):
                pan[i * nw:(i + 1) * nw].copy_(t, non_blocking=True)
        return a_pan, bt_pan, ev_bt, ev_a

    def gemm(self, M: int, N: int, K: int, gemm_fn: Callable, split: Optional[bool] = None) -> torch.Tensor:
        """C block of this rank = gemm_fn(*a_panel_tensors, *bt_panel_tensors, out=block).  split=True computes the rows that come from the rank's
        own A shard while the other A shards are still in flight (default: when the per-rank problem is large enough to pay for the extra
        launches); every part is written straight into its rows of the result (gemm_fn must accept `out=`)."""
        g = self.grid
        if g.world_size == 1:
            return gemm_fn(*self.a_mine, *self.bt_mine, out=None)
        W, R, Cn = g.world_size, g.R, g.C
        mw = M // W
        cur = torch.cuda.current_stream(self.dev)
        a_pan, bt_pan, ev_bt, ev_a = self._pull(M, N)
        if split is None:
            split = (2.0 * (M // R) * (N // Cn) * K) > 2.0e11            # ~> 0.1 ms of tensor-core work per rank
        for e in ev_bt:
            cur.wait_event(e)
        if not split or Cn == 1:
            j = g.c
            for pan, t in zip(a_pan, self.a_mine):
                pan[j * mw:(j + 1) * mw].copy_(t, non_blocking=True)
            for e in ev_a:
                cur.wait_event(e)
            return gemm_fn(*a_pan, *bt_pan, out=None)
        Cblk = torch.empty((M // R, N // Cn), dtype=torch.float32, device=self.dev)
        gemm_fn(*self.a_mine, *bt_pan, out=Cblk[g.c * mw:(g.c + 1) * mw])
        for e in ev_a:
            cur.wait_event(e)
        for j in range(Cn):
            if j != g.c:
                gemm_fn(*[pan[j * mw:(j + 1) * mw] for pan in a_pan], *bt_pan, out=Cblk[j * mw:(j + 1) * mw])
        return Cblk

    def begin_update(self) -> None:
        """Enqueue (current stream) a wait until every peer has finished pulling the shards of the latest epoch: call it BEFORE rewriting the
        registered shards in place (the next gemm() publishes the new contents)."""
        if self.grid.world_size == 1 or self.epoch == 0:
            return
        cur = torch.cuda.current_stream(self.dev)
        self._lib.check(self._lib.lib().lof_p2p_post(self.flags.data_ptr(), self.epoch, self.epoch * self.n_consumers, cur.cuda_stream), "lof_p2p_post")

    def error_flag(self) -> int:
        """Nonzero if a bounded spin of the hand-shake expired on this rank (protocol error: unequal call sequences)."""
        return int(self.flags[3].item())
