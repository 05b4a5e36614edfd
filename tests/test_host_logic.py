"""CPU: the Python mirror of the reference interface (names, argument meaning, error behaviour) and
the multi-GPU sharding layer (world_size-2 gloo)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as L


@pytest.fixture(scope="module")
def lof():
    import __graft_entry__ as g
    from lofloat_b200 import _lib
    if not _lib.LIB_PATH.exists():
        g.build()
    import lofloat_b200
    return lofloat_b200


def test_enum_values_match_reference(lof):
    # fp_tools.hpp:13-50 / SURVEY.md Appendix D
    want = dict(RoundToNearestEven=0, RoundTowardsZero=1, RoundAwayFromZero=2, StochasticRoundingA=3,
                RoundToNearestOdd=4, RoundDown=5, RoundUp=6, RoundTiesToAway=7, StochasticRoundingB=8,
                StochasticRoundingC=9, True_StochasticRounding=10, ProbabilisticRounding=11, RoundToOdd=12,
                RoundTiesTowardsZero=13, StochasticRoundingD=14)
    for k, v in want.items():
        assert int(getattr(lof.RoundingMode, k)) == v
    assert int(lof.InfBehavior.Extended) == 0 and int(lof.InfBehavior.Saturating) == 1
    assert int(lof.Signedness.Signed) == 0 and int(lof.Signedness.Unsigned) == 1
    assert int(lof.NaNBehavior._3109) == 0 and int(lof.NaNBehavior._754) == 1


def test_format_factories(lof):
    p = lof.create_p3109_params(8, 4)                      # bias 2^(k-p-1) = 8 (formats.py:58)
    assert (p.total_bits, p.mantissa_bits, p.bias) == (8, 3, 8)
    assert p.inf_behavior == lof.InfBehavior.Saturating and p.signedness == lof.Signedness.Signed
    assert lof.create_p3109_params(8, 4, bias=7).bias == 7
    e = lof.create_e8m0_params()
    assert (e.total_bits, e.mantissa_bits, e.bias, e.inf_behavior, e.signedness) == \
        (8, 0, 127, lof.InfBehavior.Extended, lof.Signedness.Unsigned)
    h, s = lof.create_half_params(), lof.create_single_params()
    assert (h.total_bits, h.mantissa_bits, h.bias) == (16, 10, 15)
    assert (s.total_bits, s.mantissa_bits, s.bias) == (32, 23, 127)
    assert s.is_nan_checker(0x7FC00000) and not s.is_nan_checker(0x7F800000) and s.is_inf_checker(0xFF800000)
    with pytest.raises(AttributeError):
        p.bias = 3                                         # def_readonly


def test_descriptor_validation(lof):
    class Ok:
        def __call__(self, b): return False
        def minNegInf(self): return 0
        def minPosInf(self): return 0
        def qNanBitPattern(self): return 0
    with pytest.raises(ValueError, match="callable"):
        lof.FloatFormatDescriptor(8, 3, 7, 1, 0, 0, object(), Ok())
    class NoPos:
        def __call__(self, b): return False
        def minNegInf(self): return 0
    with pytest.raises(ValueError, match="minPosInf"):
        lof.FloatFormatDescriptor(8, 3, 7, 1, 0, 0, NoPos(), Ok())
    class NoNan:
        def __call__(self, b): return False
    with pytest.raises(ValueError, match="qNanBitPattern"):
        lof.FloatFormatDescriptor(8, 3, 7, 1, 0, 0, Ok(), NoNan())


def test_custom_ops_registered_with_reference_schemas(lof):
    s = str(torch.ops.lofloat.virtual_round.default._schema).replace("SymInt", "int")
    assert "Tensor input, int format_id, int round_mode, int stoch_len, float scale" in s
    s = str(torch.ops.lofloat.virtual_mx_round.default._schema).replace("SymInt", "int")
    assert "int block_size, int element_format_id, int scale_format_id, int round_mode, int scale_round_mode, int stoch_len" in s
    s = str(torch.ops.lofloat.lof_gemm.default._schema).replace("SymInt", "int")
    assert "Tensor A, Tensor B, int accum_mant_bits, int round_mode, int stochastic_rounding_bits, float scale_a, float scale_b" in s
    # fake (meta) kernels: _custom_ops.py:51-53, 79-82, 104-108
    p = lof.create_p3109_params(8, 4)
    fid = lof.register_format(p)
    assert lof.register_format(p) == fid and lof.get_format(fid) is p
    x = torch.empty(4, 64, device="meta")
    assert torch.ops.lofloat.virtual_round(x, fid, 0, 0, 1.0).shape == (4, 64)
    assert torch.ops.lofloat.virtual_mx_round(x, 32, fid, fid, 0, 0, 0).shape == (4, 64)
    a, b = torch.empty(8, 16, device="meta"), torch.empty(16, 24, device="meta")
    assert torch.ops.lofloat.lof_gemm(a, b, 10, 0, 0, 1.0, 1.0).shape == (8, 24)


def test_argument_errors_mirror_reference(lof):
    p = lof.create_p3109_params(8, 4)
    e8 = lof.create_e8m0_params()
    with pytest.raises(RuntimeError, match="float32 tensors only"):           # LoPy_bind.cpp:172
        lof.virtual_mx_round(torch.zeros(64, dtype=torch.float64), 32, p, e8)
    with pytest.raises(RuntimeError, match="block_size must be positive"):    # LoPy_bind.cpp:173
        lof.virtual_mx_round(torch.zeros(64), 0, p, e8)
    with pytest.raises(RuntimeError, match="must be 2-D"):                    # Lof_kernel.cu:106
        lof.lof_gemm(torch.zeros(4), torch.zeros(4, 4), 10)
    with pytest.raises(RuntimeError, match=r"A.cols \(3\) != B.rows \(4\)"):  # Lof_kernel.cu:107
        lof.lof_gemm(torch.zeros(4, 3), torch.zeros(4, 4), 10)
    with pytest.raises(RuntimeError, match="only float32"):                   # Lof_kernel.cu:109
        lof.lof_gemm(torch.zeros(4, 4, dtype=torch.float64), torch.zeros(4, 4, dtype=torch.float64), 10)
    with pytest.raises(RuntimeError, match="requires CUDA tensors"):          # Lof_kernel.cu:111
        lof.lof_gemm(torch.zeros(4, 4), torch.zeros(4, 4), 10)
    with pytest.raises(RuntimeError, match="not implemented for"):            # AT_DISPATCH, LoPy_bind.cpp:83
        lof.virtual_round(torch.zeros(4, dtype=torch.int32), 10)
    with pytest.raises(TypeError, match="incompatible function arguments"):
        lof.virtual_round(torch.zeros(4), "e4m3")


def test_row_partition(lof):
    sh = lof.sharding
    for n in (0, 1, 7, 8, 1000, 16384):
        for w in (1, 2, 3, 4, 8):
            parts = [sh.row_partition(n, w, r) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [e - s for s, e in parts]
            assert max(sizes) - min(sizes) <= 1
    assert sh.process_grid(8) == (2, 4) and sh.process_grid(4) == (2, 2) and sh.process_grid(2) == (1, 2)
    assert sh.process_grid(1) == (1, 1)
    with pytest.raises(ValueError):
        sh.row_partition(4, 2, 2)
    s = sh.RowShard.make(10, 64, 4, 1)
    assert (s.row_start, s.row_end, s.first_index, s.local_rows) == (3, 6, 192, 3)
    with pytest.raises(ValueError, match="divisible by block_size"):
        sh.sharded_virtual_mx_round(torch.zeros(3, 48), sh.RowShard.make(10, 48, 4, 1), 32, None, None, mx_fn=lambda *a, **k: None)


# ---------------------------------------------------------------------------------- gloo, world 2
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _np_round(x, params, mode, stoch_len, scale, rng=None, first_index=0):
    """Stand-in compute for the CPU test of the sharding layer: the oracle (tests only)."""
    f = L.fmt(params.total_bits, params.mantissa_bits, params.bias, int(params.inf_behavior), int(params.signedness))
    seed, off = rng if rng is not None else (0, 0)
    y = L.o_round_params(x.numpy().reshape(-1), f, int(mode), stoch_len, scale, seed=seed, offset=off,
                         first_index=first_index)
    return torch.from_numpy(y).reshape(x.shape)


def _np_mx(x, bs, ef, sf, mode, smode, stoch_len, rng=None, first_index=0):
    e = L.fmt(ef.total_bits, ef.mantissa_bits, ef.bias, int(ef.inf_behavior), int(ef.signedness))
    s = L.fmt(sf.total_bits, sf.mantissa_bits, sf.bias, int(sf.inf_behavior), int(sf.signedness))
    y = L.o_mx_round(x.numpy().reshape(-1), bs, s, e, int(mode), int(smode), stoch_len, first_index=first_index)
    return torch.from_numpy(y).reshape(x.shape)


def _np_gemm(A, B, bits, mode, sr, sa, sb, **kw):
    return torch.from_numpy(L.o_gemm(A.numpy(), B.numpy(), bits, int(mode), sr, kc=256, alpha=1.0 / (sa * sb)))


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import lofloat_b200 as lof
        sh = lof.sharding
        g = torch.Generator().manual_seed(2024)
        full = torch.randn(10, 64, generator=g) * torch.exp2(torch.randint(-12, 13, (10, 64), generator=g).float())
        p = lof.create_p3109_params(8, 4, bias=7)
        e8 = lof.create_e8m0_params()
        local, shard = sh.shard_rows(full, world, rank)
        # stochastic rounding keyed on the global index: sharded == unsharded
        y_loc = sh.sharded_virtual_round(local, shard, p, 3, 8, 1.0, rng=(7, 1), round_fn=_np_round)
        y = sh.gather_rows(y_loc, shard, world)
        y_full = _np_round(full, p, 3, 8, 1.0, rng=(7, 1))
        ok_round = torch.equal(y.view(torch.int32), y_full.view(torch.int32))
        m_loc = sh.sharded_virtual_mx_round(local, shard, 32, p, e8, mx_fn=_np_mx)
        m = sh.gather_rows(m_loc, shard, world)
        ok_mx = torch.equal(m.view(torch.int32), _np_mx(full, 32, p, e8, 0, 0, 0).view(torch.int32))
        # 2-D GEMM: 1x2 grid at world 2, 2x2 at world 4
        M, N, K = 8, 12, 300
        A = torch.rand(M, K, generator=g) * 2 - 1
        B = torch.rand(K, N, generator=g) * 2 - 1
        grid = sh.GemmGrid.create(world, rank)
        A_loc = A[rank * (M // world):(rank + 1) * (M // world)].contiguous()
        B_loc = B[rank * (K // world):(rank + 1) * (K // world)].contiguous()
        C_blk = sh.gemm_2d(A_loc, B_loc, grid, M, N, K, 10, gemm_fn=_np_gemm)
        C_full = _np_gemm(A, B, 10, 0, 0, 1.0, 1.0)
        r, c = grid.r, grid.c
        want = C_full[r * (M // grid.R):(r + 1) * (M // grid.R), c * (N // grid.C):(c + 1) * (N // grid.C)]
        ok_gemm = torch.equal(C_blk.view(torch.int32), want.contiguous().view(torch.int32))
        # the same product through the packed-operand exchange (operands row-sharded as A [M/W, K] and B^T [N/W, K])
        M2, N2 = 4 * world, 6 * world
        A2 = torch.rand(M2, K, generator=g) * 2 - 1
        Bt2 = torch.rand(N2, K, generator=g) * 2 - 1
        C2_blk = sh.gemm_2d_packed(A2[rank * 4:(rank + 1) * 4].contiguous(), Bt2[rank * 6:(rank + 1) * 6].contiguous(), grid, M2, N2, K, 10,
                                   gemm_fn=lambda a, bt, bits, mode, sa, sb, **kw: _np_gemm(a, bt.t().contiguous(), bits, mode, 0, sa, sb))
        C2_full = _np_gemm(A2, Bt2.t().contiguous(), 10, 0, 0, 1.0, 1.0)
        want2 = C2_full[r * (M2 // grid.R):(r + 1) * (M2 // grid.R), c * (N2 // grid.C):(c + 1) * (N2 // grid.C)]
        ok_gemm = ok_gemm and torch.equal(C2_blk.view(torch.int32), want2.contiguous().view(torch.int32))
        C2_seq = sh.gemm_2d_packed(A2[rank * 4:(rank + 1) * 4].contiguous(), Bt2[rank * 6:(rank + 1) * 6].contiguous(), grid, M2, N2, K, 10,
                                   overlap=False,
                                   gemm_fn=lambda a, bt, bits, mode, sa, sb, **kw: _np_gemm(a, bt.t().contiguous(), bits, mode, 0, sa, sb))
        ok_gemm = ok_gemm and torch.equal(C2_seq.view(torch.int32), C2_blk.view(torch.int32))
        # MX storage exchange (element bytes + scale bytes): stand-in GEMM = decode on the host, then the oracle GEMM
        K3 = 64
        Ae = torch.randint(0, 120, (M2, K3), generator=g, dtype=torch.uint8)
        As = torch.randint(120, 134, (M2, K3 // 32), generator=g, dtype=torch.uint8)
        Be = torch.randint(0, 120, (N2, K3), generator=g, dtype=torch.uint8)
        Bs = torch.randint(120, 134, (N2, K3 // 32), generator=g, dtype=torch.uint8)

        def _dec(e, sc):
            return (e.float().view(e.size(0), -1, 32) * torch.exp2(sc.float() - 127.0).unsqueeze(-1)).reshape(e.shape)

        def _mx_fn(ae, asc, be, bsc, bits, mode, **kw):
            return _np_gemm(_dec(ae, asc), _dec(be, bsc).t().contiguous(), bits, mode, 0, 1.0, 1.0)
        C3_blk = sh.gemm_2d_mx(Ae[rank * 4:(rank + 1) * 4], As[rank * 4:(rank + 1) * 4], Be[rank * 6:(rank + 1) * 6], Bs[rank * 6:(rank + 1) * 6],
                               grid, M2, N2, K3, 10, gemm_fn=_mx_fn)
        C3_seq = sh.gemm_2d_mx(Ae[rank * 4:(rank + 1) * 4], As[rank * 4:(rank + 1) * 4], Be[rank * 6:(rank + 1) * 6], Bs[rank * 6:(rank + 1) * 6],
                               grid, M2, N2, K3, 10, gemm_fn=_mx_fn, overlap=False)
        C3_full = _mx_fn(Ae, As, Be, Bs, 10, 0)
        want3 = C3_full[r * (M2 // grid.R):(r + 1) * (M2 // grid.R), c * (N2 // grid.C):(c + 1) * (N2 // grid.C)]
        ok_gemm = ok_gemm and torch.equal(C3_blk.view(torch.int32), want3.contiguous().view(torch.int32)) and torch.equal(C3_seq, C3_blk)
        q.put((rank, ok_round, ok_mx, ok_gemm, (grid.R, grid.C)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,want_grid", [(2, (1, 2)), (4, (2, 2))])
def test_sharding_gloo(lof, world, want_grid):
    """world 2 (1 x 2 grid) and world 4 (2 x 2: row AND column panels are real unions of shards)."""
    L.build_oracle()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_round, ok_mx, ok_gemm, grid in res:
        assert ok_round and ok_mx and ok_gemm, (rank, ok_round, ok_mx, ok_gemm)
        assert grid == want_grid
