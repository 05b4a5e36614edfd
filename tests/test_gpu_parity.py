"""GPU: parity of the CUDA path (through the C ABI, liblofloat_b200.so) against the CPU oracle on
the same seeded inputs, against the committed golden vectors (generated from the compiled
reference), and -- at BASELINE.json's full sizes -- through size-independent properties.

Bar: BIT-EXACT for every deterministic mode and for the stochastic modes against the oracle's
Philox stream (integer/bit work); the chunked tensor-core GEMM states its tolerance inline."""
import ctypes as C
from pathlib import Path

import numpy as np
import pytest
import torch

import oracle_lib as L

pytestmark = pytest.mark.gpu
G = Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="module")
def lof(cuda_dev):
    import lofloat_b200
    return lofloat_b200


def _desc(lof, f: L.Format):
    chk = lof.create_single_params()
    return lof.FloatFormatDescriptor(f.bitwidth, f.mantissa_bits, f.bias, f.ov_behavior, f.na_behavior, f.is_signed,
                                     chk.is_inf_checker, chk.is_nan_checker)


def _gpu(x, dev):
    return torch.from_numpy(np.ascontiguousarray(x)).to(dev)


def _np(t):
    return t.detach().cpu().numpy()


def _bits_t(t):
    return {2: torch.int16, 4: torch.int32, 8: torch.int64}[t.element_size()]


@pytest.fixture(scope="module")
def xs():
    return np.concatenate([L.special_values_f32(), L.wide_random_f32(300000, 31)])


# ------------------------------------------------------------------------------ virtual_round
def test_round_params_all_formats_all_modes(lof, cuda_dev, xs):
    xt = _gpu(xs, cuda_dev)
    for name, f in L.FORMATS.items():
        d = _desc(lof, f)
        r = 23 - f.mantissa_bits
        for mode in range(15):
            sl = min(8, max(r, 1)) if mode in (L.SR_A, L.SR_B, L.SR_C, L.SR_D) else 0
            got = _np(lof.virtual_round(xt, d, mode, sl, rng=(2024, mode)))
            want = L.o_round_params(xs, f, mode, stoch_len=sl, seed=2024, offset=mode)
            assert L.bits_equal(got, want), (name, L.MODE_NAMES[mode], L.first_mismatch(xs, got, want))


def test_round_params_golden_from_reference(lof, cuda_dev):
    """GPU output == what the COMPILED REFERENCE produced (tests/golden/round_params_f32.npz)."""
    g = np.load(G / "round_params_f32.npz")
    x = g["x"]
    xt = _gpu(x, cuda_dev)
    for name, f in L.FORMATS.items():
        for mode in L.DETERMINISTIC_MODES:
            got = _np(lof.virtual_round(xt, _desc(lof, f), mode))
            want = g[f"{name}__{L.MODE_NAMES[mode]}"]
            assert L.bits_equal(got, want), (name, L.MODE_NAMES[mode], L.first_mismatch(x, got, want))


def test_round_mantissa_overload(lof, cuda_dev, xs):
    xt = _gpu(xs, cuda_dev)
    g = np.load(G / "round_mantissa_f32.npz")
    gt = _gpu(g["x"], cuda_dev)
    for m in (0, 2, 3, 7, 10, 15, 22, 23, 30):
        for mode in range(15):
            sl = min(8, 23 - m) if mode in (L.SR_A, L.SR_B, L.SR_C, L.SR_D) and m < 23 else 0
            got = _np(lof.virtual_round(xt, m, mode, sl, rng=(5, 9)))
            want = L.o_round_mantissa(xs, m, mode, stoch_len=sl, seed=5, offset=9)
            assert L.bits_equal(got, want), (m, L.MODE_NAMES[mode], L.first_mismatch(xs, got, want))
        if m in (0, 2, 3, 7, 10, 15, 22):
            for mode in L.DETERMINISTIC_MODES:
                got = _np(lof.virtual_round(gt, m, mode))
                assert L.bits_equal(got, g[f"m{m}__{L.MODE_NAMES[mode]}"]), (m, mode)


def test_round_fused_scale(lof, cuda_dev, xs):
    xt = _gpu(xs, cuda_dev)
    f = L.FORMATS["e4m3_b7"]
    for scale in (0.37, 448.0 / 3.3, 2.0 ** -5):
        got = _np(lof.virtual_round(xt, _desc(lof, f), 0, 0, scale))
        want = L.o_round_params(xs, f, L.RNE, scale=scale)
        assert L.bits_equal(got, want), scale
        got = _np(lof.virtual_round(xt, 10, 0, 0, scale))
        assert L.bits_equal(got, L.o_round_mantissa(xs, 10, L.RNE, scale=scale))


@pytest.mark.parametrize("dtype,code,tdt", [("f16", L.F16, torch.float16), ("bf16", L.BF16, torch.bfloat16),
                                            ("f64", L.F64, torch.float64)])
def test_round_other_containers(lof, cuda_dev, dtype, code, tdt):
    rng = np.random.default_rng(12)
    if code == L.F64:
        x = np.concatenate([L.special_values_f32().astype(np.float64),
                            rng.standard_normal(100000) * np.exp2(rng.integers(-160, 160, 100000))])
        xt = _gpu(x, cuda_dev)
        x_bits = x
    else:
        bits = np.concatenate([np.arange(0, 65536, dtype=np.uint32),
                               rng.integers(0, 65536, 70001, dtype=np.uint32)]).astype(np.uint16)
        xt = _gpu(bits.view(np.int16), cuda_dev).view(tdt)
        x_bits = bits.view(np.float16) if code == L.F16 else bits
    for name in ("e4m3_b7", "e5m2", "e2m1", "e3m2", "half", "bf16_like", "e8m0"):
        f = L.FORMATS[name]
        for mode in list(L.DETERMINISTIC_MODES) + [L.TRUE_SR, L.PROB]:
            got = lof.virtual_round(xt, _desc(lof, f), mode, 0, rng=(3, 4))
            want = L.o_round_params(x_bits, f, mode, dtype=code, seed=3, offset=4)
            gb = _np(got.view(_bits_t(got)))
            wb = want.view({2: np.int16, 8: np.int64}[want.dtype.itemsize])
            assert np.array_equal(gb, wb), (dtype, name, L.MODE_NAMES[mode], int((gb != wb).sum()))
    for m in (2, 5, 7):
        got = lof.virtual_round(xt, m, L.RNE, 0, 0.75)
        want = L.o_round_mantissa(x_bits, m, L.RNE, scale=0.75, dtype=code)
        assert np.array_equal(_np(got.view(_bits_t(got))), want.view({2: np.int16, 8: np.int64}[want.dtype.itemsize]))


def test_round_edge_shapes_alignment_inplace(lof, cuda_dev):
    """empty, ragged (n % 4 != 0), unaligned views, in-place through the raw C ABI, shard offsets."""
    from lofloat_b200 import _lib
    f = L.FORMATS["e4m3_b7"]
    d = _desc(lof, f)
    assert lof.virtual_round(torch.empty(0, device=cuda_dev), d).numel() == 0
    x = L.wide_random_f32(4103, 77)
    xt = _gpu(x, cuda_dev)
    for n in (1, 2, 3, 5, 255, 1023, 1025, 4103):
        for off in (0, 1, 2, 3):
            if off + n > x.size:
                continue
            view = xt[off:off + n]                               # data_ptr misaligned by 4*off bytes
            got = _np(lof.virtual_round(view, d, L.SR_A, 6, rng=(1, 2), first_index=1000 + off))
            want = L.o_round_params(x[off:off + n], f, L.SR_A, stoch_len=6, seed=1, offset=2, first_index=1000 + off)
            assert L.bits_equal(got, want), (n, off)
            got = _np(lof.virtual_round(view, d))
            assert L.bits_equal(got, L.o_round_params(x[off:off + n], f)), (n, off)
    # non-contiguous input: result is the elementwise op on the logical tensor
    m2 = xt[:4096].reshape(64, 64).t()
    got = _np(lof.virtual_round(m2, d))
    assert L.bits_equal(got, L.o_round_params(x[:4096].reshape(64, 64).T.copy(), f).reshape(64, 64))
    # raw C ABI, in place (out == in), as the reference's round_fp_params does
    lib = _lib.lib()
    buf = xt.clone()
    fmt = _lib.CFormat(8, 3, 7, 1, 0, 0, 0)
    ps = _lib.CProjSpec(0, 0, 0)
    st = lib.lof_round_params(buf.data_ptr(), buf.data_ptr(), buf.numel(), 0, C.byref(fmt), C.byref(ps), 1.0, None, 0,
                              torch.cuda.current_stream().cuda_stream)
    assert st == 0
    assert L.bits_equal(_np(buf), L.o_round_params(x, f))
    # 2-D shape preserved, new tensor returned, input untouched
    y = lof.virtual_round(xt[:4096].reshape(64, 64), d)
    assert y.shape == (64, 64) and y.data_ptr() != xt.data_ptr()
    assert L.bits_equal(_np(xt), x)


def test_host_buffer_entry_points(lof, cuda_dev):
    """CPU tensors are staged through the GPU by the library's chunked H2D/kernel/D2H pipeline."""
    from lofloat_b200 import _lib
    x = L.wide_random_f32(3 * (1 << 20) + 77, 5)
    f = L.FORMATS["e5m2"]
    old = _lib.get_option("host.chunk_bytes")
    _lib.set_option("host.chunk_bytes", 1 << 20)                 # force many chunks
    try:
        xt = torch.from_numpy(x).pin_memory()
        got = lof.virtual_round(xt, _desc(lof, f), L.SR_D, 4, rng=(8, 8)).numpy()
        assert L.bits_equal(got, L.o_round_params(x, f, L.SR_D, stoch_len=4, seed=8, offset=8))
        got = lof.virtual_round(torch.from_numpy(x), 7).numpy()  # pageable memory works too
        assert L.bits_equal(got, L.o_round_mantissa(x, 7))
        xm = np.nan_to_num(x[: 3 * (1 << 20)], nan=1.0, posinf=2.0, neginf=-2.0)
        got = lof.virtual_mx_round(torch.from_numpy(xm), 32, _desc(lof, L.FORMATS["e4m3_b7"]),
                                   lof.create_e8m0_params()).numpy()
        assert L.bits_equal(got, L.o_mx_round(xm, 32, L.FORMATS["e8m0"], L.FORMATS["e4m3_b7"]))
    finally:
        _lib.set_option("host.chunk_bytes", old)


def test_round_unroll_and_grid_variants(lof, cuda_dev, xs):
    from lofloat_b200 import _lib
    xt = _gpu(xs, cuda_dev)
    f = L.FORMATS["e4m3_b7"]
    want = L.o_round_params(xs, f, L.RNE)
    try:
        for unroll in (1, 2, 4, 8):
            for ctas in (1, 4, 16):
                _lib.set_option("round.unroll", unroll)
                _lib.set_option("round.ctas_per_sm", ctas)
                assert L.bits_equal(_np(lof.virtual_round(xt, _desc(lof, f))), want), (unroll, ctas)
    finally:
        _lib.set_option("round.unroll", 0)
        _lib.set_option("round.ctas_per_sm", 0)


def test_stochastic_statistics_on_device(lof, cuda_dev):
    """The reference's probability tests (test_prob_round.cpp:176-236) on the device: 2e5 trials."""
    n = 200000
    x = torch.full((n,), 1.0 + 1.3 / 8, device=cuda_dev)
    lo = 1.125
    xf = 0.3
    def p_up(mode, sl):
        y = lof.virtual_round(x, 3, mode, sl, rng=(99, mode * 16 + sl))
        return float((y > lo).float().mean())
    assert abs(p_up(L.PROB, 0) - 0.5) <= 0.02
    assert abs(p_up(L.TRUE_SR, 0) - xf) <= 0.02
    for sl in (1, 2, 3, 4):
        q = 2 ** sl
        assert abs(p_up(L.SR_D, sl) - 0.5 * (np.floor(q * xf) / q + np.ceil(q * xf) / q)) <= 0.02
    assert abs(p_up(L.SR_A, 10) - xf) <= 0.02
    # successive calls on the global stream draw fresh randomness; manual_seed rewinds it
    lof.manual_seed(123)
    a = lof.virtual_round(x, 3, L.TRUE_SR)
    b = lof.virtual_round(x, 3, L.TRUE_SR)
    lof.manual_seed(123)
    c = lof.virtual_round(x, 3, L.TRUE_SR)
    assert not torch.equal(a, b) and torch.equal(a, c)


# ---------------------------------------------------------------------------------- MX
MX_CONFIGS = [("e4m3_b7", "e8m0"), ("p3109_8p4", "e8m0"), ("e5m2", "e8m0"), ("e3m2", "e8m0"), ("e2m3", "e8m0"),
              ("e2m1", "e8m0"), ("e4m3_b7", "e5m2"), ("e5m2", "e5m2")]


def _mx_input(n, seed):
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal(n) * np.exp2(rng.integers(-8, 9, n))).astype(np.float32)
    if n >= 100:
        x[rng.integers(0, n, max(1, n // 1000))] *= 64
    return x


@pytest.mark.parametrize("variant", [0, 1, 2])
def test_mx_bit_exact(lof, cuda_dev, variant):
    """Methodology of test/test_virtual_mx_round_cuda.cu:76-152: GPU vs CPU scalar oracle, bit-exact,
    block sizes x formats x modes x n in {bs, 3bs, 3bs+1, 1000, 4096, 5003, ...}."""
    from lofloat_b200 import _lib
    _lib.set_option("mx.variant", variant)
    try:
        for bs in (16, 32, 64, 128, 256):
            for n in (bs, 3 * bs, 3 * bs + 1, 1000, 4096, 5003, 1 << 17, (1 << 17) + 4 * bs + 3):
                x = _mx_input(n, 1234 + n)
                xt = _gpu(x, cuda_dev)
                for en, sn in MX_CONFIGS:
                    for mode in (L.RNE, L.RTZ, L.RAZ, L.RTO, L.RNO):
                        got = _np(lof.virtual_mx_round(xt, bs, _desc(lof, L.FORMATS[en]), _desc(lof, L.FORMATS[sn]), mode))
                        want = L.o_mx_round(x, bs, L.FORMATS[sn], L.FORMATS[en], mode)
                        assert L.bits_equal(got, want), (variant, bs, n, en, sn, mode, L.first_mismatch(x, got, want))
    finally:
        _lib.set_option("mx.variant", 0)


@pytest.mark.parametrize("variant", [0, 1, 2])
def test_mx_modes_scale_modes_and_edge_cases(lof, cuda_dev, variant):
    from lofloat_b200 import _lib
    _lib.set_option("mx.variant", variant)
    try:
        e4, e8 = L.FORMATS["e4m3_b7"], L.FORMATS["e8m0"]
        x = _mx_input(1 << 15, 9)
        x[64:96] = 0.0                                           # all-zero block -> 0, not NaN
        x[128:160] = 1e-42                                       # amax underflows the scale format -> block = 0
        x[200] = 3e38                                            # scale overflows E8M0 (Extended -> +inf)
        x[300] = np.nan                                          # NaN element: dropped from the amax, +qNaN in the output
        x[420] = np.inf                                          # infinite amax -> infinite scale
        x[520] = -np.inf
        x[640:672] = np.float32(2.0 ** 127)                      # scale 2^127 / 2^126 boundary of the reciprocal shortcut
        x[700] = np.float32(1.7e38)
        x[800:832] = np.float32(1e-38) * np.arange(32, dtype=np.float32)   # subnormal-ish block: tiny scales
        x[900:932] = np.float32(3e-36)
        xt = _gpu(x, cuda_dev)
        for mode in range(15):
            sl = 8 if mode in (L.SR_A, L.SR_B, L.SR_C, L.SR_D) else 0
            for smode in (L.RNE, L.RU, L.RD, L.RTZ, L.TRUE_SR):
                got = _np(lof.virtual_mx_round(xt, 32, _desc(lof, e4), _desc(lof, e8), mode, smode, sl, rng=(6, 7)))
                want = L.o_mx_round(x, 32, e8, e4, mode, smode, sl, seed=6, offset=7)
                assert L.bits_equal(got, want), (variant, L.MODE_NAMES[mode], smode, L.first_mismatch(x, got, want))
        z = lof.virtual_mx_round(torch.zeros(4096, device=cuda_dev), 32, _desc(lof, e4), _desc(lof, e8))
        assert not torch.isnan(z).any() and not z.any()
        # KAT block of test/test_mx_round_py.py:37-40
        kat = torch.tensor([448, 224, -112, 1, 0.5, -2, 8, 0] * 4, device=cuda_dev)
        out = lof.virtual_mx_round(kat, 32, lof.create_p3109_params(8, 4), lof.create_e8m0_params())
        assert torch.equal(out, kat)
        # golden vectors generated from the compiled reference
        g = np.load(G / "mx_round_f32.npz")
        for n in (32, 96, 97, 1000, 5003):
            xx = g[f"x_{n}"]
            xg = _gpu(xx, cuda_dev)
            for bs in (16, 32, 64, 128, 256):
                for en in ("e4m3_b7", "e2m1"):
                    got = _np(lof.virtual_mx_round(xg, bs, _desc(lof, L.FORMATS[en]), _desc(lof, e8)))
                    assert L.bits_equal(got, g[f"mx_{n}_{bs}_{en}_e8m0_RNE"]), (n, bs, en)
        # narrow m = 0 scale formats whose Saturating overflow returns the UNROUNDED quotient (not a power of two), signed and
        # unsigned, plus an Extended one whose overflow gives +inf: every block size and every deterministic element mode
        xw = _mx_input(1 << 14, 21) * np.float32(64.0)           # amax / pmn far beyond 2^max_exp of the narrow formats
        xw[:512] *= np.float32(2.0 ** -9)                        # ... and some blocks inside their range
        xwt = _gpu(xw, cuda_dev)
        for sf in (L.fmt(4, 0, 4, L.SATURATING, L.UNSIGNED), L.fmt(4, 0, 3, L.SATURATING, L.SIGNED), L.fmt(5, 0, 8, L.EXTENDED, L.UNSIGNED)):
            for bs in (16, 32, 64, 128, 256):
                for mode in L.DETERMINISTIC_MODES:
                    for smode in (L.RNE, L.RU):
                        got = _np(lof.virtual_mx_round(xwt, bs, _desc(lof, e4), _desc(lof, sf), mode, smode))
                        want = L.o_mx_round(xw, bs, sf, e4, mode, smode)
                        assert L.bits_equal(got, want), (variant, sf.key(), bs, L.MODE_NAMES[mode], smode, L.first_mismatch(xw, got, want))
        # stochastic element modes on the fast path (one Philox block per 16 elements, runs of 4096 elements): every block size, a shard whose
        # first_index is run-aligned (fast kernel) and one that is not (generic kernel), a length that leaves a remainder for the generic kernel
        xs = _mx_input((1 << 15) + 4096 + 512, 33)
        xst = _gpu(xs, cuda_dev)
        for bs in (16, 32, 64, 128, 256):
            for mode, sl in ((L.SR_A, 8), (L.SR_A, 3), (L.SR_D, 7)):
                for fi in (0, 3 * 4096, 1024):
                    got = _np(lof.virtual_mx_round(xst, bs, _desc(lof, e4), _desc(lof, e8), mode, L.RNE, sl, rng=(9, 2), first_index=fi))
                    want = L.o_mx_round(xs, bs, e8, e4, mode, L.RNE, sl, seed=9, offset=2, first_index=fi)
                    assert L.bits_equal(got, want), (variant, bs, L.MODE_NAMES[mode], sl, fi, L.first_mismatch(xs, got, want))
        # shape is preserved, blocks run over the flattened contiguous tensor (LoPy_bind.cpp:175,216)
        x2 = xt[:4096].reshape(8, 512)
        assert lof.virtual_mx_round(x2, 32, _desc(lof, e4), _desc(lof, e8)).shape == (8, 512)
        # unaligned view falls back to the generic kernel and is still exact
        view = xt[1:1 + 4096]
        got = _np(lof.virtual_mx_round(view, 32, _desc(lof, e4), _desc(lof, e8)))
        assert L.bits_equal(got, L.o_mx_round(x[1:1 + 4096], 32, e8, e4))
    finally:
        _lib.set_option("mx.variant", 0)


# ---------------------------------------------------------------------------------- GEMM
def test_mx_division_shortcut_exhaustive(lof, cuda_dev):
    """The fast MX kernel forms amax / priv_max_normal with a 3-FMA sequence instead of an IEEE division:
    checked on the device against __fdiv_rn for EVERY non-negative finite fp32 amax, for every MX element format."""
    from lofloat_b200 import _lib
    lib = _lib.lib()
    for name in ("e4m3_b7", "p3109_8p4", "e5m2", "e3m2", "e2m3", "e2m1"):
        f = L.FORMATS[name]
        cf = _lib.CFormat(f.bitwidth, f.mantissa_bits, f.bias, f.ov_behavior, f.na_behavior, f.is_signed, 0)
        bad = C.c_uint64(123)
        _lib.check(lib.lof_selftest_mx_division(C.byref(cf), C.byref(bad), None), "selftest")
        assert bad.value == 0, (name, bad.value)


# --------------------------------------------------------------------------- chunked tensor-core GEMM
def _ints(shape, lo, hi, seed):
    return np.random.default_rng(seed).integers(lo, hi + 1, shape).astype(np.float32)


def _e4m3_uniform(shape, seed):
    """U(-1,1) rounded to OCP e4m3 -- the C4 operand distribution (test/gemm_test.cpp:82-86)."""
    x = np.random.default_rng(seed).uniform(-1, 1, shape).astype(np.float32)
    return L.o_round_params(x.ravel(), L.FORMATS["e4m3_b7"], L.RNE).reshape(shape)


def _tc(lof, dev, A, B, mant, k_chunk, **kw):
    return _np(lof.lof_gemm(_gpu(A, dev), _gpu(B, dev), mant, k_chunk=k_chunk, **kw))


def test_gemm_chunked_tcgen05_exactly_summable_data(lof, cuda_dev):
    """tcgen05 chunked GEMM vs the CPU statement of its semantics (oracle/lof_oracle.c lofo_gemm_chunked_f32) on
    small-integer operands, where every k-chunk partial is exact in any accumulator: BIT-EXACT for every
    deterministic mode, both fold semantics, mantissa-only and full-format accumulators, e4m3 / e5m2 / bf16
    operands, ragged shapes and the alpha scaling."""
    A, B = _ints((256, 1024), -8, 8, 5), _ints((1024, 384), -8, 8, 6)
    assert L.bits_equal(_tc(lof, cuda_dev, A, B, 23, 1024), A @ B)                       # plain GEMM, one chunk, no rounding
    for mant in (10, 7, 4):
        assert L.bits_equal(_tc(lof, cuda_dev, A, B, mant, 32, allow_packed=False), L.o_gemm_chunked(A, B, mant, 32)), mant
    for mode in L.DETERMINISTIC_MODES:
        got = _tc(lof, cuda_dev, A, B, 6, 32, round_mode=mode, allow_packed=False)
        assert L.bits_equal(got, L.o_gemm_chunked(A, B, 6, 32, mode=mode)), L.MODE_NAMES[mode]
        got = _tc(lof, cuda_dev, A, B, 5, 64, round_mode=mode, partial_round=True)
        assert L.bits_equal(got, L.o_gemm_chunked(A, B, 5, 64, mode=mode, partial_round=True)), L.MODE_NAMES[mode]
    e5m10 = L.fmt(16, 10, 15, L.SATURATING, L.SIGNED)
    got = _tc(lof, cuda_dev, A, B, 0, 32, accum_format=_desc(lof, e5m10), allow_packed=False)
    assert L.bits_equal(got, L.o_gemm_chunked(A, B, 0, 32, acc_fmt=e5m10))
    e4m3acc = L.FORMATS["e4m3_b7"]                                                          # tiny accumulator: saturates / overflows
    got = _tc(lof, cuda_dev, A, B, 0, 32, accum_format=_desc(lof, e4m3acc), round_mode=L.RTZ)
    assert L.bits_equal(got, L.o_gemm_chunked(A, B, 0, 32, acc_fmt=e4m3acc, mode=L.RTZ))
    # bf16 and e5m2 operands
    Ab, Bb = _ints((256, 512), -100, 100, 7), _ints((512, 256), -100, 100, 8)
    assert L.bits_equal(_tc(lof, cuda_dev, Ab, Bb, 8, 16, operand_dtype="bf16"), L.o_gemm_chunked(Ab, Bb, 8, 16))
    got = _tc(lof, cuda_dev, Ab, Bb, 7, 32, operand_dtype="bf16", partial_round=True)               # packed bfloat162 epilogue
    assert L.bits_equal(got, L.o_gemm_chunked(Ab, Bb, 7, 32, partial_round=True))
    A5, B5 = _ints((128, 256), -3, 3, 9) * np.float32(0.25), _ints((256, 128), -3, 3, 10) * np.float32(4.0)
    assert L.bits_equal(_tc(lof, cuda_dev, A5, B5, 10, 32, operand_dtype="e5m2"), L.o_gemm_chunked(A5, B5, 10, 32))
    # ragged M, N, K (K % 128 != 0, last chunk short), alpha
    Ar, Br = _ints((200, 400), -6, 6, 11), _ints((400, 136), -6, 6, 12)
    assert L.bits_equal(_tc(lof, cuda_dev, Ar, Br, 10, 32, allow_packed=False), L.o_gemm_chunked(Ar, Br, 10, 32))
    assert L.bits_equal(_tc(lof, cuda_dev, Ar, Br, 10, 96, allow_packed=False), L.o_gemm_chunked(Ar, Br, 10, 96))
    got = _tc(lof, cuda_dev, Ar, Br, 10, 32, scale_a=2.0, scale_b=4.0, allow_packed=False)
    assert L.bits_equal(got, L.o_gemm_chunked(Ar, Br, 10, 32, alpha=0.125))


def test_gemm_chunked_tcgen05_c4_distribution_and_packed_epilogue(lof, cuda_dev):
    """C4 operand distribution (U(-1,1) rounded to E4M3), k_chunk 32, accumulator mantissa 10 / E5M10.
    Tolerance vs the oracle: the tensor core may form a chunk's fp32 partial with a different internal alignment
    than the oracle's exactly rounded sum, which can move an accumulator by one unit of the ACCUMULATOR format per
    disagreement; stated bound: |C - C_oracle| <= 4 ulp of the accumulator format (2^-10 relative) at the result's
    magnitude (in practice the two agree bit for bit on this data).  The packed half2 epilogue must equal the
    generic epilogue BIT FOR BIT (same tensor-core partials)."""
    A, B = _e4m3_uniform((512, 2048), 13), _e4m3_uniform((2048, 256), 14)

    def ulp_err(got, want, mant=10):
        ulp = np.exp2(np.floor(np.log2(np.maximum(np.abs(want), 1e-30))) - mant)
        return float((np.abs(got.astype(np.float64) - want) / ulp).max())
    want2 = L.o_gemm_chunked(A, B, 10, 32, partial_round=True)
    gen2 = _tc(lof, cuda_dev, A, B, 10, 32, partial_round=True, allow_packed=False)
    assert ulp_err(gen2, want2) <= 4.0
    assert L.bits_equal(_tc(lof, cuda_dev, A, B, 10, 32, partial_round=True), gen2)           # packed half2 == generic
    assert lof.lof_gemm.last_status == 0
    e5m10 = L.fmt(16, 10, 15, L.SATURATING, L.SIGNED)
    assert L.bits_equal(_tc(lof, cuda_dev, A, B, 0, 32, accum_format=_desc(lof, e5m10), partial_round=True), gen2)
    assert ulp_err(_tc(lof, cuda_dev, A, B, 10, 32), L.o_gemm_chunked(A, B, 10, 32)) <= 4.0  # single-rounding fold
    # leaving the fp16 range: the packed kernel must flag it and the generic kernel recompute
    Ab, Bb = (_ints((128, 512), 200, 448, 15) // 32) * 32, (_ints((512, 128), 200, 448, 16) // 32) * 32
    got = _tc(lof, cuda_dev, Ab, Bb, 10, 32, partial_round=True)
    assert L.bits_equal(got, L.o_gemm_chunked(Ab, Bb, 10, 32, partial_round=True))
    assert lof.lof_gemm.last_status & 4
    # operands that are not exactly representable are reported, not silently rounded
    with pytest.raises(lof.LoFloatError):
        _tc(lof, cuda_dev, np.full((128, 128), 0.3, np.float32), np.ones((128, 128), np.float32), 10, 32)
    with pytest.raises(lof.LoFloatError):
        lof.lof_gemm(_gpu(A, cuda_dev), _gpu(B, cuda_dev), 10, k_chunk=48)                          # k_chunk % 32 != 0


def test_gemm_chunked_full_size_properties(lof, cuda_dev):
    """C4 at full size (8192^3, E4M3 operands, k_chunk 32, accumulator mantissa 10): size-independent properties.
    (1) sampled rows x columns against the oracle, (2) rows of A that are zero give zero rows, (3) scaling A by 2
    (exact in every format involved) scales C by exactly 2, (4) packed == generic on a sampled band."""
    n = 8192
    g = torch.Generator(device=cuda_dev).manual_seed(42)
    e4m3 = lof.create_p3109_params(8, 4, bias=7)
    A = lof.virtual_round(torch.rand(n, n, device=cuda_dev, generator=g) * 2 - 1, e4m3)
    B = lof.virtual_round(torch.rand(n, n, device=cuda_dev, generator=g) * 2 - 1, e4m3)
    A[100:164] = 0
    Cm = lof.lof_gemm(A, B, 10, k_chunk=32, partial_round=True)
    assert not Cm[100:164].any()
    rows, cols = [0, 1, 777, 4095, 8191], slice(4000, 4128)
    want = L.o_gemm_chunked(_np(A[rows]), _np(B[:, cols]), 10, 32, partial_round=True)
    got = _np(Cm[rows][:, cols])
    ulp = np.exp2(np.floor(np.log2(np.maximum(np.abs(want), 1e-30))) - 10)
    assert float((np.abs(got.astype(np.float64) - want) / ulp).max()) <= 4.0
    C2 = lof.lof_gemm(A * 2, B, 10, k_chunk=32, partial_round=True)
    assert torch.equal(C2, Cm * 2)
    Cg = lof.lof_gemm(A[:256], B, 10, k_chunk=32, partial_round=True, allow_packed=False)
    assert torch.equal(Cg, Cm[:256])


def test_gemm_persistent_fold_kernel_equals_first_generation_and_oracle(lof, cuda_dev):
    """Third-generation persistent per-chunk fold kernel (lof_gemm_tc3.cuh: one CTA per SM walking a tile schedule, 128 x 256
    tiles with packed half2 / bfloat162 accumulators, 128 x 128 tiles with fp32 accumulators) must be BIT-IDENTICAL to the
    first-generation one-tile-per-CTA kernel (same tensor-core partials, same fold) for every accumulator semantics, on
    shapes with fewer tiles than SMs, several tiles per CTA, ragged edges and short last chunks; and bit-exact against the
    oracle on exactly summable data."""
    dev = cuda_dev
    e5m10 = _desc(lof, L.fmt(16, 10, 15, L.SATURATING, L.SIGNED))
    shapes = ((128, 256, 128), (200, 136, 400), (2560, 2560, 512), (1000, 520, 1024 + 32), (384, 4096, 256), (4000, 300, 96))
    for (M, N, K) in shapes:
        A, B = _e4m3_uniform((M, K), 3 * M + K), _e4m3_uniform((K, N), N + 7 * K + 1)
        At, Bt = _gpu(A, dev), _gpu(B, dev)
        cases = [(10, dict(k_chunk=32, partial_round=True)),                         # packed half2 fold (+ gated generic recompute)
                 (7, dict(k_chunk=32, partial_round=True)),                          # packed bfloat162 fold
                 (7, dict(k_chunk=64, partial_round=True, operand_dtype="bf16")),    # ... with bf16 operands (C5)
                 (7, dict(k_chunk=32)),                                              # generic single-rounding fold, mantissa-only
                 (5, dict(k_chunk=32, round_mode=L.RTZ, partial_round=True)),
                 (4, dict(k_chunk=96, round_mode=L.RTO)),
                 (0, dict(k_chunk=32, accum_format=e5m10, round_mode=L.RAZ)),        # full-format accumulator, non-RNE: no hardware path
                 (10, dict(k_chunk=32, allow_packed=False, partial_round=True))]
        for mant, kw in cases:
            got = lof.lof_gemm(At, Bt, mant, **kw)
            ref = lof.lof_gemm(At, Bt, mant, allow_persistent=False, **kw)
            assert torch.equal(got.view(torch.int32), ref.view(torch.int32)), (M, N, K, mant, kw)
    A, B = _ints((300, 1024), -8, 8, 5), _ints((1024, 520), -8, 8, 6)
    for mode in L.DETERMINISTIC_MODES:
        assert L.bits_equal(_tc(lof, dev, A, B, 6, 32, round_mode=mode, allow_packed=False), L.o_gemm_chunked(A, B, 6, 32, mode=mode)), mode
        assert L.bits_equal(_tc(lof, dev, A, B, 5, 64, round_mode=mode, partial_round=True),
                            L.o_gemm_chunked(A, B, 5, 64, mode=mode, partial_round=True)), mode
    Ab, Bb = _ints((384, 640), -100, 100, 7), _ints((640, 300), -100, 100, 8)
    assert L.bits_equal(_tc(lof, dev, Ab, Bb, 7, 32, operand_dtype="bf16", partial_round=True), L.o_gemm_chunked(Ab, Bb, 7, 32, partial_round=True))
    # leaving the fp16 range: the packed half2 fold flags it and the gated generic kernel recomputes C
    Ao, Bo = (_ints((128, 512), 200, 448, 15) // 32) * 32, (_ints((512, 128), 200, 448, 16) // 32) * 32
    assert L.bits_equal(_tc(lof, dev, Ao, Bo, 10, 32, partial_round=True), L.o_gemm_chunked(Ao, Bo, 10, 32, partial_round=True))
    assert lof.lof_gemm.last_status & 4


def test_gemm_pair_kernel_equals_single_cta_kernel_and_oracle(lof, cuda_dev):
    """Persistent CTA-pair main loop (tcgen05 cta_group::2, 256 x 256 tiles, lof_gemm_tc2.cuh): the accumulator
    semantics it serves -- fp16 accumulation inside the tensor core (k_chunk 32, E5M10 / mantissa 10, RNE) and
    one-chunk fp32 accumulation with a final rounding -- must be BIT-IDENTICAL to the single-CTA kernel (same
    tensor-core arithmetic, different tiling) on aligned, ragged and tiny shapes, and bit-exact against the oracle on
    exactly summable data."""
    e5m10 = _desc(lof, L.fmt(16, 10, 15, L.SATURATING, L.SIGNED))
    for (M, N, K) in ((256, 256, 512), (200, 136, 400), (1000, 520, 4096 + 32), (128, 64, 128), (768, 1280, 1024)):
        A, B = _e4m3_uniform((M, K), M + K), _e4m3_uniform((K, N), N + K + 1)
        for kw, mant in ((dict(k_chunk=32, accum_format=e5m10), 0), (dict(k_chunk=32), 10), (dict(k_chunk=0), 10),
                         (dict(k_chunk=0, accum_format=e5m10), 0), (dict(k_chunk=0), 23)):
            got = _np(lof.lof_gemm(_gpu(A, cuda_dev), _gpu(B, cuda_dev), mant, **kw))
            ref = _np(lof.lof_gemm(_gpu(A, cuda_dev), _gpu(B, cuda_dev), mant, allow_pair=False, **kw))
            assert L.bits_equal(got, ref), (M, N, K, kw)
    A, B = _ints((300, 1024), -8, 8, 5), _ints((1024, 520), -8, 8, 6)
    assert L.bits_equal(_tc(lof, cuda_dev, A, B, 23, 0), A @ B)
    assert L.bits_equal(_tc(lof, cuda_dev, A, B, 10, 32), L.o_gemm_chunked(A, B, 10, 32))
    assert L.bits_equal(_tc(lof, cuda_dev, A, B, 0, 32, accum_format=e5m10),
                        L.o_gemm_chunked(A, B, 0, 32, acc_fmt=L.fmt(16, 10, 15, L.SATURATING, L.SIGNED)))
    assert L.bits_equal(_tc(lof, cuda_dev, A, B, 4, 0), L.o_gemm_chunked(A, B, 4, 1024))
    Ab, Bb = _ints((384, 640), -60, 60, 7), _ints((640, 300), -60, 60, 8)
    assert L.bits_equal(_tc(lof, cuda_dev, Ab, Bb, 8, 0, operand_dtype="bf16"), L.o_gemm_chunked(Ab, Bb, 8, 640))
    # leaving the fp16 range: flagged, and the generic kernel recomputes C
    Ao, Bo = (_ints((128, 512), 200, 448, 15) // 32) * 32, (_ints((512, 128), 200, 448, 16) // 32) * 32
    assert L.bits_equal(_tc(lof, cuda_dev, Ao, Bo, 10, 32), L.o_gemm_chunked(Ao, Bo, 10, 32))
    assert lof.lof_gemm.last_status & 4


def test_gemm_pair_kernel_c4_full_size_properties(lof, cuda_dev):
    """C4 at full size through the CTA-pair kernel (8192^3, E4M3 operands, E5M10 accumulator re-rounded every k = 32
    by the tensor core): sampled rows against the oracle within the stated 4 ulp-of-accumulator tolerance, zero rows,
    exact power-of-two scaling, and pair == single-CTA kernel on a band."""
    n = 8192
    g = torch.Generator(device=cuda_dev).manual_seed(42)
    e4m3 = lof.create_p3109_params(8, 4, bias=7)
    A = lof.virtual_round(torch.rand(n, n, device=cuda_dev, generator=g) * 2 - 1, e4m3)
    B = lof.virtual_round(torch.rand(n, n, device=cuda_dev, generator=g) * 2 - 1, e4m3)
    A[300:364] = 0
    e5m10_o = L.fmt(16, 10, 15, L.SATURATING, L.SIGNED)
    Cm = lof.lof_gemm(A, B, 0, k_chunk=32, accum_format=_desc(lof, e5m10_o))
    assert lof.lof_gemm.last_status == 0
    assert not Cm[300:364].any()
    rows, cols = [0, 1, 255, 256, 4097, 8191], slice(3968, 4224)
    want = L.o_gemm_chunked(_np(A[rows]), _np(B[:, cols]), 0, 32, acc_fmt=e5m10_o)
    got = _np(Cm[rows][:, cols])
    ulp = np.exp2(np.floor(np.log2(np.maximum(np.abs(want), 2.0 ** -14))) - 10)
    assert float((np.abs(got.astype(np.float64) - want) / ulp).max()) <= 4.0
    C2 = lof.lof_gemm(A * 2, B, 0, k_chunk=32, accum_format=_desc(lof, e5m10_o))
    assert torch.equal(C2, Cm * 2)
    Cs = lof.lof_gemm(A[:512], B, 0, k_chunk=32, accum_format=_desc(lof, e5m10_o), allow_pair=False)
    assert torch.equal(Cs, Cm[:512])


def test_gemm_exact_vs_oracle_and_golden(lof, cuda_dev):
    g = np.load(G / "gemm_f32.npz")
    for (M, N, K) in ((2, 2, 4), (17, 33, 64), (64, 48, 300), (40, 24, 513)):
        A, B = g[f"A_{M}_{N}_{K}"], g[f"B_{M}_{N}_{K}"]
        At, Bt = _gpu(A, cuda_dev), _gpu(B, cuda_dev)
        for mode in L.DETERMINISTIC_MODES:
            for bits in (10, 7):
                got = _np(lof.lof_gemm(At, Bt, bits, mode, kc=256))          # goldens: bli_sgemm + lof_sgemm_ukr, KC = 256
                assert L.bits_equal(got, g[f"C_{M}_{N}_{K}_{L.MODE_NAMES[mode]}_{bits}"]), (M, N, K, mode, bits)
    A, B = g["kat_A"], g["kat_B"]                                # exact-representable KAT, every mode == fp64 product
    for mode in L.DETERMINISTIC_MODES:
        got = _np(lof.lof_gemm(_gpu(A, cuda_dev), _gpu(B, cuda_dev), 10, mode, kc=256))
        assert L.bits_equal(got, g["kat_C"])
    rng = np.random.default_rng(21)
    for (M, N, K) in ((1, 1, 1), (65, 63, 129), (128, 192, 512), (200, 72, 1031)):
        A = rng.uniform(-1, 1, (M, K)).astype(np.float32)
        B = rng.uniform(-1, 1, (K, N)).astype(np.float32)
        At, Bt = _gpu(A, cuda_dev), _gpu(B, cuda_dev)
        for kc in (256, 0, 100):
            got = _np(lof.lof_gemm(At, Bt, 10, L.RNE, kc=kc))
            assert L.bits_equal(got, L.o_gemm(A, B, 10, L.RNE, kc=kc)), (M, N, K, kc)
        # the DEFAULT call is the reference's CUDA leaf: one accumulator over all of K (cutlass_gemms.cuh:97-168), not the BLIS panels
        assert L.bits_equal(_np(lof.lof_gemm(At, Bt, 10)), L.o_gemm(A, B, 10, L.RNE, kc=0)), (M, N, K, "default kc")
        got = _np(lof.lof_gemm(At, Bt, 5, L.RAZ, 0, 3.0, 0.25, kc=256))          # output / (scale_a * scale_b)
        assert L.bits_equal(got, L.o_gemm(A, B, 5, L.RAZ, kc=256, alpha=np.float32(1.0 / 0.75)))
        for mode, sl in ((L.SR_A, 6), (L.TRUE_SR, 0), (L.SR_D, 3)):
            got = _np(lof.lof_gemm(At, Bt, 10, mode, sl, kc=256, rng=(11, 12)))
            assert L.bits_equal(got, L.o_gemm(A, B, 10, mode, sl, kc=256, seed=11, offset=12)), (M, N, K, mode)
    # wide-range operands: cancellation, subnormal and overflowing partial sums exercise round_to_odd_fma
    A = (rng.standard_normal((48, 96)) * np.exp2(rng.integers(-70, 60, (48, 96)))).astype(np.float32)
    B = (rng.standard_normal((96, 40)) * np.exp2(rng.integers(-70, 60, (96, 40)))).astype(np.float32)
    got = _np(lof.lof_gemm(_gpu(A, cuda_dev), _gpu(B, cuda_dev), 10, L.RNE, kc=256))
    want = L.o_gemm(A, B, 10, L.RNE, kc=256)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_gemm_via_torch_op_and_linear_composition(lof, cuda_dev):
    """torch.ops.lofloat.* end to end: the LoF_Linear composition of layers.py:361-398
    (round activations and weights, transpose, lof_gemm)."""
    rng = np.random.default_rng(4)
    x = rng.standard_normal((32, 128)).astype(np.float32)
    w = rng.standard_normal((48, 128)).astype(np.float32)
    p = lof.create_p3109_params(8, 4, bias=7)
    fid = lof.register_format(p)
    xq = torch.ops.lofloat.virtual_round(_gpu(x, cuda_dev), fid, 0, 0, 1.0)
    wq = torch.ops.lofloat.virtual_round(_gpu(w, cuda_dev), fid, 0, 0, 1.0)
    out = torch.ops.lofloat.lof_gemm(xq, wq.t().contiguous(), 10, 0, 0, 1.0, 1.0)
    f = L.FORMATS["e4m3_b7"]
    xo, wo = L.o_round_params(x, f).reshape(x.shape), L.o_round_params(w, f).reshape(w.shape)
    want = L.o_gemm(xo, wo.T.copy(), 10, L.RNE, kc=256)
    assert L.bits_equal(_np(out), want)
    e8 = lof.register_format(lof.create_e8m0_params())
    xm = torch.ops.lofloat.virtual_mx_round(_gpu(x, cuda_dev), 32, fid, e8, 0, 0, 0)
    assert L.bits_equal(_np(xm), L.o_mx_round(x, 32, L.FORMATS["e8m0"], f).reshape(x.shape))


# ------------------------------------------------------------------ full-size properties
def test_full_size_properties_round(lof, cuda_dev):
    """BASELINE config C2 size (1 GiB fp32 = 2^28 elements): idempotence, sign symmetry, agreement with
    the oracle on sampled windows, and partition invariance of the stochastic stream."""
    n = 1 << 28
    g = torch.Generator(device=cuda_dev).manual_seed(2024)
    x = torch.randn(n, device=cuda_dev, generator=g)
    x *= torch.exp2(torch.randint(-12, 13, (n,), device=cuda_dev, generator=g).float())
    for name in ("e4m3_b7", "e2m1", "bf16_like"):
        f = L.FORMATS[name]
        d = _desc(lof, f)
        for mode in (L.RNE, L.RTZ, L.RTO):
            y = lof.virtual_round(x, d, mode)
            if f.ov_behavior == L.SATURATING:
                # idempotence holds on everything the format can represent (overflowing inputs are
                # returned unrounded by the reference)
                y2 = lof.virtual_round(y, d, mode)
                assert torch.equal(y2.view(torch.int32), y.view(torch.int32))
            ym = lof.virtual_round(-x, d, mode)
            if f.ov_behavior == L.SATURATING:
                assert torch.equal((-ym).view(torch.int32), y.view(torch.int32))      # odd symmetry of these modes
            for start in (0, 12345678, n - (1 << 16)):
                w = slice(start, start + (1 << 16))
                assert L.bits_equal(_np(y[w]), L.o_round_params(_np(x[w]), f, mode)), (name, mode, start)
            del y, ym
    # stochastic: rounding two halves separately with their global offsets == rounding the whole
    d = _desc(lof, L.FORMATS["e4m3_b7"])
    whole = lof.virtual_round(x, d, L.SR_A, 8, rng=(42, 1))
    half = n // 2
    lo = lof.virtual_round(x[:half], d, L.SR_A, 8, rng=(42, 1), first_index=0)
    hi = lof.virtual_round(x[half:], d, L.SR_A, 8, rng=(42, 1), first_index=half)
    assert torch.equal(whole[:half], lo) and torch.equal(whole[half:], hi)
    w = slice(777777, 777777 + 4096)
    assert L.bits_equal(_np(whole[w]), L.o_round_params(_np(x[w]), L.FORMATS["e4m3_b7"], L.SR_A, stoch_len=8,
                                                        seed=42, offset=1, first_index=777777))


@pytest.mark.parametrize("variant", [0, 1])
def test_full_size_properties_mx(lof, cuda_dev, variant):
    """BASELINE config C3 size (16384 x 16384 fp32, block 32): fake-quant is idempotent for power-of-two
    scales, rows are independent, and sampled rows agree with the oracle bit for bit."""
    from lofloat_b200 import _lib
    _lib.set_option("mx.variant", variant)
    try:
        R = Ccols = 16384
        g = torch.Generator(device=cuda_dev).manual_seed(7)
        x = torch.randn(R, Ccols, device=cuda_dev, generator=g)
        x *= torch.exp2(torch.randint(-8, 9, (R, 1), device=cuda_dev, generator=g).float())
        e8 = lof.create_e8m0_params()
        for en in ("e4m3_b7", "e2m1"):
            d = _desc(lof, L.FORMATS[en])
            y = lof.virtual_mx_round(x, 32, d, e8)
            for r0 in (0, 5000, R - 4):
                want = L.o_mx_round(_np(x[r0:r0 + 4]).reshape(-1), 32, L.FORMATS["e8m0"], L.FORMATS[en])
                assert L.bits_equal(_np(y[r0:r0 + 4]).reshape(-1), want), (en, r0)
            # row shards computed separately (first_index = row offset * row length) == whole-tensor result
            ys = lof.virtual_mx_round(x[4096:8192], 32, d, e8, first_index=4096 * Ccols)
            assert torch.equal(ys, y[4096:8192])
            # |y| never exceeds the block's representable maximum and zero stays zero
            assert torch.equal(y == 0, (y == 0) | (x == 0))
            del y, ys
    finally:
        _lib.set_option("mx.variant", 0)


def test_fast_kernel_equals_exact_kernel_every_fp32_pattern(lof, cuda_dev):
    """The branch-free fp32 kernels (lof_fast_f32.cuh) against the exact integer kernel (round.fast = 0) on EVERY
    fp32 bit pattern, every named format, all 15 modes, with and without the fused scale: round.fast = 2 (the
    default: magic-add path for the six IEEE-like modes, first-generation fast path for the others) and, for
    those six modes, also round.fast = 1.  Compared on the device; the exact kernel itself is pinned to the
    oracle above."""
    from lofloat_b200 import _lib
    chunk = 1 << 28
    checked = 0
    magic_modes = (L.RNE, L.RTZ, L.RAZ, L.RU, L.RD, L.RTO)

    def run(x, d, mode, sl, scale, base, fast):
        _lib.set_option("round.fast", fast)
        return lof.virtual_round(x, d, mode, sl, scale, rng=(17, 3), first_index=base)

    def check(a, b, what, base):
        if not torch.equal(a.view(torch.int32), b.view(torch.int32)):
            idx = int((a.view(torch.int32) != b.view(torch.int32)).nonzero()[0])
            raise AssertionError(f"{what}: in=0x{(base + idx) & 0xFFFFFFFF:08x} "
                                 f"fast=0x{int(a.view(torch.int32)[idx]) & 0xFFFFFFFF:08x} "
                                 f"exact=0x{int(b.view(torch.int32)[idx]) & 0xFFFFFFFF:08x}")
    try:
        for name, f in L.FORMATS.items():
            d = _desc(lof, f)
            r = 23 - f.mantissa_bits
            for mode in range(15):
                sl = min(7, max(r, 1)) if mode in (L.SR_A, L.SR_B, L.SR_C, L.SR_D) else 0
                for base in range(0, 1 << 32, chunk):
                    bits = torch.arange(base, base + chunk, device=cuda_dev, dtype=torch.int64).to(torch.int32)
                    x = bits.view(torch.float32)
                    b = run(x, d, mode, sl, 1.0, base, 0)
                    check(run(x, d, mode, sl, 1.0, base, 2), b, f"{name} {L.MODE_NAMES[mode]} fast=2", base)
                    if mode in magic_modes:
                        check(run(x, d, mode, sl, 1.0, base, 1), b, f"{name} {L.MODE_NAMES[mode]} fast=1", base)
                    checked += chunk
                    del bits, x, b
        # fused scale (scale * x is formed inside the kernel)
        for name in ("e4m3_b7", "bf16_like"):
            d = _desc(lof, L.FORMATS[name])
            for mode in (L.RNE, L.RAZ, L.RNO, L.SR_A, L.RD):
                sl = 4 if mode == L.SR_A else 0
                for base in range(0, 1 << 32, chunk):
                    x = torch.arange(base, base + chunk, device=cuda_dev, dtype=torch.int64).to(torch.int32).view(torch.float32)
                    b = run(x, d, mode, sl, 0.3, 0, 0)
                    check(run(x, d, mode, sl, 0.3, 0, 2), b, f"{name} {mode} scaled fast=2", base)
                    check(run(x, d, mode, sl, 0.3, 0, 1), b, f"{name} {mode} scaled fast=1", base)
                    del x, b
    finally:
        _lib.set_option("round.fast", 2)
    assert checked == len(L.FORMATS) * 15 * (1 << 32)


@pytest.mark.slow
def test_exhaustive_fp32_sweep_8bit_formats(lof, cuda_dev):
    """Every fp32 bit pattern (2^32) for the 8-bit benchmark format under RNE, plus every 16th
    pattern for the other <= 8-bit formats and deterministic modes (north star: exhaustive check for
    <= 8-bit formats).  The full cross product -- all 2^32 patterns x 7 formats x the 9 deterministic modes against the
    COMPILED REFERENCE on the host cores -- is run by scripts/exhaustive_sweep.py (7.4 minutes on a B200 box with 16 host threads); its
    log is committed as profiles/r02_exhaustive_sweep.log (270 582 939 648 comparisons, 0 mismatches)."""
    chunk = 1 << 26
    def sweep(fname, mode, stride):
        f = L.FORMATS[fname]
        d = _desc(lof, f)
        for base in range(0, 1 << 32, chunk * stride):
            bits = torch.arange(base, base + chunk * stride, stride, device=cuda_dev, dtype=torch.int64).to(torch.int32)
            got = _np(lof.virtual_round(bits.view(torch.float32), d, mode))
            xb = _np(bits).view(np.float32)
            want = L.o_round_params(xb, f, mode)
            assert L.bits_equal(got, want), (fname, L.MODE_NAMES[mode], hex(base), L.first_mismatch(xb, got, want))
    sweep("e4m3_b7", L.RNE, 1)
    for fname in ("p3109_8p4", "e5m2", "e3m2", "e2m3", "e2m1", "e8m0"):
        for mode in L.DETERMINISTIC_MODES:
            sweep(fname, mode, 64)


# ------------------------------------------------------------------------------ storage codec (lo_pack.hpp), row f-1
def _abi_encode(cuda_dev, x, f, mode=0, sat=0, stoch_len=0, endian=0, seed=0, offset=0, first_index=0, misalign=0):
    from lofloat_b200 import _lib
    lib = _lib.lib()
    xt = _gpu(x, cuda_dev)
    nbytes = L.packed_size(f.fmt.bitwidth, x.size)
    buf = torch.full((nbytes + 8,), 0xAB, dtype=torch.uint8, device=cuda_dev)
    ps, rng = _lib.CProjSpec(mode, sat, stoch_len), _lib.CRng(seed, offset)
    st = lib.lof_encode_f32(xt.data_ptr(), buf.data_ptr() + misalign, x.size, C.cast(C.byref(f), C.POINTER(_lib.CCodecFormat)),
                            C.byref(ps), endian, C.byref(rng), first_index, torch.cuda.current_stream().cuda_stream)
    assert st == 0, lib.lof_last_error_string().decode()
    out = _np(buf)
    assert np.all(out[:misalign] == 0xAB) and np.all(out[misalign + nbytes:] == 0xAB)        # nothing written outside the packed range
    return out[misalign:misalign + nbytes]


def _abi_decode(cuda_dev, packed, count, f, endian=0):
    from lofloat_b200 import _lib
    lib = _lib.lib()
    pt = _gpu(packed, cuda_dev)
    out = torch.empty(count, dtype=torch.float32, device=cuda_dev)
    st = lib.lof_decode_f32(pt.data_ptr(), out.data_ptr(), count, C.cast(C.byref(f), C.POINTER(_lib.CCodecFormat)), endian,
                            torch.cuda.current_stream().cuda_stream)
    assert st == 0, lib.lof_last_error_string().decode()
    return _np(out)


def test_pack_codec_bit_exact_all_formats_modes(lof, cuda_dev):
    """lof_encode_f32 / lof_decode_f32 through the C ABI against the oracle (itself pinned to the compiled reference's
    lo_pack.hpp): every codec format of test/test_pack.cpp (+ MX element / scale formats) x every deterministic mode
    x every saturation mode x both slot orders on the probe set; decode over every representation; ragged counts
    (partial groups, partial 32-bit words, tiles), misaligned destination; committed golden bytes."""
    g = np.load(G / "pack_codec.npz")
    for name, f in L.CODEC_FORMATS.items():
        x = L.codec_probe_floats(f)
        n = f.fmt.bitwidth
        for mode in L.DETERMINISTIC_MODES:
            for sat in ((0, 1, 2) if mode == L.RNE else (0,)):
                for endian in (0, 1):
                    got = _abi_encode(cuda_dev, x, f, mode, sat, 0, endian)
                    assert np.array_equal(got, L.o_encode(x, f, mode, sat, 0, endian)), (name, L.MODE_NAMES[mode], sat, endian)
        reps = np.arange(1 << n, dtype=np.uint32)
        for endian in (0, 1):
            packed = L.o_pack_reps(reps, n, endian)
            assert L.bits_equal(_abi_decode(cuda_dev, packed, reps.size, f, endian), L.o_rep_to_float(reps, f)), (name, endian)
        # every representation again over several whole tiles + a ragged tail: the register-lane decode kernels (pack.fast = 2,
        # 4 / 6 / 8 / 16-bit formats: arithmetic decode, infinity / NaN codes through the rare path) and the table kernel (pack.fast = 1)
        from lofloat_b200 import _lib as _l
        many = np.resize(np.roll(reps, 5), 5 * 4096 + 37).astype(np.uint32)
        try:
            for gen in (2, 1):
                _l.set_option("pack.fast", gen)
                for endian in (0, 1):
                    packed = L.o_pack_reps(many, n, endian)
                    assert L.bits_equal(_abi_decode(cuda_dev, packed, many.size, f, endian), L.o_rep_to_float(many, f)), (name, endian, gen)
        finally:
            _l.set_option("pack.fast", 2)
        for count in (1, 3, 15, 16, 17, 1001, 4095, 4096, 4097, 3 * 4096 + 5):
            xs_ = np.resize(x, count)
            want = L.o_encode(xs_, f, endian=count & 1)
            assert np.array_equal(_abi_encode(cuda_dev, xs_, f, endian=count & 1), want), (name, count)
            assert np.array_equal(_abi_encode(cuda_dev, xs_, f, endian=count & 1, misalign=1), want), (name, count, "misaligned")
            assert L.bits_equal(_abi_decode(cuda_dev, want, count, f, count & 1), L.o_decode(want, count, f, count & 1)), (name, count)
        assert np.array_equal(_abi_encode(cuda_dev, g[f"{name}__x"][:1001], f, endian=1), g[f"{name}__packed1"]), name
    # stochastic modes draw the same Philox event words as the oracle (global element index keyed)
    f = L.CODEC_FORMATS["e4m3fn"]
    x = L.wide_random_f32(50000, 3) * np.float32(2.0 ** -3)
    for mode, sl in ((L.SR_A, 4), (L.SR_B, 3), (L.SR_C, 5), (L.SR_D, 6), (L.TRUE_SR, 0), (L.PROB, 0)):
        got = _abi_encode(cuda_dev, x, f, mode, 0, sl, 0, seed=9, offset=2, first_index=4096)
        assert np.array_equal(got, L.o_encode(x, f, mode, 0, sl, 0, seed=9, offset=2, first_index=4096)), L.MODE_NAMES[mode]


def test_pack_python_surface_and_full_size_properties(lof, cuda_dev):
    """lof.lof_encode / lof.lof_decode (checker callables folded into masks) == oracle for the factory formats; at
    2^26 elements: decode(encode(x)) is idempotent under re-encoding (the codec is a projection), equals the oracle on
    a sampled window, and the packed size is exact."""
    pairs = ((lof.create_p3109_params(8, 4), L.CODEC_FORMATS["p3109_8p4"]),
             (lof.create_p3109_params(8, 4, saturating=False), L.CODEC_FORMATS["p3109_8p4_ext"]),
             (lof.create_p3109_params(5, 2), L.CODEC_FORMATS["p3109_5p2"]),
             (lof.create_p3109_params(11, 5), L.CODEC_FORMATS["p3109_11p5"]),
             (lof.create_half_params(), L.codec_half()))
    for d, f in pairs:
        x = L.codec_probe_floats(f)
        for endian in (lof.PackEndian.Forward, lof.PackEndian.Reverse):
            packed = lof.lof_encode(_gpu(x, cuda_dev), d, endian=endian)
            assert np.array_equal(_np(packed), L.o_encode(x, f, endian=int(endian))), (d, endian)
            back = lof.lof_decode(packed, x.size, d, endian)
            assert L.bits_equal(_np(back), L.o_decode(_np(packed), x.size, f, int(endian))), (d, endian)
    n = 1 << 26
    g = torch.Generator(device=cuda_dev).manual_seed(3)
    x = torch.randn(n, device=cuda_dev, generator=g) * torch.exp2(torch.randint(-10, 9, (n,), device=cuda_dev, generator=g).float())
    for d, f in ((lof.create_p3109_params(8, 4), L.CODEC_FORMATS["p3109_8p4"]), (lof.create_p3109_params(6, 3, bias=3), None),
                 (lof.create_p3109_params(4, 2, bias=1), None)):
        packed = lof.lof_encode(x, d)
        assert packed.numel() == lof.packed_size(d.total_bits, n)
        y = lof.lof_decode(packed, n, d)
        assert torch.equal(lof.lof_encode(y, d), packed)                                   # projection: re-encoding is the identity
        if f is not None:
            lo = 12345 * 16
            assert np.array_equal(_np(packed[lo:lo + 4096]), L.o_encode(_np(x[lo:lo + 4096]), f))
    with pytest.raises(RuntimeError):
        lof.lof_encode(x.double(), lof.create_p3109_params(8, 4))
    with pytest.raises(ValueError):
        lof.lof_encode(x, lof.create_single_params())                                       # 32-bit formats cannot be packed (lo_pack.hpp:4-9)


def test_gemm_on_packed_operands_equals_fp32_container_path(lof, cuda_dev):
    """lof_gemm_chunked_packed: A and B^T stored as e4m3 bytes by lof_encode -> the same C, bit for bit, as the
    fp32-container entry point (which converts inside the call)."""
    from lofloat_b200 import formats
    A, B = _e4m3_uniform((512, 2048), 13), _e4m3_uniform((2048, 384), 14)
    ocp = lof.FloatFormatDescriptor(8, 3, 7, lof.InfBehavior.Saturating, lof.NaNBehavior._3109, lof.Signedness.Signed,
                                    formats._P3109Inf(8, True, True), formats._BitPatternNaN(0xFF))   # param_float8_e4m3fn
    Ap = lof.lof_encode(_gpu(A, cuda_dev), ocp).view(512, 2048)
    Btp = lof.lof_encode(_gpu(np.ascontiguousarray(B.T), cuda_dev), ocp).view(384, 2048)
    e5m10 = _desc(lof, L.fmt(16, 10, 15, L.SATURATING, L.SIGNED))
    for kw, mant in ((dict(k_chunk=32, accum_format=e5m10), 0), (dict(k_chunk=32), 10), (dict(k_chunk=0), 10),
                     (dict(k_chunk=128), 7), (dict(k_chunk=32, partial_round=True), 10)):
        want = lof.lof_gemm(_gpu(A, cuda_dev), _gpu(B, cuda_dev), mant, **kw)
        got = lof.lof_gemm_packed(Ap, Btp, mant, **kw)
        assert torch.equal(got, want), kw
    # bf16 words as the packed operand type (how MX-rounded operands travel between GPUs, sharding.gemm_2d_packed)
    Ab, Btb = _gpu(A, cuda_dev).bfloat16(), _gpu(np.ascontiguousarray(B.T), cuda_dev).bfloat16()
    for kw, mant in ((dict(k_chunk=32, partial_round=True), 7), (dict(k_chunk=64), 7), (dict(k_chunk=0), 10)):
        want = lof.lof_gemm(_gpu(A, cuda_dev), _gpu(B, cuda_dev), mant, operand_dtype="bf16", **kw)
        assert torch.equal(lof.lof_gemm_packed(Ab, Btb, mant, **kw), want), ("bf16", kw)
    with pytest.raises(RuntimeError):
        lof.lof_gemm_packed(Ab, Btp, 7)


def test_pack_fast_path_equals_exact_every_fp32_pattern(lof, cuda_dev):
    """The magic-add encode paths (pack.fast = 2, default: register-lane kernels for 4 / 6 / 8 / 16-bit formats, the
    second-generation kernel for the other widths; pack.fast = 1: second-generation kernel everywhere) against the exact
    integer routine (pack.fast = 0, pinned to the oracle / compiled reference above) on EVERY fp32 bit pattern, for every
    eligible codec format, the four IEEE-like modes and all three saturation modes under RNE.  Compared on the device."""
    from lofloat_b200 import _lib
    lib = _lib.lib()
    chunk = 1 << 28
    stream = torch.cuda.current_stream().cuda_stream
    checked = 0

    def enc(x, f, mode, sat, out):
        ps = _lib.CProjSpec(mode, sat, 0)
        st = lib.lof_encode_f32(x.data_ptr(), out.data_ptr(), x.numel(), C.cast(C.byref(f), C.POINTER(_lib.CCodecFormat)),
                                C.byref(ps), 0, None, 0, stream)
        assert st == 0, lib.lof_last_error_string().decode()
    try:
        for name in ("e2m1", "e3m2", "e2m3", "p3109_5p2", "e5m2", "e4m3fn", "p3109_8p4", "p3109_8p4_ext", "p3109_8p3_unsigned",
                     "p3109_8p3_unsigned_ext_nan", "p3109_11p5", "half"):
            f = L.CODEC_FORMATS[name]
            nbytes = L.packed_size(f.fmt.bitwidth, chunk)
            a = torch.empty(nbytes, dtype=torch.uint8, device=cuda_dev)
            b = torch.empty(nbytes, dtype=torch.uint8, device=cuda_dev)
            for mode, sat in ((L.RNE, 0), (L.RNE, 1), (L.RNE, 2), (L.RTZ, 0), (L.RAZ, 0), (L.RTO, 0)):
                for base in range(0, 1 << 32, chunk):
                    x = torch.arange(base, base + chunk, device=cuda_dev, dtype=torch.int64).to(torch.int32).view(torch.float32)
                    _lib.set_option("pack.fast", 0)
                    enc(x, f, mode, sat, b)
                    for gen in ((2, 1) if (name in ("e4m3fn", "e2m1", "e3m2", "half") and mode == L.RNE and sat == 0) else (2,)):
                        _lib.set_option("pack.fast", gen)
                        enc(x, f, mode, sat, a)
                        if not torch.equal(a, b):
                            idx = int((a != b).nonzero()[0])
                            raise AssertionError(f"{name} {L.MODE_NAMES[mode]} sat{sat} pack.fast={gen}: packed byte {idx} of chunk 0x{base:08x} "
                                                 f"fast=0x{int(a[idx]):02x} exact=0x{int(b[idx]):02x}")
                    checked += chunk
                    del x
            del a, b
    finally:
        _lib.set_option("pack.fast", 2)
    assert checked == 12 * 6 * (1 << 32)


def test_strict_lane_kernel_equals_exact_every_fp32_pattern(lof, cuda_dev):
    """lof_round_strict_f32 on the register-lane kernel (pack.fast = 2: magic-add encode + arithmetic decode, overflow / negative
    codes as launch constants) against the exact integer routine + table decode (pack.fast = 0, pinned to the oracle in
    test_strict_typed_rounding_op) on EVERY fp32 bit pattern: signed / unsigned / no -0 formats, with and without infinities,
    widths 4 .. 16, the four IEEE-like modes, all saturation modes under RNE.  Compared on the device, NaN == NaN by bit pattern."""
    from lofloat_b200 import _lib
    lib = _lib.lib()
    chunk = 1 << 28
    stream = torch.cuda.current_stream().cuda_stream
    a = torch.empty(chunk, dtype=torch.float32, device=cuda_dev)
    b = torch.empty(chunk, dtype=torch.float32, device=cuda_dev)

    def strict(x, f, mode, sat, out):
        ps, rng = _lib.CProjSpec(mode, sat, 0), _lib.CRng(0, 0)
        st = lib.lof_round_strict_f32(x.data_ptr(), out.data_ptr(), x.numel(), C.cast(C.byref(f), C.POINTER(_lib.CCodecFormat)),
                                      C.byref(ps), C.byref(rng), 0, stream)
        assert st == 0, lib.lof_last_error_string().decode()
    checked = 0
    try:
        for name in ("e2m1", "e3m2", "p3109_5p2", "e5m2", "e4m3fn", "p3109_8p4", "p3109_8p4_ext", "p3109_8p3_unsigned",
                     "p3109_8p3_unsigned_ext_nan", "p3109_11p5", "half"):
            f = L.CODEC_FORMATS[name]
            for mode, sat in ((L.RNE, 0), (L.RNE, 1), (L.RNE, 2), (L.RTZ, 0), (L.RAZ, 0), (L.RTO, 0)):
                for base in range(0, 1 << 32, chunk):
                    x = torch.arange(base, base + chunk, device=cuda_dev, dtype=torch.int64).to(torch.int32).view(torch.float32)
                    _lib.set_option("pack.fast", 0)
                    strict(x, f, mode, sat, b)
                    _lib.set_option("pack.fast", 2)
                    strict(x, f, mode, sat, a)
                    if not torch.equal(a.view(torch.int32), b.view(torch.int32)):
                        idx = int((a.view(torch.int32) != b.view(torch.int32)).nonzero()[0])
                        raise AssertionError(f"{name} {L.MODE_NAMES[mode]} sat{sat}: x bits 0x{base + idx:08x} lanes={float(a[idx])!r} exact={float(b[idx])!r}")
                    checked += chunk
                    del x
    finally:
        _lib.set_option("pack.fast", 2)
    assert checked == 11 * 6 * (1 << 32)


def test_strict_typed_rounding_op(lof, cuda_dev):
    """Row f-2: lof_round_strict_f32 == decode(encode(x)) of the oracle codec (pinned to the compiled reference's lo_pack.hpp,
    itself validated against Templated_Float<Fp> by test/test_pack.cpp) for every codec format x deterministic mode x
    saturation mode, stochastic modes on the oracle's Philox stream, the behaviours test/test_sat_mode.cpp:11-18 spells
    out, and agreement with the committed typed ground truth."""
    from lofloat_b200 import _lib
    lib = _lib.lib()

    def strict(x, f, mode=0, sat=0, sl=0, seed=0, offset=0, first_index=0):
        xt = _gpu(x, cuda_dev)
        out = torch.empty_like(xt)
        ps, rng = _lib.CProjSpec(mode, sat, sl), _lib.CRng(seed, offset)
        st = lib.lof_round_strict_f32(xt.data_ptr(), out.data_ptr(), x.size, C.cast(C.byref(f), C.POINTER(_lib.CCodecFormat)),
                                      C.byref(ps), C.byref(rng), first_index, torch.cuda.current_stream().cuda_stream)
        assert st == 0, lib.lof_last_error_string().decode()
        return _np(out)
    for name, f in L.CODEC_FORMATS.items():
        x = L.codec_probe_floats(f)
        for mode in L.DETERMINISTIC_MODES:
            for sat in ((0, 1, 2) if mode in (L.RNE, L.RTZ) else (0,)):
                want = L.o_rep_to_float(L.o_float_to_rep(x, f, mode, sat), f)
                assert L.bits_equal(strict(x, f, mode, sat), want), (name, L.MODE_NAMES[mode], sat)
    f = L.CODEC_FORMATS["e5m2"]
    x = L.wide_random_f32(40001, 8)[:40001]
    for mode, sl in ((L.SR_A, 7), (L.SR_D, 5), (L.TRUE_SR, 0), (L.PROB, 0)):
        want = L.o_rep_to_float(L.o_float_to_rep(x, f, mode, 0, sl, seed=4, offset=1, first_index=64), f)
        assert L.bits_equal(strict(x, f, mode, 0, sl, seed=4, offset=1, first_index=64), want), L.MODE_NAMES[mode]
    # test_sat_mode.cpp:11-18: finite overflow -> inf under OvfInf, max finite under SatFinite / SatPropagate; inf input -> inf
    # under OvfInf / SatPropagate, max finite under SatFinite (formats with infinities); saturating formats always clamp
    ext, sat8 = L.CODEC_FORMATS["p3109_8p4_ext"], L.CODEC_FORMATS["p3109_8p4"]
    big = np.array([1e30, -1e30, np.inf, -np.inf], np.float32)
    mx = float(L.o_rep_to_float(np.array([0x7E], np.uint32), ext)[0])
    assert np.array_equal(strict(big, ext, 0, 0), np.array([np.inf, -np.inf, np.inf, -np.inf], np.float32))
    assert np.array_equal(strict(big, ext, 0, 1), np.array([mx, -mx, mx, -mx], np.float32))
    assert np.array_equal(strict(big, ext, 0, 2), np.array([mx, -mx, np.inf, -np.inf], np.float32))
    smx = float(L.o_rep_to_float(np.array([0x7F], np.uint32), sat8)[0])
    for sat in (0, 1, 2):
        assert np.array_equal(strict(big, sat8, 0, sat), np.array([smx, -smx, smx, -smx], np.float32))
    # negatives into unsigned formats: NegtoZero / NegtoNaN (test/test_unsigned.cpp)
    neg = np.array([-1.0, -0.0, -np.inf, 2.0], np.float32)
    got = strict(neg, L.CODEC_FORMATS["p3109_8p3_unsigned"])
    assert np.array_equal(got[:3], np.zeros(3, np.float32)) and got[3] == 2.0
    got = strict(neg, L.CODEC_FORMATS["p3109_8p3_unsigned_ext_nan"])
    assert np.all(np.isnan(got[:3])) and got[3] == 2.0
    # typed ground truth (Templated_Float<Fp>) committed with the golden vectors
    g = np.load(G / "pack_codec.npz")
    for tag, key in (("typed_e4m3fn", "e4m3fn"), ("typed_p3109_8p4", "p3109_8p4")):
        x, back = g[f"{tag}__x"], g[f"{tag}__back"]
        got = strict(x, L.CODEC_FORMATS[key])
        keep = ~np.isnan(back)
        assert L.bits_equal(got[keep], back[keep]) and np.all(np.isnan(got[~keep])), tag
    # Python surface + full size: idempotent projection at 2^28 elements
    d = lof.create_p3109_params(8, 4, saturating=False)
    xt = torch.randn(1 << 28, device=cuda_dev, generator=torch.Generator(device=cuda_dev).manual_seed(1887)) * 37.0
    xt[:4] = torch.tensor([1e9, -1e9, 2.0 * mx, -2.0 * mx], device=cuda_dev)     # overflow is present whatever the draw
    y = lof.virtual_round_strict(xt, d, saturation_mode=lof.SaturationMode.SatFinite)
    assert torch.equal(lof.virtual_round_strict(y, d, saturation_mode=lof.SaturationMode.SatFinite), y)
    assert float(y.abs().max()) == mx                       # nothing exceeds the largest finite value ...
    assert y[:4].tolist() == [mx, -mx, mx, -mx]              # ... and overflow clamps to it with its sign


def test_ste_backward_and_cuda_graph_capture(lof, cuda_dev):
    """Row f-4: the straight-through wrappers (layers.py:8-57) -- forward == the op, backward = grad * scale / identity --
    and CUDA-graph safety: rounding, MX, strict rounding and the storage codec captured once and replayed on new data
    give the eager result bit for bit (single async launches on the capturing stream, no allocation, no sync)."""
    from lofloat_b200 import _lib
    e4m3, e8m0 = lof.create_p3109_params(8, 4, bias=7), lof.create_e8m0_params()
    x = torch.randn(64, 256, device=cuda_dev, requires_grad=True)
    y = lof.STERound.apply(x, e4m3, lof.RoundingMode.RoundToNearestEven, 4.0)
    assert torch.equal(y, lof.virtual_round(x.detach(), e4m3, scale=4.0))
    g = torch.randn_like(y)
    y.backward(g)
    assert torch.equal(x.grad, g * 4.0)
    x.grad = None
    fid = lof.register_format(e4m3)
    y = lof.STERound.apply(x, fid)
    assert torch.equal(y, lof.virtual_round(x.detach(), e4m3))
    y.sum().backward()
    assert torch.equal(x.grad, torch.ones_like(x))
    x.grad = None
    y = lof.STEMXRound.apply(x, e4m3, e8m0, 32)
    assert torch.equal(y, lof.virtual_mx_round(x.detach(), 32, e4m3, e8m0))
    y.backward(g)
    assert torch.equal(x.grad, g)
    # CUDA graph: raw C-ABI launches on the capturing stream into static buffers
    lib = _lib.lib()
    n = 1 << 20
    xin = torch.randn(n, device=cuda_dev)
    out_r, out_m, out_s = torch.empty_like(xin), torch.empty_like(xin), torch.empty_like(xin)
    out_p = torch.empty(n, dtype=torch.uint8, device=cuda_dev)
    fmt, fe8 = e4m3._c(), e8m0._c()
    cf = lof.codec_format(e4m3)
    ps = _lib.CProjSpec(0, 0, 0)

    def launch_all(stream):
        assert lib.lof_round_params(xin.data_ptr(), out_r.data_ptr(), n, 0, C.byref(fmt), C.byref(ps), 1.0, None, 0, stream) == 0
        assert lib.lof_mx_round_f32(xin.data_ptr(), out_m.data_ptr(), n, 32, C.byref(fe8), C.byref(fmt), C.byref(ps), C.byref(ps), None, 0, stream) == 0
        assert lib.lof_round_strict_f32(xin.data_ptr(), out_s.data_ptr(), n, C.byref(cf), C.byref(ps), None, 0, stream) == 0
        assert lib.lof_encode_f32(xin.data_ptr(), out_p.data_ptr(), n, C.byref(cf), C.byref(ps), 0, None, 0, stream) == 0
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        launch_all(s.cuda_stream)                                                # warm-up outside capture
    s.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=s):
        launch_all(torch.cuda.current_stream().cuda_stream)
    for seed in (1, 2):
        xin.copy_(torch.randn(n, device=cuda_dev, generator=torch.Generator(device=cuda_dev).manual_seed(seed)) * 3)
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(out_r, lof.virtual_round(xin, e4m3))
        assert torch.equal(out_m, lof.virtual_mx_round(xin, 32, e4m3, e8m0))
        assert torch.equal(out_s, lof.virtual_round_strict(xin, e4m3))
        assert torch.equal(out_p, lof.lof_encode(xin, e4m3))


def test_mx_storage_encode_decode(lof, cuda_dev):
    """MX quantisation with materialised E8M0 scale bytes + packed MXFP8 / MXFP6 / MXFP4 elements (row f-1, second half):
    bit-exact against the oracle composition (virtual_mx_round's scale / element rounding + lo_pack.hpp codec, both pinned
    to the compiled reference) for every block size, element format, slot order, deterministic element / scale modes and
    stochastic element rounding; decode(encode(x)) == virtual_mx_round(x) wherever the fake-quant value is representable;
    16384 x 16384 (C3 shape) round trip."""
    e8 = L.codec_e8m0()
    elem = {"mxfp8_e4m3": (L.codec_p3109(8, 4, bias=7), lof.create_p3109_params(8, 4, bias=7), L.FORMATS["e4m3_b7"]),
            "mxfp8_e5m2": (L.codec_p3109(8, 3, bias=15), lof.create_p3109_params(8, 3, bias=15), L.FORMATS["e5m2"]),
            "mxfp6_e3m2": (L.codec_p3109(6, 3, bias=3), lof.create_p3109_params(6, 3, bias=3), L.FORMATS["e3m2"]),
            "mxfp6_e2m3": (L.codec_p3109(6, 4, bias=1), lof.create_p3109_params(6, 4, bias=1), L.FORMATS["e2m3"]),
            "mxfp4_e2m1": (L.codec_p3109(4, 2, bias=1), lof.create_p3109_params(4, 2, bias=1), L.FORMATS["e2m1"])}
    from lofloat_b200 import formats
    d8 = lof.FloatFormatDescriptor(8, 0, 127, lof.InfBehavior.Saturating, lof.NaNBehavior._3109, lof.Signedness.Unsigned,
                                   formats._P3109Inf(8, False, True), formats._BitPatternNaN(0xFF))          # param_ocp_e8m0
    rng = np.random.default_rng(21)
    n = 128 * 300
    x = (rng.standard_normal(n) * np.exp2(rng.integers(-20, 21, n))).astype(np.float32)
    x[:128] = 0.0
    x[128:256] = 2.0 ** -140                                                              # fp32-subnormal block
    x[300] = np.float32(3e38)
    x[128 * 7 + 5] = np.nan                                                              # a NaN element: dropped from the block maximum, encodes as the NaN code
    x[128 * 9 + 17] = np.inf                                                             # infinite elements: the block's scale saturates
    x[128 * 11 + 3] = -np.inf
    x[128 * 13 + 1], x[128 * 13 + 2] = np.inf, -np.inf                                   # both signs in one block
    xt = _gpu(x, cuda_dev)
    from lofloat_b200 import _lib as _l
    try:
        for gen in (2, 3, 1):       # block-per-thread encode + lane decode (default) / warp-step lane encode / second-generation kernels
            _l.set_option("pack.fast", gen)
            for name, (cf, d, ff) in elem.items():
                for bs in (16, 32, 64, 128):
                    for endian in (0, 1):
                        sc, el = lof.mx_encode(xt, bs, d, d8, endian=endian)
                        wsc, wel = L.o_mx_encode(x, bs, e8, cf, endian=endian)
                        assert np.array_equal(_np(sc), wsc), (gen, name, bs, endian, "scales")
                        assert np.array_equal(_np(el), wel), (gen, name, bs, endian, "elements")
                        back = _np(lof.mx_decode(sc, el, n, bs, d, d8, endian))
                        assert L.bits_equal(back, L.o_mx_decode(wsc, wel, n, bs, e8, cf, endian)), (gen, name, bs, endian, "decode")
    finally:
        _l.set_option("pack.fast", 2)
    for name, (cf, d, ff) in elem.items():
        for mode, smode, sl in ((L.RTZ, L.RNE, 0), (L.RAZ, L.RU, 0), (L.RTO, L.RD, 0), (L.RNE, L.RAZ, 0), (L.SR_A, L.RNE, 6), (L.TRUE_SR, L.RNE, 0)):
            sc, el = lof.mx_encode(xt, 32, d, d8, mode, smode, sl, rng=(5, 7))
            wsc, wel = L.o_mx_encode(x, 32, e8, cf, mode, smode, sl, seed=5, offset=7)
            assert np.array_equal(_np(sc), wsc) and np.array_equal(_np(el), wel), (name, L.MODE_NAMES[mode], L.MODE_NAMES[smode])
        # against the fake-quant op: identical wherever virtual_mx_round's value is representable in the element format
        xr = x.copy()
        xr[128:256] = 1.5                                  # (a 2^-127 scale has no nonzero E8M0 code in lo_pack.hpp's decoder: exp field 0 reads as 0)
        xr[~np.isfinite(xr)] = 1.0                         # (NaN / infinite elements are covered by the byte comparison above)
        xrt = _gpu(xr, cuda_dev)
        sc, el = lof.mx_encode(xrt, 32, d, d8, scale_round_mode=lof.RoundingMode.RoundUp)
        fake = _np(lof.virtual_mx_round(xrt, 32, d, d8, scale_round_mode=lof.RoundingMode.RoundUp))
        back = _np(lof.mx_decode(sc, el, n, 32, d, d8))
        assert np.array_equal(back, fake), name            # a scale rounded UP never lets x / scale overflow -> always representable
                                                           # (value compare: the P3109-style element formats have no -0, the fake-quant op keeps it)
    # C3 shape: 16384 x 16384, MXFP8 and MXFP4
    g = torch.Generator(device=cuda_dev).manual_seed(7)
    big = torch.randn(16384, 16384, device=cuda_dev, generator=g) * torch.exp2(torch.randint(-8, 9, (16384, 1), device=cuda_dev, generator=g).float())
    for name in ("mxfp8_e4m3", "mxfp4_e2m1"):
        cf, d, ff = elem[name]
        sc, el = lof.mx_encode(big, 32, d, d8, scale_round_mode=lof.RoundingMode.RoundUp)
        assert sc.numel() == big.numel() // 32 and el.numel() == lof.packed_size(d.total_bits, big.numel())
        back = lof.mx_decode(sc, el, big.numel(), 32, d, d8).view_as(big)
        assert torch.equal(back, lof.virtual_mx_round(big, 32, d, d8, scale_round_mode=lof.RoundingMode.RoundUp)), name
        rows = slice(5000, 5002)
        wsc, wel = L.o_mx_encode(_np(big[rows]).ravel(), 32, e8, cf, scale_mode=L.RU)
        lo = 5000 * 16384
        assert np.array_equal(_np(sc[lo // 32:(lo + 2 * 16384) // 32]), wsc), name
    with pytest.raises(lof.LoFloatError):
        lof.mx_encode(xt[:100], 32, elem["mxfp8_e4m3"][1], d8)                            # n % block_size != 0


def test_bf16_container_fast_path_equals_exact_all_patterns(lof, cuda_dev):
    """bf16 containers run on the fp32 branch-free routines applied to (bits << 16) (round.fast = 1 / 2); they must equal
    the exact bf16-container routine (round.fast = 0, pinned to the oracle by test_round_other_containers) on ALL 65536
    bf16 bit patterns for every named format and deterministic mode, at every vector alignment."""
    from lofloat_b200 import _lib
    allbits = torch.arange(0, 65536, device=cuda_dev, dtype=torch.int32).to(torch.int16).view(torch.bfloat16).repeat(9)[3:]
    try:
        for name, f in L.FORMATS.items():
            if f.bitwidth > 16:
                continue
            d = _desc(lof, f)
            for mode in L.DETERMINISTIC_MODES:
                _lib.set_option("round.fast", 0)
                want = lof.virtual_round(allbits, d, mode).view(torch.int16)
                for fast in (1, 2):
                    _lib.set_option("round.fast", fast)
                    assert torch.equal(lof.virtual_round(allbits, d, mode).view(torch.int16), want), (name, L.MODE_NAMES[mode], fast)
            r16 = 7 - f.mantissa_bits
            for mode in (L.SR_A, L.SR_D):                      # stochastic: same Philox events, same offsets in both containers
                for sl in {1, max(1, min(r16, 4)), max(1, r16 - (1 if mode == L.SR_D else 0))}:
                    if r16 < 1 or sl > r16:
                        continue
                    _lib.set_option("round.fast", 0)
                    want = lof.virtual_round(allbits, d, mode, sl, rng=(77, 5), first_index=8).view(torch.int16)
                    _lib.set_option("round.fast", 2)
                    got = lof.virtual_round(allbits, d, mode, sl, rng=(77, 5), first_index=8).view(torch.int16)
                    assert torch.equal(got, want), (name, L.MODE_NAMES[mode], sl)
    finally:
        _lib.set_option("round.fast", 2)


def test_fp16_container_fast_path_equals_exact_all_patterns(lof, cuda_dev):
    """fp16 containers (a reference dtype: AT_DISPATCH_FLOATING_TYPES_AND_HALF, LoPy_bind.cpp:83,133) take the fp32 magic-add
    routine on the widened values where the host proves every result is an fp16 value again; it must equal the exact
    fp16-container routine (round.fast = 0, pinned to the oracle by test_round_other_containers) on ALL 65536 patterns for
    every named format (eligible or not) and deterministic mode, at every vector alignment."""
    from lofloat_b200 import _lib
    allbits = torch.arange(0, 65536, device=cuda_dev, dtype=torch.int32).to(torch.int16).view(torch.float16).repeat(9)[5:]
    try:
        for name, f in L.FORMATS.items():
            if f.bitwidth > 16:
                continue
            d = _desc(lof, f)
            for mode in L.DETERMINISTIC_MODES:
                _lib.set_option("round.fast", 0)
                want = lof.virtual_round(allbits, d, mode).view(torch.int16)
                _lib.set_option("round.fast", 2)
                assert torch.equal(lof.virtual_round(allbits, d, mode).view(torch.int16), want), (name, L.MODE_NAMES[mode])
    finally:
        _lib.set_option("round.fast", 2)


def test_fused_hadamard_rounding(lof, cuda_dev):
    """Row f-3: the block Walsh-Hadamard rotation of LoF_Linear (layers.py:182-216) fused with the rounding that follows it.
    Plain transform == the oracle's restatement of `_fwht` and == the same formula evaluated with torch ops on the device
    (what the reference computes), bit for bit, for every block size 2 .. 1024; the fused variants == the separate ops."""
    def torch_fwht(x, block):                      # same dataflow as layers.py:182-216, torch float32 ops on the device
        flat = x.reshape(-1, block)
        M, h = flat.shape[0], 1
        while h < block:
            v = flat.view(M, block // (2 * h), 2, h)
            flat = torch.stack([v[:, :, 0, :] + v[:, :, 1, :], v[:, :, 0, :] - v[:, :, 1, :]], dim=2).reshape(M, block)
            h *= 2
        return (flat * (1.0 / float(np.sqrt(block)))).view(x.shape)
    rng = np.random.default_rng(4)
    e4m3, e8m0 = lof.create_p3109_params(8, 4, bias=7), lof.create_e8m0_params()
    for block in (2, 4, 8, 16, 32, 64, 128, 256, 512, 1024):
        cols = max(block, 128) * 3
        x = (rng.standard_normal((37, cols)) * np.exp2(rng.integers(-6, 7, (37, 1)))).astype(np.float32)
        if (37 * cols) % 128:
            x = x[:32]
        xt = _gpu(x, cuda_dev)
        y = lof.fwht(xt, block)
        assert L.bits_equal(_np(y), L.o_fwht(x, block)), block
        assert torch.equal(y, torch_fwht(xt, block)), block
        for mode, sl, scale in ((L.RNE, 0, 1.0), (L.RTO, 0, 0.5), (L.SR_A, 5, 3.0), (L.RD, 0, 1.0)):
            got = lof.fwht(xt, block, params=e4m3, round_mode=mode, stoch_len=sl, scale=scale, rng=(3, 1))
            assert torch.equal(got, lof.virtual_round(y, e4m3, mode, sl, scale, rng=(3, 1))), (block, mode)
        for mxb in (16, 32, 128):
            if cols % mxb:
                continue
            for mode in (L.RNE, L.RTZ, L.TRUE_SR):
                got = lof.fwht(xt, block, params=e4m3, round_mode=mode, mx_block_size=mxb, scale_format=e8m0, rng=(9, 2))
                want = lof.virtual_mx_round(y, mxb, e4m3, e8m0, mode, rng=(9, 2))
                assert torch.equal(got.view(torch.int32), want.view(torch.int32)), (block, mxb, mode)
    # involution (H_B / sqrt(B) is orthonormal and symmetric) up to rounding, on the C3 shape with the whole-row default
    big = torch.randn(4096, 1024, device=cuda_dev)
    back = lof.fwht(lof.fwht(big), None)
    assert float((back - big).abs().max()) < 1e-5
    with pytest.raises(ValueError):
        lof.fwht(big, 48)
    # blocks of 2048 ... 32768 (block_size=None on an LLM-sized in_features): two-pass transform, bit-identical to `_fwht`;
    # with rounding parameters the Python surface composes it with the rounding op
    for block in (2048, 4096, 8192, 16384, 32768):
        xt = torch.randn(24, block * 2, device=cuda_dev) * 3
        y = lof.fwht(xt, block)
        assert torch.equal(y, torch_fwht(xt, block)), block
        assert L.bits_equal(_np(y[:2]), L.o_fwht(_np(xt[:2]), block)), block
        assert torch.equal(lof.fwht(xt.clone(), None), torch_fwht(xt, block * 2)) if block < 32768 else True
        got = lof.fwht(xt, block, params=e4m3, scale=0.5)
        assert torch.equal(got, lof.virtual_round(y, e4m3, L.RNE, 0, 0.5)), block
        got = lof.fwht(xt, block, params=e4m3, mx_block_size=32, scale_format=e8m0)
        assert torch.equal(got.view(torch.int32), lof.virtual_mx_round(y, 32, e4m3, e8m0).view(torch.int32)), block
    with pytest.raises(lof.LoFloatError):
        lof.fwht(torch.randn(4, 131072, device=cuda_dev), 65536)                            # > 32768: not handled


def test_batched_gemm_equals_per_slice_calls(lof, cuda_dev):
    """Row f-3: lof_gemm_batched (one launch, grid.z = batch) == lof_gemm on every (batch, head) slice, which is what the
    reference's Python loop computes (layers.py:142-149), bit for bit; attention-like shapes, ragged sizes, the BLIS-panel
    and single-panel accumulation modes, a non-RNE mode."""
    g = torch.Generator(device=cuda_dev).manual_seed(12)
    for (bh, M, K, N) in (((2, 3), 64, 96, 64), ((5,), 100, 300, 36), ((4, 2), 17, 33, 129)):
        A = torch.rand(*bh, M, K, device=cuda_dev, generator=g) * 2 - 1
        B = torch.rand(*bh, K, N, device=cuda_dev, generator=g) * 2 - 1
        for mant, mode, kc in ((10, L.RNE, None), (7, L.RTZ, 0), (4, L.RTO, 64)):
            got = lof.lof_gemm_batched(A, B, mant, mode, kc=kc)
            Af, Bf = A.reshape(-1, M, K), B.reshape(-1, K, N)
            want = torch.stack([lof.lof_gemm(Af[i], Bf[i], mant, mode, kc=kc) for i in range(Af.shape[0])]).reshape(*bh, M, N)
            assert torch.equal(got, want), (bh, M, K, N, mant)
    A0, B0 = _np(A.reshape(-1, M, K)[3]), _np(B.reshape(-1, K, N)[3])
    assert L.bits_equal(_np(lof.lof_gemm_batched(A, B, 10).reshape(-1, M, N)[3]), L.o_gemm(A0, B0, 10, L.RNE, kc=256))
    # tensor-core chunked mode per slice (k_chunk keyword, or the global opt-in the reference's positional call sites use)
    Aq = _gpu(_e4m3_uniform((3, 2, 128, 256), 31), cuda_dev)
    Bq = _gpu(_e4m3_uniform((3, 2, 256, 160), 32), cuda_dev)
    want = torch.stack([lof.lof_gemm(Aq.reshape(-1, 128, 256)[i], Bq.reshape(-1, 256, 160)[i], 7, k_chunk=32, partial_round=True)
                        for i in range(6)]).reshape(3, 2, 128, 160)
    assert torch.equal(lof.lof_gemm_batched(Aq, Bq, 7, k_chunk=32, partial_round=True), want)
    lof.set_gemm_tensor_core(32, partial_round=True)
    try:
        assert torch.equal(lof.lof_gemm_batched(Aq, Bq, 7), want)
    finally:
        lof.set_gemm_tensor_core(None)
    with pytest.raises(TypeError):
        lof.lof_gemm_batched(Aq, Bq, 7, partial_round=True)
    # lof_linear on the tensor-core chunked mode == lof_gemm(k_chunk) on the transposed weight + bias
    W = _gpu(_e4m3_uniform((160, 256), 33), cuda_dev)
    bias = torch.rand(160, device=cuda_dev, generator=g)
    want = lof.lof_gemm(Aq.reshape(-1, 256), W.t().contiguous(), 7, k_chunk=32, partial_round=True) + bias
    assert torch.equal(lof.lof_linear(Aq, W, bias, 7, k_chunk=32, partial_round=True), want.reshape(3, 2, 128, 160))


def test_exact_gemm_more_than_65535_row_tiles(lof, cuda_dev):
    """lof_gemm / lof_linear flatten every leading dimension into M: more than 65535 tiles of 64 rows (4.19 M tokens) put the row
    tiles on grid.x.  Row blocks of the big product == the same rows computed on their own (deterministic mode)."""
    M, K, N = 65535 * 64 + 70, 8, 16
    g = torch.Generator(device=cuda_dev).manual_seed(77)
    A = torch.rand(M, K, device=cuda_dev, generator=g) * 2 - 1
    B = torch.rand(K, N, device=cuda_dev, generator=g) * 2 - 1
    Cb = lof.lof_gemm(A, B, 10)
    for lo in (0, 2_000_000, M - 135):
        assert torch.equal(Cb[lo:lo + 135], lof.lof_gemm(A[lo:lo + 135].contiguous(), B, 10)), lo
    assert L.bits_equal(_np(Cb[M - 70:]), L.o_gemm(_np(A[M - 70:]), _np(B), 10, L.RNE, kc=0))


def test_conv2d_implicit_gemm_and_fused_linear(lof, cuda_dev):
    """Row f-3: lof_conv2d (implicit GEMM, no im2col matrix) == F.unfold + one lof_gemm per image and group, which is what
    LoF_Conv2d's im2col backend computes (layers.py:710-745), bit for bit -- strides, dilation, padding, groups, ragged
    tiles, both reduction orders (the (r, s, c) order is the same GEMM on the permuted K axis), the fused bias epilogue,
    the output rescale; one image also against the CPU oracle's GEMM.  lof_linear (weight read as stored, bias fused)
    == lof_gemm on the transposed copy + bias (layers.py:152-168)."""
    import torch.nn.functional as F
    g = torch.Generator(device=cuda_dev).manual_seed(31)
    e4m3 = lof.create_p3109_params(8, 4, bias=7)
    cases = (  # B, C_in, H, W, C_out, R, S, pad, stride, dil, groups
        (2, 8, 12, 12, 16, 3, 3, (1, 1), (1, 1), (1, 1), 1),
        (3, 6, 17, 13, 10, 3, 2, (2, 0), (2, 1), (1, 2), 2),
        (1, 32, 20, 20, 70, 1, 1, (0, 0), (1, 1), (1, 1), 1),
        (2, 4, 9, 31, 5, 5, 3, (3, 4), (3, 2), (2, 1), 1),
    )
    for (B, Ci, H, W, Co, R, S, pad, st, dil, grp) in cases:
        x = lof.virtual_round(torch.randn(B, Ci, H, W, device=cuda_dev, generator=g), e4m3)
        w = lof.virtual_round(torch.randn(Co, Ci // grp, R, S, device=cuda_dev, generator=g) * 0.5, e4m3)
        bias = lof.virtual_round(torch.randn(Co, device=cuda_dev, generator=g), e4m3)
        cols = F.unfold(x, (R, S), dilation=dil, padding=pad, stride=st)                      # (B, Ci*R*S, L), k = (c, r, s)
        Lout = cols.shape[-1]
        Cg, Mg = Ci // grp, Co // grp
        cols = cols.view(B, grp, Cg * R * S, Lout)
        wcol = w.view(grp, Mg, Cg * R * S)
        for mant, mode, kc, ws, xs_ in ((10, L.RNE, 0, 1.0, 1.0), (7, L.RTZ, 256, 2.0, 4.0), (4, L.RTO, 16, 1.0, 1.0)):
            for order in ("crs", "rsc"):
                if order == "crs":
                    wc, xc = wcol, cols
                else:                                                                         # same chains with k = (r, s, c)
                    wc = wcol.view(grp, Mg, Cg, R * S).transpose(2, 3).reshape(grp, Mg, -1)
                    xc = cols.view(B, grp, Cg, R * S, Lout).transpose(2, 3).reshape(B, grp, -1, Lout)
                want = torch.stack([torch.cat([lof.lof_gemm(wc[gi].contiguous(), xc[b, gi].contiguous(), mant, mode, 0, ws, xs_, kc=kc)
                                               for gi in range(grp)], 0) for b in range(B)])
                want = want.view(B, Co, *[(d + 2 * p - dl * (k - 1) - 1) // s_ + 1 for d, p, dl, k, s_ in zip((H, W), pad, dil, (R, S), st)])
                got = lof.lof_conv2d(x, w, pad[0], pad[1], st[0], st[1], dil[0], dil[1], mant, mode, 0, ws, xs_, groups=grp, k_order=order, kc=kc)
                assert got.shape == want.shape and torch.equal(got, want), (B, Ci, H, W, Co, R, S, order, mant)
                got_b = lof.lof_conv2d(x, w, pad[0], pad[1], st[0], st[1], dil[0], dil[1], mant, mode, 0, ws, xs_, groups=grp, k_order=order,
                                       kc=kc, bias=bias)
                assert torch.equal(got_b, want + bias.view(1, -1, 1, 1))
    # against the CPU oracle (pinned to bli_sgemm + lof_sgemm_ukr): last case, image 0
    w0, x0 = _np(wcol[0]), _np(cols[0, 0])
    got = lof.lof_conv2d(x, w, pad[0], pad[1], st[0], st[1], dil[0], dil[1], 10, k_order="crs", kc=256)
    assert L.bits_equal(_np(got[0].reshape(Co, -1)), L.o_gemm(w0, x0, 10, L.RNE, kc=256))
    # accum_mant_bits = 23 on exactly summable data == torch's conv
    xi = torch.randint(-4, 5, (2, 8, 10, 10), device=cuda_dev, generator=g).float()
    wi = torch.randint(-4, 5, (12, 8, 3, 3), device=cuda_dev, generator=g).float()
    assert torch.equal(lof.lof_conv2d(xi, wi, 1, 1, 1, 1, 1, 1, 23), F.conv2d(xi.double(), wi.double(), padding=1).float())
    with pytest.raises(RuntimeError):
        lof.lof_conv2d(xi, wi[:, :4], 1, 1, 1, 1, 1, 1, 10)                                   # C_in mismatch (Lof_kernel.cu:175)
    with pytest.raises(RuntimeError):
        lof.lof_conv2d(xi[:, :, :2, :2], wi, 0, 0, 1, 1, 1, 1, 10)                            # window does not fit (Lof_kernel.cu:196)

    # fused linear
    for (lead, K, N) in (((4, 9), 96, 50), ((130,), 257, 64), ((1,), 16, 3)):
        x = torch.rand(*lead, K, device=cuda_dev, generator=g) * 2 - 1
        wt = torch.rand(N, K, device=cuda_dev, generator=g) * 2 - 1
        bias = torch.randn(N, device=cuda_dev, generator=g)
        for mant, mode, kc in ((10, L.RNE, None), (5, L.RTZ, 0)):
            want = lof.lof_gemm(x.reshape(-1, K), wt.t().contiguous(), mant, mode, 0, 2.0, 0.5, kc=kc).reshape(*lead, N)
            assert torch.equal(lof.lof_linear(x, wt, None, mant, mode, 0, 2.0, 0.5, kc=kc), want)
            assert torch.equal(lof.lof_linear(x, wt, bias, mant, mode, 0, 2.0, 0.5, kc=kc), want + bias)


def test_exact_gemm_optimistic_chain_equals_exact_routines(lof, cuda_dev):
    """The per-FMA GEMM's default kernel runs a branch-free chain (RZ FMA | inexact, rounding on the signed pattern) and
    falls back per tile to the exact routines when an intermediate reaches 2^127 / inf / NaN.  It must equal the exact
    routines (gemm.fast = 0, the kernel pinned to bli_sgemm + lof_sgemm_ukr) bit for bit: all 15 modes, accumulator widths
    incl. r = 0, panels, data with exact cancellations, subnormal products, overflowing sums, infinities and NaNs."""
    g = torch.Generator(device=cuda_dev).manual_seed(77)
    M, K, N = 96, 200, 130

    def operands(kind):
        A = torch.randn(M, K, device=cuda_dev, generator=g)
        B = torch.randn(K, N, device=cuda_dev, generator=g)
        if kind == "cancel":                                  # small integers: exact sums, exact zeros of both signs
            A, B = torch.round(A * 2), torch.round(B * 2)
            A[::7] = -0.0
        elif kind == "tiny":                                  # products in the fp32 subnormal range
            A, B = A * 1e-22, B * 1e-22
        elif kind == "tiny_some":                             # a few operands below 2^-40: only their tiles fall back
            A[5, ::3] *= 1e-30
            B[::5, 77] *= 1e-25
        elif kind == "huge":                                  # sums beyond FLT_MAX in some tiles only
            A[:10] *= 1e19
            B[:, :10] *= 1e19
        elif kind == "special":
            A[3, 5], A[40, 7], B[9, 100] = float("inf"), float("nan"), float("-inf")
            A[70, :] *= 3e19
            B[:, 64:70] *= 3e19
        return A, B

    try:
        for kind in ("plain", "cancel", "tiny", "tiny_some", "huge", "special"):
            A, B = operands(kind)
            for mode in range(15):
                sl = 3 if mode in (L.SR_A, L.SR_B, L.SR_C, L.SR_D) else 0
                for mant, kc in ((10, 256), (23, 0), (2, 64), (0, 0), (1, 32)) if mode in (L.RNE, L.RTZ, L.SR_A, L.RU) else ((7, 0),):
                    lof.manual_seed(5)
                    lof.set_option("gemm.fast", 1)
                    fast = lof.lof_gemm(A, B, mant, mode, sl, kc=kc)
                    lof.manual_seed(5)
                    lof.set_option("gemm.fast", 0)
                    exact = lof.lof_gemm(A, B, mant, mode, sl, kc=kc)
                    assert L.bits_equal(_np(fast), _np(exact)), (kind, mode, mant, kc)
                    if mode == L.RNE:                                              # the two-FMA flavour of the RNE chain
                        lof.set_option("gemm.fast", 2)
                        assert L.bits_equal(_np(lof.lof_gemm(A, B, mant, mode, sl, kc=kc)), _np(exact)), (kind, mant, kc, "fast=2")
    finally:
        lof.set_option("gemm.fast", 1)


def test_exact_gemm_float_domain_rne_selftest(lof, cuda_dev):
    """(t + K) - K == round_bits<RNE> + mask for every normal fp32 pattern x every accumulator width, and every out-of-range
    pattern surfaces as NaN (the fallback trigger): device self-test over 2^32 x 22 cases (the chain is not used for a 0-bit mantissa, where it differs)."""
    from lofloat_b200 import _lib
    bad = torch.zeros(1, dtype=torch.int64, device=cuda_dev)
    _lib.check(_lib.lib().lof_selftest_gemm_rne_round(bad.data_ptr(), None), "selftest")
    torch.cuda.synchronize()
    assert int(bad.item()) == 0


# ------------------------------------------------------------------------- block-scaled MX GEMM (tcgen05.mma.kind::mxf8f6f4.block_scale)
def _e4m3fn_decode_table(dev):
    t = np.zeros(256, dtype=np.float32)
    for c in range(256):
        e, m = (c >> 3) & 15, c & 7
        v = np.nan if (e == 15 and m == 7) else (m * 2.0 ** -9 if e == 0 else (1 + m / 8.0) * 2.0 ** (e - 7))
        t[c] = -v if c & 0x80 else v
    return torch.from_numpy(t).to(dev)


def _mx_storage_decode(elems, scales, table):
    v = table[elems.long()]
    return (v.view(elems.size(0), -1, 32) * torch.exp2(scales.float() - 127.0).unsqueeze(-1)).view(elems.shape)


def test_mx_pack_operand_is_an_exact_split_of_the_mx_rounded_tensor(lof, cuda_dev):
    """mx_pack_operand: fp32 containers holding virtual_mx_round's output -> e4m3fn element codes + E8M0 scale bytes.  The split is
    exact: decode(pack(y)) == y bit for bit on every element that lies on the block's element grid -- for MXFP6 / MXFP4 element
    formats and for MXFP8 e4m3 that is every element except the block maxima the reference's saturating virtual_round leaves
    UNROUNDED and the reference's 480 (the hardware e4m3fn ends at 448); those are snapped and flagged."""
    torch.manual_seed(11)
    tab = _e4m3fn_decode_table(cuda_dev)
    e8 = lof.create_e8m0_params()
    fmts = {"e4m3": (lof.create_p3109_params(8, 4, bias=7), L.FORMATS["e4m3_b7"]), "e3m2": (lof.create_p3109_params(6, 3, bias=3), L.FORMATS["e3m2"]),
            "e2m3": (lof.create_p3109_params(6, 4, bias=1), L.FORMATS["e2m3"]), "e2m1": (lof.create_p3109_params(4, 2, bias=1), L.FORMATS["e2m1"])}
    for name, (d, f) in fmts.items():
        x = torch.randn(256, 1024, device=cuda_dev) * torch.exp2(torch.randint(-20, 21, (256, 1), device=cuda_dev).float())
        x[5, 64:96] = 0.0                                            # an all-zero block
        y = lof.virtual_mx_round(x, 32, d, e8)
        assert L.bits_equal(_np(y).reshape(-1), L.o_mx_round(_np(x).reshape(-1), 32, L.FORMATS["e8m0"], f))     # the quantisation itself: reference semantics
        el, sc, st = lof.mx_pack_operand(y)
        assert el.shape == y.shape and sc.shape == (256, 32) and el.dtype == torch.uint8 and sc.dtype == torch.uint8
        dec = _mx_storage_decode(el, sc, tab)
        # elements of y that are on the element grid of their block (value / block scale representable in the element format)
        yb, xb = y.view(256, 32, 32), x.view(256, 32, 32)
        amax = xb.abs().amax(dim=-1, keepdim=True)
        pmn = float(lof._lib.lib().lof_mx_priv_max_normal(C.byref(d._c())))
        scale = lof.virtual_round(amax / pmn, e8)
        q = torch.where(scale > 0, yb / scale, torch.zeros_like(yb))
        # e4m3 only: a block whose largest |q| exceeds 448 (the reference's 480, or an unrounded maximum) is stored one binade up (480 -> 240 x 2
        # exactly), which costs the block's elements below 2^-5 their last bit; every other block / format converts exactly
        hot = q.abs().amax(dim=-1, keepdim=True) > 448.0
        on_grid = (lof.virtual_round(q.contiguous(), d) == q) & (q.abs() <= pmn) & ~hot     # (an overflowing q comes back from virtual_round unchanged)
        assert torch.equal(dec.view(256, 32, 32)[on_grid].view(torch.int32), yb[on_grid].view(torch.int32)), name
        assert name == "e4m3" or not bool(hot.any().item())
        if bool((~on_grid).any().item()) and bool((dec != y).any().item()):
            assert bool(st.item() & 1), name
        # snapped values: rounded to nearest on the e4m3fn grid of their block (4 significant bits; absolute floor = the grid's smallest step)
        err = (dec - y).abs().view(256, 32, 32)
        bound = torch.maximum(yb.abs() / 16.0, yb.abs().amax(dim=-1, keepdim=True) * 2.0 ** -16)
        assert bool((err <= bound).all().item()), name


def test_mx_gemm_block_scaled_tensor_core_vs_oracle(lof, cuda_dev):
    """lof_mx_gemm (tcgen05.mma.kind::mxf8f6f4.block_scale on MX storage) against the chunked-accumulation oracle evaluated on the DECODED
    operands, for every accumulator semantics the kernel serves.  BIT-EXACT wherever each chunk's partial sum is exactly representable in
    the tensor core's fp32 accumulator (k_chunk = one MX block of e4m3 products: 32 products of <= 8 significant bits on a common
    power-of-two grid), and equal to the bf16-word tensor-core path (lof_gemm_chunked_packed) on the same decoded operands; with fp32
    accumulation over several blocks the tensor core's internal alignment truncates: tolerance = 1 ulp of the ACCUMULATOR format at the result
    plus (K / 32) * 2^-23 * sum |a_ik b_jk| (stated inline)."""
    torch.manual_seed(12)
    tab = _e4m3fn_decode_table(cuda_dev)
    e8 = lof.create_e8m0_params()
    fm = {"e4m3": lof.create_p3109_params(8, 4, bias=7), "e3m2": lof.create_p3109_params(6, 3, bias=3), "e2m1": lof.create_p3109_params(4, 2, bias=1)}
    e5m10 = lof.FloatFormatDescriptor(16, 10, 15, lof.InfBehavior.Saturating, lof.NaNBehavior._3109, lof.Signedness.Signed,
                                      lof.create_half_params().is_inf_checker, lof.create_half_params().is_nan_checker)

    def run(M, N, K, name, k_chunk, mant, partial, mode=L.RNE, acc_format=None, acc_fmt_oracle=None, exact=True, row_scale=True):
        A = torch.randn(M, K, device=cuda_dev)
        Bt = torch.randn(N, K, device=cuda_dev)
        if row_scale:                                               # block scales spread over 2^-6 .. 2^6: the scale path is exercised
            A = A * torch.exp2(torch.randint(-6, 7, (M, K // 32), device=cuda_dev).float()).repeat_interleave(32, dim=1)
            Bt = Bt * torch.exp2(torch.randint(-6, 7, (N, K // 32), device=cuda_dev).float()).repeat_interleave(32, dim=1)
        ae, asc, _ = lof.mx_quantize_operand(A, fm[name], e8)
        be, bsc, _ = lof.mx_quantize_operand(Bt, fm[name], e8)
        Ad, Bd = _mx_storage_decode(ae, asc, tab), _mx_storage_decode(be, bsc, tab)
        got = lof.mx_gemm(ae, asc, be, bsc, mant, mode, k_chunk=k_chunk, partial_round=partial, accum_format=acc_format)
        want = L.o_gemm_chunked(_np(Ad), _np(Bd.t().contiguous()), mant, k_chunk if k_chunk > 0 else K, mode=mode, acc_fmt=acc_fmt_oracle,
                                partial_round=partial)
        g = _np(got)
        what = (M, N, K, name, k_chunk, mant, partial, mode)
        if exact:
            assert L.bits_equal(g, want), (what, L.first_mismatch(g.reshape(-1), g.reshape(-1), want.reshape(-1)))
        else:
            # tolerance: 1 ulp of the accumulator format at the result + the tensor core's fp32 alignment error, which scales with the sum of
            # the |products| (cancellation): K / 32 block additions of relative error <= 2^-23 each
            w64 = want.astype(np.float64)
            S = _np(Ad.abs().double() @ Bd.abs().double().t())
            tol = np.abs(w64) * 2.0 ** -mant + S * (K // 32) * 2.0 ** -23
            if 0 < k_chunk < K:          # chunked: a partial that sits on a rounding boundary may round the other way, once per chunk at most
                tol = tol + S * 2.0 ** -mant
            assert bool((np.abs(g.astype(np.float64) - w64) <= tol).all()), (what, float((np.abs(g - w64) / tol).max()))
        # the other tensor-core route to the same numbers: bf16 words of the decoded operands, chunked GEMM
        if exact and k_chunk > 0 and acc_format is None:
            ref_tc = lof.lof_gemm_packed(Ad.bfloat16(), Bd.bfloat16(), mant, mode, k_chunk=k_chunk, partial_round=partial)
            assert torch.equal(ref_tc, got), what
        return got

    for name in fm:                                                  # C5 semantics: bf16-like accumulator re-rounded every MX block, packed fold
        run(256, 448, 256, name, 32, 7, True)
    run(200, 300, 160, "e4m3", 32, 7, True)                          # ragged M / N (TMA zero fill, partial tiles), K not a multiple of 128
    run(128, 224, 4096, "e4m3", 32, 7, True)                         # long K: 128 chunks per tile
    run(384, 672, 512, "e4m3", 64, 7, True, row_scale=False)         # two MX blocks per chunk, blocks of similar magnitude: still exact
    run(384, 672, 512, "e4m3", 64, 7, True, exact=False)             # ... with block scales 2^-6 .. 2^6 the fp32 sum of two block sums rounds
    run(256, 224, 512, "e4m3", 32, 7, False)                         # single rounding: generic fold (round-to-odd add)
    for mode in (L.RTZ, L.RD, L.RU, L.RAZ, L.RTO, L.RNO, L.RTA, L.RTTZ):
        run(128, 224, 256, "e4m3", 32, 5, True, mode=mode)
        run(128, 224, 256, "e4m3", 32, 10, False, mode=mode)
    run(256, 320, 512, "e4m3", 32, 10, False, acc_format=e5m10, acc_fmt_oracle=L.fmt(16, 10, 15), row_scale=False)   # full-format accumulator (E5M10)
    run(128, 224, 256, "e4m3", 0, 7, False)                          # ONE chunk: plain MXFP8 GEMM, one final rounding (bf16-like output)
    run(256, 448, 2048, "e4m3", 0, 10, False, exact=False)           # fp32 accumulation over 64 blocks: tolerance, not bit-equality
    run(256, 448, 1024, "e4m3", 256, 7, True, exact=False)
    # argument validation
    ae = torch.zeros(128, 64, dtype=torch.uint8, device=cuda_dev); sc = torch.full((128, 2), 127, dtype=torch.uint8, device=cuda_dev)
    with pytest.raises(RuntimeError):
        lof.mx_gemm(ae, sc, ae[:, :32].contiguous(), sc, 7)          # K mismatch
    with pytest.raises(RuntimeError):
        lof.mx_gemm(ae, sc, ae, sc, 7, k_chunk=48)                   # k_chunk not a multiple of the MX block
    out = lof.mx_gemm(ae, sc, ae, sc, 7, k_chunk=32, partial_round=True)
    assert out.shape == (128, 128) and bool((out == 0).all().item())
    # a NaN scale byte poisons exactly its rows
    A = torch.randn(128, 64, device=cuda_dev)
    ae, asc, _ = lof.mx_quantize_operand(A, fm["e4m3"], e8)
    asc2 = asc.clone(); asc2[3, 1] = 0xFF
    o2 = lof.mx_gemm(ae, asc2, ae, asc, 7, k_chunk=0)
    assert bool(torch.isnan(o2[3]).all().item()) and not bool(torch.isnan(o2[[0, 1, 2, 4]]).any().item())


def test_mx_gemm_full_size_properties(lof, cuda_dev):
    """8192^3 on MX storage (C5's per-GPU shape class), bf16-like accumulator re-rounded every MX block: equals the bf16-word tensor-core
    path on the decoded operands, linear in a power-of-two rescaling of one operand's SCALE bytes (an exact exponent shift), every output is
    a bf16 value, and rows / columns of zero elements give exact zeros."""
    torch.manual_seed(13)
    n = 8192
    tab = _e4m3fn_decode_table(cuda_dev)
    e8 = lof.create_e8m0_params()
    f = lof.create_p3109_params(8, 4, bias=7)
    A = torch.randn(n, n, device=cuda_dev); Bt = torch.randn(n, n, device=cuda_dev)
    A[100] = 0.0; Bt[7] = 0.0
    ae, asc, _ = lof.mx_quantize_operand(A, f, e8)
    be, bsc, _ = lof.mx_quantize_operand(Bt, f, e8)
    del A, Bt
    Cm = lof.mx_gemm(ae, asc, be, bsc, 7, k_chunk=32, partial_round=True)
    assert bool(torch.isfinite(Cm).all().item())
    assert bool(((Cm.view(torch.int32) & 0xFFFF) == 0).all().item())             # bf16 values
    assert bool((Cm[100] == 0).all().item()) and bool((Cm[:, 7] == 0).all().item())
    C2 = lof.mx_gemm(ae, (asc + 3).contiguous(), be, bsc, 7, k_chunk=32, partial_round=True)
    assert torch.equal(C2, Cm * 8.0)
    Ad16 = _mx_storage_decode(ae, asc, tab).bfloat16()
    Bd16 = _mx_storage_decode(be, bsc, tab).bfloat16()
    Cb = lof.lof_gemm_packed(Ad16, Bd16, 7, k_chunk=32, partial_round=True)
    neq = int((Cb != Cm).sum().item())
    # the two tensor-core routes agree except where one chunk's exact partial sum is a tie of the fp32 accumulator's internal alignment
    assert neq <= n * n // 1_000_000, neq


def test_positional_lof_gemm_on_the_tensor_core_path(lof, cuda_dev):
    """set_gemm_tensor_core(k_chunk): the reference's positional call sites (torch.ops.lofloat.lof_gemm, LoFloat/_custom_ops.py:92-99; _lof_linear,
    layers.py:152-168) run the chunked tensor-core mode without being changed; None restores the per-FMA reference semantics."""
    torch.manual_seed(21)
    e4m3 = lof.create_p3109_params(8, 4, bias=7)
    A = lof.virtual_round(torch.rand(256, 512, device=cuda_dev) * 2 - 1, e4m3)
    B = lof.virtual_round(torch.rand(512, 384, device=cuda_dev) * 2 - 1, e4m3)
    exact = lof.lof_gemm(A, B, 10)
    try:
        lof.set_gemm_tensor_core(32)
        via_op = torch.ops.lofloat.lof_gemm(A, B, 10, 0, 0, 1.0, 1.0)
        assert torch.equal(via_op, lof.lof_gemm(A, B, 10, k_chunk=32))
        assert L.bits_equal(_np(via_op), L.o_gemm_chunked(_np(A), _np(B), 10, 32))
        assert torch.equal(lof.lof_gemm(A, B, 10, kc=0), exact)                  # an explicit panel length keeps the exact mode
    finally:
        lof.set_gemm_tensor_core(None)
    assert torch.equal(torch.ops.lofloat.lof_gemm(A, B, 10, 0, 0, 1.0, 1.0), exact)


def test_chunked_gemm_stochastic_accumulator_rounding(lof, cuda_dev):
    """The chunked tensor-core GEMM with STOCHASTIC accumulator rounding (the reference leaf honours every mode, blis_lof_gemm.hpp:217-228):
    bit-exact against the oracle's Philox stream (one dense word per rounding event, index ((which * n_chunks + chunk) * M + i) * N + j) for all
    six stochastic modes, single and partial rounding, ragged shapes (N % 4 != 0 takes the per-element word path), and unbiased on average."""
    torch.manual_seed(31)
    e4m3 = lof.create_p3109_params(8, 4, bias=7)
    for (M, N, K) in ((192, 256, 512), (100, 130, 96)):
        A = lof.virtual_round(torch.rand(M, K, device=cuda_dev) * 2 - 1, e4m3)
        B = lof.virtual_round(torch.rand(K, N, device=cuda_dev) * 2 - 1, e4m3)
        for mode, sl in ((L.SR_A, 6), (L.SR_B, 4), (L.SR_C, 5), (L.SR_D, 3), (L.TRUE_SR, 0), (L.PROB, 0)):
            for partial in (False, True):
                got = lof.lof_gemm(A, B, 7, mode, sl, k_chunk=32, partial_round=partial, rng=(5, 9))
                want = L.o_gemm_chunked(_np(A), _np(B), 7, 32, mode=mode, partial_round=partial, stoch_len=sl, seed=5, offset=9)
                assert L.bits_equal(_np(got), want), (M, N, K, L.MODE_NAMES[mode], partial, L.first_mismatch(_np(got).reshape(-1), _np(got).reshape(-1), want.reshape(-1)))
    # unbiased: the mean over many seeds approaches the fp32-accumulated product, where RNE into 4 mantissa bits every chunk drifts
    A = lof.virtual_round(torch.rand(128, 1024, device=cuda_dev), e4m3)
    B = lof.virtual_round(torch.rand(1024, 128, device=cuda_dev), e4m3)
    ref = A.double() @ B.double()
    mean = torch.zeros_like(ref)
    reps = 48
    for s in range(reps):
        mean += lof.lof_gemm(A, B, 4, L.TRUE_SR, 0, k_chunk=32, rng=(100 + s, 0)).double()
    mean /= reps
    err_sr = float(((mean - ref).abs() / ref.abs()).mean())
    err_rne = float(((lof.lof_gemm(A, B, 4, k_chunk=32).double() - ref).abs() / ref.abs()).mean())
    assert err_sr < 0.5 * err_rne, (err_sr, err_rne)          # (48 repetitions leave ~ sigma / 7 of sampling noise in the mean)
