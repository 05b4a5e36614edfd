"""CPU: pins the oracle against the reference ITSELF (oracle/_ref/libref_shim.so, compiled from
/root/reference by oracle/Makefile).  Skipped where the shim does not exist; the committed golden
vectors (test_oracle_golden.py) cover those machines."""
import numpy as np
import pytest

import oracle_lib as L


@pytest.fixture(scope="module", autouse=True)
def _need_ref(ref_available):
    if not ref_available:
        pytest.skip("compiled reference (oracle/_ref) not available here")


@pytest.fixture(scope="module")
def xs():
    return np.concatenate([L.special_values_f32(), L.wide_random_f32(60000, 17)])


def test_params_all_formats_all_deterministic_modes(xs):
    for name, f in L.FORMATS.items():
        for mode in L.DETERMINISTIC_MODES:
            a, b = L.o_round_params(xs, f, mode), L.r_round_params(xs, f, mode)
            assert L.bits_equal(a, b), (name, L.MODE_NAMES[mode], L.first_mismatch(xs, a, b))


def test_params_exhaustive_strided_e4m3():
    """Every 257th fp32 bit pattern (2^24 values covering all exponents and sign) for the benchmark format."""
    bits = np.arange(0, 2 ** 32, 257, dtype=np.uint64).astype(np.uint32)
    x = bits.view(np.float32)
    for name in ("e4m3_b7", "e2m1"):
        for mode in (L.RNE, L.RTZ, L.RTO):
            a, b = L.o_round_params(x, L.FORMATS[name], mode), L.r_round_params(x, L.FORMATS[name], mode)
            assert L.bits_equal(a, b), (name, mode, L.first_mismatch(x, a, b))


def test_mantissa_overload(xs):
    for m in (0, 1, 3, 7, 10, 15, 22, 23):
        for mode in L.DETERMINISTIC_MODES:
            a, b = L.o_round_mantissa(xs, m, mode), L.r_round_mantissa(xs, m, mode)
            assert L.bits_equal(a, b), (m, L.MODE_NAMES[mode], L.first_mismatch(xs, a, b))


def test_stochastic_recipes_with_reference_draws(xs):
    x = xs[:20000]
    for name in ("e4m3_b7", "e5m2", "e2m1", "bf16_like", "e8m0"):
        f = L.FORMATS[name]
        r = 23 - f.mantissa_bits
        for mode in L.STOCHASTIC_MODES:
            for sl in (0, 1, 3, 8, min(r, 12)):
                if sl > r:
                    continue
                for seed in (12345, 99):
                    w = L.ref_fixed_word(mode, sl, seed)
                    a = L.o_round_params(x, f, mode, stoch_len=sl, fixed_word=w)
                    b = L.r_round_params(x, f, mode, stoch_len=sl, seed=seed)
                    assert L.bits_equal(a, b), (name, L.MODE_NAMES[mode], sl, seed, L.first_mismatch(x, a, b))


def test_double_source(xs):
    rng = np.random.default_rng(8)
    xd = np.concatenate([xs[:5000].astype(np.float64),
                         rng.standard_normal(5000) * np.exp2(rng.integers(-200, 200, 5000))])
    for name in ("e4m3_b7", "e5m2", "bf16_like", "half", "single", "e8m0"):
        for mode in L.DETERMINISTIC_MODES:
            a = L.o_round_params(xd, L.FORMATS[name], mode, dtype=L.F64)
            b = L.r_round_params(xd, L.FORMATS[name], mode, dtype=L.F64)
            assert L.bits_equal(a, b), (name, L.MODE_NAMES[mode])
    for m in (3, 10, 23, 40):
        a, b = L.o_round_mantissa(xd, m, L.RNE, dtype=L.F64), L.r_round_mantissa(xd, m, L.RNE, dtype=L.F64)
        assert L.bits_equal(a, b), m


def test_mx(xs):
    rng = np.random.default_rng(1234)      # test_virtual_mx_round_cuda.cu uses srand(1234)
    for n in (32, 96, 97, 1000, 4096, 5003):
        x = (rng.standard_normal(n) * np.exp2(rng.integers(-8, 9, n))).astype(np.float32)
        for bs in (16, 32, 64, 128, 256):
            for en in ("e4m3_b7", "e5m2", "e2m1"):
                for sn in ("e8m0", "e5m2"):
                    for mode in (L.RNE, L.RAZ, L.RTO, L.RNO):
                        a = L.o_mx_round(x, bs, L.FORMATS[sn], L.FORMATS[en], mode)
                        b = L.r_mx_round(x, bs, L.FORMATS[sn], L.FORMATS[en], mode)
                        assert L.bits_equal(a, b), (n, bs, en, sn, mode, L.first_mismatch(x, a, b))


def test_gemm_and_rto_fma():
    rng = np.random.default_rng(2)
    for (M, N, K) in ((3, 5, 7), (33, 17, 256), (20, 36, 257), (48, 40, 600)):
        A = rng.uniform(-1, 1, (M, K)).astype(np.float32)
        B = rng.uniform(-1, 1, (K, N)).astype(np.float32)
        for mode in L.DETERMINISTIC_MODES:
            a, b = L.o_gemm(A, B, 10, mode, kc=256), L.r_gemm(A, B, 10, mode)
            assert L.bits_equal(a, b), (M, N, K, L.MODE_NAMES[mode])
        a, b = L.o_gemm(A, B, 5, L.RNE, kc=256, alpha=0.37), L.r_gemm(A, B, 5, L.RNE, alpha=0.37)
        assert L.bits_equal(a, b)
    o, r = L.oracle(), L.ref()
    a = (rng.standard_normal(20000) * np.exp2(rng.integers(-60, 60, 20000))).astype(np.float32)
    b = (rng.standard_normal(20000) * np.exp2(rng.integers(-60, 60, 20000))).astype(np.float32)
    c = (rng.standard_normal(20000) * np.exp2(rng.integers(-60, 60, 20000))).astype(np.float32)
    for p, q, s in zip(a, b, c):
        u = np.float32(o.lofo_round_to_odd_fma(p, q, s)).view(np.uint32)
        v = np.float32(r.ref_round_to_odd_fma(p, q, s)).view(np.uint32)
        assert u == v, (p, q, s)


def test_pack_codec_vs_reference():
    """Storage codec: oracle == compiled reference (lo_pack.hpp) for every codec format x deterministic mode x
    saturation mode on the full probe set, the exhaustive decode table, and packed bytes in both slot orders."""
    for name, f in L.CODEC_FORMATS.items():
        x = L.codec_probe_floats(f)
        n = f.fmt.bitwidth
        for mode in L.DETERMINISTIC_MODES:
            for sat in (0, 1, 2):
                assert np.array_equal(L.o_float_to_rep(x, f, mode, sat), L.r_float_to_rep(x, f, mode, sat)), (name, mode, sat)
        reps = np.arange(1 << n, dtype=np.uint32)
        assert L.bits_equal(L.o_rep_to_float(reps, f), L.r_rep_to_float(reps, f)), name
        for endian in (0, 1):
            for count in (1, 7, 1001, 4096 + 17):
                xs = x[:count]
                ea, eb = L.o_encode(xs, f, endian=endian), L.r_encode(xs, f, endian=endian)
                assert np.array_equal(ea, eb), (name, endian, count)
                assert L.bits_equal(L.o_decode(ea, xs.size, f, endian), L.r_decode(eb, xs.size, f, endian))
