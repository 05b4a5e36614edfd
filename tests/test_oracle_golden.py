"""CPU: the oracle (oracle/lof_oracle.c) against the committed golden vectors that were generated
from the COMPILED REFERENCE (tests/golden/make_golden.py), against the known-answer vectors of
SURVEY.md Appendix C, and against the property checks the reference's own tests make."""
from pathlib import Path

import numpy as np
import pytest

import oracle_lib as L

G = Path(__file__).resolve().parent / "golden"


def _load(name):
    return np.load(G / name)


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    assert L.o_philox([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert L.o_philox([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert L.o_philox([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == \
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_rand_word_layout():
    # word i of element index: block = index >> 2, lane = index & 3, offset in ctr[2:4], stream xor-ed into ctr[3]
    seed, off = 0x0123456789ABCDEF, 0x1122334455667788
    blk = L.o_philox([5, 0, off & 0xFFFFFFFF, off >> 32], [seed & 0xFFFFFFFF, seed >> 32])
    for lane in range(4):
        assert L.o_rand_word(seed, off, 20 + lane) == blk[lane]
    blk_s = L.o_philox([5, 0, off & 0xFFFFFFFF, (off >> 32) ^ L.STREAM_SCALE], [seed & 0xFFFFFFFF, seed >> 32])
    assert L.o_rand_word(seed, off, 21, L.STREAM_SCALE) == blk_s[1]


def test_rand_event_packing():
    """Per-event random words (oracle/lof_oracle.c lofo_rand_event == lofloat_b200/csrc/lof_core.cuh philox_group4):
    > 16 bits -> dense word i; <= 16 bits -> half of dense word i >> 1; <= 8 bits -> byte (i & 3) of the word owned by
    the 4-element group, where the groups v, v + 256, v + 512, v + 768 of an aligned 4096-element run share a block."""
    seed, off = 77, 5
    for i in (0, 1, 5, 1023, 1024, 4095, 4096, 4097 + 3 * 1024, 10 ** 7 + 3):
        assert L.o_rand_event(seed, off, i, 32) == L.o_rand_word(seed, off, i)
        assert L.o_rand_event(seed, off, i, 17) == L.o_rand_word(seed, off, i)
        assert L.o_rand_event(seed, off, i, 16) == (L.o_rand_word(seed, off, i >> 1) >> (16 * (i & 1))) & 0xFFFF
        v = i >> 2
        q = 4 * (((v >> 10) << 8) | (v & 255)) + ((v >> 8) & 3)
        assert L.o_rand_event(seed, off, i, 8) == (L.o_rand_word(seed, off, q) >> (8 * (i & 3))) & 0xFF
        assert L.o_rand_event(seed, off, i, 1) == L.o_rand_event(seed, off, i, 8)
    # every byte of every word of a block is used exactly once inside an aligned 4096-element run
    owners = {}
    for i in range(4096):
        v = i >> 2
        q = 4 * (((v >> 10) << 8) | (v & 255)) + ((v >> 8) & 3)
        owners.setdefault((q, i & 3), []).append(i)
    assert len(owners) == 4096 and all(len(o) == 1 for o in owners.values())
    assert L.oracle().lofo_rand_bits_needed(L.SR_A, 8) == 8 and L.oracle().lofo_rand_bits_needed(L.SR_D, 8) == 9
    assert L.oracle().lofo_rand_bits_needed(L.TRUE_SR, 0) == 32 and L.oracle().lofo_rand_bits_needed(L.PROB, 0) == 1


def test_round_params_golden_f32():
    g = _load("round_params_f32.npz")
    x = g["x"]
    n = 0
    for name, f in L.FORMATS.items():
        for mode in L.DETERMINISTIC_MODES:
            want = g[f"{name}__{L.MODE_NAMES[mode]}"]
            got = L.o_round_params(x, f, mode)
            assert L.bits_equal(got, want), (name, L.MODE_NAMES[mode], L.first_mismatch(x, got, want))
            n += 1
    assert n == len(L.FORMATS) * 9


def test_round_mantissa_golden_f32():
    g = _load("round_mantissa_f32.npz")
    x = g["x"]
    for m in (0, 2, 3, 7, 10, 15, 22):
        for mode in L.DETERMINISTIC_MODES:
            want = g[f"m{m}__{L.MODE_NAMES[mode]}"]
            got = L.o_round_mantissa(x, m, mode)
            assert L.bits_equal(got, want), (m, L.MODE_NAMES[mode], L.first_mismatch(x, got, want))


def test_round_stochastic_golden_f32():
    """Stochastic bit recipes: the oracle fed the reference's own draw reproduces the reference."""
    g = _load("round_stochastic_f32.npz")
    x = g["x"]
    keys, fmts, modes, lens, words = g["meta_keys"], g["meta_fmt"], g["meta_mode"], g["meta_len"], g["meta_word"]
    assert len(keys) > 150
    changed = 0
    for key, fname, mode, sl, w in zip(keys, fmts, modes, lens, words):
        want = g[str(key)]
        got = L.o_round_params(x, L.FORMATS[str(fname)], int(mode), stoch_len=int(sl), fixed_word=int(w))
        assert L.bits_equal(got, want), (key, L.first_mismatch(x, got, want))
        trunc = L.o_round_params(x, L.FORMATS[str(fname)], L.RTZ)
        changed += int(not L.bits_equal(got, trunc))
    assert changed > 50      # the draws actually moved results away from truncation


def test_round_params_golden_f64():
    g = _load("round_params_f64.npz")
    x = g["x"]
    for name in ("e4m3_b7", "e5m2", "bf16_like", "half", "single"):
        for mode in (L.RNE, L.RTZ, L.RAZ, L.RTO, L.RU, L.RD):
            want = g[f"{name}__{L.MODE_NAMES[mode]}"]
            got = L.o_round_params(x, L.FORMATS[name], mode, dtype=L.F64)
            assert L.bits_equal(got, want), (name, mode)


def test_mx_golden():
    g = _load("mx_round_f32.npz")
    cnt = 0
    for n in (32, 96, 97, 1000, 5003):
        xx = g[f"x_{n}"]
        for bs in (16, 32, 64, 128, 256):
            for en in ("e4m3_b7", "p3109_8p4", "e5m2", "e3m2", "e2m3", "e2m1"):
                for sn in ("e8m0", "e5m2"):
                    for mode in (L.RNE, L.RTZ, L.RAZ):
                        want = g[f"mx_{n}_{bs}_{en}_{sn}_{L.MODE_NAMES[mode]}"]
                        got = L.o_mx_round(xx, bs, L.FORMATS[sn], L.FORMATS[en], mode)
                        assert L.bits_equal(got, want), (n, bs, en, sn, mode, L.first_mismatch(xx, got, want))
                        cnt += 1
    assert cnt == 5 * 5 * 6 * 2 * 3
    # KAT block of test/test_mx_round_py.py:37-40 round-trips unchanged
    assert L.bits_equal(g["kat_y"], g["kat_x"])
    got = L.o_mx_round(g["kat_x"], 32, L.FORMATS["e8m0"], L.FORMATS["p3109_8p4"], L.RNE)
    assert L.bits_equal(got, g["kat_x"])
    # all-zero block stays zero, never NaN (test_virtual_mx_round_cuda.cu:156-172)
    z = L.o_mx_round(np.zeros(64, np.float32), 32, L.FORMATS["e8m0"], L.FORMATS["e4m3_b7"])
    assert not np.isnan(z).any() and not z.any()


def test_gemm_golden():
    g = _load("gemm_f32.npz")
    for (M, N, K) in ((2, 2, 4), (17, 33, 64), (64, 48, 300), (40, 24, 513)):
        A, B = g[f"A_{M}_{N}_{K}"], g[f"B_{M}_{N}_{K}"]
        for mode in L.DETERMINISTIC_MODES:
            for bits in (10, 7):
                want = g[f"C_{M}_{N}_{K}_{L.MODE_NAMES[mode]}_{bits}"]
                got = L.o_gemm(A, B, bits, mode, kc=256)
                assert L.bits_equal(got, want), (M, N, K, mode, bits)
    # exact-representable KAT (test/test_blis_lof_accum.cpp:205-245): equals the fp64 product in every mode
    A, B = g["kat_A"], g["kat_B"]
    exact = (A.astype(np.float64) @ B.astype(np.float64)).astype(np.float32)
    assert L.bits_equal(g["kat_C"], exact)
    for mode in L.DETERMINISTIC_MODES:
        assert L.bits_equal(L.o_gemm(A, B, 10, mode), exact)
    # round-to-odd FMA
    o = L.oracle()
    got = np.array([o.lofo_round_to_odd_fma(float(a), float(b), float(c))
                    for a, b, c in zip(g["fma_a"], g["fma_b"], g["fma_c"])], dtype=np.float32)
    assert L.bits_equal(got, g["fma_r"])


def test_gemm_gamma_bound():
    """|C - C_exact| <= gamma_K (|A||B|) + 1e-6 with eps = 2^-10 (test_blis_lof_accum.cpp:126-177)."""
    rng = np.random.default_rng(3)
    M, N, K = 24, 20, 64
    A = rng.uniform(-1, 1, (M, K)).astype(np.float32)
    B = rng.uniform(-1, 1, (K, N)).astype(np.float32)
    exact = A.astype(np.float64) @ B.astype(np.float64)
    absab = np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64)
    eps = 2.0 ** -10
    gamma = K * eps / (1 - K * eps)
    err32 = np.abs(L.o_gemm(A, B, 23, L.RNE) - exact).max()
    for mode in L.DETERMINISTIC_MODES:
        Cq = L.o_gemm(A, B, 10, mode)
        err = np.abs(Cq - exact)
        # directed modes accumulate a one-sided error of up to 2 eps per step
        k = 2.0 if mode in (L.RTZ, L.RAZ, L.RU, L.RD, L.RTO) else 1.0
        assert (err <= k * gamma * absab + 1e-6).all(), L.MODE_NAMES[mode]
    assert np.abs(L.o_gemm(A, B, 10, L.RNE) - exact).max() > 50 * err32     # rounding is engaged (:182-198)


# ---- known answers extracted from the compiled reference ---------------
def _r(vals, fname, mode=L.RNE):
    return L.o_round_params(np.array(vals, np.float32), L.FORMATS[fname], mode)


def test_appendix_c_e4m3():
    x = [1.0625, 1.1875, 447, 464, 480, 0.001, 0.0009765625, 0.00146484375, 1e-30]
    want = [1.0, 1.25, 448, 448, 480, 0.001953125, 0.0, 0.001953125, 0.0]
    assert np.array_equal(_r(x, "e4m3_b7"), np.array(want, np.float32))
    # saturating overflow returns the UNROUNDED input; extended returns +inf even for negatives
    assert np.array_equal(_r([496, 512, -1e6], "e4m3_b7"), np.array([496, 512, -1e6], np.float32))
    ext = _r([496, 512, -1e6, -np.inf], "e4m3_b7_ext")
    assert (ext.view(np.uint32) == 0x7F800000).all()
    assert _r([-np.inf], "e4m3_b7")[0] == -np.inf
    assert _r([np.nan], "e4m3_b7").view(np.uint32)[0] == 0x7FC00000
    assert _r([-0.0], "e4m3_b7").view(np.uint32)[0] == 0x80000000


def test_appendix_c_e8m0():
    x = [1.4, 1.5, 2.9, 3.0, 3.1, 0.75, 1e-38, 1e-39, 3e38]
    want = [1, 2, 2, 2, 4, 0.5, 2.0 ** -126, 0, np.inf]
    assert np.array_equal(_r(x, "e8m0"), np.array(want, np.float32))
    assert _r([3e38], "e8m0", L.RD)[0] == np.float32(2.0 ** 127)
    assert _r([1e-39], "e8m0", L.RU)[0] == np.float32(2.0 ** -126)


def test_appendix_c_e2m1_table():
    cols = [L.RNE, L.RTZ, L.RAZ, L.RNO, L.RD, L.RU, L.RTA, L.RTO, L.RTTZ]
    table = {
        0.1: [0, 0, 0.5, 0, 0, 0.5, 0, 0.5, 0],
        0.25: [0, 0, 0.5, 0.5, 0, 0.5, 0.5, 0.5, 0],
        0.3: [0.5, 0, 0.5, 0.5, 0, 0.5, 0.5, 0.5, 0.5],
        0.75: [1, 0.5, 1, 0.5, 0.5, 1, 1, 0.5, 0.5],
        1.25: [1, 1, 1.5, 1.5, 1, 1.5, 1.5, 1.5, 1],
        -1.25: [-1, -1, -1.5, -1.5, -1.5, -1, -1.5, -1.5, -1],
        2.5: [2, 2, 3, 3, 2, 3, 3, 3, 2],
        -2.5: [-2, -2, -3, -3, -3, -2, -3, -3, -2],
        5: [4, 4, 6, 6, 4, 6, 6, 6, 4],
        6: [6] * 9,
        6.9: [6, 6, 6.9, 6, 6, 6.9, 6, 6, 6],
        7: [7, 6, 7, 6, 6, 7, 7, 6, 6],
        7.5: [7.5, 6, 7.5, 7.5, 6, 7.5, 7.5, 6, 7.5],
        8: [8] * 9,
        -0.2: [-0.0, -0.0, -0.5, -0.0, -0.5, -0.0, -0.0, -0.5, -0.0],
    }
    for xv, row in table.items():
        for mode, want in zip(cols, row):
            got = _r([xv], "e2m1", mode)
            assert got.view(np.uint32)[0] == np.array([want], np.float32).view(np.uint32)[0], (xv, L.MODE_NAMES[mode], got)


def test_bf16_like_bogus_hidden_bit():
    # fp32-subnormal input to a target with ToMin_exp = -126
    got = _r([2.0 ** -130], "bf16_like")
    assert got.view(np.uint32)[0] == 0x00440000


# ---- property checks the reference makes (test/test_virtual_round.cpp:95-215) -----------------
@pytest.mark.parametrize("m", [7, 10, 15])
def test_mantissa_properties(m):
    rng = np.random.default_rng(m)
    x = (rng.standard_normal(20000) * np.exp2(rng.integers(-30, 30, 20000))).astype(np.float32)
    xd = x.astype(np.float64)
    rd = L.o_round_mantissa(x, m, L.RD).astype(np.float64)
    ru = L.o_round_mantissa(x, m, L.RU).astype(np.float64)
    assert (rd <= xd).all() and (ru >= xd).all()
    ulp = np.exp2(np.floor(np.log2(np.abs(xd))) - m)
    for mode in (L.RD, L.RU, L.RTZ, L.RAZ, L.RTO):
        assert (np.abs(L.o_round_mantissa(x, m, mode).astype(np.float64) - xd) <= ulp * (1 + 1e-12)).all()
    for mode in (L.RNE, L.RNO, L.RTA, L.RTTZ):
        assert (np.abs(L.o_round_mantissa(x, m, mode).astype(np.float64) - xd) <= ulp / 2 * (1 + 1e-12)).all()
    # ties: RNE -> even, RNO -> odd, RTA -> away, RTTZ -> toward zero (:139-184)
    base = (rng.integers(1 << 23, 1 << 24, 2000).astype(np.uint32) >> (23 - m)) << (23 - m)
    tie = ((base | (1 << (22 - m))) & 0x7FFFFF) | (120 << 23)
    t = tie.astype(np.uint32).view(np.float32)
    lsb = lambda a: (a.view(np.uint32) >> (23 - m)) & 1
    assert (lsb(L.o_round_mantissa(t, m, L.RNE)) == 0).all()
    assert (lsb(L.o_round_mantissa(t, m, L.RNO)) == 1).all()
    assert (L.o_round_mantissa(t, m, L.RTA) > t).all()
    assert (L.o_round_mantissa(t, m, L.RTTZ) < t).all()
    # stochastic results are one of the two neighbours (:189-215)
    for mode, sl in ((L.SR_A, 8), (L.SR_B, 8), (L.SR_C, 8), (L.SR_D, 8), (L.TRUE_SR, 0), (L.PROB, 0)):
        s = L.o_round_mantissa(x, m, mode, stoch_len=sl, seed=99, offset=1).astype(np.float64)
        lo, hi = np.minimum(rd, ru), np.maximum(rd, ru)
        trunc = L.o_round_mantissa(x, m, L.RTZ).astype(np.float64)
        away = L.o_round_mantissa(x, m, L.RAZ).astype(np.float64)
        ok = (s == trunc) | (s == away)
        if mode == L.SR_D:      # SR-D may bump an exactly representable value by one ulp (R + coin == 2^len)
            ok |= np.abs(s - xd) <= ulp * (1 + 1e-12)
        assert ok.all(), L.MODE_NAMES[mode]
        assert (s >= lo - ulp).all() and (s <= hi + ulp).all()


def test_stochastic_probabilities():
    """P(up) model of each SR variant (test/test_prob_round.cpp:176-236, stoch_prob_test.cpp),
    with Philox draws: |P(up) - ideal| <= 0.02 at 2e5 trials."""
    n = 200000
    m = 3
    frac = 0.3
    x = np.full(n, 1.0 + (1 + frac) / 8, np.float32)          # between 1.125 and 1.25, fraction 0.3 of the gap
    lo = np.float32(1.125)
    xf = float((x[0] - lo) * 8)
    def p_up(mode, sl):
        y = L.o_round_mantissa(x, m, mode, stoch_len=sl, seed=2024, offset=mode * 16 + sl)
        return float((y > lo).mean())
    assert abs(p_up(L.PROB, 0) - 0.5) <= 0.02
    assert abs(p_up(L.TRUE_SR, 0) - xf) <= 0.02
    for sl in (1, 2, 3, 4):
        q = 2 ** sl
        ideal_d = 0.5 * (np.floor(q * xf) / q + np.ceil(q * xf) / q)
        assert abs(p_up(L.SR_D, sl) - ideal_d) <= 0.02, sl
        assert abs(p_up(L.SR_A, sl) - np.floor(q * xf) / q) <= 0.02, sl
    assert abs(p_up(L.SR_A, 10) - xf) <= 0.02
    assert abs(p_up(L.SR_B, 10) - xf) <= 0.02
    assert abs(p_up(L.SR_C, 10) - xf) <= 0.02
    # unbiasedness of true SR: mean of rounded == x within 3 sigma
    y = L.o_round_mantissa(x, m, L.TRUE_SR, seed=5).astype(np.float64)
    assert abs(y.mean() - float(x[0])) < 3 * (0.125 * 0.5) / np.sqrt(n)


def test_mx_properties():
    """Scale is the RNE E8M0 neighbour of amax/priv_max; every representable value round-trips under all
    deterministic modes (test/test_virtual_mx_round.cpp:154-193, 234-254)."""
    for l in range(3, 9):
        for p in range(1, l):
            k, m = l, p - 1
            bias = 1 << (k - p - 1) if k - p - 1 >= 0 else 0
            f = L.fmt(k, m, bias)
            ebits = k - m - 1
            if ebits < 1:
                continue
            pmn = float(L.oracle().lofo_mx_priv_max_normal(f))
            # all representable positive values of the element format
            vals = []
            for e in range(0, 1 << ebits):
                for mm in range(0, 1 << m):
                    v = (mm / (1 << m)) * 2.0 ** (1 - bias) if e == 0 else (1 + mm / (1 << m)) * 2.0 ** (e - bias)
                    vals.append(v)
            vals = np.array(sorted(set(vals)), np.float32)
            blk = np.zeros(((len(vals) + 30) // 31) * 32, np.float32)
            # put the max normal into every block so that scale == 1 exactly
            blk.reshape(-1, 32)[:, 0] = pmn
            blk.reshape(-1, 32)[:, 1:].flat[: len(vals)] = vals
            for mode in L.DETERMINISTIC_MODES:
                y = L.o_mx_round(blk, 32, L.FORMATS["e8m0"], f, mode)
                assert L.bits_equal(y, blk), (l, p, L.MODE_NAMES[mode])
    # scale = RNE E8M0 neighbour
    rng = np.random.default_rng(0)
    f = L.FORMATS["e4m3_b7"]
    pmn = 480.0
    for _ in range(200):
        amax = float(np.float32(np.exp2(rng.uniform(-20, 20))))
        blk = np.zeros(32, np.float32)
        blk[3] = -amax
        y = L.o_mx_round(blk, 32, L.FORMATS["e8m0"], f)
        scale = L.o_round_params(np.array([np.float32(amax) / np.float32(pmn)], np.float32), L.FORMATS["e8m0"])[0]
        lo = 2.0 ** np.floor(np.log2(amax / pmn))
        assert scale in (np.float32(lo), np.float32(2 * lo))
        q = L.o_round_params(np.array([-np.float32(amax) / scale], np.float32), f)[0] * scale
        assert y[3] == q


# ------------------------------------------------------------------------------ storage codec (lo_pack.hpp)
def test_pack_geometry_against_independent_lcm_oracle():
    """test/test_pack.cpp:30-38: group bytes / counts / packed sizes for every width, via an independent lcm."""
    from math import gcd
    lib = L.oracle()
    for n in range(1, 17):
        al = 8 if n <= 8 else 16
        lcm = n * al // gcd(n, al)
        assert lib.lofo_pack_group_count(n) == lcm // n and lib.lofo_pack_group_bytes(n) == lcm // 8
        for count in (0, 1, lcm // n - 1, lcm // n, lcm // n + 1, 1000, 4097):
            assert L.packed_size(n, count) == -(-count // (lcm // n)) * (lcm // 8)


def test_pack_codec_golden_from_reference():
    """Oracle codec == the compiled reference's convert_float_to_rep / convert_rep_to_float / LoF_encode / LoF_decode
    on the probe set of test/test_pack.cpp (grid values, fp32 neighbours, midpoints, overflow, fp32 subnormals) for
    every format that test runs (+ MX element / scale formats), every deterministic mode and saturation mode."""
    g = _load("pack_codec.npz")
    for name, f in L.CODEC_FORMATS.items():
        x = g[f"{name}__x"]
        for mode in L.DETERMINISTIC_MODES:
            assert np.array_equal(L.o_float_to_rep(x, f, mode, 0), g[f"{name}__rep__{L.MODE_NAMES[mode]}__sat0"]), (name, mode)
        for sat in (1, 2):
            assert np.array_equal(L.o_float_to_rep(x, f, L.RNE, sat), g[f"{name}__rep__RNE__sat{sat}"]), (name, sat)
        assert L.bits_equal(L.o_rep_to_float(g[f"{name}__dec_reps"].astype(np.uint32), f), g[f"{name}__dec"]), name
        xs = x[:1001]
        for endian in (0, 1):
            packed = L.o_encode(xs, f, endian=endian)
            assert np.array_equal(packed, g[f"{name}__packed{endian}"]), (name, endian)
            assert L.bits_equal(L.o_decode(packed, xs.size, f, endian), g[f"{name}__unpacked{endian}"]), (name, endian)


def test_pack_codec_against_typed_ground_truth():
    """test/test_pack.cpp:64-115: the runtime codec must agree with the compile-time Templated_Float<Fp> path
    (rep of TF(x), and (float)TF(x) after a round trip); NaN reps only need to be NaN."""
    g = _load("pack_codec.npz")
    for tag, key in (("typed_e4m3fn", "e4m3fn"), ("typed_p3109_8p4", "p3109_8p4")):
        f = L.CODEC_FORMATS[key]
        x, rep, back = g[f"{tag}__x"], g[f"{tag}__rep"].astype(np.uint32), g[f"{tag}__back"]
        got = L.o_float_to_rep(x, f)
        is_nan = np.isnan(back)
        assert np.array_equal(got[~is_nan], rep[~is_nan]), tag
        assert np.all(np.isnan(L.o_rep_to_float(got[is_nan], f))), tag
        assert L.bits_equal(L.o_rep_to_float(got, f)[~is_nan], back[~is_nan]), tag


def test_pack_roundtrip_properties():
    """test/test_pack.cpp:118-227: packing is a lossless bijection on reps for every width and both slot orders
    (incl. a partial trailing group), the two orders differ iff a group holds more than one element, and encoding
    a decoded value reproduces its rep (idempotence) for every non-NaN rep of the small formats."""
    rng = np.random.default_rng(5)
    lib = L.oracle()
    for n in range(1, 17):
        gc = lib.lofo_pack_group_count(n)
        for count in (2 * gc + 1 if gc > 1 else 5, 1 << min(n, 10)):
            reps = rng.integers(0, 1 << n, count, dtype=np.uint32)
            fwd, rev = L.o_pack_reps(reps, n, 0), L.o_pack_reps(reps, n, 1)
            assert fwd.size == L.packed_size(n, count)
            assert np.array_equal(L.o_unpack_reps(fwd, count, n, 0), reps)
            assert np.array_equal(L.o_unpack_reps(rev, count, n, 1), reps)
            if count > gc:
                assert (gc > 1) == (not np.array_equal(fwd, rev)) or gc == 1
    for name in ("p3109_3p2", "e2m1", "p3109_5p2", "e3m2", "e2m3", "p3109_7p3", "e5m2", "e4m3fn", "p3109_8p4", "p3109_8p4_ext"):
        f = L.CODEC_FORMATS[name]
        reps = np.arange(1 << f.fmt.bitwidth, dtype=np.uint32)
        vals = L.o_rep_to_float(reps, f)
        keep = ~np.isnan(vals)
        again = L.o_float_to_rep(vals[keep], f)
        assert np.array_equal(again, reps[keep]), name


def test_mx_storage_oracle_consistent_with_virtual_mx_round():
    """The MX storage oracle (scale bytes + packed elements) is a composition of two pinned pieces; tie it to them:
    decode(encode(x)) equals the (reference-pinned) virtual_mx_round oracle wherever the fake-quant value is
    representable -- always when the scale is rounded UP -- the scale byte of an E8M0 scale is its biased exponent, and
    the element codes are the codec's codes of the unscaled quantised values."""
    rng = np.random.default_rng(8)
    n = 64 * 400
    x = (rng.standard_normal(n) * np.exp2(rng.integers(-15, 16, n))).astype(np.float32)
    x[:64] = 0
    e8c, e8f = L.codec_e8m0(), L.fmt(8, 0, 127, L.SATURATING, L.UNSIGNED)
    for cf, ff in ((L.codec_p3109(8, 4, bias=7), L.FORMATS["e4m3_b7"]), (L.codec_p3109(6, 3, bias=3), L.FORMATS["e3m2"]),
                   (L.codec_p3109(4, 2, bias=1), L.FORMATS["e2m1"])):
        for bs in (16, 32, 64):
            sc, el = L.o_mx_encode(x, bs, e8c, cf, scale_mode=L.RU)
            back = L.o_mx_decode(sc, el, n, bs, e8c, cf)
            fake = L.o_mx_round(x, bs, e8f, ff, scale_mode=L.RU)
            assert np.array_equal(back, fake), (cf.fmt.bitwidth, bs)                 # value compare: these formats have no -0
            scale = np.abs(x).reshape(-1, bs).max(axis=1) / L.oracle().lofo_mx_priv_max_normal(C_byref(ff))
            live = scale > 0
            want_exp = np.ceil(np.log2(scale[live])).astype(np.int64) + 127              # RoundUp to a power of two
            assert np.array_equal(sc[live].astype(np.int64), want_exp), (cf.fmt.bitwidth, bs)
            assert not sc[~live].any()
            reps = L.o_unpack_reps(el, n, cf.fmt.bitwidth)
            q = (x.reshape(-1, bs) / np.exp2(sc.astype(np.float64) - 127)[:, None].astype(np.float32)).astype(np.float32)
            assert np.array_equal(reps.reshape(-1, bs)[live], L.o_float_to_rep(q[live].ravel(), cf).reshape(-1, bs))


def C_byref(obj):
    import ctypes
    return ctypes.byref(obj)
