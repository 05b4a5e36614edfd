"""Generates tests/golden/*.npz from the COMPILED REFERENCE (oracle/_ref/libref_shim.so, built
from /root/reference by oracle/Makefile).  Run in the build container only:

    python tests/golden/make_golden.py

The vectors pin oracle/lof_oracle.c on machines where the reference tree does not exist
(tests/test_oracle_golden.py).  Inputs are seeded; outputs are whatever the reference produced.
"""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
import oracle_lib as L  # noqa: E402


def main():
    assert L.build_ref() is not None and L.ref() is not None, "reference shim unavailable"
    x = np.concatenate([L.special_values_f32(), L.wide_random_f32(6000, 2024)])

    # 1) format-params overload, every named format x every deterministic mode
    out = {"x": x}
    for name, f in L.FORMATS.items():
        for mode in L.DETERMINISTIC_MODES:
            out[f"{name}__{L.MODE_NAMES[mode]}"] = L.r_round_params(x, f, mode)
    np.savez_compressed(HERE / "round_params_f32.npz", **out)

    # 2) mantissa-only overload
    out = {"x": x}
    for m in (0, 2, 3, 7, 10, 15, 22):
        for mode in L.DETERMINISTIC_MODES:
            out[f"m{m}__{L.MODE_NAMES[mode]}"] = L.r_round_mantissa(x, m, mode)
    np.savez_compressed(HERE / "round_mantissa_f32.npz", **out)

    # 3) stochastic recipes: the reference re-seeds mt19937 per call, so every element sees the
    #    same draw; the draw itself is stored (as the oracle's event word) next to the outputs.
    out = {"x": x}
    meta = []
    for name in ("e4m3_b7", "e5m2", "e2m1", "bf16_like"):
        f = L.FORMATS[name]
        r = 23 - f.mantissa_bits
        for mode in L.STOCHASTIC_MODES:
            for sl in (0, 1, 4, min(r, 10)):
                for seed in (12345, 7):
                    w = L.ref_fixed_word(mode, sl, seed)
                    key = f"{name}__{L.MODE_NAMES[mode]}__l{sl}__s{seed}"
                    out[key] = L.r_round_params(x, f, mode, stoch_len=sl, seed=seed)
                    meta.append((key, name, mode, sl, w))
    out["meta_keys"] = np.array([m[0] for m in meta])
    out["meta_fmt"] = np.array([m[1] for m in meta])
    out["meta_mode"] = np.array([m[2] for m in meta], dtype=np.int32)
    out["meta_len"] = np.array([m[3] for m in meta], dtype=np.int32)
    out["meta_word"] = np.array([m[4] for m in meta], dtype=np.uint32)
    np.savez_compressed(HERE / "round_stochastic_f32.npz", **out)

    # 4) double source
    xd = np.concatenate([x.astype(np.float64), (np.random.default_rng(5).standard_normal(2000) *
                                               np.exp2(np.random.default_rng(6).integers(-40, 40, 2000)))])
    out = {"x": xd}
    for name in ("e4m3_b7", "e5m2", "bf16_like", "half", "single"):
        for mode in (L.RNE, L.RTZ, L.RAZ, L.RTO, L.RU, L.RD):
            out[f"{name}__{L.MODE_NAMES[mode]}"] = L.r_round_params(xd, L.FORMATS[name], mode, dtype=L.F64)
    np.savez_compressed(HERE / "round_params_f64.npz", **out)

    # 5) MX
    rng = np.random.default_rng(7)
    out = {}
    for n in (32, 96, 97, 1000, 5003):
        xx = (rng.standard_normal(n) * np.exp2(rng.integers(-8, 9, n))).astype(np.float32)
        xx[rng.integers(0, n, max(1, n // 100))] *= 64
        if n == 96:
            xx[32:64] = 0.0            # all-zero block -> 0, not NaN (test_virtual_mx_round_cuda.cu:156-172)
        out[f"x_{n}"] = xx
        for bs in (16, 32, 64, 128, 256):
            for en in ("e4m3_b7", "p3109_8p4", "e5m2", "e3m2", "e2m3", "e2m1"):
                for sn in ("e8m0", "e5m2"):
                    for mode in (L.RNE, L.RTZ, L.RAZ):
                        out[f"mx_{n}_{bs}_{en}_{sn}_{L.MODE_NAMES[mode]}"] = L.r_mx_round(
                            xx, bs, L.FORMATS[sn], L.FORMATS[en], mode)
    # explicit KAT block of test/test_mx_round_py.py:37-40
    kat = np.array([448, 224, -112, 1, 0.5, -2, 8, 0] * 4, dtype=np.float32)
    out["kat_x"] = kat
    out["kat_y"] = L.r_mx_round(kat, 32, L.FORMATS["e8m0"], L.FORMATS["p3109_8p4"], L.RNE)
    np.savez_compressed(HERE / "mx_round_f32.npz", **out)

    # 6) GEMM through BLIS + lof_sgemm_ukr, and the round-to-odd FMA
    rng = np.random.default_rng(42)
    out = {}
    for (M, N, K) in ((2, 2, 4), (17, 33, 64), (64, 48, 300), (40, 24, 513)):
        A = rng.uniform(-1, 1, (M, K)).astype(np.float32)
        B = rng.uniform(-1, 1, (K, N)).astype(np.float32)
        out[f"A_{M}_{N}_{K}"], out[f"B_{M}_{N}_{K}"] = A, B
        for mode in L.DETERMINISTIC_MODES:
            for bits in (10, 7):
                out[f"C_{M}_{N}_{K}_{L.MODE_NAMES[mode]}_{bits}"] = L.r_gemm(A, B, bits, mode)
    # exact-representable KAT of test/test_blis_lof_accum.cpp:205-245 (column-major there)
    A = np.array([.5, -.5, 1, 0, .5, .5, -.5, 1], np.float32).reshape(4, 2).T.copy()
    B = np.array([.5, 0, .5, 1, 1, -.5, .5, .5], np.float32).reshape(2, 4).T.copy()
    out["kat_A"], out["kat_B"] = A, B
    out["kat_C"] = L.r_gemm(A, B, 10, L.RNE)
    a = (rng.standard_normal(4000) * np.exp2(rng.integers(-20, 20, 4000))).astype(np.float32)
    b = (rng.standard_normal(4000) * np.exp2(rng.integers(-20, 20, 4000))).astype(np.float32)
    c = (rng.standard_normal(4000) * np.exp2(rng.integers(-20, 20, 4000))).astype(np.float32)
    c[:500] = (-a[:500].astype(np.float64) * b[:500].astype(np.float64)).astype(np.float32)   # heavy cancellation
    lib = L.ref()
    out["fma_a"], out["fma_b"], out["fma_c"] = a, b, c
    out["fma_r"] = np.array([lib.ref_round_to_odd_fma(float(p), float(q), float(s)) for p, q, s in zip(a, b, c)],
                            dtype=np.float32)
    np.savez_compressed(HERE / "gemm_f32.npz", **out)
    # 7) storage codec (lo_pack.hpp): reps of the probe set of test/test_pack.cpp under every deterministic mode
    #    (OvfInf) and under RNE for the other two saturation modes, the full decode table, packed bytes for both
    #    slot orders, and the typed ground truth (Templated_Float<Fp>) for the two presets the shim instantiates.
    out = {}
    for name, f in L.CODEC_FORMATS.items():
        x = L.codec_probe_floats(f)
        x = x[:: max(1, x.size // 3000)]
        out[f"{name}__x"] = x
        for mode in L.DETERMINISTIC_MODES:
            out[f"{name}__rep__{L.MODE_NAMES[mode]}__sat0"] = L.r_float_to_rep(x, f, mode, 0).astype(np.uint16)
        for sat in (1, 2):
            out[f"{name}__rep__RNE__sat{sat}"] = L.r_float_to_rep(x, f, L.RNE, sat).astype(np.uint16)
        n = f.fmt.bitwidth
        reps = np.arange(1 << n, dtype=np.uint32) if n <= 12 else np.arange(0, 1 << n, 7, dtype=np.uint32)
        out[f"{name}__dec_reps"] = reps.astype(np.uint16)
        out[f"{name}__dec"] = L.r_rep_to_float(reps, f)
        xs = x[:1001]
        for endian in (0, 1):
            out[f"{name}__packed{endian}"] = L.r_encode(xs, f, L.RNE, 0, 0, endian)
            out[f"{name}__unpacked{endian}"] = L.r_decode(out[f"{name}__packed{endian}"], xs.size, f, endian)
    lib = L.ref()
    for tag, fn, f in (("typed_e4m3fn", lib.ref_typed_e4m3fn, L.CODEC_FORMATS["e4m3fn"]),
                       ("typed_p3109_8p4", lib.ref_typed_p3109_8_4, L.CODEC_FORMATS["p3109_8p4"])):
        x = L.codec_probe_floats(f)
        rep, back = np.empty(x.size, np.uint32), np.empty(x.size, np.float32)
        fn(L._ptr(x), L._ptr(rep), L._ptr(back), x.size)
        out[f"{tag}__x"], out[f"{tag}__rep"], out[f"{tag}__back"] = x, rep.astype(np.uint16), back
    np.savez_compressed(HERE / "pack_codec.npz", **out)
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
