"""Bring-up / characterisation driver for the tcgen05 chunked GEMM (runs on the GPU box).  TEST INFRASTRUCTURE: it checks
against the CPU oracle, which is why it lives under tests/.

    python tests/gpu_bringup/tc_bringup.py            # every step, each in its own subprocess with a timeout
    python tests/gpu_bringup/tc_bringup.py <step>     # one step in this process

Each step prints one line starting with PASS / FAIL / INFO so the gpurun tail stays readable.
"""
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def ints(shape, lo, hi, seed):
    import numpy as np
    return np.random.default_rng(seed).integers(lo, hi + 1, shape).astype(np.float32)


def e4m3_uniform(shape, seed):
    """U(-1,1) rounded to OCP e4m3 (bias 7) -- the C4 operand distribution (test/gemm_test.cpp:82-86)."""
    import numpy as np
    import oracle_lib as L
    x = np.random.default_rng(seed).uniform(-1, 1, shape).astype(np.float32)
    return L.o_round_params(x.ravel(), L.FORMATS["e4m3_b7"], L.RNE).reshape(shape)


def compare(name, got, want, exact=True, tol_ulps=0.0, mant=10):
    import numpy as np
    if exact:
        same = np.array_equal(got.view(np.uint32), want.view(np.uint32))
        nbad = int((got.view(np.uint32) != want.view(np.uint32)).sum())
        idx = np.argwhere(got.view(np.uint32) != want.view(np.uint32))
        extra = ""
        if nbad:
            i, j = idx[0]
            extra = f" first@({i},{j}) got={got[i, j]!r} want={want[i, j]!r}; bad rows {sorted(set(idx[:, 0].tolist()))[:8]} cols {sorted(set(idx[:, 1].tolist()))[:8]}"
        print(("PASS" if same else "FAIL") + f" {name}: {nbad}/{got.size} mismatching{extra}", flush=True)
        return same
    ulp = np.exp2(np.floor(np.log2(np.maximum(np.abs(want), 1e-30))) - mant)
    err = np.abs(got.astype(np.float64) - want.astype(np.float64)) / ulp
    frac = float((got.view(np.uint32) != want.view(np.uint32)).mean())
    ok = float(err.max()) <= tol_ulps
    print(("PASS" if ok else "FAIL") + f" {name}: max err {err.max():.2f} ulp_target (tol {tol_ulps}), mean {err.mean():.4f}, "
          f"{100 * frac:.3f}% elements differ", flush=True)
    return ok


def run_gemm(A, B, mant, k_chunk, **kw):
    import torch
    import lofloat_b200 as lof
    dev = torch.device("cuda:0")
    out = lof.lof_gemm(torch.from_numpy(A).to(dev), torch.from_numpy(B).to(dev), mant, k_chunk=k_chunk, **kw)
    torch.cuda.synchronize()
    return out.cpu().numpy()


def step_plain_1tile():
    A, B = ints((128, 128), -4, 4, 1), ints((128, 128), -4, 4, 2)
    got = run_gemm(A, B, 23, 128)
    compare("plain 128x128x128 e4m3, one chunk, no rounding", got, A @ B)


def step_plain_k512():
    A, B = ints((128, 512), -4, 4, 3), ints((512, 128), -4, 4, 4)
    got = run_gemm(A, B, 23, 512)
    compare("plain 128x128x512 e4m3 (4 smem stages), one chunk", got, A @ B)
    got = run_gemm(A, B, 23, 32)
    compare("plain 128x128x512 e4m3, 16 chunks of 32, no rounding", got, A @ B)


def step_multitile():
    import oracle_lib as L
    A, B = ints((256, 1024), -8, 8, 5), ints((1024, 384), -8, 8, 6)
    got = run_gemm(A, B, 23, 1024)
    compare("plain 256x384x1024 e4m3", got, A @ B)
    for mant in (10, 7, 4):
        got = run_gemm(A, B, mant, 32, allow_packed=False)
        compare(f"chunk 32, mant {mant} RNE, S1, integer data", got, L.o_gemm_chunked(A, B, mant, 32))
    got = run_gemm(A, B, 10, 64, allow_packed=False, partial_round=True)
    compare("chunk 64, mant 10 RNE, S2 generic", got, L.o_gemm_chunked(A, B, 10, 64, partial_round=True))
    for mode in (1, 2, 4, 5, 6, 7, 12, 13):
        got = run_gemm(A, B, 6, 32, round_mode=mode, allow_packed=False)
        compare(f"chunk 32, mant 6, mode {mode}", got, L.o_gemm_chunked(A, B, 6, 32, mode=mode))


def step_bf16():
    import numpy as np
    import oracle_lib as L
    A, B = ints((256, 512), -100, 100, 7), ints((512, 256), -100, 100, 8)
    got = run_gemm(A, B, 23, 512, operand_dtype="bf16")
    compare("plain 256x256x512 bf16 operands", got, A @ B)
    got = run_gemm(A, B, 8, 16, operand_dtype="bf16")
    compare("bf16 operands, chunk 16, mant 8", got, L.o_gemm_chunked(A, B, 8, 16))
    got = run_gemm(A, B, 7, 32, operand_dtype="bf16", partial_round=True)
    compare("bf16 operands, chunk 32, mant 7, S2 packed bfloat162", got, L.o_gemm_chunked(A, B, 7, 32, partial_round=True))
    A5, B5 = ints((128, 256), -3, 3, 9) * 0.25, ints((256, 128), -3, 3, 10) * 4.0
    got = run_gemm(A5, B5, 10, 32, operand_dtype="e5m2")
    compare("e5m2 operands, chunk 32, mant 10", got, L.o_gemm_chunked(A5, B5, 10, 32))


def step_tails():
    import oracle_lib as L
    A, B = ints((200, 400), -6, 6, 11), ints((400, 136), -6, 6, 12)      # K*1 % 16 == 0, ragged M, N, K % 128 != 0
    got = run_gemm(A, B, 10, 32, allow_packed=False)
    compare("ragged 200x136x400, chunk 32 (last chunk 16), mant 10", got, L.o_gemm_chunked(A, B, 10, 32))
    got = run_gemm(A, B, 10, 96, allow_packed=False)
    compare("ragged, chunk 96", got, L.o_gemm_chunked(A, B, 10, 96))
    got = run_gemm(A, B, 10, 32, scale_a=2.0, scale_b=4.0, allow_packed=False)
    compare("alpha = 1/8", got, L.o_gemm_chunked(A, B, 10, 32, alpha=0.125))


def step_packed():
    import oracle_lib as L
    import lofloat_b200 as lof
    A, B = e4m3_uniform((512, 2048), 13), e4m3_uniform((2048, 256), 14)
    want = L.o_gemm_chunked(A, B, 10, 32, partial_round=True)
    got_g = run_gemm(A, B, 10, 32, partial_round=True, allow_packed=False)
    compare("U(-1,1) e4m3 512x256x2048, chunk 32, mant 10, S2 GENERIC vs oracle", got_g, want, exact=False, tol_ulps=4.0)
    got_p = run_gemm(A, B, 10, 32, partial_round=True)
    compare("  S2 PACKED half2 vs S2 generic (same tensor-core partials -> bit-exact)", got_p, got_g)
    print("INFO packed status bits:", lof.lof_gemm.last_status, flush=True)
    e5m10 = lof.FloatFormatDescriptor(16, 10, 15, lof.InfBehavior.Saturating, lof.NaNBehavior._3109, lof.Signedness.Signed,
                                      lof.create_half_params().is_inf_checker, lof.create_half_params().is_nan_checker)
    got_f = run_gemm(A, B, 0, 32, accum_format=e5m10, partial_round=True, allow_packed=False)
    compare("  E5M10 full format generic vs mantissa-only generic", got_f, got_g)
    got_fp = run_gemm(A, B, 0, 32, accum_format=e5m10, partial_round=True)
    compare("  E5M10 full format packed vs generic", got_fp, got_f)
    got_s1 = run_gemm(A, B, 10, 32)
    compare("  S1 generic vs oracle S1", got_s1, L.o_gemm_chunked(A, B, 10, 32), exact=False, tol_ulps=4.0)
    # overflow -> the packed kernel must flag it and the generic kernel must recompute
    Ab, Bb = ints((128, 512), 200, 448, 15), ints((512, 128), 200, 448, 16)
    Ab, Bb = (Ab // 32) * 32, (Bb // 32) * 32
    got = run_gemm(Ab, Bb, 10, 32, partial_round=True)
    compare("  overflow of the fp16 range -> generic recompute", got, L.o_gemm_chunked(Ab, Bb, 10, 32, partial_round=True))
    print("INFO status bits after overflow case:", lof.lof_gemm.last_status, flush=True)


def step_hw16():
    """fp16 accumulation inside the tensor core (c_format = F16) vs the oracle's single-rounding fold, E5M10 / RNE / k = 32."""
    import numpy as np
    import oracle_lib as L
    import lofloat_b200 as lof
    e5m10 = L.fmt(16, 10, 15, L.SATURATING, L.SIGNED)
    d = lof.FloatFormatDescriptor(16, 10, 15, lof.InfBehavior.Saturating, lof.NaNBehavior._3109, lof.Signedness.Signed,
                                  lof.create_half_params().is_inf_checker, lof.create_half_params().is_nan_checker)
    A, B = ints((256, 1024), -8, 8, 5), ints((1024, 384), -8, 8, 6)
    compare("hw16: integer data, E5M10", run_gemm(A, B, 0, 32, accum_format=d), L.o_gemm_chunked(A, B, 0, 32, acc_fmt=e5m10))
    A, B = e4m3_uniform((512, 2048), 13), e4m3_uniform((2048, 256), 14)
    want = L.o_gemm_chunked(A, B, 0, 32, acc_fmt=e5m10)
    compare("hw16: U(-1,1) e4m3, E5M10 vs oracle (bit-exact?)", run_gemm(A, B, 0, 32, accum_format=d), want)
    compare("hw16: mantissa-only 10 vs oracle (bit-exact?)", run_gemm(A, B, 10, 32), L.o_gemm_chunked(A, B, 10, 32))
    compare("generic (no hw) E5M10 vs oracle", run_gemm(A, B, 0, 32, accum_format=d, allow_hw_accum=False, allow_packed=False), want)
    # wide-range operands: every e4m3 magnitude class, random signs -> partial sums with large exponent spread
    rng = np.random.default_rng(99)
    def wide(shape):
        e = rng.integers(-6, 5, shape)
        m = rng.integers(8, 16, shape) / 8.0
        return (np.where(rng.random(shape) < 0.5, -1, 1) * m * np.exp2(e)).astype(np.float32)
    A, B = wide((256, 4096)), wide((4096, 256))
    want = L.o_gemm_chunked(A, B, 0, 32, acc_fmt=e5m10)
    compare("hw16: wide-range e4m3 operands, K = 4096 (bit-exact?)", run_gemm(A, B, 0, 32, accum_format=d), want)
    compare("hw16: wide-range, tolerance view", run_gemm(A, B, 0, 32, accum_format=d), want, exact=False, tol_ulps=4.0)
    compare("generic: wide-range (bit-exact?)", run_gemm(A, B, 0, 32, accum_format=d, allow_hw_accum=False, allow_packed=False), want)
    Ab, Bb = ints((128, 512), -60, 60, 7), ints((512, 128), -60, 60, 8)
    compare("hw16: bf16 operands, k_chunk 16", run_gemm(Ab, Bb, 0, 16, accum_format=d, operand_dtype="bf16"),
            L.o_gemm_chunked(Ab, Bb, 0, 16, acc_fmt=e5m10))
    Ao, Bo = (ints((128, 512), 200, 448, 15) // 32) * 32, (ints((512, 128), 200, 448, 16) // 32) * 32
    compare("hw16: overflow -> generic recompute", run_gemm(Ao, Bo, 0, 32, accum_format=d), L.o_gemm_chunked(Ao, Bo, 0, 32, acc_fmt=e5m10))
    print("INFO status after overflow:", lof.lof_gemm.last_status, flush=True)


def step_inexact():
    import numpy as np
    import lofloat_b200 as lof
    A = np.full((128, 128), 0.3, np.float32)
    B = np.ones((128, 128), np.float32)
    try:
        run_gemm(A, B, 10, 32)
        print("FAIL inexact operands were not reported", flush=True)
    except lof.LoFloatError as e:
        print("PASS inexact operands reported:", e, flush=True)


def step_perf():
    import torch
    import lofloat_b200 as lof
    dev = torch.device("cuda:0")
    for n in (4096, 8192):
        A = torch.from_numpy(e4m3_uniform((n, n), 21)).to(dev)
        B = torch.from_numpy(e4m3_uniform((n, n), 22)).to(dev)
        for label, kw in (("k32 mant10 S1 hw fp16 accumulate", dict(k_chunk=32)),
                          ("k32 mant10 S2 packed", dict(k_chunk=32, partial_round=True)),
                          ("k32 mant10 S1 generic", dict(k_chunk=32, allow_hw_accum=False)),
                          ("k128 mant10 S1 generic", dict(k_chunk=128)),
                          ("k512 mant10 S1 generic", dict(k_chunk=512)),
                          ("k512 mant10 S2 packed", dict(k_chunk=512, partial_round=True)),
                          ("one chunk, no rounding", dict(k_chunk=n, accum_mant_bits=23))):
            mant = kw.pop("accum_mant_bits", 10)
            lof.lof_gemm(A, B, mant, check=False, **kw)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 3
            a.record()
            for _ in range(reps):
                lof.lof_gemm(A, B, mant, check=False, **kw)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / reps
            print(f"INFO perf {n}^3 {label}: {ms:.3f} ms  {2.0 * n ** 3 / ms / 1e9:.1f} TFLOP/s (incl. operand conversion)", flush=True)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        Ab, Bb = A.bfloat16(), B.bfloat16()
        torch.matmul(Ab, Bb)
        a.record()
        for _ in range(5):
            torch.matmul(Ab, Bb)
        b.record()
        torch.cuda.synchronize()
        print(f"INFO perf {n}^3 cuBLAS bf16 (context): {2.0 * n ** 3 / (a.elapsed_time(b) / 5) / 1e9:.1f} TFLOP/s", flush=True)


def step_pair():
    """Persistent CTA-pair kernel (lof_gemm_tc2.cuh) against the first-generation kernel: same tensor-core arithmetic
    -> bit-identical C, on aligned, ragged and tiny shapes; then its speed."""
    import numpy as np
    import torch
    import lofloat_b200 as lof
    d = lof.FloatFormatDescriptor(16, 10, 15, lof.InfBehavior.Saturating, lof.NaNBehavior._3109, lof.Signedness.Signed,
                                  lof.create_half_params().is_inf_checker, lof.create_half_params().is_nan_checker)
    for (M, N, K) in ((256, 256, 512), (512, 768, 2048), (200, 136, 400), (1000, 520, 4096 + 32), (128, 64, 128), (1536, 2048, 1024)):
        A, B = e4m3_uniform((M, K), M + K), e4m3_uniform((K, N), N + K + 1)
        for label, kw in (("hw16 E5M10 k32", dict(k_chunk=32, accum_format=d)), ("hw16 mant10 k32", dict(k_chunk=32)),
                          ("one chunk mant10", dict(k_chunk=0)), ("one chunk E5M10", dict(k_chunk=0, accum_format=d)),
                          ("one chunk mant23", dict(k_chunk=0, mant=23))):
            kw = dict(kw)
            mant = kw.pop("mant", 0 if "accum_format" in kw else 10)
            got = run_gemm(A, B, mant, allow_pair=True, **kw)
            want = run_gemm(A, B, mant, allow_pair=False, **kw)
            compare(f"pair vs single-CTA kernel {M}x{N}x{K} {label}", got, want)
    Ab, Bb = ints((384, 640), -60, 60, 7), ints((640, 300), -60, 60, 8)
    compare("pair vs single-CTA, bf16 operands one chunk mant 8", run_gemm(Ab, Bb, 8, 640, operand_dtype="bf16"),
            run_gemm(Ab, Bb, 8, 640, operand_dtype="bf16", allow_pair=False))
    dev = torch.device("cuda:0")
    for n in (4096, 8192):
        A = torch.from_numpy(e4m3_uniform((n, n), 21)).to(dev)
        B = torch.from_numpy(e4m3_uniform((n, n), 22)).to(dev)
        for label, kw in (("pair hw16 k32", dict(k_chunk=32)), ("single-CTA hw16 k32", dict(k_chunk=32, allow_pair=False)),
                          ("pair one chunk", dict(k_chunk=n)), ("single-CTA one chunk", dict(k_chunk=n, allow_pair=False))):
            lof.lof_gemm(A, B, 10, check=False, **kw)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                lof.lof_gemm(A, B, 10, check=False, **kw)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 5
            print(f"INFO perf {n}^3 {label}: {ms:.3f} ms  {2.0 * n ** 3 / ms / 1e9:.1f} TFLOP/s (incl. operand conversion)", flush=True)


STEPS = {"pair": step_pair, "plain_1tile": step_plain_1tile, "plain_k512": step_plain_k512, "multitile": step_multitile, "bf16": step_bf16,
         "tails": step_tails, "packed": step_packed, "hw16": step_hw16, "inexact": step_inexact, "perf": step_perf}


def main():
    if len(sys.argv) > 1:
        for name in sys.argv[1:]:
            STEPS[name]()
        return
    for name in STEPS:
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, __file__, name], timeout=420, capture_output=True, text=True)
            out = (p.stdout or "") + "\n".join(l for l in (p.stderr or "").splitlines()[-6:])
            print(f"--- {name} (rc {p.returncode}, {time.time() - t0:.0f}s)\n{out}", flush=True)
        except subprocess.TimeoutExpired:
            print(f"--- {name}: TIMEOUT", flush=True)


if __name__ == "__main__":
    main()
