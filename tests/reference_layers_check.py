"""Runs the reference's UNMODIFIED LoFloat Python package (LoF_Quantize / LoF_Linear / LoF_Conv2d, STE autograd functions,
torch.ops.lofloat.*) on top of lofloat_b200's compiled extension and compares every forward bit for bit with the CPU oracle's
composition of the same steps.  Executed in its own interpreter by tests/test_reference_layers.py, because the reference's
_custom_ops.py and lofloat_b200/_custom_ops.py register the same torch.ops.lofloat.* names.

    python tests/reference_layers_check.py surface      # no GPU: import, module surface, factories, layer construction
    python tests/reference_layers_check.py gpu          # on a B200: forwards / backwards against the oracle
"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent))
import oracle_lib as L          # noqa: E402
import reference_pkg            # noqa: E402


def ofmt(d) -> L.Format:
    """oracle format of a FloatFormatDescriptor built by the reference's factories"""
    return L.Format(d.total_bits, d.mantissa_bits, d.bias, int(d.inf_behavior), int(d.nan_behavior), int(d.signedness), 0)


def surface():
    lof, ext = reference_pkg.load()
    # every name LoFloat/__init__.py:1 imports from the compiled module exists and the package imported on top of it
    for name in ("virtual_round", "virtual_mx_round", "RoundingMode", "InfBehavior", "NaNBehavior", "Signedness",
                 "FloatFormatDescriptor", "lof_gemm", "lof_conv2d", "pwl_silu"):
        assert hasattr(ext, name), name
        assert getattr(lof, name) is getattr(ext, name), name
    # enum values travel as ints through torch.ops.lofloat.* 
    want = dict(RoundToNearestEven=0, RoundTowardsZero=1, RoundAwayFromZero=2, StochasticRoundingA=3, RoundDown=5, RoundUp=6,
                StochasticRoundingB=8, StochasticRoundingC=9, True_StochasticRounding=10, ProbabilisticRounding=11, RoundToOdd=12,
                RoundTiesTowardsZero=13, StochasticRoundingD=14)
    for k, v in want.items():
        assert int(getattr(lof.RoundingMode, k)) == v and getattr(ext, k) == getattr(lof.RoundingMode, k), k
    assert int(lof.InfBehavior.Saturating) == 1 and int(lof.Signedness.Unsigned) == 1 and int(lof.NaNBehavior._754) == 1
    # the reference's factories build the extension's descriptor class
    p = lof.create_p3109_params(8, 4)
    assert isinstance(p, ext.FloatFormatDescriptor)
    assert (p.total_bits, p.mantissa_bits, p.bias, p.inf_behavior, p.signedness) == (8, 3, 8, lof.InfBehavior.Saturating, lof.Signedness.Signed)
    e8 = lof.create_e8m0_params()
    assert (e8.total_bits, e8.mantissa_bits, e8.bias, e8.signedness) == (8, 0, 127, lof.Signedness.Unsigned)
    assert lof.create_half_params().mantissa_bits == 10 and lof.create_single_params().bias == 127
    try:
        p.bias = 3
        raise AssertionError("descriptor attributes must be read-only")
    except AttributeError:
        pass
    for bad in (dict(is_inf_checker=object()), dict(is_nan_checker=object())):
        kw = dict(is_inf_checker=p.is_inf_checker, is_nan_checker=p.is_nan_checker)
        kw.update(bad)
        try:
            ext.FloatFormatDescriptor(8, 3, 7, lof.InfBehavior.Saturating, lof.NaNBehavior._3109, lof.Signedness.Signed, **kw)
            raise AssertionError("malformed checker accepted")
        except ValueError:
            pass
    # the three custom ops are registered with the reference's schemas (LoFloat/_custom_ops.py:34-108)
    for op in ("virtual_round", "virtual_mx_round", "lof_gemm"):
        assert hasattr(torch.ops.lofloat, op), op
    meta = torch.ops.lofloat.lof_gemm(torch.empty(5, 7, device="meta"), torch.empty(7, 3, device="meta"), 10, 0, 0, 1.0, 1.0)
    assert tuple(meta.shape) == (5, 3)
    # layers construct (formats registered into the reference's registry)
    lin = lof.LoF_Linear(64, 32, act_params=p, weight_params=p, bias_params=lof.create_half_params(), accum_mant_bits=10)
    assert lin.weight.shape == (32, 64) and lin._act_fid == lin._weight_fid
    lof.LoF_Linear(64, 32, act_params=p, weight_params=p, bias_params=p, scaling="mx", hadamard_transform=True)
    lof.LoF_Conv2d(8, 4, 3, act_params=p, weight_params=p, bias_params=p, padding=1)
    lof.LoF_Quantize(p, scaling="mx")
    # argument errors keep the reference's exception types
    try:
        lof.virtual_mx_round(torch.zeros(32, dtype=torch.float64), 32, p, e8)
        raise AssertionError("float64 MX input accepted")
    except RuntimeError as e:
        assert "float32 tensors only" in str(e)
    try:
        lof.lof_gemm(torch.zeros(4, 4), torch.zeros(4, 4), 10)
        raise AssertionError("CPU GEMM accepted")
    except RuntimeError as e:
        assert "requires CUDA tensors" in str(e)
    print("reference surface ok")


def gpu():
    lof, ext = reference_pkg.load()
    assert torch.cuda.is_available()
    dev = torch.device("cuda:0")
    torch.manual_seed(20260)
    rng = np.random.default_rng(5)
    RNE = L.RNE
    launches0 = ext.kernel_launch_count()

    def t(a):
        return torch.from_numpy(np.ascontiguousarray(a)).to(dev)

    def same(got, want, what):
        got = got.detach().cpu().numpy()
        assert got.shape == want.shape, (what, got.shape, want.shape)
        assert L.bits_equal(got, want), (what, L.first_mismatch(want, got, want))

    e4m3 = lof.create_p3109_params(8, 4)                 # P3109 binary8p4: bias 8 (formats.py:58)
    e5m2 = lof.create_p3109_params(8, 3)
    half = lof.create_half_params()
    e8m0 = lof.create_e8m0_params()
    x = (rng.standard_normal((48, 256)) * np.exp2(rng.integers(-6, 7, (48, 1)))).astype(np.float32)

    # ---- LoF_Quantize: per-tensor and MX, every deterministic mode the reference binds; STE backward
    for mode_name, mode in (("RoundToNearestEven", L.RNE), ("RoundTowardsZero", L.RTZ), ("RoundUp", L.RU), ("RoundToOdd", L.RTO)):
        rm = getattr(lof.RoundingMode, mode_name)
        q = lof.LoF_Quantize(e4m3, rounding_mode=rm)
        same(q(t(x)), L.o_round_params(x, ofmt(e4m3), mode), f"LoF_Quantize per_tensor {mode_name}")
        q = lof.LoF_Quantize(e4m3, rounding_mode=rm, scaling="mx", mx_block_size=32)
        same(q(t(x)), L.o_mx_round(x.ravel(), 32, ofmt(e8m0), ofmt(e4m3), mode).reshape(x.shape), f"LoF_Quantize mx {mode_name}")
    xt = t(x).requires_grad_(True)
    y = lof.LoF_Quantize(e5m2, scaling="mx", mx_block_size=16)(xt)
    g = torch.randn_like(y)
    y.backward(g)
    assert torch.equal(xt.grad, g), "STEMXRound backward must be the identity"
    try:
        lof.LoF_Quantize(e4m3, scaling="mx", mx_block_size=48)(t(x))
        raise AssertionError("indivisible MX block accepted")
    except ValueError:
        pass

    # ---- the traceable path: torch.ops.lofloat.* with registered format ids
    from LoFloat._custom_ops import register_format
    fid, sid = register_format(e4m3), register_format(e8m0)
    same(torch.ops.lofloat.virtual_round(t(x), fid, 0, 0, 4.0), L.o_round_params(x, ofmt(e4m3), RNE, scale=4.0), "ops.virtual_round")
    same(torch.ops.lofloat.virtual_mx_round(t(x), 32, fid, sid, 0, 0, 0),
         L.o_mx_round(x.ravel(), 32, ofmt(e8m0), ofmt(e4m3)).reshape(x.shape), "ops.virtual_mx_round")
    same(lof.virtual_round(t(x), 7), L.o_round_mantissa(x, 7, RNE), "virtual_round mantissa overload")

    # ---- LoF_Linear, per-tensor scaling (layers.py:376-398): round(x * sa), round(w * sw), lof_gemm / (sa * sw), + round(b * sb) / sb
    K, N = 256, 96
    for accum, sa, sw, sb in ((10, 1.0, 1.0, 1.0), (7, 4.0, 2.0, 8.0), (23, 1.0, 0.5, 1.0)):
        lin = lof.LoF_Linear(K, N, act_params=e4m3, weight_params=e4m3, bias_params=half, accum_mant_bits=accum,
                             act_scale_factor=sa, w_scale_factor=sw, b_scale_factor=sb).to(dev)
        with torch.no_grad():
            lin.bias.copy_(torch.randn(N, device=dev))
        w, b = lin.weight.detach().cpu().numpy(), lin.bias.detach().cpu().numpy()
        xq = L.o_round_params(x, ofmt(e4m3), RNE, scale=sa)
        wq = L.o_round_params(w, ofmt(e4m3), RNE, scale=sw)
        alpha = np.float32(1.0 / (sa * sw))
        acc = L.o_gemm(xq, np.ascontiguousarray(wq.T), accum, RNE, kc=0, alpha=float(alpha))
        bq = L.o_round_params(b, ofmt(half), RNE, scale=sb)
        if sb != 1.0:
            bq = bq / np.float32(sb)
        out = lin(t(x))
        same(out, (acc + bq[None, :]).astype(np.float32), f"LoF_Linear per_tensor accum={accum} scales=({sa},{sw},{sb})")
        out.sum().backward()                      # the GEMM is not differentiable in the reference either; the bias path is (STERound)
        assert torch.equal(lin.bias.grad, torch.full((N,), float(x.shape[0]), device=dev)), "STERound backward through the bias"
    # 3-D activations: (batch, tokens, K)
    lin = lof.LoF_Linear(K, N, act_params=e4m3, weight_params=e4m3, bias=False, bias_params=half, accum_mant_bits=10).to(dev)
    x3 = x[:24].reshape(2, 12, K)
    wq = L.o_round_params(lin.weight.detach().cpu().numpy(), ofmt(e4m3), RNE)
    want = L.o_gemm(L.o_round_params(x3.reshape(-1, K), ofmt(e4m3), RNE), np.ascontiguousarray(wq.T), 10, RNE, kc=0).reshape(2, 12, N)
    same(lin(t(x3)), want, "LoF_Linear 3-D input")

    # ---- LoF_Linear, scaling="mx" (layers.py:366-377), Hadamard rotation on / off (layers.py:363-365)
    for had, hb in ((False, None), (True, 64), (True, None)):
        lin = lof.LoF_Linear(K, N, act_params=e4m3, weight_params=e5m2, bias_params=half, accum_mant_bits=7, scaling="mx",
                             mx_block_size=32, hadamard_transform=had, hadamard_block_size=hb).to(dev)
        with torch.no_grad():
            lin.bias.copy_(torch.randn(N, device=dev))
        w, b = lin.weight.detach().cpu().numpy(), lin.bias.detach().cpu().numpy()
        xr = L.o_fwht(x, hb if hb is not None else K) if had else x
        xq = L.o_mx_round(xr.ravel(), 32, ofmt(e8m0), ofmt(e4m3)).reshape(x.shape)
        wq = L.o_mx_round(w.ravel(), 32, ofmt(e8m0), ofmt(e5m2)).reshape(w.shape)
        acc = L.o_gemm(xq, np.ascontiguousarray(wq.T), 7, RNE, kc=0)
        want = (acc + L.o_round_params(b, ofmt(half), RNE)[None, :]).astype(np.float32)
        same(lin(t(x)), want, f"LoF_Linear mx hadamard={had}/{hb}")

    # ---- LoF_Conv2d: im2col backend (per-tensor, groups, mx) and the native backend (layers.py:621-751)
    xi = (rng.standard_normal((2, 8, 9, 10))).astype(np.float32)

    def conv_oracle(xq, wq, groups, pad, stride, accum, alpha=1.0):
        cols = torch.nn.functional.unfold(torch.from_numpy(xq), wq.shape[2:], padding=pad, stride=stride).numpy()     # (B, K, L): pure data movement
        B_, Co = xq.shape[0], wq.shape[0]
        cg_out, kg = Co // groups, cols.shape[1] // groups
        outs = []
        for b_ in range(B_):
            per_g = [L.o_gemm(wq.reshape(Co, -1)[g_ * cg_out:(g_ + 1) * cg_out], np.ascontiguousarray(cols[b_, g_ * kg:(g_ + 1) * kg]),
                              accum, RNE, kc=0, alpha=alpha) for g_ in range(groups)]
            outs.append(np.concatenate(per_g, axis=0))
        return np.stack(outs)

    for groups, pad, stride in ((1, 1, 1), (2, 0, 2)):
        conv = lof.LoF_Conv2d(8, 6, 3, act_params=e4m3, weight_params=e4m3, bias_params=half, padding=pad, stride=stride, groups=groups,
                              accum_mant_bits=10, act_scale_factor=2.0).to(dev)
        with torch.no_grad():
            conv.bias.copy_(torch.randn(6, device=dev))
        w, b = conv.weight.detach().cpu().numpy(), conv.bias.detach().cpu().numpy()
        xq, wq = L.o_round_params(xi, ofmt(e4m3), RNE, scale=2.0), L.o_round_params(w, ofmt(e4m3), RNE)
        got = conv(t(xi))
        acc = conv_oracle(xq, wq, groups, pad, stride, 10, alpha=0.5).reshape(got.shape)
        same(got, (acc + L.o_round_params(b, ofmt(half), RNE).reshape(1, -1, 1, 1)).astype(np.float32), f"LoF_Conv2d im2col groups={groups}")
    conv = lof.LoF_Conv2d(8, 6, 2, act_params=e4m3, weight_params=e4m3, bias_params=half, accum_mant_bits=10, scaling="mx",
                          mx_block_size=32).to(dev)          # K = 8 * 2 * 2 = 32
    w = conv.weight.detach().cpu().numpy()
    cols = torch.nn.functional.unfold(torch.from_numpy(xi), (2, 2)).numpy()                                            # (B, K, L)
    cq = L.o_mx_round(np.ascontiguousarray(cols.transpose(0, 2, 1)).ravel(), 32, ofmt(e8m0), ofmt(e4m3)).reshape(2, -1, 32).transpose(0, 2, 1)
    wq = L.o_mx_round(w.reshape(6, -1).ravel(), 32, ofmt(e8m0), ofmt(e4m3)).reshape(6, 32)
    got = conv(t(xi))
    want = np.stack([L.o_gemm(wq, np.ascontiguousarray(cq[b_]), 10, RNE, kc=0) for b_ in range(2)]).reshape(got.shape)
    same(got, (want + L.o_round_params(conv.bias.detach().cpu().numpy(), ofmt(half), RNE).reshape(1, -1, 1, 1)).astype(np.float32),
         "LoF_Conv2d im2col mx")
    # native backend: the reference reduces in NHWC / KRSC order (Lof_kernel.cu:156-228), i.e. k = (r, s, c)
    conv = lof.LoF_Conv2d(8, 6, 3, act_params=e4m3, weight_params=e4m3, bias=False, bias_params=half, padding=1, accum_mant_bits=10,
                          backend="cutlass").to(dev)
    w = conv.weight.detach().cpu().numpy()
    xq, wq = L.o_round_params(xi, ofmt(e4m3), RNE), L.o_round_params(w, ofmt(e4m3), RNE)
    cols = torch.nn.functional.unfold(torch.from_numpy(xq), (3, 3), padding=1).numpy().reshape(2, 8, 9, -1)             # (B, c, rs, L)
    cols_rsc = np.ascontiguousarray(cols.transpose(0, 2, 1, 3)).reshape(2, 72, -1)                                     # k = (r, s, c)
    w_rsc = np.ascontiguousarray(wq.reshape(6, 8, 9).transpose(0, 2, 1)).reshape(6, 72)
    got = conv(t(xi))
    want = np.stack([L.o_gemm(w_rsc, np.ascontiguousarray(cols_rsc[b_]), 10, RNE, kc=0) for b_ in range(2)]).reshape(got.shape)
    same(got, want, "LoF_Conv2d native backend")

    # ---- the same LoF_Linear call on the tensor cores: set_gemm_policy routes lof_gemm to the chunked tcgen05 kernels
    ext.set_gemm_policy(k_chunk=32, operand_dtype="e4m3")
    lin = lof.LoF_Linear(512, 256, act_params=lof.create_p3109_params(8, 4, bias=7), weight_params=lof.create_p3109_params(8, 4, bias=7),
                         bias=False, bias_params=half, accum_mant_bits=10).to(dev)
    xs = rng.uniform(-1, 1, (128, 512)).astype(np.float32)
    f7 = L.FORMATS["e4m3_b7"]
    want = L.o_gemm_chunked(L.o_round_params(xs, f7, RNE), np.ascontiguousarray(L.o_round_params(lin.weight.detach().cpu().numpy(), f7, RNE).T), 10, 32)
    same(lin(t(xs)), want, "LoF_Linear with the chunked tensor-core GEMM policy")
    ext.set_gemm_policy()
    torch.cuda.synchronize()
    n_launch = ext.kernel_launch_count() - launches0
    assert n_launch > 50, n_launch
    print(f"reference layers ok: {n_launch} kernel launches through lofloat_b200/LoFloat.so -> liblofloat_b200.so")


if __name__ == "__main__":
    {"surface": surface, "gpu": gpu}[sys.argv[1]]()
