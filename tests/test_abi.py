"""CPU: the C-ABI library loads and exports every symbol include/lofloat_b200.h declares; argument
validation works without a GPU; compute without a GPU fails loudly (no CPU fallback)."""
import ctypes as C
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    from lofloat_b200 import _lib
    if not _lib.LIB_PATH.exists():
        g.build()
    return _lib.lib()


def _declared_symbols():
    text = (ROOT / "include" / "lofloat_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lof_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    from lofloat_b200 import _lib
    declared = _declared_symbols()
    assert len(declared) >= 15
    assert set(declared) == set(_lib.ABI_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name


def test_version_and_strings(lib):
    assert lib.lof_abi_version() == 1
    assert lib.lof_status_string(0) == b"ok"
    assert b"CPU fallback" in lib.lof_status_string(4)


def test_struct_layout_matches_header():
    from lofloat_b200._lib import CFormat, CProjSpec, CRng
    assert C.sizeof(CFormat) == 28 and C.sizeof(CProjSpec) == 12 and C.sizeof(CRng) == 16


def test_priv_max_normal(lib):
    from lofloat_b200._lib import CFormat
    want = {(8, 3, 7): 480.0, (8, 3, 8): 240.0, (8, 2, 15): 114688.0, (6, 2, 3): 28.0, (6, 3, 1): 7.5, (4, 1, 1): 6.0}
    for (bw, m, b), v in want.items():
        f = CFormat(bw, m, b, 1, 0, 0, 0)
        assert lib.lof_mx_priv_max_normal(C.byref(f)) == v


def test_options(lib):
    from lofloat_b200 import _lib
    _lib.set_option("round.unroll", 2)
    assert _lib.get_option("round.unroll") == 2
    _lib.set_option("round.unroll", 0)
    with pytest.raises(_lib.LoFloatError):
        _lib.set_option("round.unroll", 3)
    with pytest.raises(_lib.LoFloatError):
        _lib.set_option("no.such.option", 1)


def test_validation_without_gpu(lib):
    from lofloat_b200._lib import (CFormat, CProjSpec, LOF_ERR_INVALID_ARG, LOF_ERR_UNSUPPORTED)
    f = CFormat(8, 3, 7, 1, 0, 0, 0)
    ps = CProjSpec(0, 0, 0)
    assert lib.lof_round_params(None, None, -1, 0, C.byref(f), C.byref(ps), 1.0, None, 0, None) == LOF_ERR_INVALID_ARG
    assert lib.lof_round_params(None, None, 8, 9, C.byref(f), C.byref(ps), 1.0, None, 0, None) == LOF_ERR_INVALID_ARG
    bad = CProjSpec(99, 0, 0)
    assert lib.lof_round_params(None, None, 8, 0, C.byref(f), C.byref(bad), 1.0, None, 0, None) == LOF_ERR_INVALID_ARG
    # stochastic bit count larger than the bits being dropped: undefined in the reference, refused here
    sr = CProjSpec(3, 0, 24)
    assert lib.lof_round_params(None, None, 8, 0, C.byref(f), C.byref(sr), 1.0, None, 0, None) == LOF_ERR_INVALID_ARG
    assert b"stoch_length" in lib.lof_last_error_string()
    # MX block sizes: Lof_kernel.cu:83-92
    e8 = CFormat(8, 0, 127, 0, 0, 1, 0)
    st = lib.lof_mx_round_f32(None, None, 96, 48, C.byref(e8), C.byref(f), C.byref(ps), C.byref(ps), None, 0, None)
    assert st == LOF_ERR_UNSUPPORTED and b"16/32/64/128/256" in lib.lof_last_error_string()
    st = lib.lof_mx_round_f32(None, None, 96, 0, C.byref(e8), C.byref(f), C.byref(ps), C.byref(ps), None, 0, None)
    assert st == LOF_ERR_INVALID_ARG
    # n == 0 is a no-op that needs no device
    assert lib.lof_round_params(None, None, 0, 0, C.byref(f), C.byref(ps), 1.0, None, 0, None) == 0


def test_conv2d_geometry_and_validation_without_gpu(lib):
    """lof_conv2d_out_size == the reference's P / Q formula (Lof_kernel.cu:194-195) == torch's conv output shape; the conv /
    linear entry points validate their arguments before touching a device."""
    import torch.nn.functional as F
    from lofloat_b200._lib import CConv2d, CProjSpec, LOF_ERR_INVALID_ARG, LOF_ERR_UNSUPPORTED, LOF_CONV_K_CRS
    out = C.c_int64(0)
    for (h, k, p, s, d) in ((12, 3, 1, 1, 1), (17, 3, 2, 2, 1), (9, 5, 3, 3, 2), (31, 3, 4, 2, 1), (224, 7, 3, 2, 1), (5, 1, 0, 1, 1)):
        assert lib.lof_conv2d_out_size(h, k, p, s, d, C.byref(out)) == 0
        want = F.conv2d(torch.zeros(1, 1, h, h), torch.zeros(1, 1, k, k), padding=p, stride=s, dilation=d).shape[-1]
        assert out.value == want == (h + 2 * p - d * (k - 1) - 1) // s + 1
    assert lib.lof_conv2d_out_size(2, 3, 0, 1, 1, C.byref(out)) == 0 and out.value == 0          # window does not fit
    assert lib.lof_conv2d_out_size(8, 3, 0, 0, 1, C.byref(out)) == LOF_ERR_INVALID_ARG
    ps = CProjSpec(0, 0, 0)
    geo = CConv2d(1, 8, 10, 10, 16, 3, 3, 1, 1, 1, 1, 1, 1, 1, LOF_CONV_K_CRS)
    assert lib.lof_conv2d_f32(None, None, None, None, None, 10, C.byref(ps), 0, 1.0, None, None) == LOF_ERR_INVALID_ARG
    bad = CConv2d(1, 8, 10, 10, 16, 3, 3, 1, 1, 1, 1, 1, 1, 3, LOF_CONV_K_CRS)                    # groups must divide C_in / C_out
    assert lib.lof_conv2d_f32(None, None, None, None, C.byref(bad), 10, C.byref(ps), 0, 1.0, None, None) == LOF_ERR_INVALID_ARG
    bad = CConv2d(1, 8, 2, 2, 16, 3, 3, 0, 0, 1, 1, 1, 1, 1, LOF_CONV_K_CRS)
    assert lib.lof_conv2d_f32(None, None, None, None, C.byref(bad), 10, C.byref(ps), 0, 1.0, None, None) == LOF_ERR_INVALID_ARG
    assert b"non-positive output spatial size" in lib.lof_last_error_string()
    bad = CConv2d(1, 8, 10, 10, 16, 3, 3, 1, 1, 1, 1, 1, 1, 1, 7)                                 # unknown reduction order
    assert lib.lof_conv2d_f32(None, None, None, None, C.byref(bad), 10, C.byref(ps), 0, 1.0, None, None) == LOF_ERR_INVALID_ARG
    assert lib.lof_conv2d_f32(None, None, None, None, C.byref(geo), 24, C.byref(ps), 0, 1.0, None, None) == LOF_ERR_INVALID_ARG
    geo0 = CConv2d(0, 8, 10, 10, 16, 3, 3, 1, 1, 1, 1, 1, 1, 1, LOF_CONV_K_CRS)                   # empty batch: no-op, no device needed
    assert lib.lof_conv2d_f32(None, None, None, None, C.byref(geo0), 10, C.byref(ps), 0, 1.0, None, None) == 0
    assert lib.lof_linear_f32(None, None, None, None, -1, 4, 4, 10, C.byref(ps), 0, 1.0, None, None) == LOF_ERR_INVALID_ARG
    assert lib.lof_linear_f32(None, None, None, None, 0, 4, 4, 10, C.byref(ps), 0, 1.0, None, None) == 0
    assert lib.lof_linear_f32(None, None, None, None, 1 << 31, 4, 4, 10, C.byref(ps), 0, 1.0, None, None) == LOF_ERR_UNSUPPORTED
    import lofloat_b200 as lof
    with pytest.raises(RuntimeError, match="CUDA tensors"):
        lof.lof_conv2d(torch.zeros(1, 8, 10, 10), torch.zeros(16, 8, 3, 3), 1, 1, 1, 1, 1, 1, 10)
    with pytest.raises(RuntimeError, match="4-D"):
        lof.lof_conv2d(torch.zeros(8, 10, 10), torch.zeros(16, 8, 3, 3), 1, 1, 1, 1, 1, 1, 10)
    with pytest.raises(RuntimeError, match="incompatible shapes"):
        lof.lof_linear(torch.zeros(4, 7), torch.zeros(5, 8), None, 10)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_compute_without_gpu_fails_loudly(lib):
    import lofloat_b200 as lof
    from lofloat_b200._lib import CFormat, CProjSpec, LOF_ERR_NO_DEVICE, LOF_ERR_CUDA
    f = CFormat(8, 3, 7, 1, 0, 0, 0)
    ps = CProjSpec(0, 0, 0)
    buf = (C.c_float * 16)()
    st = lib.lof_round_params(buf, buf, 16, 0, C.byref(f), C.byref(ps), 1.0, None, 0, None)
    assert st in (LOF_ERR_NO_DEVICE, LOF_ERR_CUDA)
    with pytest.raises(RuntimeError):
        lof.virtual_round(torch.ones(16), lof.create_p3109_params(8, 4))
