"""The drop-in claim, tested literally: the reference's unmodified LoFloat/*.py (LoF_Quantize, LoF_Linear per-tensor and
scaling="mx" with / without the Hadamard rotation, LoF_Conv2d im2col / mx / native backends, STE backward,
torch.ops.lofloat.*) run on top of the compiled extension lofloat_b200/LoFloat.so and match the oracle composition bit for
bit.  The checks live in tests/reference_layers_check.py and run in their own interpreter (the reference package and
lofloat_b200 register the same torch.ops.lofloat.* names)."""
import subprocess
import sys
from pathlib import Path

import pytest

import reference_pkg

HERE = Path(__file__).resolve().parent


def _run(which: str, timeout: int):
    proc = subprocess.run([sys.executable, str(HERE / "reference_layers_check.py"), which], stdout=subprocess.PIPE,
                          stderr=subprocess.STDOUT, text=True, timeout=timeout)
    assert proc.returncode == 0, proc.stdout[-4000:]
    return proc.stdout


def _need_package():
    if not reference_pkg.available():
        pytest.skip("oracle/_ref/lofloat_ref_py.zip (make -C oracle refpy, needs /root/reference) or lofloat_b200/LoFloat.so not built")


def test_reference_package_imports_on_the_compiled_extension(oracle_mod):
    _need_package()
    assert "reference surface ok" in _run("surface", 300)


@pytest.mark.gpu
def test_reference_layers_bit_exact_on_gpu(oracle_mod, cuda_dev):
    _need_package()
    assert "reference layers ok" in _run("gpu", 900)
