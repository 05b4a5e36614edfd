import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run on the GPU box with -m gpu)")
    config.addinivalue_line("markers", "slow: long-running sweep (still bounded to a few minutes)")


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle_lib
    oracle_lib.build_oracle()
    return oracle_lib


@pytest.fixture(scope="session")
def ref_available():
    """The compiled reference only exists where /root/reference was present at build time."""
    import oracle_lib
    if Path("/root/reference/src/lo_float.h").exists():
        try:
            oracle_lib.build_ref()
        except Exception:
            pass
    return oracle_lib.ref() is not None


@pytest.fixture(scope="session")
def cuda_dev():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    return torch.device("cuda:0")


@pytest.fixture(autouse=True)
def _deterministic_torch_rng(request):
    """Every test starts from a fixed torch seed (CPU and CUDA generators), derived from the test's name: no assertion in
    this suite may depend on an unseeded draw (round 1: an unseeded randn made `pytest -m gpu` flaky on the driver's box)."""
    try:
        import torch
    except ImportError:
        return
    import zlib
    torch.manual_seed(zlib.crc32(request.node.nodeid.encode()) & 0x7FFFFFFF)
