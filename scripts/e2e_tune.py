"""Host-buffer entry point: chunk size sweep of the 3-stream H2D / kernel / D2H pipeline (pinned 1 GiB in, 1 GiB out)."""
import ctypes as C
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from lofloat_b200 import _lib  # noqa: E402

lib = _lib.lib()
n = 1 << 28
xh = torch.randn(n).pin_memory()
yh = torch.empty(n).pin_memory()
fmt, ps = _lib.CFormat(8, 3, 7, 1, 0, 0, 0), _lib.CProjSpec(0, 0, 0)
torch.cuda.init()
for mib in (4, 8, 16, 32, 64, 128):
    _lib.set_option("host.chunk_bytes", mib << 20)
    lib.lof_round_params_host(xh.data_ptr(), yh.data_ptr(), n, 0, C.byref(fmt), C.byref(ps), 1.0, None, 0, 0)
    t0 = time.perf_counter()
    for _ in range(4):
        lib.lof_round_params_host(xh.data_ptr(), yh.data_ptr(), n, 0, C.byref(fmt), C.byref(ps), 1.0, None, 0, 0)
    dt = (time.perf_counter() - t0) / 4
    print(f"chunk {mib:4d} MiB: {8.0 * n / dt / 1e9:6.1f} GB/s algorithmic ({4.0 * n / dt / 1e9:5.1f} GB/s per direction)")
