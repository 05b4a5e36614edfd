"""C1 (n = 2^24, 128 MiB of traffic per launch): launch geometry sweep with an L2 flush between launches."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import lofloat_b200 as lof  # noqa: E402

dev = torch.device("cuda:0")
n = 1 << 24
x = torch.randn(n, device=dev) * 4
e4m3 = lof.create_p3109_params(8, 4, bias=7)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def best(fn, reps=20):
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[0], ts[len(ts) // 2]


for unroll in (4, 8):
    for ctas in (2, 3, 4, 6, 8, 12, 16):
        lof.set_option("round.unroll", unroll)
        lof.set_option("round.ctas_per_sm", ctas)
        lo, med = best(lambda: lof.virtual_round(x, e4m3))
        print(f"unroll {unroll} ctas {ctas:2d}: best {lo * 1e3:6.1f} us = {8.0 * n / lo / 1e6:6.0f} GB/s   median {med * 1e3:6.1f} us", flush=True)
