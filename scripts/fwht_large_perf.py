import sys, torch
sys.path.insert(0, '/root/repo')
import lofloat_b200 as lof
x = torch.randn(16384, 16384, device='cuda')
for b in (1024, 4096, 8192):
    lof.fwht(x, b); torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): lof.fwht(x, b)
    e.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(e) / 5
    print(f"fwht block {b}: {ms:.3f} ms  ({8.0 * x.numel() / ms / 1e6:.0f} GB/s at 8 B/element)")
