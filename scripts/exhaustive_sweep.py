#!/usr/bin/env python
"""Exhaustive parity sweep: EVERY fp32 bit pattern x the <= 8-bit formats x the 9 deterministic rounding modes, the
CUDA path (default kernels, through lofloat_b200.virtual_round) against the REFERENCE'S OWN scalar virtual_round
(lo_float.h:2843-2939, compiled unmodified into oracle/_ref/libref_shim.so, looped under OpenMP on the host cores;
falls back to the oracle port oracle/liblof_oracle.so when the shim is absent and says so).

    python scripts/exhaustive_sweep.py [--log FILE] [--chunk-log2 26] [--formats a,b,...] [--modes 0,1,...]

TEST INFRASTRUCTURE (it executes oracle/): never imported by the package.  Writes one line per (format, mode) with the
number of patterns compared and of mismatches (bit patterns compared as uint32, NaNs included), and a final verdict.
The committed log of the full run is profiles/r02_exhaustive_sweep.log.
"""
from __future__ import annotations

import argparse
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log", default="")
    ap.add_argument("--chunk-log2", type=int, default=26)
    ap.add_argument("--formats", default="e4m3_b7,p3109_8p4,e5m2,e3m2,e2m3,e2m1,e8m0")
    ap.add_argument("--modes", default="")
    ap.add_argument("--stride-log2", type=int, default=0, help="compare only every 2^s-th chunk (quick runs)")
    args = ap.parse_args()

    import torch
    import oracle_lib as L
    import lofloat_b200 as lof

    dev = torch.device("cuda", 0)
    threads = os.cpu_count() or 1
    use_ref = L.ref() is not None
    if not use_ref:
        L.build_oracle()
    modes = [int(m) for m in args.modes.split(",")] if args.modes else list(L.DETERMINISTIC_MODES)
    names = args.formats.split(",")
    chk = lof.create_single_params()
    descs = {n: lof.FloatFormatDescriptor(L.FORMATS[n].bitwidth, L.FORMATS[n].mantissa_bits, L.FORMATS[n].bias, L.FORMATS[n].ov_behavior,
                                          L.FORMATS[n].na_behavior, L.FORMATS[n].is_signed, chk.is_inf_checker, chk.is_nan_checker) for n in names}
    out = open(args.log, "w") if args.log else None

    def say(s):
        print(s, flush=True)
        if out:
            out.write(s + "\n")
            out.flush()

    say(f"# exhaustive sweep: CUDA virtual_round (default kernels) vs {'the compiled reference (oracle/_ref/libref_shim.so)' if use_ref else 'the oracle port (liblof_oracle.so)'}"
        f", OpenMP x{threads}; GPU {torch.cuda.get_device_name(0)}")
    say(f"# formats {names}; modes {[L.MODE_NAMES[m] for m in modes]}; chunk 2^{args.chunk_log2}; stride 2^{args.stride_log2}")
    chunk = 1 << args.chunk_log2
    nchunks = (1 << 32) // chunk
    step = 1 << args.stride_log2
    counts = {(n, m): [0, 0] for n in names for m in modes}
    first_bad = {}
    t0 = time.time()
    cpu_s = 0.0
    pinned = torch.empty(chunk, dtype=torch.int32).pin_memory()
    for c in range(0, nchunks, step):
        base = c * chunk
        bits = np.arange(base, base + chunk, dtype=np.uint64).astype(np.uint32)
        x_np = bits.view(np.float32)
        x = torch.from_numpy(bits.view(np.int32)).to(dev).view(torch.float32)
        for n in names:
            f = L.FORMATS[n]
            for m in modes:
                y = lof.virtual_round(x, descs[n], m, 0)
                tc = time.time()
                want = L.r_round_params(x_np, f, m, 0, threads=threads) if use_ref else L.o_round_params(x_np, f, m, 0)
                cpu_s += time.time() - tc
                pinned.copy_(torch.from_numpy(want.view(np.int32)))
                w = pinned.to(dev, non_blocking=True)
                ne = (y.view(torch.int32) != w)
                bad = int(ne.sum().item())
                counts[(n, m)][0] += chunk
                counts[(n, m)][1] += bad
                if bad and (n, m) not in first_bad:
                    i = int(ne.nonzero()[0])
                    first_bad[(n, m)] = (f"in=0x{(base + i) & 0xFFFFFFFF:08x} cuda=0x{int(y.view(torch.int32)[i]) & 0xFFFFFFFF:08x} "
                                         f"ref=0x{int(w[i]) & 0xFFFFFFFF:08x}")
                del y, w, ne
        if (c // step) % 8 == 7:
            print(f"  chunk {c + 1}/{nchunks}  elapsed {time.time() - t0:.0f} s (cpu {cpu_s:.0f} s)", flush=True)
    total_bad = 0
    for n in names:
        for m in modes:
            cnt, bad = counts[(n, m)]
            total_bad += bad
            say(f"{n:10s} {L.MODE_NAMES[m]:5s} patterns {cnt:>11d}  mismatches {bad}" + (f"  first: {first_bad[(n, m)]}" if bad else ""))
    say(f"# total: {sum(v[0] for v in counts.values())} comparisons, {total_bad} mismatches, {time.time() - t0:.0f} s wall ({cpu_s:.0f} s in the CPU reference)")
    say("# VERDICT: " + ("BIT-EXACT on every pattern" if total_bad == 0 else "MISMATCHES FOUND"))
    return 0 if total_bad == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
