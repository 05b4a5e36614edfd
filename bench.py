#!/usr/bin/env python
"""bench.py -- headline benchmark of the LoFloat virtual-rounding hot path on B200.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (BASELINE.json configs[1], "C2"): virtual_round of a 1 GiB fp32 tensor (2^28 elements,
randn * 2^U{-12..12}, seed 2024) over the format x mode sweep
    formats  E5M2 (8,2,15)  E4M3 (8,3,7)  E3M2 (6,2,3)  E2M1 (4,1,1)  bf16-like (16,7,127,Extended)
    modes    RNE, RTZ, RoundToOdd, StochasticRoundingA(len 8)
One STEP = the 20 (format, mode) passes over the tensor = 20 kernel launches, 40 GiB of algorithmic
traffic (8 B/element: 4 read + 4 written).  With N GPUs the tensor is N GiB, row-sharded, one
1 GiB shard per rank, no collective (weak scaling).  Metric: aggregate GB/s of algorithmic bytes.

The line also carries: `roofline` for the dominant kernel (CUDA-event time per launch inside the
timed region), `e2e` (same sweep through the C ABI's HOST-buffer entry point, H2D/D2H inside the
timed region), `cpu_baseline` (the reference's own scalar CPU path on this box's cores, bounded
sample) and short side measurements of mx_round and lof_gemm (`other`, outside the timed region).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N_ELEMS = 1 << 28
FORMATS = [("e5m2", (8, 2, 15, 1, 0)), ("e4m3", (8, 3, 7, 1, 0)), ("e3m2", (6, 2, 3, 1, 0)),
           ("e2m1", (4, 1, 1, 1, 0)), ("bf16like", (16, 7, 127, 0, 0))]          # (bw, m, bias, ov, signed)
MODES = [("RNE", 0, 0), ("RTZ", 1, 0), ("RoundToOdd", 12, 0), ("SR_A8", 3, 8)]   # (name, mode, stoch_len)
KERNEL_OF_MODE = {"RNE": "round_vec_kernel<f32,RoundToNearestEven>", "RTZ": "round_vec_kernel<f32,RoundTowardsZero>",
                  "RoundToOdd": "round_vec_kernel<f32,RoundToOdd>", "SR_A8": "round_vec_kernel<f32,StochasticRoundingA>"}
WORKLOAD = "C2: virtual_round fp32 1GiB/GPU, 5 formats x 4 modes (RNE,RTZ,RoundToOdd,SR-A8), row-sharded"


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def mark(self):
        """Wall-clock marker: rows received after mark_start and before mark_end are 'under load'."""
        return time.time()

    def stop(self, t_begin=None, t_end=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for (ts, r) in self.rows if (t_begin is None or ts >= t_begin) and (t_end is None or ts <= t_end + 0.05)]
        if not rows:                                     # region shorter than one sample: take the nearest samples
            rows = [r for (_, r) in self.rows[-3:]]
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def make_structs():
    from lofloat_b200 import _lib
    combos = []
    for fname, (bw, m, bias, ov, sg) in FORMATS:
        for mname, mode, sl in MODES:
            combos.append((f"{fname}/{mname}", mname, _lib.CFormat(bw, m, bias, ov, 0, sg, 0), _lib.CProjSpec(mode, 0, sl)))
    return combos


# ----------------------------------------------------------------------------------------- own arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from lofloat_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU; lofloat_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # stdout carries exactly ONE JSON line (NCCL prints its version banner there otherwise)
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.lib()
    combos = make_structs()
    n = N_ELEMS
    first_index = rank * n                                   # this rank's row shard of the N GiB tensor
    g = torch.Generator(device=dev).manual_seed(2024 + rank)
    x = torch.randn(n, device=dev, generator=g)
    x *= torch.exp2(torch.randint(-12, 13, (n,), device=dev, generator=g).float())
    y = torch.empty_like(x)
    stream = torch.cuda.current_stream().cuda_stream
    rng = _lib.CRng(2024, 0)

    def one_pass(i, fmt, ps, evs=None):
        if evs is not None:
            evs[0].record()
        st = lib.lof_round_params(x.data_ptr(), y.data_ptr(), n, 0, C.byref(fmt), C.byref(ps), 1.0, C.byref(rng),
                                  first_index, stream)
        if evs is not None:
            evs[1].record()
        if st != 0:
            raise RuntimeError(lib.lof_last_error_string().decode())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                                      # nvidia-smi needs a moment to come up: start before warm-up
    for _ in range(args.warmup):
        for i, (_, _, fmt, ps) in enumerate(combos):
            one_pass(i, fmt, ps)
    barrier()
    per_launch = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in combos]
                  for _ in range(args.steps)]
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = _lib.kernel_launch_count()
    barrier()
    w_begin = time.time()
    t0.record()
    for s in range(args.steps):
        for i, (_, _, fmt, ps) in enumerate(combos):
            one_pass(i, fmt, ps, per_launch[s][i])
    t1.record()
    barrier()
    launches = _lib.kernel_launch_count() - launches0
    clocks = sampler.stop(w_begin, time.time()) if rank == 0 else None
    ms = torch.tensor([t0.elapsed_time(t1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    bytes_per_pass = 8.0 * n
    value = bytes_per_pass * len(combos) * args.steps * world / (total_ms * 1e-3) / 1e9

    # per-kernel launch durations (CUDA events inside the timed region)
    by_kernel, by_combo = {}, {}
    for s in range(args.steps):
        for i, (cname, mname, _, _) in enumerate(combos):
            d = per_launch[s][i][0].elapsed_time(per_launch[s][i][1])
            by_kernel.setdefault(mname, []).append(d)
            by_combo.setdefault(cname, []).append(d)
    peak, peak_src = peaks()
    kernels = {KERNEL_OF_MODE[k]: {"avg_ms": statistics.mean(v), "GBps": bytes_per_pass / (statistics.mean(v) * 1e-3) / 1e9,
                                   "launches": len(v)} for k, v in by_kernel.items()}
    dom_mode = max(by_kernel, key=lambda k: sum(by_kernel[k]))
    dom = kernels[KERNEL_OF_MODE[dom_mode]]
    traffic = None
    tp = ROOT / "profiles" / "traffic.json"                      # dram bytes per launch from the ncu --set full capture
    if tp.exists():
        try:
            traffic = json.loads(tp.read_text()).get(KERNEL_OF_MODE[dom_mode])
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": KERNEL_OF_MODE[dom_mode], "achieved": dom["GBps"], "peak": peak, "unit": "GB/s",
                "frac": dom["GBps"] / peak, "traffic": traffic, "algorithmic_bytes_per_launch": bytes_per_pass,
                "peak_source": peak_src, "share_of_step": sum(by_kernel[dom_mode]) / sum(sum(v) for v in by_kernel.values())}

    # ---- e2e: same sweep through the C ABI's host-buffer entry point (pinned host memory) ----
    e2e = None
    try:
        n_h = n
        xh = torch.empty(n_h, dtype=torch.float32).pin_memory()
        xh.copy_(x[:n_h])
        yh = torch.empty(n_h, dtype=torch.float32).pin_memory()
        if args.e2e_steps <= 0:
            raise RuntimeError("skipped (--e2e-steps 0)")
        e2e_steps = min(args.steps, args.e2e_steps)

        def host_sweep():
            for _, _, fmt, ps in combos:
                st = lib.lof_round_params_host(xh.data_ptr(), yh.data_ptr(), n_h, 0, C.byref(fmt), C.byref(ps), 1.0,
                                               C.byref(rng), first_index, local)
                if st != 0:
                    raise RuntimeError(lib.lof_last_error_string().decode())
        host_sweep()                                             # warm-up (allocates the staging buffers)
        barrier()
        l0 = _lib.kernel_launch_count()
        w0 = time.perf_counter()
        for _ in range(e2e_steps):
            host_sweep()
        torch.cuda.synchronize()
        w1 = time.perf_counter()
        e2e_launches = _lib.kernel_launch_count() - l0
        barrier()
        el = torch.tensor([w1 - w0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(el, op=dist.ReduceOp.MAX)
        e2e = {"value": 8.0 * n_h * len(combos) * e2e_steps * world / float(el.item()) / 1e9, "unit": "GB/s",
               "h2d_bytes_per_step": 4 * n_h * len(combos), "d2h_bytes_per_step": 4 * n_h * len(combos),
               "steps": e2e_steps, "api": "lof_round_params_host (C ABI, pinned host buffers, chunked 3-stream pipeline)",
               "gpu_launches": int(e2e_launches)}
        del xh, yh
    except Exception as exc:  # pinned allocation can fail on a small host
        e2e = {"value": None, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "error": str(exc)}

    workloads = {}
    if not args.no_side:
        try:
            workloads = named_workloads(torch, dist, dev, rank, world, args)
        except Exception as exc:
            workloads = {"error": str(exc)}
    other = {}
    cpu_baseline = None
    if rank == 0 and not args.no_side:
        other = side_measurements(torch, lib, _lib, dev, x, y, stream)
        if world == 1:
            cpu_baseline = cpu_reference_sample(x[: 1 << 26].cpu().numpy(), budget_s=20.0)
            try:
                other.update(c1_and_cpu_side(torch, dev))
            except Exception as exc:
                other["cpu_side_error"] = str(exc)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    line = {
        "metric": "virtual_round_GBps", "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32 bit arithmetic on fp32 containers", "data": "synthetic",
        "config": {"workload": WORKLOAD, "elements_per_gpu": n, "passes_per_step": len(combos),
                   "bytes_per_element": 8, "l2": "inputs (1 GiB) larger than L2 (126 MB); no flush needed",
                   "parallelism": f"row-shard x{world}, no collective"},
        "roofline": roofline, "kernels": kernels,
        "sweep_GBps": {k: bytes_per_pass / (statistics.mean(v) * 1e-3) / 1e9 for k, v in by_combo.items()},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "cpu_baseline": cpu_baseline, "workloads": workloads, "other": other,
    }
    print(json.dumps(line))


def bf16_peak():
    """(measured dense bf16 TFLOP/s burst, sustained) from MEASURED_PEAKS.json, else the profiling guide's fallback."""
    p = ROOT / "MEASURED_PEAKS.json"
    try:
        d = json.loads(p.read_text())
        return float(d["bf16_tflops"]), float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 1590.0, 1400.0, "fallback (B200_PROFILING.md)"


def named_workloads(torch, dist, dev, rank, world, args):
    """BASELINE.json configs C3 / C4 / C5 (and C1 on rank 0) measured at THIS world size, inside the same JSON line as the
    headline sweep: every rank takes part (row shards for MX, the R x C grid with its NCCL exchange for the GEMMs), device
    time by CUDA events, max over ranks.  Each entry carries its own roofline fraction per GPU."""
    import lofloat_b200 as lof
    from lofloat_b200 import sharding as S
    hbm_peak, hbm_src = peaks()
    bf16_burst, bf16_sust, bf16_src = bf16_peak()
    out = {}

    def sync_max(ms):
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def timed(fn, reps, warm=2):
        keep = None
        for _ in range(warm):
            keep = fn()
        del keep
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            keep = fn()
        b.record()
        torch.cuda.synchronize()
        return sync_max(a.elapsed_time(b) / reps), keep

    e8m0 = lof.create_e8m0_params()
    # ------------------------------------------------------------------ C1: fp32 -> E4M3 RNE saturating, 2^24 elements (the reference's own CPU case)
    try:
        n1 = 1 << 24
        NB = 16                                                                   # 16 input / output pairs = 2 GiB: nothing survives in the 126 MB L2
        g = torch.Generator(device=dev).manual_seed(12345 + rank)
        xs = [torch.randn(n1, device=dev, generator=g) * 4 for _ in range(NB)]
        ys = [torch.empty(n1, device=dev) for _ in range(NB)]
        e4 = lof.create_p3109_params(8, 4, bias=7)
        from lofloat_b200 import _lib
        lib = _lib.lib()
        fmt_c, ps_c, rng_c = e4._c(), _lib.CProjSpec(0, 0, 0), _lib.CRng(0, 0)
        stream = torch.cuda.current_stream().cuda_stream

        def sweep():
            for xi, yi in zip(xs, ys):
                lib.lof_round_params(xi.data_ptr(), yi.data_ptr(), n1, 0, C.byref(fmt_c), C.byref(ps_c), 1.0, C.byref(rng_c), 0, stream)
        ms, _ = timed(sweep, 5, warm=2)
        per_launch_us = ms * 1e3 / NB
        gbps = 8.0 * n1 / (per_launch_us * 1e-6) / 1e9

        def copy_sweep():
            for xi, yi in zip(xs, ys):
                yi.copy_(xi)
        ms_cp, _ = timed(copy_sweep, 5, warm=2)
        gbps_cp = 8.0 * n1 / (ms_cp * 1e-3 / NB) / 1e9
        # one launch at a time, L2 flushed by a 256 MiB fill before each (events around the single launch: includes the launch latency)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        best = 1e9
        for _ in range(20):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            lib.lof_round_params(xs[0].data_ptr(), ys[0].data_ptr(), n1, 0, C.byref(fmt_c), C.byref(ps_c), 1.0, C.byref(rng_c), 0, stream)
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        del flush
        c1 = {"metric": "virtual_round_GBps", "value": gbps * world, "unit": "GB/s", "scaling": "weak", "elements_per_gpu": n1,
              "format": "E4M3 (8, 3, bias 7), saturating, RNE", "us_per_launch": per_launch_us,
              "timing": "16 distinct 64 MiB input / output pairs (2 GiB working set > L2) launched back to back, 5 repetitions, CUDA events",
              "single_launch_l2_flushed": {"us_best_of_20": best * 1e3, "GBps": 8.0 * n1 / (best * 1e-3) / 1e9,
                                           "note": "events around ONE launch after a 256 MiB fill: includes launch latency and the cold ramp"},
              "roofline": {"bound": "hbm", "kernel": "round_vec_kernel<f32,RoundToNearestEven>", "achieved": gbps, "peak": hbm_peak, "unit": "GB/s",
                           "frac": gbps / hbm_peak, "algorithmic_bytes_per_launch": 8.0 * n1, "peak_source": hbm_src,
                           "size_matched_copy_GBps": gbps_cp, "frac_of_size_matched_copy": gbps / gbps_cp,
                           "size_matched_copy_note": "torch Tensor.copy_ over the same 16 buffer pairs, same timing: what a plain copy reaches at 128 MiB per launch"}}
        if rank == 0 and args.e2e_steps > 0:
            xh = torch.empty(n1, dtype=torch.float32).pin_memory()
            xh.copy_(xs[0])
            yh = torch.empty_like(xh).pin_memory()
            call = lambda: lib.lof_round_params_host(xh.data_ptr(), yh.data_ptr(), n1, 0, C.byref(fmt_c), C.byref(ps_c), 1.0, C.byref(rng_c), 0, dev.index)
            assert call() == 0
            t0 = time.perf_counter()
            for _ in range(5):
                assert call() == 0
            dt = (time.perf_counter() - t0) / 5
            c1["e2e"] = {"value": 8.0 * n1 / dt / 1e9, "unit": "GB/s", "h2d_bytes_per_step": 4 * n1, "d2h_bytes_per_step": 4 * n1,
                         "api": "lof_round_params_host (pinned host buffers, copies inside the timed region)"}
            del xh, yh
        out["C1_virtual_round_2^24"] = c1
        del xs, ys
    except Exception as exc:
        out["C1_virtual_round_2^24"] = {"error": str(exc)}
    # ------------------------------------------------------------------ C3: mx_round, 16384 x 16384 fp32 per GPU, block 32, E8M0 scale
    try:
        R = Ccols = 16384
        g = torch.Generator(device=dev).manual_seed(7 + rank)
        x = torch.randn(R, Ccols, device=dev, generator=g)
        x *= torch.exp2(torch.randint(-8, 9, (R, 1), device=dev, generator=g).float())
        x[torch.rand(R, Ccols, device=dev, generator=g) < 1e-3] *= 64.0                      # 0.1 % outliers
        shard = S.RowShard.make(R * world, Ccols, world, rank)                                 # this rank's rows of the (world * 16384) x 16384 tensor
        fmts = (("mxfp8_e4m3", lof.create_p3109_params(8, 4, bias=7)), ("mxfp8_e5m2", lof.create_p3109_params(8, 3, bias=15)),
                ("mxfp6_e3m2", lof.create_p3109_params(6, 3, bias=3)), ("mxfp6_e2m3", lof.create_p3109_params(6, 4, bias=1)),
                ("mxfp4_e2m1", lof.create_p3109_params(4, 2, bias=1)))
        per, tot_ms = {}, 0.0
        for name, d in fmts:
            ms, _ = timed(lambda: S.sharded_virtual_mx_round(x, shard, 32, d, e8m0), 5)
            per[name] = {"ms": ms, "GBps_aggregate": 8.0 * x.numel() * world / (ms * 1e-3) / 1e9}
            tot_ms += ms
        agg = 8.0 * x.numel() * world * len(fmts) / (tot_ms * 1e-3) / 1e9
        blocks = {}
        for bs in (16, 64, 128, 256):
            ms, _ = timed(lambda: S.sharded_virtual_mx_round(x, shard, bs, fmts[0][1], e8m0), 3)
            blocks[f"bs{bs}"] = 8.0 * x.numel() / (ms * 1e-3) / 1e9
        ms_sr, _ = timed(lambda: S.sharded_virtual_mx_round(x, shard, 32, fmts[0][1], e8m0, lof.RoundingMode.StochasticRoundingA, 0, 8, rng=(7, 1)), 3)
        c3 = {"metric": "mx_round_GBps", "value": agg, "unit": "GB/s", "scaling": "weak", "elements_per_gpu": x.numel(), "block_size": 32,
              "scale_format": "E8M0", "formats": per, "collective": None,
              "roofline": {"bound": "hbm", "kernel": "mx_fast_kernel<32,RNE>", "achieved": agg / world, "peak": hbm_peak, "unit": "GB/s",
                           "frac": agg / world / hbm_peak, "algorithmic_bytes_per_launch": 8.0 * x.numel(), "peak_source": hbm_src},
              "per_gpu_GBps_other_block_sizes_e4m3": blocks,
              "per_gpu_GBps_stochastic_SR_A8_e4m3": 8.0 * x.numel() / (ms_sr * 1e-3) / 1e9}
        if rank == 0 and args.e2e_steps > 0:
            # e2e (rank 0's shard): pinned host tensor in, pinned host tensor out, through the C ABI's host entry point
            from lofloat_b200 import _lib
            lib = _lib.lib()
            xh = torch.empty(x.numel(), dtype=torch.float32).pin_memory()
            xh.copy_(x.view(-1))
            yh = torch.empty_like(xh).pin_memory()
            fe, fs, ps = fmts[0][1]._c(), e8m0._c(), _lib.CProjSpec(0, 0, 0)
            call = lambda: lib.lof_mx_round_f32_host(xh.data_ptr(), yh.data_ptr(), xh.numel(), 32, C.byref(fs), C.byref(fe), C.byref(ps),
                                                     C.byref(ps), None, 0, dev.index)
            assert call() == 0
            t0 = time.perf_counter()
            assert call() == 0
            dt = time.perf_counter() - t0
            c3["e2e"] = {"value": 8.0 * xh.numel() / dt / 1e9, "unit": "GB/s", "h2d_bytes_per_step": 4 * xh.numel(), "d2h_bytes_per_step": 4 * xh.numel(),
                         "api": "lof_mx_round_f32_host (rank 0's shard)"}
            del xh, yh
        out["C3_mx_round"] = c3
        del x
    except Exception as exc:
        out["C3_mx_round"] = {"error": str(exc)}

    # ------------------------------------------------------------------ C4: lof-GEMM 8192^3 (strong scaling over the R x C grid)
    grid = S.GemmGrid.create(world, rank)
    try:
        n = 8192
        M = N = K = n
        e4m3 = lof.create_p3109_params(8, 4, bias=7)
        e5m10 = lof.FloatFormatDescriptor(16, 10, 15, lof.InfBehavior.Saturating, lof.NaNBehavior._3109, lof.Signedness.Signed,
                                          lof.create_half_params().is_inf_checker, lof.create_half_params().is_nan_checker)
        from lofloat_b200 import formats as _formats
        ocp = lof.FloatFormatDescriptor(8, 3, 7, lof.InfBehavior.Saturating, lof.NaNBehavior._3109, lof.Signedness.Signed,
                                        _formats._P3109Inf(8, True, True), _formats._BitPatternNaN(0xFF))
        g = torch.Generator(device=dev).manual_seed(42 + rank)
        A_loc = lof.virtual_round(torch.rand(M // world, K, device=dev, generator=g) * 2 - 1, e4m3)      # this rank's rows of A
        Wt_loc = lof.virtual_round(torch.rand(N // world, K, device=dev, generator=g) * 2 - 1, e4m3)     # ... of B^T ([out, in] layout)
        A8, Wt8 = lof.lof_encode(A_loc, ocp).view(M // world, K), lof.lof_encode(Wt_loc, ocp).view(N // world, K)
        fn_nccl = lambda: S.gemm_2d_packed(A8, Wt8, grid, M, N, K, 0, k_chunk=32, accum_format=e5m10)
        ms_nccl, _ = timed(fn_nccl, 10, warm=3)
        ms, exchange = ms_nccl, "nccl"
        if world > 1:
            # the same operand broadcast as PULLS over NVLink peer memory (sharding.PeerPanels: CUDA IPC mappings, cudaMemcpyAsync on side streams,
            # device-side flag hand-shake): no collective, no rendez-vous latency
            try:
                pp4 = S.PeerPanels(grid)
                pp4.register([A8], [Wt8])
                g4 = lambda a_, bt_, out=None: lof.lof_gemm_packed(a_, bt_, 0, k_chunk=32, accum_format=e5m10, out=out)
                same = torch.equal(pp4.gemm(M, N, K, g4), fn_nccl())
                ms_p2p, _ = timed(lambda: pp4.gemm(M, N, K, g4), 10, warm=3)
                ok = torch.tensor([1 if (same and pp4.error_flag() == 0) else 0], device=dev)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN)
                if int(ok.item()) == 1:
                    ms, exchange = ms_p2p, "p2p"
            except Exception as exc:
                exchange = f"nccl (PeerPanels unavailable: {exc!r})"[:200]
        tf = 2.0 * M * N * K / (ms * 1e-3) / 1e12
        a_gather = (M // grid.R) * K * (grid.C - 1) // grid.C if world > 1 else 0                       # bytes RECEIVED per rank
        b_recv = (N // grid.C) * K * (grid.R * grid.C - 1) // (grid.R * grid.C) if world > 1 else 0
        c4 = {"metric": "lof_gemm_TFLOPs", "value": tf, "unit": "TFLOP/s", "scaling": "strong", "M": M, "N": N, "K": K, "ms": ms,
              "operands": "E4M3 bytes, row-sharded (A rows, B^T rows)", "accumulator": "E5M10 re-rounded every k = 32 (RNE), in the tensor core (fp16 TMEM accumulator)",
              "grid": f"{grid.R}x{grid.C}", "timed_region": "operand exchange + GEMM (every call pulls / gathers the panels again)",
              "comm": {"exchange": ("peer-to-peer pulls over NVLink (CUDA IPC mappings, cudaMemcpyAsync on side streams, device-side flag hand-shake; no collective)"
                                    if exchange == "p2p" else exchange) if world > 1 else None,
                       "collectives": ([] if exchange == "p2p" else ["all_gather_into_tensor (A row panel, row group)", "all_to_all_single (B^T column panel, world)"]) if world > 1 else [],
                       "bytes_received_per_rank": int(a_gather + b_recv),
                       "nccl_exchange_ms": ms_nccl if world > 1 else None,
                       "nccl_exchange_TFLOPs": (2.0 * M * N * K / (ms_nccl * 1e-3) / 1e12) if world > 1 else None},
              "roofline": {"bound": "tensor", "achieved": tf / world, "peak": 4500.0, "unit": "TFLOP/s", "frac": tf / world / 4500.0,
                           "peak_source": "nominal dense FP8 (B200_PROFILING.md)", "frac_of_2x_measured_bf16": tf / world / (2.0 * bf16_burst),
                           "measured_bf16_source": bf16_src,
                           "kernel": "gemm_tc_pair_kernel<Hw16,f8> (tcgen05 cta_group::2, 256x256 tiles)"}}
        if rank == 0 and world == 1 and args.e2e_steps > 0:
            # e2e: fp32 containers in pinned host memory -> device -> lof_gemm (conversion + GEMM) -> C back to pinned host memory
            Ah, Bh = A_loc.cpu().pin_memory(), Wt_loc.t().contiguous().cpu().pin_memory()
            Ch = torch.empty(M, N, dtype=torch.float32).pin_memory()

            def e2e_call():
                Ad, Bd = Ah.to(dev, non_blocking=True), Bh.to(dev, non_blocking=True)
                Ch.copy_(lof.lof_gemm(Ad, Bd, 0, k_chunk=32, accum_format=e5m10, check=False), non_blocking=True)
                torch.cuda.synchronize()
            e2e_call()
            t0 = time.perf_counter()
            e2e_call()
            dt = time.perf_counter() - t0
            c4["e2e"] = {"value": 2.0 * M * N * K / dt / 1e12, "unit": "TFLOP/s", "h2d_bytes_per_step": 8 * M * K, "d2h_bytes_per_step": 4 * M * N,
                         "api": "lof_gemm(k_chunk=32, accum_format=E5M10) on fp32 host tensors (pinned), copies inside the timed region"}
            del Ah, Bh, Ch
        out["C4_lof_gemm_8192"] = c4
        del A_loc, Wt_loc, A8, Wt8
    except Exception as exc:
        out["C4_lof_gemm_8192"] = {"error": str(exc)}

    # ------------------------------------------------------------------ C5: MXFP8 GEMM 32768^3, bf16-rounded accumulation
    try:
        n = args.c5_size
        M = N = K = n
        e4m3 = lof.create_p3109_params(8, 4, bias=7)
        g = torch.Generator(device=dev).manual_seed(99 + rank)
        A_loc = torch.randn(M // world, K, device=dev, generator=g)
        Wt_loc = torch.randn(N // world, K, device=dev, generator=g)           # weight layout [out, in]: MX blocks run along K
        # the two MX passes: virtual_mx_round (reference semantics, E4M3 elements / E8M0 scales, block 32 along K) + the exact re-split into
        # the tensor core's MX storage (e4m3fn bytes + E8M0 scale bytes); timed together, warm allocator
        quant = lambda: (lof.mx_quantize_operand(A_loc, e4m3, e8m0), lof.mx_quantize_operand(Wt_loc, e4m3, e8m0))
        ms_mx, ((ae, asc, _sa), (be, bsc, _sb)) = timed(quant, 3, warm=2)
        ms_round, _ = timed(lambda: (lof.virtual_mx_round(A_loc, 32, e4m3, e8m0), lof.virtual_mx_round(Wt_loc, 32, e4m3, e8m0)), 3, warm=2)
        del A_loc, Wt_loc
        torch.cuda.empty_cache()
        fn = lambda: S.gemm_2d_mx(ae, asc, be, bsc, grid, M, N, K, 7, k_chunk=32, partial_round=True)
        ms_nccl5, Cblk = timed(fn, 2, warm=1)
        finite = bool(torch.isfinite(Cblk).all().item())
        ms, exchange5 = ms_nccl5, "nccl"
        if world > 1:
            try:
                pp5 = S.PeerPanels(grid)
                pp5.register([ae, asc], [be, bsc])
                g5 = lambda a_e, a_s, b_e, b_s, out=None: lof.mx_gemm(a_e, a_s, b_e, b_s, 7, k_chunk=32, partial_round=True, out=out)
                same = torch.equal(pp5.gemm(M, N, K, g5), Cblk)
                ms_p2p5, _ = timed(lambda: pp5.gemm(M, N, K, g5), 2, warm=1)
                ok = torch.tensor([1 if (same and pp5.error_flag() == 0) else 0], device=dev)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN)
                if int(ok.item()) == 1:
                    ms, exchange5 = ms_p2p5, "p2p"
                    fn = lambda: pp5.gemm(M, N, K, g5)
            except Exception as exc:
                exchange5 = f"nccl (PeerPanels unavailable: {exc!r})"[:200]
        del Cblk
        # the same product with fp32 accumulation over all of K and ONE final bf16-like rounding (plain MXFP8 GEMM): how much the per-block
        # re-rounding costs
        if world > 1 and exchange5 == "p2p":
            g5c = lambda a_e, a_s, b_e, b_s, out=None: lof.mx_gemm(a_e, a_s, b_e, b_s, 7, k_chunk=0, out=out)
            ms_1c, Cblk = timed(lambda: pp5.gemm(M, N, K, g5c), 2, warm=1)
        else:
            ms_1c, Cblk = timed(lambda: S.gemm_2d_mx(ae, asc, be, bsc, grid, M, N, K, 7, k_chunk=0), 2, warm=1)
        del Cblk
        tf = 2.0 * M * N * K / (ms * 1e-3) / 1e12
        per_elem = 1.0 + 1.0 / 32.0
        a_gather = per_elem * (M // grid.R) * K * (grid.C - 1) // grid.C if world > 1 else 0
        b_recv = per_elem * (N // grid.C) * K * (world - 1) // world if world > 1 else 0
        out["C5_mxfp8_gemm"] = {
            "metric": "mx_gemm_TFLOPs", "value": tf, "unit": "TFLOP/s", "scaling": "strong", "M": M, "N": N, "K": K, "ms": ms,
            "composition": "LoF_Linear(scaling='mx') (layers.py:366-377): mx_round(E4M3 / E8M0, block 32 along K) of both operands, then the GEMM "
                           "with the accumulator rounded to 7 mantissa bits (bf16-like) every k = 32 (= one MX block)",
            "operands": "MX storage: e4m3fn bytes + E8M0 scale bytes, row-sharded; consumed by tcgen05.mma.kind::mxf8f6f4.block_scale", "grid": f"{grid.R}x{grid.C}",
            "timed_region": "operand exchange + GEMM; the two MX passes are timed separately",
            "mx_passes": {"ms": ms_mx, "GBps_aggregate": (8.0 + 4.0 + per_elem) * (M * K + N * K) / (ms_mx * 1e-3) / 1e9,
                          "bytes_per_element": "8 (virtual_mx_round: 4 read + 4 written) + 4 + 1.03 (re-split into bytes + scales)",
                          "virtual_mx_round_only_ms": ms_round, "virtual_mx_round_only_GBps_aggregate": 8.0 * (M * K + N * K) / (ms_round * 1e-3) / 1e9},
            "fp32_accumulation_one_final_rounding": {"ms": ms_1c, "TFLOPs": 2.0 * M * N * K / (ms_1c * 1e-3) / 1e12},
            "comm": {"exchange": ("peer-to-peer pulls over NVLink (CUDA IPC mappings, cudaMemcpyAsync on side streams, device-side flag hand-shake; no collective)"
                                  if exchange5 == "p2p" else exchange5) if world > 1 else None,
                     "collectives": ([] if exchange5 == "p2p" else ["all_gather_into_tensor x2 (A row panel: element bytes, scale bytes; row group)",
                                     "all_to_all_single x2 (B^T column panel: element bytes, scale bytes; world)"]) if world > 1 else [],
                     "bytes_received_per_rank": int(a_gather + b_recv),
                     "nccl_exchange_ms": ms_nccl5 if world > 1 else None,
                     "nccl_exchange_TFLOPs": (2.0 * M * N * K / (ms_nccl5 * 1e-3) / 1e12) if world > 1 else None},
            "finite": finite,
            "roofline": {"bound": "tensor", "achieved": tf / world, "peak": 4500.0, "unit": "TFLOP/s", "frac": tf / world / 4500.0,
                         "peak_source": "nominal dense FP8 (B200_PROFILING.md)", "frac_of_2x_measured_bf16": tf / world / (2.0 * bf16_burst),
                         "measured_bf16_source": bf16_src,
                         "kernel": "gemm_mx_kernel<RNE,Bf162,partial> (tcgen05 block-scaled MMA + per-block bfloat162 fold from TMEM)",
                         "note": "the per-block re-rounding makes this kernel bound by the MMA -> TMEM read-back -> fold hand-off, not by the tensor pipe "
                                 "(profiles/r02j_probe_*.txt); the fp32-accumulation figure next to it is bound by L2 -> SM operand bandwidth"}}
        del ae, asc, be, bsc
    except Exception as exc:
        out["C5_mxfp8_gemm"] = {"error": str(exc)}
    torch.cuda.empty_cache()
    return out


def side_measurements(torch, lib, _lib, dev, x, y, stream):
    """mx_round (C3 shape) and lof_gemm, outside the timed region: reported next to the headline."""
    out = {}
    peak, _ = peaks()
    try:
        e8 = _lib.CFormat(8, 0, 127, 0, 0, 1, 0)
        ps = _lib.CProjSpec(0, 0, 0)
        n = 16384 * 16384
        for variant in (0, 1, 2):                                   # 0 fast kernel (default), 1 TMA-staged, 2 first-generation shuffle
            _lib.set_option("mx.variant", variant)
            for fname, (bw, m, bias, ov, sg) in (("mxfp8_e4m3", (8, 3, 7, 1, 0)), ("mxfp6_e3m2", (6, 2, 3, 1, 0)),
                                                 ("mxfp4_e2m1", (4, 1, 1, 1, 0))):
                fmt = _lib.CFormat(bw, m, bias, ov, 0, sg, 0)
                call = lambda: lib.lof_mx_round_f32(x.data_ptr(), y.data_ptr(), n, 32, C.byref(e8), C.byref(fmt),
                                                    C.byref(ps), C.byref(ps), None, 0, stream)
                for _ in range(3):
                    assert call() == 0
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(10):
                    call()
                b.record()
                torch.cuda.synchronize()
                gbps = 8.0 * n * 10 / (a.elapsed_time(b) * 1e-3) / 1e9
                out[f"mx_round_{fname}_bs32_16384x16384_variant{variant}"] = {"GBps": gbps, "frac_of_hbm_peak": gbps / peak}
        _lib.set_option("mx.variant", 0)
    except Exception as exc:
        out["mx_round_error"] = str(exc)
    try:
        # binding overhead on a small tensor (4096 elements: the kernel is ~3 us): calls per second through the ctypes mirror (lofloat_b200.ops)
        # and through the compiled extension (lofloat_b200/LoFloat.so, the reference's module surface), launches queued back to back
        import importlib.machinery
        import importlib.util
        import lofloat_b200 as lof
        xs = x[:4096].contiguous()
        e4 = lof.create_p3109_params(8, 4, bias=7)
        reps = 2000

        def per_call_us(fn):
            for _ in range(50):
                fn()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            torch.cuda.synchronize()
            return (time.perf_counter() - t0) / reps * 1e6
        entry = {"elements": 4096, "calls": reps, "ctypes_ops_us_per_call": per_call_us(lambda: lof.virtual_round(xs, e4)),
                 "torch_copy_us_per_call": per_call_us(lambda: xs.clone())}
        ext_path = ROOT / "lofloat_b200" / "LoFloat.so"
        if ext_path.exists():
            loader = importlib.machinery.ExtensionFileLoader("LoFloat", str(ext_path))
            ext = importlib.util.module_from_spec(importlib.util.spec_from_loader("LoFloat", loader))
            loader.exec_module(ext)
            fd = ext.FloatFormatDescriptor(8, 3, 7, ext.InfBehavior.Saturating, ext.NaNBehavior._3109, ext.Signedness.Signed, e4.is_inf_checker, e4.is_nan_checker)
            entry["compiled_extension_us_per_call"] = per_call_us(lambda: ext.virtual_round(xs, fd))
            entry["results_identical"] = bool(torch.equal(ext.virtual_round(xs, fd), lof.virtual_round(xs, e4)))
        out["binding_overhead_virtual_round_4096"] = entry
    except Exception as exc:
        out["binding_overhead_error"] = str(exc)
    try:
        # same-box GPU baseline: the reference's OWN device code (lo_float.h virtual_round + mx_round.cuh, compiled unmodified for sm_100a into
        # oracle/_ref/libref_gpu.so by `make -C oracle refgpu`; test / bench infrastructure) in the reference's launch shapes, next to ours
        so = ROOT / "oracle" / "_ref" / "libref_gpu.so"
        if so.exists():
            R = C.CDLL(str(so))
            R.refgpu_round_params_f32.argtypes = [C.c_void_p, C.c_int64] + [C.c_int] * 8 + [C.c_float, C.c_void_p]
            R.refgpu_mx_round_f32.argtypes = [C.c_void_p, C.c_int64, C.c_int] + [C.c_int] * 12 + [C.c_void_p]
            import lofloat_b200 as lof
            n = x.numel()

            def time_it(fn, reps=5):
                fn()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(reps):
                    fn()
                b.record()
                torch.cuda.synchronize()
                return a.elapsed_time(b) / reps
            for label, (bw, m, bias, ov), mode, sl in (("e4m3_RNE", (8, 3, 7, 1), 0, 0), ("e4m3_RTZ", (8, 3, 7, 1), 1, 0), ("e4m3_RoundToOdd", (8, 3, 7, 1), 12, 0),
                                                       ("e4m3_SR_A8", (8, 3, 7, 1), 3, 8), ("bf16like_RNE", (16, 7, 127, 0), 0, 0)):
                y.copy_(x)
                assert R.refgpu_round_params_f32(y.data_ptr(), n, bw, m, bias, ov, 0, 0, mode, sl, 1.0, stream) == 0
                entry = {}
                if sl == 0:         # deterministic: the reference's device result on this box == ours, bit for bit
                    d = lof.FloatFormatDescriptor(bw, m, bias, ov, 0, 0, lof.create_single_params().is_inf_checker, lof.create_single_params().is_nan_checker)
                    ours = lof.virtual_round(x, d, mode)
                    ne = y.view(torch.int32) != ours.view(torch.int32)
                    entry["bit_identical_to_lofloat_b200"] = not bool(ne.any().item())
                    if not entry["bit_identical_to_lofloat_b200"]:      # the reference's DEVICE helpers vs its own HOST path (ours is pinned to the host path)
                        i = int(ne.view(-1).nonzero()[0])
                        entry["differing_elements"] = int(ne.sum().item())
                        entry["first_difference"] = {"x": float(x.view(-1)[i]), "reference_gpu": float(y.view(-1)[i]), "lofloat_b200_and_reference_cpu": float(ours.view(-1)[i])}
                    del ours, ne
                ms = time_it(lambda: R.refgpu_round_params_f32(y.data_ptr(), n, bw, m, bias, ov, 0, 0, mode, sl, 1.0, stream))
                entry.update({"GBps": 8.0 * n / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": 8.0 * n / (ms * 1e-3) / 1e9 / peak,
                              "launch": "round_fp_params<float><<<n / 256, 256>>> in place (Lof_kernel.cu:49-60), 2^28 elements"})
                out[f"reference_gpu_virtual_round_{label}"] = entry
            nmx = 16384 * 16384
            for bs in (32, 128):
                y.copy_(x)
                call = lambda: R.refgpu_mx_round_f32(y.data_ptr(), nmx, bs, 8, 0, 127, 0, 1, 8, 3, 7, 1, 0, 0, 0, stream)
                assert call() == 0
                e8p, e4p = lof.create_e8m0_params(), lof.create_p3109_params(8, 4, bias=7)
                ours = lof.virtual_mx_round(x, bs, e4p, e8p)
                ne = y.view(torch.int32) != ours.view(torch.int32)
                same = not bool(ne.any().item())
                extra = {}
                if not same:
                    i = int(ne.view(-1).nonzero()[0])
                    extra = {"differing_elements": int(ne.sum().item()),
                             "first_difference": {"x": float(x.view(-1)[i]), "reference_gpu": float(y.view(-1)[i]), "lofloat_b200_and_reference_cpu": float(ours.view(-1)[i])}}
                del ours, ne
                ms = time_it(call)
                out[f"reference_gpu_mx_round_mxfp8_e4m3_bs{bs}"] = {
                    **extra, "GBps": 8.0 * nmx / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": 8.0 * nmx / (ms * 1e-3) / 1e9 / peak, "bit_identical_to_lofloat_b200": same,
                    "launch": "virtual_mx_round<float, 256, MX, NX><<<min(65535, n / bs), 256>>> in place (mx_round.cuh:27-214), 16384 x 16384"}
        else:
            out["reference_gpu"] = "oracle/_ref/libref_gpu.so not built"
    except Exception as exc:
        out["reference_gpu_error"] = str(exc)
    try:
        # C2's bf16-in / bf16-out variant: 2^29 bf16 containers (1 GiB), 4 B / element of algorithmic traffic
        import lofloat_b200 as lof
        nb = 1 << 29
        xb = x.to(torch.bfloat16).repeat(2)[:nb].contiguous()
        for label, d, mode, sl in (("e4m3_RNE", lof.create_p3109_params(8, 4, bias=7), 0, 0), ("e2m1_RNE", lof.create_p3109_params(4, 2, bias=1), 0, 0),
                                   ("e4m3_SR_A4", lof.create_p3109_params(8, 4, bias=7), 3, 4)):
            for _ in range(2):
                lof.virtual_round(xb, d, mode, sl)
            torch.cuda.synchronize()
            # one event pair per launch, median: a host-side stall of the caching allocator (1 GiB result per call, late in a long
            # bench process) between two launches would otherwise be counted as kernel time (r02x recorded 3.05 TB/s for e4m3 RNE that way)
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(7)]
            for a, b in evs:
                a.record()
                lof.virtual_round(xb, d, mode, sl)
                b.record()
            torch.cuda.synchronize()
            ms = sorted(a.elapsed_time(b) for a, b in evs)[len(evs) // 2]
            gbps = 4.0 * nb / (ms * 1e-3) / 1e9
            out[f"C2_virtual_round_bf16_container_{label}_2^29"] = {"GBps": gbps, "frac_of_hbm_peak": gbps / peak, "timing": "median of 7 launches, one event pair each"}
        del xb
    except Exception as exc:
        out["bf16_container_error"] = str(exc)
    try:
        M = N = K = 2048
        A = torch.rand(M, K, device=dev) * 2 - 1
        B = torch.rand(K, N, device=dev) * 2 - 1
        Cm = torch.empty(M, N, device=dev)
        ps = _lib.CProjSpec(0, 0, 0)
        call = lambda: lib.lof_gemm_f32(A.data_ptr(), B.data_ptr(), Cm.data_ptr(), M, N, K, 10, C.byref(ps), 256, 1.0, 0.0,
                                        None, stream)
        assert call() == 0
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            call()
        b.record()
        torch.cuda.synchronize()
        out["lof_gemm_exact_2048^3_accum10_RNE_kc256"] = {"TFLOPs": 2.0 * M * N * K * 3 / (a.elapsed_time(b) * 1e-3) / 1e12,
                                                        "note": "per-FMA rounding, SIMT (bit-exact vs BLIS reference); bound of the chain: "
                                                                "12 clk per warp-MAC and SMSP = 6.06 TFLOP/s at 1.92 GHz"}
        # row f-3: Conv2d as an implicit GEMM on the same kernel (ResNet-style 3x3 layer) vs the reference's composition
        import lofloat_b200 as lof
        import torch.nn.functional as F
        Bn, Ci, Hh, Co = 32, 64, 56, 64
        e4m3 = lof.create_p3109_params(8, 4, bias=7)
        cx = lof.virtual_round(torch.randn(Bn, Ci, Hh, Hh, device=dev), e4m3)
        cw = lof.virtual_round(torch.randn(Co, Ci, 3, 3, device=dev) * 0.1, e4m3)
        cbias = torch.randn(Co, device=dev)

        def composed():                                  # LoF_Conv2d im2col backend, LoFloat/layers.py:710-751
            cols = F.unfold(cx, (3, 3), padding=1)
            wc = cw.view(Co, -1)
            o = torch.stack([lof.lof_gemm(wc, cols[i].contiguous(), 10, kc=0) for i in range(Bn)]).view(Bn, Co, Hh, Hh)
            return o + cbias.view(1, -1, 1, 1)

        fused = lambda: lof.lof_conv2d(cx, cw, 1, 1, 1, 1, 1, 1, 10, k_order="crs", bias=cbias)
        same = bool(torch.equal(composed(), fused()))
        for name, fn in (("unfold_plus_per_image_lof_gemm", composed), ("lof_conv2d_implicit_gemm", fused)):
            fn()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(3):
                fn()
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 3
            out[f"conv3x3_32x64x56x56_to_64_accum10_{name}"] = {"ms": ms, "TFLOPs": 2.0 * Bn * Co * Hh * Hh * Ci * 9 / ms / 1e9,
                                                                "bit_identical": same}
    except Exception as exc:
        out["lof_gemm_error"] = str(exc)
    try:
        # C4: tcgen05 chunked GEMM, 8192^3, E4M3 operands (U(-1,1) rounded to E4M3), accumulator re-rounded every k = 32
        import lofloat_b200 as lof
        n = 8192
        g = torch.Generator(device=dev).manual_seed(42)
        e4m3 = lof.create_p3109_params(8, 4, bias=7)
        A = lof.virtual_round(torch.rand(n, n, device=dev, generator=g) * 2 - 1, e4m3)
        B = lof.virtual_round(torch.rand(n, n, device=dev, generator=g) * 2 - 1, e4m3)
        e5m10 = lof.FloatFormatDescriptor(16, 10, 15, lof.InfBehavior.Saturating, lof.NaNBehavior._3109, lof.Signedness.Signed,
                                          lof.create_half_params().is_inf_checker, lof.create_half_params().is_nan_checker)
        cases = (("k32_E5M10_RNE_singleround_hw_fp16_accumulator", dict(k_chunk=32, accum_format=e5m10)),
                 ("k32_E5M10_RNE_partialround_packed", dict(k_chunk=32, accum_format=e5m10, partial_round=True)),
                 ("k32_mant10_RNE_singleround_generic", dict(k_chunk=32, allow_hw_accum=False)),
                 ("k128_mant10_RNE_singleround_generic", dict(k_chunk=128)),
                 ("k512_mant10_RNE_singleround_generic", dict(k_chunk=512)))
        for label, kw in cases:
            lof.lof_gemm(A, B, 10, check=False, **kw)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(3):
                lof.lof_gemm(A, B, 10, check=False, **kw)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 3
            out[f"lof_gemm_tcgen05_8192^3_e4m3_{label}"] = {
                "ms": ms, "TFLOPs": 2.0 * n ** 3 / (ms * 1e-3) / 1e12,
                "note": "includes the fp32 -> e4m3 operand conversion kernels"}
        # C4 on operands that are already stored as e4m3 bytes (lof_encode_f32 output; B^T = [N, K] weight layout):
        # the GEMM kernel alone, no conversion pass.  Tensor roofline: FP8 dense nominal 4500 TFLOP/s (B200_PROFILING.md)
        # and 2 x the measured dense bf16 rate of MEASURED_PEAKS.json.
        from lofloat_b200 import formats as _formats
        ocp = lof.FloatFormatDescriptor(8, 3, 7, lof.InfBehavior.Saturating, lof.NaNBehavior._3109, lof.Signedness.Signed,
                                        _formats._P3109Inf(8, True, True), _formats._BitPatternNaN(0xFF))
        Ap = lof.lof_encode(A, ocp).view(n, n)
        Btp = lof.lof_encode(B.t().contiguous(), ocp).view(n, n)
        lof.lof_gemm_packed(Ap, Btp, 0, k_chunk=32, accum_format=e5m10)
        reps = 10
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            lof.lof_gemm_packed(Ap, Btp, 0, k_chunk=32, accum_format=e5m10)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        tf = 2.0 * n ** 3 / (ms * 1e-3) / 1e12
        bf16_meas = None
        try:
            bf16_meas = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["bf16_tflops"])
        except Exception:
            pass
        out["C4_lof_gemm_packed_e4m3_8192^3_k32_E5M10"] = {
            "ms": ms, "TFLOPs": tf,
            "roofline": {"bound": "tensor", "achieved": tf, "peak": 4500.0, "unit": "TFLOP/s", "frac": tf / 4500.0,
                         "peak_source": "nominal dense FP8 (B200_PROFILING.md)",
                         "frac_of_2x_measured_bf16": (tf / (2.0 * bf16_meas)) if bf16_meas else None,
                         "kernel": "gemm_tc_pair_kernel<Hw16,f8> (tcgen05 cta_group::2, 256x256 tiles, fp16 accumulation in TMEM "
                                   "re-rounded every k = 32)"},
            "note": "back-to-back launches incl. the status / fallback-gate launches; operands resident as e4m3 bytes"}
        del A, B, Ap, Btp
        # storage codec (row f-1): 2^28 fp32 -> packed n-bit and back; algorithmic bytes 4 + n/8 per element
        for fname, d in (("e4m3_8bit", lof.create_p3109_params(8, 4, bias=7)), ("e3m2_6bit", lof.create_p3109_params(6, 3, bias=3)),
                         ("e2m1_4bit", lof.create_p3109_params(4, 2, bias=1))):
            nel = x.numel()
            packed = lof.lof_encode(x, d)
            back = lof.lof_decode(packed, nel, d)
            for what, fn in (("encode", lambda: lof.lof_encode(x, d)), ("decode", lambda: lof.lof_decode(packed, nel, d))):
                fn()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(5):
                    fn()
                b.record()
                torch.cuda.synchronize()
                bytes_ = (4.0 + d.total_bits / 8.0) * nel
                gbps = bytes_ * 5 / (a.elapsed_time(b) * 1e-3) / 1e9
                out[f"pack_{what}_{fname}_2^28"] = {"GBps": gbps, "frac_of_hbm_peak": gbps / peak,
                                                    "algorithmic_bytes_per_element": 4.0 + d.total_bits / 8.0}
            del packed, back
        # strict typed-semantics rounding (row f-2), fused Hadamard (+ MX) (row f-3), MX storage (row f-1): 2^28 elements
        from lofloat_b200 import formats as _f
        d8 = lof.FloatFormatDescriptor(8, 0, 127, lof.InfBehavior.Saturating, lof.NaNBehavior._3109, lof.Signedness.Unsigned,
                                       _f._P3109Inf(8, False, True), _f._BitPatternNaN(0xFF))
        e4 = lof.create_p3109_params(8, 4, bias=7)
        x2 = x.view(16384, 16384)
        sc, el = lof.mx_encode(x2, 32, e4, d8)
        extra = (("strict_round_e4m3_RNE", lambda: lof.virtual_round_strict(x, e4), 8.0),
                 ("fwht128", lambda: lof.fwht(x2, 128), 8.0),
                 ("fwht128_fused_virtual_round_e4m3", lambda: lof.fwht(x2, 128, params=e4), 8.0),
                 ("fwht128_fused_mx_round_mxfp8", lambda: lof.fwht(x2, 128, params=e4, mx_block_size=32, scale_format=lof.create_e8m0_params()), 8.0),
                 ("mx_encode_mxfp8_scale_bytes_plus_packed", lambda: lof.mx_encode(x2, 32, e4, d8), 4.0 + 1.0 + 1.0 / 32),
                 ("mx_decode_mxfp8", lambda: lof.mx_decode(sc, el, x.numel(), 32, e4, d8), 4.0 + 1.0 + 1.0 / 32))
        for label, fn, bpe in extra:
            fn()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                fn()
            b.record()
            torch.cuda.synchronize()
            gbps = bpe * x.numel() * 5 / (a.elapsed_time(b) * 1e-3) / 1e9
            out[f"{label}_2^28"] = {"GBps": gbps, "frac_of_hbm_peak": gbps / peak, "algorithmic_bytes_per_element": bpe}
        del sc, el
    except Exception as exc:
        out["lof_gemm_tc_error"] = str(exc)
    return out


def c1_and_cpu_side(torch, dev):
    """C1 (BASELINE.json configs[0]: fp32 -> E4M3 RNE saturating on 2^24 elements, the reference's own CPU-runnable
    case, benchmark/vec_round.cpp) on the GPU next to the reference's xsimd array loop and scalar loop on this box's
    cores, plus bounded CPU samples of C3 (mx_round) and C4 (BLIS lof GEMM) so every GPU number above has its CPU
    counterpart measured in the same run.  Reported baselines, not targets."""
    import numpy as np
    import lofloat_b200 as lof
    L, kind = _load_cpu_impl()
    threads = os.cpu_count() or 1
    out = {}
    n = 1 << 24
    g = torch.Generator(device=dev).manual_seed(12345)
    x = torch.randn(n, device=dev, generator=g) * 4
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)               # > L2 (126 MB): C1's 64 MiB input would otherwise stay cached
    xn = x.cpu().numpy()
    for label, d, f in (("e4m3_bias7", lof.create_p3109_params(8, 4, bias=7), L.FORMATS["e4m3_b7"]),
                        ("p3109_8p4_bias8", lof.create_p3109_params(8, 4), L.FORMATS["p3109_8p4"])):
        lof.virtual_round(x, d)
        best = 1e9
        for _ in range(20):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            lof.virtual_round(x, d)
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        entry = {"gpu_GBps": 8.0 * n / (best * 1e-3) / 1e9, "gpu_ms_best_of_20": best, "l2": "flushed between launches (256 MiB fill)"}
        if kind == "reference":
            lib = L.ref()
            buf_in, buf_out = xn.copy(), np.empty_like(xn)
            bt = 1e9
            for _ in range(7):
                t0 = time.perf_counter()
                lib.ref_round_params_array_f32(L._ptr(buf_in), L._ptr(buf_out), n, f.bitwidth, f.mantissa_bits, f.bias, f.ov_behavior,
                                               f.na_behavior, f.is_signed, 0, 0)
                bt = min(bt, time.perf_counter() - t0)
            entry["cpu_xsimd_array_GBps"] = 8.0 * n / bt / 1e9
            entry["cpu_xsimd_note"] = (f"lo_float.h:2942 xsimd loop, -O3 -march=x86-64-v3 -fopenmp -D_LOFOPENMP, {threads} threads, best of 7 "
                                       "(not an oracle: drops signs / flushes subnormals, SURVEY B.10)")
            bt = 1e9
            for _ in range(3):
                t0 = time.perf_counter()
                L.r_round_params(xn, f, 0, 0, threads=threads)
                bt = min(bt, time.perf_counter() - t0)
            entry["cpu_scalar_loop_GBps"] = 8.0 * n / bt / 1e9
        out[f"C1_virtual_round_2^24_{label}"] = entry
    del flush
    # C3 on the CPU: the reference's virtual_mx_round is a serial loop (lo_float.h:4311-4383)
    m = 1 << 22
    xm = (np.random.default_rng(7).standard_normal(m)).astype(np.float32)
    t0 = time.perf_counter()
    if kind == "reference":
        L.r_mx_round(xm, 32, L.FORMATS["e8m0"], L.FORMATS["e4m3_b7"])
    else:
        L.o_mx_round(xm, 32, L.FORMATS["e8m0"], L.FORMATS["e4m3_b7"])
    dt = time.perf_counter() - t0
    out["C3_cpu_mx_round_mxfp8"] = {"GBps": 8.0 * m / dt / 1e9, "cores": 1 if kind == "reference" else threads, "kind": kind,
                                    "sample": f"2^22 elements, block 32, {dt:.2f} s"}
    # C4 on the CPU: BLIS generic + lof_sgemm_ukr (per-FMA rounding to 10 mantissa bits), bounded size
    k = 1536
    rng = np.random.default_rng(42)
    A = L.o_round_params(rng.uniform(-1, 1, (k * k)).astype(np.float32), L.FORMATS["e4m3_b7"]).reshape(k, k)
    B = L.o_round_params(rng.uniform(-1, 1, (k * k)).astype(np.float32), L.FORMATS["e4m3_b7"]).reshape(k, k)
    t0 = time.perf_counter()
    if kind == "reference" and L.ref().ref_has_blis():
        L.r_gemm(A, B, 10, threads=threads)
        gk = "reference"
    else:
        L.o_gemm(A, B, 10)
        gk = "port"
    dt = time.perf_counter() - t0
    out["C4_cpu_lof_gemm_accum10"] = {"GFLOPs": 2.0 * k ** 3 / dt / 1e9, "cores": threads, "kind": gk,
                                      "sample": f"{k}^3, bli_sgemm + lof_sgemm_ukr (BLIS generic, KC = 256), {dt:.1f} s"}
    return out


# -------------------------------------------------------------------------------- CPU reference
def _load_cpu_impl():
    """The reference's own CPU path (oracle/_ref, compiled from /root/reference) if the prebuilt shim
    loads and runs on this host, else the oracle port.  Probed in a subprocess so an illegal
    instruction on a different CPU cannot take the benchmark down."""
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle_lib as L
    kind = "port"
    if L.REF_SO.exists():
        probe = ("import sys; sys.path.insert(0, %r); import numpy as np, oracle_lib as L; "
                 "x = np.linspace(-3, 3, 4096, dtype=np.float32); "
                 "a = L.r_round_params(x, L.FORMATS['e4m3_b7']); b = L.o_round_params(x, L.FORMATS['e4m3_b7']); "
                 "assert L.bits_equal(a, b)") % str(ROOT / "tests")
        try:
            if subprocess.run([sys.executable, "-c", probe], timeout=120, capture_output=True).returncode == 0:
                kind = "reference"
        except Exception:
            pass
    L.build_oracle()
    return L, kind


def _cpu_sweep_rates(L, kind, x_np, det_n, sto_n, threads):
    """GB/s of the reference's scalar virtual_round for each of the 20 (format, mode) passes of the C2 sweep, on samples of
    the workload's data: deterministic modes on the first det_n elements, the stochastic ones on the first sto_n (the
    reference constructs and seeds a std::mt19937 per element there, lo_float.h:606, which is ~100x slower)."""
    rates = {}
    for fname, (bw, m, bias, ov, sg) in FORMATS:
        f = L.fmt(bw, m, bias, ov, sg)
        for mname, mode, sl in MODES:
            xs = x_np[:sto_n] if mname.startswith("SR") else x_np[:det_n]
            t0 = time.perf_counter()
            if kind == "reference":
                L.r_round_params(xs, f, mode, sl, threads=threads)
            else:
                L.o_round_params(xs, f, mode, sl, seed=1)
            rates[f"{fname}/{mname}"] = 8.0 * xs.size / (time.perf_counter() - t0) / 1e9
    return rates


def _sweep_rate(rates):
    """Rate of the whole sweep when every pass moves the SAME number of bytes, as it does in the workload: the harmonic mean
    of the per-pass rates (a bytes-weighted mean over unequal samples would hide the slow stochastic passes)."""
    return len(rates) / sum(1.0 / r for r in rates.values())


def cpu_reference_sample(x_np, budget_s: float):
    """Time the reference's scalar virtual_round (lo_float.h:2843-2939 looped by oracle/ref_shim.cpp, OpenMP over all host
    threads) on a bounded sample of the same workload; every (format, mode) pass weighs the same, as in the workload."""
    import numpy as np
    L, kind = _load_cpu_impl()
    threads = os.cpu_count() or 1
    det_n, sto_n = x_np.size, min(x_np.size, 1 << 20)
    t0 = time.perf_counter()
    rates = _cpu_sweep_rates(L, kind, x_np, det_n, sto_n, threads)
    total_s = time.perf_counter() - t0
    det = [v for k, v in rates.items() if "SR" not in k]
    return {"value": _sweep_rate(rates), "unit": "GB/s", "cores": threads, "kind": kind,
            "deterministic_passes_GBps": len(det) / sum(1.0 / r for r in det),
            "sample": f"the 20 passes of the sweep, each timed once: deterministic modes on the first 2^{int(np.log2(det_n))} elements, SR-A8 on "
                      f"2^{int(np.log2(sto_n))}; value = harmonic mean of the per-pass rates (equal bytes per pass, as in the workload); scalar "
                      f"virtual_round loop, OpenMP x{threads}, {total_s:.1f} s of CPU work"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation on this box's host cores, same metric/config.  One step = the
    20-pass sweep on a 2^26-element sample of the workload's data (SR-A8 passes on 2^20 elements); the step's rate is the
    harmonic mean of the per-pass rates, i.e. the rate of the real sweep in which every pass moves the same bytes."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    import numpy as np
    rng = np.random.default_rng(2024)
    n = 1 << 26
    x = (rng.standard_normal(n, dtype=np.float32) * np.exp2(rng.integers(-12, 13, n)).astype(np.float32)).astype(np.float32)
    L, kind = _load_cpu_impl()
    threads = os.cpu_count() or 1
    sto_n = 1 << 20
    for _ in range(min(args.warmup, 1)):
        _cpu_sweep_rates(L, kind, x, 1 << 22, 1 << 16, threads)
    t0 = time.perf_counter()
    steps = []
    for _ in range(args.steps):
        steps.append(_cpu_sweep_rates(L, kind, x, n, sto_n, threads))
        if time.perf_counter() - t0 > 240.0:          # bounded: a slow host stops after ~4 minutes
            break
    dt = time.perf_counter() - t0
    mean_rates = {k: len(steps) / sum(1.0 / st[k] for st in steps) for k in steps[0]}
    value = _sweep_rate(mean_rates)
    sample = (f"each step = the 20-pass sweep on a 2^26-element sample (SR-A8 passes on 2^20: the reference seeds a std::mt19937 per "
              f"element), value = harmonic mean of the per-pass rates (every pass weighs the same, as in the workload), scalar virtual_round "
              f"loop, OpenMP x{threads}")
    print(json.dumps({
        "impl": "reference", "metric": "virtual_round_GBps", "value": value, "unit": "GB/s", "n_gpus": world,
        "steps": len(steps), "warmup": args.warmup, "ms_per_step": dt / len(steps) * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32 bit arithmetic on fp32 containers", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "per_pass_GBps": mean_rates,
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--c5-size", type=int, default=32768, help="M = N = K of the C5 MXFP8 GEMM (BASELINE.json configs[4]: 32768)")
    ap.add_argument("--no-side", action="store_true", help="skip the mx/gemm side measurements and the CPU sample (profiling runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
