#!/usr/bin/env python
"""Summarise `ncu --page raw --csv` exports (one row per profiled kernel) into the handful of metrics the
roofline discussion needs.  Usage: python scripts/ncu_summary.py gpurun_out/prof_r01/*.raw.csv"""
import csv
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed.sum", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum", "smsp__cycles_active.avg",
    "sm__cycles_elapsed.max", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "local_load_bytes", "smsp__inst_executed_op_local_ld.sum",
]


def main():
    for path in sys.argv[1:]:
        rows = list(csv.reader(open(path)))
        hdr = None
        for i, r in enumerate(rows):
            if "Kernel Name" in r:
                hdr = i
                break
        if hdr is None:
            print(path, ": no header")
            continue
        names, units = rows[hdr], rows[hdr + 1]
        for r in rows[hdr + 2:]:
            if len(r) != len(names):
                continue
            d = dict(zip(names, r))
            u = dict(zip(names, units))
            print(f"== {path}\n   kernel: {d.get('Kernel Name', '?')[:150]}")
            for k in KEYS:
                if k in d:
                    print(f"   {k:75s} {d[k]:>18s} {u.get(k, '')}")


if __name__ == "__main__":
    main()
