#!/usr/bin/env python
"""Opcode histogram of an `ncu --page source --csv` export weighted by executed warp-instructions, plus the
top stall lines.  Usage: python scripts/ncu_sass_hist.py file.source.csv [n_elements]"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    nelem = float(sys.argv[2]) if len(sys.argv) > 2 else None
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    ix = {n: i for i, n in enumerate(hdr)}
    ops = collections.Counter()
    total = 0
    samples = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        sass = r[ix["Source"]].strip()
        ex = int(r[ix["Instructions Executed"]] or 0)
        st = int(r[ix["Warp Stall Sampling (All Samples)"]] or 0)
        op = sass.split()[0] if not sass.startswith("@") else sass.split()[1]
        op = op.split(".")[0]
        ops[op] += ex
        total += ex
        samples.append((st, ex, r[ix["Address"]], sass))
    print(f"total warp-instructions executed: {total}")
    if nelem:
        print(f"thread-instructions per element: {total * 32 / nelem:.1f}")
    for op, c in ops.most_common(25):
        print(f"  {op:12s} {c:12d} {100.0 * c / total:5.1f}%" + (f"  {c * 32 / nelem:6.2f}/elem" if nelem else ""))
    print("top stall lines:")
    for st, ex, addr, sass in sorted(samples, reverse=True)[:12]:
        print(f"  {st:6d} samples  ex={ex:9d}  {sass}")


if __name__ == "__main__":
    main()
