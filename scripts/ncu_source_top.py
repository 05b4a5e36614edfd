"""Top SASS instructions of an `ncu --page source --csv` export by warp-stall samples, with their dominant stall reasons, and an
instruction-mix histogram of the executed instructions.  Usage: python scripts/ncu_source_top.py <source.csv> [top]"""
import csv
import re
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[2:]:
    try:
        data.append((float(r[col["# Samples"]]), float(r[col["Instructions Executed"]]), r))
    except (ValueError, IndexError):
        pass
tot = sum(s for s, _, _ in data) or 1.0
print(f"kernel: {rows[0][1][:140]}\ntotal samples {tot:.0f}, SASS instructions {len(data)}")
mix = Counter()
for s, n, r in data:
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[col["Source"]])
    if m:
        mix[m.group(2).split(".")[0]] += n
tn = sum(mix.values()) or 1.0
print("executed instruction mix: " + ", ".join(f"{k} {v / tn * 100:.1f}%" for k, v in mix.most_common(14)))
for i, (s, n, r) in enumerate(sorted(data, key=lambda t: -t[0])[:top]):
    why = sorted(((float(r[col[h]] or 0), h[6:]) for h in stalls), reverse=True)[:3]
    print(f"{s / tot * 100:5.1f}%  exec {n:10.0f}  {r[col['Source']].strip()[:84]:84s} " + " ".join(f"{h}:{v:.0f}" for v, h in why if v > 0))
