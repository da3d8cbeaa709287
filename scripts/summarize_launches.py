"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel.
usage: python scripts/summarize_launches.py gpurun_out/launches.csv [out.md]"""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
for r in rd:
    try:
        rows.append((r[ki], float(r[vi].replace(",", "")), r[gi], r[bi]))
    except (ValueError, IndexError):
        pass
agg = defaultdict(lambda: [0, 0.0])
for name, ns, _, _ in rows:
    short = re.sub(r"\(.*", "", name).replace("mtv::", "")
    agg[short][0] += 1
    agg[short][1] += ns
tot = sum(v[1] for v in agg.values())
out = [f"# launch list summary ({len(rows)} launches, {tot / 1e3:.1f} us total, cold-cache serialised ncu times: compare SHARES)", "",
       "| kernel | launches | total us | share | avg us |", "|---|---:|---:|---:|---:|"]
for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| `{k}` | {n} | {ns / 1e3:.1f} | {100 * ns / tot:.1f}% | {ns / n / 1e3:.2f} |")
txt = "\n".join(out) + "\n"
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(txt)
