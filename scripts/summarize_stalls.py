"""Top warp-stall sampling sites (SASS level) per kernel of an `ncu --set full --import-source on` report.
usage: python scripts/summarize_stalls.py rep.ncu-rep [out.md] [title] [topN]"""
import csv
import io
import re
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
title = sys.argv[3] if len(sys.argv) > 3 else sys.argv[1]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 10
out = [f"# warp-stall sampling by SASS instruction — {title}", "",
       "`ncu -i ... --page source --csv`; samples = 'Warp Stall Sampling (All Samples)'; the instruction a warp is stalled AT waits for the work named next to it.", ""]
blocks = re.split(r'(?m)^"Kernel Name",', raw)
seen = set()
for blk in blocks[1:]:
    lines = blk.splitlines()
    kname = re.sub(r"\(mtv::.*", "", lines[0].strip('",')).replace("mtv::", "").replace("void ", "")
    rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
    if not rows:
        continue
    hdr = rows[0]
    try:
        si, ai = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
    except ValueError:
        continue
    data = []
    for r in rows[1:]:
        try:
            data.append((int(r[ai]), r[si].strip()))
        except (ValueError, IndexError):
            pass
    tot = sum(d[0] for d in data) or 1
    key = (kname, tot)
    if kname in seen:
        continue            # one instance per kernel name
    seen.add(kname)
    out += [f"## `{kname}` ({tot} samples)", "", "| share | samples | SASS |", "|---:|---:|---|"]
    for n, src in sorted(data, reverse=True)[:topn]:
        out.append(f"| {100 * n / tot:.1f}% | {n} | `{src[:110]}` |")
    out.append("")
txt = "\n".join(out) + "\n"
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(txt)
