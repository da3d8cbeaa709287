"""Per-kernel-family DRAM traffic and time of one forward from an ncu launch list
(`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`).
usage: python scripts/summarize_traffic.py launches.csv B [out.md] -> merges into profiles/r02_traffic.json"""
import csv
import json
import os
import re
import sys
from collections import defaultdict

FAMILY = [("k_conv_tc", "conv_tc"), ("k_attn_tc", "attn_tc"), ("k_apply", "apply"), ("k_conv_simt", "conv"),
          ("k_splitk_epilogue", "conv"), ("k_gn_stats", "gn_stats"), ("k_linear_warp", "emb"), ("k_temb", "emb")]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1.0, "usecond": 1e3, "msecond": 1e6, "ns": 1.0, "us": 1e3}

with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ii, ki, mi, ui, vi = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
per = defaultdict(dict)
names = {}
for r in rd:
    try:
        per[r[ii]][r[mi]] = float(r[vi].replace(",", "")) * UNIT.get(r[ui], 1.0)
        names[r[ii]] = re.sub(r"\(.*", "", r[ki]).replace("mtv::", "")
    except (ValueError, IndexError):
        pass
fam = defaultdict(lambda: {"launches": 0, "ns": 0.0, "dram_bytes": 0.0})
kern = defaultdict(lambda: {"launches": 0, "ns": 0.0, "dram_bytes": 0.0})
for i, m in per.items():
    nm = names[i]
    f = next((fam_ for pre, fam_ in FAMILY if nm.startswith("void " + pre) or nm.startswith(pre)), "other")
    for d in (fam[f], kern[nm]):
        d["launches"] += 1
        d["ns"] += m.get("gpu__time_duration.sum", 0.0)
        d["dram_bytes"] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
B = sys.argv[2]
tot_ns = sum(v["ns"] for v in fam.values())
out = [f"# ncu launch list, one forward + DDIM step, B={B}: {sum(v['launches'] for v in fam.values())} launches, {tot_ns / 1e3:.1f} us summed "
       "(serialised ncu times, warm L2: compare SHARES), DRAM bytes = dram__bytes_read.sum + dram__bytes_write.sum", "",
       "| kernel | launches | total us | share | avg us | DRAM MB | MB / launch |", "|---|---:|---:|---:|---:|---:|---:|"]
for k, v in sorted(kern.items(), key=lambda kv: -kv[1]["ns"]):
    out.append(f"| `{k}` | {v['launches']} | {v['ns'] / 1e3:.1f} | {100 * v['ns'] / tot_ns:.1f}% | {v['ns'] / v['launches'] / 1e3:.2f} | "
               f"{v['dram_bytes'] / 1e6:.1f} | {v['dram_bytes'] / v['launches'] / 1e6:.2f} |")
out += ["", "| family | launches | total us | share | DRAM MB |", "|---|---:|---:|---:|---:|"]
for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ns"]):
    out.append(f"| {k} | {v['launches']} | {v['ns'] / 1e3:.1f} | {100 * v['ns'] / tot_ns:.1f}% | {v['dram_bytes'] / 1e6:.1f} |")
txt = "\n".join(out) + "\n"
print(txt)
if len(sys.argv) > 3:
    open(sys.argv[3], "w").write(txt)
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tp = os.path.join(root, "profiles", "r02_traffic.json")
data = json.load(open(tp)) if os.path.exists(tp) else {}
data[f"B{B}"] = {k: {"launches": v["launches"], "dram_bytes": v["dram_bytes"], "ncu_us": v["ns"] / 1e3} for k, v in fam.items()}
json.dump(data, open(tp, "w"), indent=1, sort_keys=True)
