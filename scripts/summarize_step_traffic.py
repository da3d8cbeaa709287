"""Whole-step DRAM traffic from `ncu --replay-mode range` CSVs (scripts/gpu_traffic2.sh) -> profiles/r02_traffic.json["B<n>"]["_step"]
usage: python scripts/summarize_step_traffic.py B range.csv [B range.csv ...]"""
import csv
import json
import os
import sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tp = os.path.join(root, "profiles", "r02_traffic.json")
data = json.load(open(tp)) if os.path.exists(tp) else {}
for B, path in zip(sys.argv[1::2], sys.argv[2::2]):
    vals = {}
    for row in csv.reader(l for l in open(path) if l.startswith('"')):
        for i, c in enumerate(row):
            if c in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"):
                vals[c] = float(row[i + 2].replace(",", ""))
    rd, wr = vals["dram__bytes_read.sum"], vals["dram__bytes_write.sum"]
    data.setdefault(f"B{B}", {})["_step"] = {
        "dram_bytes": rd + wr, "dram_read_bytes": rd, "dram_write_bytes": wr, "range_ns_under_ncu": vals.get("gpu__time_duration.sum"),
        "how": "ncu --replay-mode range (cudaProfilerStart/Stop around ONE step: forward + DDIM update, eager launches with PDL, "
               "MTV_NO_GRAPH=1 because range replay cannot capture cuGraphLaunch), dram__bytes_read.sum + dram__bytes_write.sum; "
               "scripts/gpu_traffic2.sh"}
    print(f"B={B}: read {rd / 1e6:.1f} MB + write {wr / 1e6:.1f} MB = {(rd + wr) / 1e6:.1f} MB")
json.dump(data, open(tp, "w"), indent=1, sort_keys=True)
