"""Key rows of an `ncu --set full` report as markdown.  usage: python scripts/summarize_full.py rep.ncu-rep [out.md] [title]"""
import csv
import io
import re
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
def col(name):
    return hdr.index(name) if name in hdr else None
want = [("Kernel Name", "kernel"), ("Grid Size", "grid"), ("Block Size", "block"), ("gpu__time_duration.sum", "time"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor (hmma) %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor insts"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem wavefronts %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
        ("launch__registers_per_thread", "regs"), ("launch__occupancy_limit_shared_mem", "occ limit smem"),
        ("sm__cycles_elapsed.max", "cycles")]
cols = [(col(n), t) for n, t in want if col(n) is not None]
out = [f"# {sys.argv[3] if len(sys.argv) > 3 else sys.argv[1]}", "", "| " + " | ".join(t + (f" [{units[c]}]" if units[c] else "") for c, t in cols) + " |",
       "|" + "---|" * len(cols)]
for r in data:
    vals = []
    for c, t in cols:
        v = r[c]
        if t == "kernel":
            v = "`" + re.sub(r"\(.*", "", v).replace("mtv::", "").replace("void ", "") + "`"
        vals.append(v)
    out.append("| " + " | ".join(vals) + " |")
txt = "\n".join(out) + "\n"
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(txt)
