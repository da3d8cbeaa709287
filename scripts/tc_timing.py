"""GPU-side diagnostic: phase breakdown of the tensor-core tap-GEMM CTAs inside a steady-state
(graph-replayed) forward.  Prints, per distinct launch shape, the mean clock64 deltas:
  setup   = kernel entry -> barriers/TMEM ready
  first   = -> first operands landed (TMA latency)
  stream  = first -> last operands landed
  drain   = last operands -> accumulator complete
  epi     = accumulator complete -> epilogue stores issued
  total   = entry -> exit,  plus the globaltimer span of the CTA."""
import ctypes
import os
import sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from moditalker_b200 import BASE_UNET_CONFIG, DiffusionWrapper, UNetModel, _lib
from moditalker_b200.synth import synth_inputs, synth_state_dict

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cfg = BASE_UNET_CONFIG
m = DiffusionWrapper(UNetModel(**cfg))
m.load_state_dict(synth_state_dict(cfg, 0, "diffusion_model."), strict=True)
m = m.to("cuda:0").eval()
x, c, ic, t = [v.cuda() for v in synth_inputs(B, seed=5)]
with torch.no_grad():
    for _ in range(4):
        m(x, c, ic, t)
torch.cuda.synchronize()
lib, h = m.diffusion_model.native_handle()
cap = 60000
buf = torch.zeros(cap * 16, dtype=torch.int64, device="cuda:0")
cnt = ctypes.c_int32(0)
_lib.check(lib.mtv_debug_tc_timing(h, ctypes.c_void_p(buf.data_ptr()), cap, ctypes.byref(cnt)), "arm")
with torch.no_grad():
    m(x, c, ic, t)
torch.cuda.synchronize()
_lib.check(lib.mtv_debug_tc_timing(h, None, 0, ctypes.byref(cnt)), "disarm")
n = min(cnt.value, cap)
rec = buf.cpu().view(cap, 16)[:n]
print(f"{n} CTA records in one forward (B={B})")
# record: [0] grid, [1] shape, [2..9] clock64 stamps: entry, setup done, first operands, last operands, accumulator
# complete, split-K tile ticket passed | arrival at the grid barrier, grid barrier passed, epilogue done; [10],[11] globaltimer
groups = defaultdict(list)
for r in rec.tolist():
    groups[(r[0], r[1], r[15])].append(r)
rows = []
for (g0, g1, fl), rs in groups.items():
    gx, gy, gz = g0 & 0xffff, (g0 >> 16) & 0xffff, (g0 >> 32) & 0xffff
    iters, taps, cin, cout = g1 & 0xffff, (g1 >> 16) & 0xff, (g1 >> 24) & 0xffff, (g1 >> 40) & 0xffff
    def mean(f):
        return sum(f(r) for r in rs) / len(rs)
    def mx(f):
        return max(f(r) for r in rs)
    setup = mean(lambda r: r[3] - r[2]); first = mean(lambda r: r[4] - r[3]); stream = mean(lambda r: r[5] - r[4])
    drain = mean(lambda r: r[6] - r[5])
    t5 = lambda r: r[7] if r[7] else r[6]
    t6 = lambda r: r[8] if r[8] else t5(r)
    e1 = mean(lambda r: t5(r) - r[6]); e2 = mean(lambda r: t6(r) - t5(r)); e3 = mean(lambda r: r[9] - t6(r))
    total = mean(lambda r: r[14] - r[2])
    span = mean(lambda r: r[11] - r[10])
    launches = len(rs) / max(1, gx * gy * gz)
    rows.append((total * launches, f"grid {gx:3d}x{gy:2d}x{gz:2d} iters {iters:3d} taps {taps} Cin {cin:4d} Cout {cout:4d} fa {fl & 1} | launches {launches:5.1f} | "
                 f"setup {setup:6.0f} first {first:6.0f} stream {stream:6.0f} drain {drain:6.0f} | tail: to-ticket/arrive {e1:6.0f} barrier {e2:6.0f} finish {e3:6.0f} | total {total:7.0f} cyc | span {span / 1e3:6.2f} us"))
for _, line in sorted(rows, reverse=True):
    print(line)

# ---- timeline of launches (globaltimer, ns): one line per launch = CTAs sharing (shape, ~start)
evs = sorted(rec.tolist(), key=lambda r: r[10])
launches = []
for r in evs:
    key = (r[0], r[1])
    if launches and launches[-1]["key"] == key and r[10] - launches[-1]["t0"] < 200000 and launches[-1]["n"] < launches[-1]["size"]:
        L = launches[-1]; L["t1"] = max(L["t1"], r[11]); L["n"] += 1
    else:
        g0 = r[0]; size = (g0 & 0xffff) * ((g0 >> 16) & 0xffff) * ((g0 >> 32) & 0xffff)
        launches.append({"key": key, "t0": r[10], "t1": r[11], "n": 1, "size": size})
t_base = launches[0]["t0"]
print(f"\n{len(launches)} tensor-core launches; timeline (us since first): start, duration, gap since previous TC launch end")
prev_end = None
tot_dur = 0
for i, L in enumerate(launches):
    g0, g1 = L["key"]
    dur = (L["t1"] - L["t0"]) / 1e3; tot_dur += dur
    gap = (L["t0"] - prev_end) / 1e3 if prev_end else 0.0
    if i < 70:
        print(f"{(L['t0'] - t_base) / 1e3:9.2f}  dur {dur:6.2f}  gap {gap:6.2f}  grid {(g0 & 0xffff)}x{(g0 >> 16) & 0xffff}x{(g0 >> 32) & 0xffff} iters {g1 & 0xffff} taps {(g1 >> 16) & 0xff} Cin {(g1 >> 24) & 0xffff} Cout {(g1 >> 40) & 0xffff}")
    prev_end = L["t1"]
print(f"sum of TC launch durations {tot_dur:.1f} us; span {(launches[-1]['t1'] - t_base) / 1e3:.1f} us")

