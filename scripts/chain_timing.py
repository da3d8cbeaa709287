"""GPU-side diagnostic: per-sub-op timing of the persistent chain kernel inside a steady-state (graph-replayed) forward.
Every CTA's thread 0 stamps %globaltimer at sub-op start, at the end of its own work (after the CTA-wide sync) and when it
leaves the grid barrier.  Usage: python scripts/chain_timing.py [B]"""
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
cap = 120000
buf = torch.zeros(cap * 16, dtype=torch.int64, device="cuda:0")
cnt = ctypes.c_int32(0)
_lib.check(lib.mtv_debug_tc_timing(h, ctypes.c_void_p(buf.data_ptr()), cap, ctypes.byref(cnt)), "arm")
with torch.no_grad():
    m(x, c, ic, t)
torch.cuda.synchronize()
_lib.check(lib.mtv_debug_tc_timing(h, None, 0, ctypes.byref(cnt)), "disarm")
n = min(cnt.value, cap)
rec = [r for r in buf.cpu().view(cap, 16)[:n].tolist() if (r[0] >> 62) & 1]
print(f"{len(rec)} chain sub-op records of {n} records in one forward (B={B})")
chains = defaultdict(lambda: defaultdict(list))
for r in rec:
    typ, idx, cta = (r[0] >> 48) & 0xff, (r[0] >> 32) & 0xffff, r[0] & 0xffffffff
    chains[(r[4], r[5])][(idx, typ)].append((cta, r[1], r[2], r[3]))
names = {0: "apply ", 1: "gemm  ", 2: "reduce"}
t0_all = min(r[1] for r in rec)
tot = defaultdict(float)
order = sorted(chains.items(), key=lambda kv: min(v[1] for vs in kv[1].values() for v in vs))
span_sum = 0.0
for (ptr, nops), subs in order:
    cstart = min(v[1] for vs in subs.values() for v in vs)
    cend = max(v[3] for vs in subs.values() for v in vs)
    span_sum += cend - cstart
    print(f"chain @{(cstart - t0_all) / 1e3:8.2f} us  nops {nops}  span {(cend - cstart) / 1e3:7.2f} us  ctas {len(next(iter(subs.values())))}")
    for (idx, typ), vs in sorted(subs.items()):
        s = min(v[1] for v in vs); w_mean = sum(v[2] - v[1] for v in vs) / len(vs); w_max = max(v[2] - v[1] for v in vs)
        e = max(v[3] for v in vs); last_arrive = max(v[2] for v in vs)
        print(f"    {idx:2d} {names.get(typ, '?')}  start {(s - cstart) / 1e3:7.2f}  work mean {w_mean / 1e3:6.2f} max {w_max / 1e3:6.2f}  "
              f"barrier exit after last arrival {(e - last_arrive) / 1e3:5.2f}  end {(e - cstart) / 1e3:7.2f}")
        tot[names.get(typ, '?')] += (e - s) / 1e3
print(f"sum of chain spans {span_sum / 1e3:.1f} us; first chain start -> last chain end {(max(r[3] for r in rec) - t0_all) / 1e3:.1f} us")
print("per type (sub-op start -> barrier exit), us:", {k: round(v, 1) for k, v in tot.items()})
