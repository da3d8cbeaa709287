"""Batch independence at B=16 (beyond the batch sizes of the pytest suite): samples of a batch vs the same samples alone."""
import sys, torch
import os; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import config_by_name, rel_l2
from moditalker_b200 import DiffusionWrapper, UNetModel
from moditalker_b200.synth import synth_inputs, synth_state_dict
cfg = config_by_name("base")
m = DiffusionWrapper(UNetModel(**cfg)); m.load_state_dict(synth_state_dict(cfg, 0, "diffusion_model."), strict=True); m = m.cuda().eval()
B = 16
x, c, ic, t = synth_inputs(B, seed=99, t=[(131 * (i + 1)) % 1000 for i in range(B)])
with torch.no_grad():
    full = m(x.cuda(), c.cuda(), ic.cuda(), t.cuda()).cpu()
    for b in (0, 7, 15):
        one = m(x[b:b+1].cuda(), c[b:b+1].cuda(), ic[b:b+1].cuda(), t[b:b+1].cuda()).cpu()
        print("B=16 sample", b, "vs alone rel-L2", f"{rel_l2(full[b:b+1], one):.2e}", "finite", bool(torch.isfinite(full).all()))
