import os, sys
sys.path.insert(0, "/root/repo")
import torch
from moditalker_b200 import BASE_UNET_CONFIG, DiffusionWrapper, UNetModel
from moditalker_b200.synth import synth_inputs, synth_state_dict
B = int(sys.argv[1])
cfg = BASE_UNET_CONFIG
m = DiffusionWrapper(UNetModel(**cfg)); m.load_state_dict(synth_state_dict(cfg, 0, "diffusion_model."), strict=True); m = m.cuda().eval()
x, c, ic, t = [v.cuda() for v in synth_inputs(B, seed=5)]
with torch.no_grad():
    m(x, c, ic, t)
torch.cuda.synchronize()
