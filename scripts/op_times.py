"""Per-op CUDA-event times of one forward (each op replayed MTV_PROFILE_REPS times back to back in a private graph).
Usage: python scripts/op_times.py [B] [config]   (MTV_B200_LIB selects another build of the library for A/B)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from moditalker_b200 import BASE_UNET_CONFIG, LONGVID_UNET_CONFIG, DiffusionWrapper, UNetModel
from moditalker_b200.synth import synth_inputs, synth_state_dict

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cfg = LONGVID_UNET_CONFIG if (len(sys.argv) > 2 and sys.argv[2] == "longvid") else BASE_UNET_CONFIG
os.environ.setdefault("MTV_PROFILE_REPS", "10")
m = DiffusionWrapper(UNetModel(**cfg))
m.load_state_dict(synth_state_dict(cfg, 0, "diffusion_model."), strict=True)
m = m.to("cuda:0").eval()
x, c, ic, t = [v.cuda() for v in synth_inputs(B, seed=5)]
with torch.no_grad():
    for _ in range(3):
        m(x, c, ic, t)
    for _ in range(2):
        _, rows = m.diffusion_model.profile_forward(x, c, ic, t)
tot = sum(r[1] for r in rows)
print(f"B={B}: {len(rows)} ops, {tot:.1f} us summed")
for name, us, flops, byts in rows:
    print(f"{us:8.2f} us  {name}")
