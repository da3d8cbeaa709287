"""One denoising step between cudaProfilerStart / Stop, for `ncu --replay-mode range`: the kernels of the step run exactly as in
production (one CUDA-graph launch, PDL overlap), so dram__bytes is the TRUE whole-step traffic (a per-kernel ncu pass serialises
kernels and counts L2 weight prefetches separately from the reads they were issued for).  usage: step_traffic.py [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from moditalker_b200 import BASE_UNET_CONFIG, DDPM, DiffusionWrapper, UNetModel, _lib
from moditalker_b200.synth import synth_inputs, synth_state_dict

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cfg = BASE_UNET_CONFIG
m = DiffusionWrapper(UNetModel(**cfg))
m.load_state_dict(synth_state_dict(cfg, 0, "diffusion_model."), strict=True)
m = m.to("cuda:0").eval()
ddpm = DDPM(m, channels=4, image_size=32, sampling_timesteps=50, w=0.0).to("cuda:0")
x, c, ic, _ = [v.cuda() if v is not None else None for v in synth_inputs(B, seed=5)]
pairs = ddpm.time_pairs()
lib, h = m.diffusion_model.native_handle("cuda:0")
noise = torch.randn_like(x)
stream = torch.cuda.current_stream()


def step(i):
    t, tn = pairs[i]
    eps = m(x, c, ic, torch.full((B,), t, device="cuda:0", dtype=torch.long))
    sr, srm1, san, cc, sigma = ddpm.step_scalars(t, tn)
    _lib.check(lib.mtv_ddim_step(h, x.data_ptr(), eps.data_ptr(), noise.data_ptr(), x.numel(), sr, srm1, san, cc, sigma, 0,
                                 stream.cuda_stream), "ddim")


with torch.no_grad():
    for i in range(5):
        step(i)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    step(5)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("done")
