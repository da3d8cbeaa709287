"""Generate tests/golden/*.npz by running the REFERENCE's own modules on CPU.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):   python oracle/make_golden.py [--only NAME]

What it does
  * imports ``models.ddpm.unet`` and ``losses.ddpm`` from /root/reference/MToV
    (read-only, unmodified),
  * neutralises the hard-coded ``.to("cuda")`` at unet.py:1024 with a runtime shim
    (``torch.Tensor.to`` ignores a literal "cuda" while the shim is active) — the
    reference file is not edited,
  * loads the by-name synthetic weights (moditalker_b200/synth.py) with
    ``strict=True`` and runs ``DiffusionWrapper.forward`` / ``DDPM.sample`` on the
    synthetic inputs,
  * writes the outputs (and the recipe needed to re-create the inputs) as small
    fixtures.  Weights and inputs are NOT stored: they are re-derived from
    (config name, seed) by the same synth module on any machine.
"""
from __future__ import annotations

import argparse
import contextlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference/MToV"

from moditalker_b200.arch import BASE_UNET_CONFIG, LONGVID_UNET_CONFIG, TINY_UNET_CONFIG  # noqa: E402
from moditalker_b200.synth import synth_inputs, synth_state_dict  # noqa: E402

CONFIGS = {"tiny": TINY_UNET_CONFIG, "base": BASE_UNET_CONFIG, "longvid": LONGVID_UNET_CONFIG}
GOLD = os.path.join(ROOT, "tests", "golden")
TAP_CSTRIDE, TAP_LSTRIDE = 4, 8


@contextlib.contextmanager
def cuda_shim():
    orig = torch.Tensor.to

    def to(self, *a, **k):
        if a and isinstance(a[0], str) and a[0] == "cuda":
            a = a[1:]
            if not a and not k:
                return self
        return orig(self, *a, **k)

    torch.Tensor.to = to
    try:
        yield
    finally:
        torch.Tensor.to = orig


def ref_modules():
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from models.ddpm.unet import DiffusionWrapper, UNetModel   # noqa
    from losses.ddpm import DDPM                               # noqa
    return UNetModel, DiffusionWrapper, DDPM


def build_ref(cfg_name: str, wseed: int):
    UNetModel, DiffusionWrapper, _ = ref_modules()
    cfg = CONFIGS[cfg_name]
    model = DiffusionWrapper(UNetModel(**cfg))
    model.load_state_dict(synth_state_dict(cfg, wseed, "diffusion_model."), strict=True)
    return model.eval()


# tap name -> (module attribute path whose per-plane outputs are captured, concatenated xy|yt|xt)
def capture_taps(model, names):
    """Forward hooks on input_attns / mid_attn / output_attns capture the joint token
    tensor [B,C,L] after each stage (the value the next stage consumes)."""
    taps, handles = {}, []
    um = model.diffusion_model
    table = {}
    for i, m in enumerate(um.input_attns):
        table[f"in{i}"] = m
    table["mid"] = um.mid_attn
    for i, m in enumerate(um.output_attns):
        table[f"out{i}"] = m
    for n in names:
        if n == "in0":
            continue
        handles.append(table[n].register_forward_hook(lambda mod, inp, out, n=n: taps.__setitem__(n, out.detach().clone())))
    return taps, handles


def gen_forward(name, cfg_name, B, t, wseed=0, iseed=2, ic_len=1024, taps=()):
    model = build_ref(cfg_name, wseed)
    x, cond, ic, tt = synth_inputs(B, iseed, ic_len, t)
    cap, handles = capture_taps(model, taps)
    t0 = time.time()
    with torch.no_grad(), cuda_shim():
        eps = model(x, cond, ic, tt)
    dt = time.time() - t0
    for h in handles:
        h.remove()
    out = {"eps": eps.numpy(), "config": cfg_name, "B": B, "t": np.asarray(tt), "wseed": wseed, "iseed": iseed,
           "ic_len": ic_len}
    for k, v in cap.items():   # strided sample of the [B,C,L] stage output (keeps fixtures small)
        out["tap_" + k] = v[:, ::TAP_CSTRIDE, ::TAP_LSTRIDE].contiguous().numpy()
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(f"{name}: eps {tuple(eps.shape)} rms={eps.pow(2).mean().sqrt():.4f} ({dt:.1f}s)")


def gen_ddim(name, cfg_name, B, S, wseed=0, iseed=2, nseed=3, noised=False, ratio=None, fix_noise=False, ic_len=1024):
    _, _, DDPM = ref_modules()
    model = build_ref(cfg_name, wseed)
    x, cond, ic, _ = synth_inputs(B, iseed, ic_len, 0)
    ddpm = DDPM(model, channels=4, image_size=32, sampling_timesteps=S, w=0.0)
    torch.manual_seed(nseed)   # the reference draws its noise from the global CPU generator
    t0 = time.time()
    with cuda_shim():
        if noised:
            x0 = torch.tanh(x)   # a latent-like start in [-1, 1]
            z = ddpm.sample(batch_size=B, cond=cond, image_cond=ic, noised_start=x0, ratio_=ratio, fix_noise=fix_noise)
        else:
            z = ddpm.sample(batch_size=B, cond=cond, image_cond=ic)
    dt = time.time() - t0
    np.savez_compressed(
        os.path.join(GOLD, name + ".npz"), z=z.numpy(), config=cfg_name, B=B, S=S, wseed=wseed, iseed=iseed,
        nseed=nseed, noised=int(noised), ratio=-1.0 if ratio is None else ratio, fix_noise=int(fix_noise), ic_len=ic_len,
    )
    print(f"{name}: z {tuple(z.shape)} rms={z.pow(2).mean().sqrt():.4f} ({dt:.1f}s)")


CASES = {
    "unet_tiny_b2": lambda: gen_forward("unet_tiny_b2", "tiny", 2, [500, 37], taps=("in1", "in2", "mid", "out0", "out3")),
    "unet_base_b1": lambda: gen_forward("unet_base_b1", "base", 1, 500, taps=("in11", "mid")),
    "unet_base_b2": lambda: gen_forward("unet_base_b2", "base", 2, [999, 3], iseed=7, ic_len=2048),
    "unet_longvid_b1": lambda: gen_forward("unet_longvid_b1", "longvid", 1, 250, iseed=5),
    "ddim_tiny_s10": lambda: gen_ddim("ddim_tiny_s10", "tiny", 2, 10),
    "ddim_tiny_noised": lambda: gen_ddim("ddim_tiny_noised", "tiny", 1, 20, noised=True, ratio=0.25, fix_noise=True),
    "ddim_base_s50": lambda: gen_ddim("ddim_base_s50", "base", 1, 50),
    "ddim_base_noised_r25": lambda: gen_ddim("ddim_base_noised_r25", "base", 1, 100, noised=True, ratio=0.25, fix_noise=True),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", nargs="*", default=None)
    args = ap.parse_args()
    if not os.path.isdir(REF):
        raise SystemExit("reference tree not present; fixtures can only be generated in the build container")
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    for n, fn in CASES.items():
        if args.only and n not in args.only:
            continue
        fn()


if __name__ == "__main__":
    main()
