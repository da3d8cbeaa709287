"""Recipe: stage the reference's OWN modules for this hot path under oracle/_ref/ (git-ignored, travels to the GPU box).

    python oracle/build_ref.py          # build container only: reads /root/reference (read-only, unmodified)

Why: /root/reference does not exist on the GPU box, but two things need the reference's unmodified Python there:
  * bench.py's ``gpu_eager_baseline`` leg — the reference UNet forward in eager PyTorch on the same B200
    ("the bar for each new kernel is what PyTorch eager dispatches on the same GPU"),
  * tests/test_reference_dropin_gpu.py — the reference's own ``DDPM`` sampler driving this repo's ``DiffusionWrapper``.
Nothing under oracle/_ref/ is committed (see .gitignore) and nothing in moditalker_b200/ imports it: it is
test / measurement infrastructure exactly like the rest of oracle/.  The files are byte-for-byte copies; the
hard-coded ``.to("cuda")`` at unet.py:1024 is harmless on a GPU box.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/MToV"
DST = os.path.join(ROOT, "oracle", "_ref", "MToV")
FILES = [
    "models/__init__.py",
    "models/ddpm/__init__.py",
    "models/ddpm/unet.py",                 # UNetModel / DiffusionWrapper (the hot function, unet.py:995-1117)
    "models/ddpm/diffusionmodules.py",     # GroupNorm32, conv_nd, timestep_embedding, checkpoint
    "losses/ddpm.py",                      # DDPM.sample / ddim_sample / model_predictions (losses/ddpm.py:338-484)
    "models/autoencoder/autoencoder_vit.py",   # ViTAutoencoder.extract / decode_from_sample (timed, not replaced)
    "models/autoencoder/vit_modules.py",
    "configs/latent-diffusion/base.yaml",
    "configs/autoencoder/base.yaml",
]


def available() -> bool:
    return os.path.isdir(REF)


def build(verbose: bool = False) -> bool:
    """Copy the listed files; returns False (and leaves any existing stage alone) when the reference tree is absent."""
    if not available():
        return False
    manifest = {}
    for rel in FILES:
        src = os.path.join(REF, rel)
        if not os.path.exists(src):
            continue
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(src, "rb").read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF, "files": manifest}, f, indent=1)
    if verbose:
        print(f"staged {len(manifest)} reference files under {DST}")
    return True


def staged_path():
    """Path to put on sys.path to import the staged reference (``models.ddpm.unet`` ...), or None."""
    return DST if os.path.exists(os.path.join(DST, "models", "ddpm", "unet.py")) else None


if __name__ == "__main__":
    ok = build(verbose=True)
    sys.exit(0 if ok else 1)
