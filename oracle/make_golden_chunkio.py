"""Generate tests/golden/chunkio_*.npz by running the REFERENCE's own chunk I/O code on CPU.

Run in the build container only (needs /root/reference):   python oracle/make_golden_chunkio.py

  * imports ``tools.dataloader_sample`` and ``tools.data_utils`` from /root/reference/MToV unmodified.  Three modules they
    import at module level but never use on this path are absent from the image (``av``, ``imageio``, ``natsort``): empty
    stand-ins are placed in ``sys.modules`` for the import (and ``torchvision.io.read_video``, which current torchvision no
    longer has, is given a placeholder); no reference file is edited,
  * calls ``EvalDataset._load_img_from_path`` (on PNG files written to a temp dir), ``_crop_lower_half``,
    ``_change_np_img_size`` and ``data_utils.resize_crop`` exactly as ``EvalDataset.__getitem__`` chains them
    (dataloader_sample.py:182-245), followed by the normalisation / layout lines of sample.py:322-325,
  * for the inline post-processing of sample.py:380-399 and the read-back of sample.py:344-358 (script code, not functions)
    executes the literal library calls the script makes (torch clamp / rearrange, cv2.cvtColor, np.rint, cv2.imwrite,
    PIL.Image.open, ToTensor),
  * stores inputs and outputs as small fixtures (key-point canvases bit-packed).
"""
from __future__ import annotations

import os
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/MToV"
GOLD = os.path.join(ROOT, "tests", "golden")


def ref_modules():
    for name in ("av", "imageio", "natsort"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            if name == "natsort":
                m.natsorted = sorted
            sys.modules[name] = m
    import torchvision.io
    if not hasattr(torchvision.io, "read_video"):      # removed from current torchvision; imported by name, unused on this path
        torchvision.io.read_video = None
    sys.path.insert(0, REF)
    cwd = os.getcwd()
    os.chdir(REF)                      # the module reads text_folders/train_id.txt at import time
    try:
        import tools.dataloader_sample as dls
        import tools.data_utils as du
    finally:
        os.chdir(cwd)
    return dls, du


def make_prep(dls, du, name, T, H, W, R, seed, masked, worker=False):
    import PIL.Image
    from einops import rearrange
    from torchvision import transforms

    rng = np.random.default_rng(seed)
    # smooth-ish content plus noise so that both flat and high-gradient regions occur; every byte value appears
    base = rng.integers(0, 256, size=(T, H // 4 + 1, W // 4 + 1, 3), dtype=np.uint8)
    frames = np.repeat(np.repeat(base, 4, axis=1), 4, axis=2)[:, :H, :W]
    frames = (frames.astype(np.int32) + rng.integers(-20, 21, size=frames.shape)).clip(0, 255).astype(np.uint8)
    frames[0, :2, :128 if W >= 128 else W, 0] = np.arange(256, dtype=np.uint8)[: (128 if W >= 128 else W)]
    kpts = rng.uniform(0, H, size=(T, 68, 2))
    kpts[0, 33, 1] = H * 0.55
    if T > 1:
        kpts[1, 33, 1] = -H * 0.25          # negative start row: numpy counts from the end
    if T > 2:
        kpts[2, 33, 1] = H + 7.5            # beyond the image: nothing masked
    stub = types.SimpleNamespace(img_resolution=R, to_tensor=transforms.ToTensor(), _open_file=lambda p: open(p, "rb"))
    with tempfile.TemporaryDirectory() as d:
        for t in range(T):
            PIL.Image.fromarray(frames[t]).save(os.path.join(d, f"{t:04d}.png"))
        load = lambda t: dls.EvalDataset._load_img_from_path(stub, d, f"{t:04d}.png")
        vid = np.stack([load(t) for t in range(T)], axis=0)                              # [T, 3, H, W] fp32 0..255
        if masked:
            vid = np.stack([dls.EvalDataset._crop_lower_half(stub, load(t), kpts[t]) for t in range(T)], axis=0)
    if worker:
        # the shipped script runs the loader in DataLoader workers (get_loaders: num_workers=4); a worker process has ONE torch
        # thread, and with one thread torch resizes 3-channel input with its other CPU kernel
        nthreads = torch.get_num_threads()
        torch.set_num_threads(1)
    out = du.resize_crop(torch.from_numpy(vid).float(), resolution=R)                    # c t h w
    if worker:
        torch.set_num_threads(nthreads)
    x = rearrange(out, "c t h w -> t c h w")[None]                                       # DataLoader batch of 1
    x = rearrange(x / 127.5 - 1, "b t c h w -> b c t h w")[0]                            # sample.py:322-325
    np.savez_compressed(os.path.join(GOLD, f"chunkio_{name}.npz"), frames=frames, kpts=kpts, R=R, masked=int(masked),
                        worker=int(worker), out=x.numpy().astype(np.float32))
    print(name, x.shape, float(x.min()), float(x.max()))


def make_landmarks(dls, name, T, N, WH, dims, dtype, flip, seed):
    from einops import rearrange

    rng = np.random.default_rng(seed)
    if dims == 3:
        lm = rng.uniform(-1.05, 1.05, size=(T, N, 3))
        lm[0, :4, :2] = [[-1, -1], [1, 1], [0.999, -0.999], [-1.02, 0.3]]
    else:
        lm = rng.uniform(-8, WH + 8, size=(T, N, 2))
        lm[0, :3] = [[0, 0], [WH - 1, WH - 1], [WH, 2.5]]
    lm = lm.astype(dtype)
    stub = types.SimpleNamespace()
    img = dls.EvalDataset._change_np_img_size(stub, lm, WH=WH, flip=flip)                # [T, 256, 256, 3] uint8
    land = torch.from_numpy(rearrange(img, "t h w c -> t c h w")).float()[None]          # dataloader_sample.py:216-222
    x_l = rearrange(land / 127.5 - 1, "b t c h w -> b c t h w")[0].numpy()               # sample.py:324
    assert set(np.unique(x_l)) <= {-1.0, 1.0}
    assert (x_l[0] == x_l[1]).all() and (x_l[0] == x_l[2]).all()
    np.savez_compressed(os.path.join(GOLD, f"chunkio_{name}.npz"), lm=lm, WH=WH, flip=int(flip),
                        canvas_bits=np.packbits(x_l[0] > 0))
    print(name, x_l.shape, int((x_l[0] > 0).sum()), "white pixels")


def make_frames_out(name, B, T, H, W, seed):
    import cv2
    import PIL.Image
    from einops import rearrange
    from torchvision import transforms

    rng = np.random.default_rng(seed)
    dec = rng.uniform(-1.2, 1.2, size=(B * T, 3, H, W)).astype(np.float32)
    # values that land exactly on .5 after (1 + x) * 127.5 (round-half-even) and on the clamp edges
    dec[-1, :, 0, :8] = np.array([-1.0, 1.0, 0.5 / 127.5 - 1, 1.5 / 127.5 - 1, 2.5 / 127.5 - 1, 0.0, 254.5 / 127.5 - 1, -5.0], dtype=np.float32)
    fake = torch.from_numpy(dec).clamp(-1, 1).cpu()                                      # sample.py:380
    fake = (1 + rearrange(fake, "(b t) c h w -> b t h w c", b=B)) * 127.5               # :381
    last_frame = fake[:, -1, :, :, :]                                                    # :386
    last_u8, next_ref = [], []
    Img2Tensor = transforms.ToTensor()
    with tempfile.TemporaryDirectory() as d:
        for idx in range(B):
            fname = os.path.join(d, f"{idx}.png")
            img = np.asarray(last_frame[idx], dtype=np.float32)                          # :391
            img = cv2.cvtColor(img, cv2.COLOR_BGR2RGB)
            img = np.rint(img).clip(0, 255).astype(np.uint8)
            cv2.imwrite(fname, img)                                                      # :394
            pil = PIL.Image.open(fname)
            last_u8.append(np.asarray(pil).copy())                                       # the file's RGB pixels
            t = Img2Tensor(pil)                                                          # :351
            t = t * 2.0 - 1.0
            t = t.unsqueeze(0)
            next_ref.append(torch.cat([t for _ in range(16)], dim=0))                    # :354-355
    frames_tensor = torch.stack(next_ref, dim=0)
    frames_tensor = rearrange(frames_tensor, "b t c h w -> b c t h w")                   # :358
    frames_u8 = fake.type(torch.uint8)                                                   # :399
    np.savez_compressed(os.path.join(GOLD, f"chunkio_{name}.npz"), dec=dec, B=B, frames_u8=frames_u8.numpy(),
                        last_u8=np.stack(last_u8), next_ref=frames_tensor.numpy().astype(np.float32))
    print(name, frames_u8.shape, frames_tensor.shape)


def main():
    os.makedirs(GOLD, exist_ok=True)
    dls, du = ref_modules()
    make_prep(dls, du, "prep_plain", T=3, H=72, W=96, R=32, seed=11, masked=False)       # landscape: crop + 2.25x down
    make_prep(dls, du, "prep_masked", T=3, H=90, W=64, R=40, seed=12, masked=True)       # portrait, masked stream, 1.6x down
    make_prep(dls, du, "prep_identity", T=2, H=32, W=32, R=32, seed=13, masked=True)     # no resampling: bit-exact case
    make_prep(dls, du, "prep_frac", T=2, H=50, W=70, R=36, seed=15, masked=False)        # inexact scale, unmasked (non-integer pixels)
    make_prep(dls, du, "prep_wide", T=1, H=101, W=117, R=128, seed=16, masked=True)      # output wider than 64: torch's other loop
    make_prep(dls, du, "prep_wide_down", T=1, H=333, W=301, R=72, seed=17, masked=False)  # ... downsampling, odd sizes
    make_prep(dls, du, "prep_worker", T=1, H=205, W=187, R=96, seed=18, masked=True, worker=True)   # as in a DataLoader worker
    make_prep(dls, du, "prep_up", T=2, H=20, W=20, R=32, seed=14, masked=False)          # upsampling
    make_landmarks(dls, "lm_norm_f32", T=3, N=40, WH=634, dims=3, dtype=np.float32, flip=False, seed=21)
    make_landmarks(dls, "lm_norm_f64", T=2, N=40, WH=256, dims=3, dtype=np.float64, flip=True, seed=22)
    make_landmarks(dls, "lm_pixel_f64", T=2, N=40, WH=726, dims=2, dtype=np.float64, flip=False, seed=23)
    make_frames_out("frames_out", B=2, T=3, H=16, W=24, seed=31)


if __name__ == "__main__":
    main()
