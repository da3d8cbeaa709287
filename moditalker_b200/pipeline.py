"""The chunk loop of MoDiTalker's sampling script with the pixel work on the GPU.

``sample_chunks`` is MToV/sample.py:305-428 for one identity as a generator: for every 16-frame chunk it prepares the four
clips (``chunkio.prep_frames`` / ``rasterize_landmarks`` instead of the CPU loader's resize / mask / cv2 drawing), calls the
first-stage autoencoders the caller passes in (the reference's ``ViTAutoencoder`` objects — they stay reference PyTorch),
samples with the drop-in ``DDPM``, turns the decoded frames into what the script writes (``chunkio.frames_out``) and chains
the last frame into the next chunk's reference (``--use_last_as_reference``) without the PNG round trip through the disk —
the tensor it hands over is bit-identical to the one the script reads back.  Files, if wanted, go through
``chunkio.AsyncFrameWriter`` under the script's names.

Every step is the same arithmetic as the script's (tests/test_pipeline_gpu.py compares a two-chunk run bit for bit against the
script's own sequence restated with the CPU oracle of the pixel work and the actual PNG round trip).
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Iterable, Iterator, Optional

import torch

from . import chunkio


@dataclass
class Chunk:
    """What ``EvalDataset.__getitem__`` reads for one chunk (tools/dataloader_sample.py:182-214), before any pixel work."""
    frames_u8: torch.Tensor          # [16, H, W, 3] uint8: the clip's decoded frames
    first_frame_u8: torch.Tensor     # [H, W, 3] uint8: frame 0 of the identity (``self.imgs[0]``), the reference image
    landmarks: torch.Tensor          # [16, N, 3] (normalised) or [16, N, 2] (pixels), fp32 / fp64 as stored: the AToM output
    keypoints: object                # [16, 68, 2] array-like: landmark 33's row is where the lower-half mask starts


@dataclass
class ChunkResult:
    index: int
    z: torch.Tensor                  # [k, 4, 2048] sampled latent
    frames_u8: torch.Tensor          # [k, 16, 256, 256, 3] uint8 video frames (what the script writes)
    last_u8: torch.Tensor            # [k, 256, 256, 3] uint8: pixels of the last-frame PNGs
    next_ref: torch.Tensor           # [k, 3, 16, 256, 256] fp32: that PNG as the script reads it back for the next chunk


@torch.no_grad()
def sample_chunks(first_stage_model, first_stage_model_ldmk, diffusion_model, chunks: Iterable[Chunk], *, device,
                  batch_size: int = 1, use_last_as_reference: bool = True, x_noisy_start: bool = True,
                  refvid_noisy_start: bool = False, ratio_: float = 0.25, fix_noise: bool = True, resolution: int = 256,
                  writer: Optional[chunkio.AsyncFrameWriter] = None, out_dir: Optional[str] = None,
                  including_ldmk_video: bool = False, loader_worker: bool = False) -> Iterator[ChunkResult]:
    """MToV/sample.py:305-428.  Keyword names are the script's command-line flags; ``batch_size`` is its ``k``.

    ``first_stage_model`` / ``first_stage_model_ldmk`` need ``extract`` and ``decode_from_sample`` (autoencoder_vit.py:212-275),
    ``diffusion_model`` is ``moditalker_b200.DDPM`` (or the reference's own).  With ``writer`` and ``out_dir`` the chunk's GIF,
    numbered frames and last-frame PNGs are written as the script writes them (``gif/``, ``frames/``, ``references/<frame>/``).
    ``loader_worker=True`` resizes as torch does inside the script's DataLoader workers (see ``chunkio.prep_frames``)."""
    dev = torch.device(device)
    k = int(batch_size)
    prev_ref = None                                                        # previous chunk's last frames, as read back
    lw = {"loader_worker": True} if loader_worker else {}                  # default: the verified direct-call form, no extra argument
    for it, ch in enumerate(chunks):
        ldmk_srt, ldmk_end = it * 16, it * 16 + 16
        frames = ch.frames_u8.to(dev, non_blocking=True)
        T, H, W, _ = frames.shape
        # the four clips of sample.py:318-325 (x_ref: frame 0 repeated, dataloader_sample.py:192-193)
        x = chunkio.prep_frames(frames, None, resolution, **lw)
        x_ref = chunkio.prep_frames(ch.first_frame_u8.to(dev, non_blocking=True).unsqueeze(0).expand(T, -1, -1, -1), None, resolution, **lw)
        rows = [chunkio.lower_half_start(H, ch.keypoints[t]) for t in range(T)]
        masked_x = chunkio.prep_frames(frames, rows, resolution, **lw)
        x_l = chunkio.rasterize_landmarks(ch.landmarks.to(dev, non_blocking=True), W)   # WH = vid.shape[-1] (dataloader_sample.py:215)
        if k > 1:                                                          # the script's loader batch is 1; k identities share the clips
            x, x_ref, masked_x, x_l = (t.expand(k, -1, -1, -1, -1) for t in (x, x_ref, masked_x, x_l))

        z_ = first_stage_model.extract(x).detach()                         # sample.py:327-331
        image_cond_ = first_stage_model.extract(x_ref).detach()
        z_l = first_stage_model_ldmk.extract(x_l).detach()
        masked_z = first_stage_model.extract(masked_x).detach()
        image_cond = image_cond_[:, :, 0:32 * 32]
        if use_last_as_reference and prev_ref is not None:                 # sample.py:340-358, without the disk
            image_cond = first_stage_model.extract(prev_ref.detach())[:, :, 0:32 * 32]

        c = torch.cat([z_l, masked_z], dim=1)                              # sample.py:365
        noised_start = None
        if x_noisy_start:
            noised_start = image_cond_.float()
        elif refvid_noisy_start:
            noised_start = z_.float()
        z = diffusion_model.sample(batch_size=k, cond=c.float(), image_cond=image_cond.float(), noised_start=noised_start,
                                   ratio_=ratio_, fix_noise=fix_noise)     # sample.py:373-380
        decoded = first_stage_model.decode_from_sample(z)                  # sample.py:381
        frames_u8, last_u8, next_ref = chunkio.frames_out(decoded.float(), k, 16)
        prev_ref = next_ref

        if writer is not None and out_dir is not None:                     # sample.py:385-396, 411-424
            writer.save_last_frames(last_u8, os.path.join(out_dir, "references", str(ldmk_end)))
            lm = x_l[:k] if including_ldmk_video else None
            os.makedirs(os.path.join(out_dir, "gif"), exist_ok=True)
            writer.save_gif(frames_u8, os.path.join(out_dir, "gif", f"generated_{it}.gif"), lm)
            writer.save_frames(ldmk_srt, frames_u8, os.path.join(out_dir, "frames"), lm)
        yield ChunkResult(it, z, frames_u8, last_u8, next_ref)
