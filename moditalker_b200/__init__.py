"""moditalker_b200 — B200-native (sm_100a) replacement for MoDiTalker's MToV
latent-video-diffusion denoising loop.

Drop-in surface (same names / signatures as the reference):
    UNetModel, DiffusionWrapper     <- MToV/models/ddpm/unet.py
    DDPM                            <- MToV/losses/ddpm.py
plus ``sample_chunks_sharded`` (chunk-level multi-GPU sampling, one NCCL
all-gather) which the reference does not have, and ``chunkio`` — the per-chunk
pixel work of MToV/sample.py and tools/dataloader_sample.py (frame preparation,
lower-half mask, landmark rasterisation, decoded frames -> uint8 / next reference)
as device kernels, and ``pipeline.sample_chunks``, the script's chunk loop
(sample.py:305-428) built on them.

Importing this package does not load the CUDA library; the first forward does,
and raises if ``libmtv_b200.so`` is missing (no fallback).
"""
from . import chunkio, pipeline
from .arch import BASE_UNET_CONFIG, LONGVID_UNET_CONFIG, TINY_UNET_CONFIG, build_arch
from .ddpm import DDPM
from .sharding import chunk_partition, sample_chunks_sharded
from .unet import DiffusionWrapper, UNetModel

__all__ = [
    "UNetModel", "DiffusionWrapper", "DDPM", "build_arch", "chunkio", "pipeline", "chunk_partition", "sample_chunks_sharded",
    "BASE_UNET_CONFIG", "LONGVID_UNET_CONFIG", "TINY_UNET_CONFIG",
]
__version__ = "0.1.0"
