"""Chunk I/O around the denoising loop on the GPU.

Host-side mirror of what MoDiTalker's sampling script does on the CPU for every 16-frame chunk, before and after
``DDPM.sample``:

    prep_frames            EvalDataset._load_img_from_path / _crop_lower_half / resize_crop
                           (MToV/tools/dataloader_sample.py:130-146, MToV/tools/data_utils.py:73-98) followed by
                           ``x / 127.5 - 1`` and the "b t c h w -> b c t h w" rearrange (MToV/sample.py:322-325)
    lower_half_start       the row ``mask[int(landmarks[33][1]):, :] = 0`` starts at (dataloader_sample.py:134-135)
    rasterize_landmarks    EvalDataset._change_np_img_size (dataloader_sample.py:153-180) + sample.py:324
    frames_out             sample.py:380-399 (clamp, (1 + x) * 127.5, uint8 video frames, the last frame as the PNG's pixels)
                           and sample.py:344-358 (that PNG read back as the next chunk's reference clip)

Every function takes and returns CUDA tensors, runs on the caller's current stream through ``libmtv_b200.so`` and is
bit-exact against the reference sequence (tests/test_chunkio_gpu.py).  There is no CPU fallback: a CPU tensor raises.
Decoding JPEG and encoding PNG / GIF stay on the host; these functions are the pixel work between the files and the
autoencoder.  ``AsyncFrameWriter`` takes the encoding off the GPU loop's critical path: pinned device-to-host copies on a side
stream and a worker thread that writes the script's GIF / numbered PNG frames / last-frame reference PNGs under the reference's
file names (sample.py:56-106, 385-396).
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

CANVAS = 256   # _change_np_img_size always draws on a 256 x 256 canvas (dataloader_sample.py:164)
MTV_IO_LOADER_WORKER = 1   # include/mtv_b200.h


def _need_cuda(t: torch.Tensor, what: str) -> None:
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise RuntimeError(f"moditalker_b200.chunkio.{what}: expects a CUDA tensor (there is no CPU fallback)")


def _stream(dev: torch.device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def lower_half_start(height: int, landmarks) -> int:
    """Row from which ``_crop_lower_half`` zeroes a frame: ``int(landmarks[33][1])`` with numpy's slice semantics
    (a negative value counts from the bottom; anything past the frame masks nothing)."""
    r = int(np.asarray(landmarks)[33][1].astype(int))
    if r < 0:
        r = max(height + r, 0)
    return min(r, height)


def prep_frames(frames: torch.Tensor, mask_rows: Optional[Sequence[int]] = None, resolution: int = 256,
                loader_worker: bool = False) -> torch.Tensor:
    """uint8 frames ``[T, H, W, 3]`` (RGB, as decoded) -> fp32 ``[1, 3, T, R, R]`` in [-1, 1]: what the sampling script
    feeds ``first_stage_model.extract``.  ``mask_rows`` (one ``lower_half_start`` per frame) selects the masked stream.

    ``loader_worker=True`` reproduces the resize as torch computes it inside a DataLoader worker process (one torch thread: its
    "vectorized" bilinear kernel at every size), which is how the shipped script runs the loader (num_workers = 4); the default
    reproduces a direct call in a multi-threaded process.  The two differ by at most one ulp of the 0..255 value and not at all
    when the interpolation weights are exactly representable (e.g. even source sizes at 256)."""
    _need_cuda(frames, "prep_frames")
    if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[-1] != 3:
        raise ValueError("prep_frames: frames must be uint8 [T, H, W, 3]")
    frames = frames.contiguous()
    T, H, W, _ = frames.shape
    dev = frames.device
    rows = None
    if mask_rows is not None:
        if len(mask_rows) != T:
            raise ValueError("prep_frames: one mask row per frame")
        rows = torch.tensor([int(r) for r in mask_rows], dtype=torch.int32).to(dev, non_blocking=False)
    out = torch.empty((1, 3, T, resolution, resolution), dtype=torch.float32, device=dev)
    lib = _lib.load_library()
    rows_p = ctypes.c_void_p(rows.data_ptr()) if rows is not None else None
    if loader_worker:
        _lib.check(lib.mtv_io_prep_frames_ex(dev.index or 0, ctypes.c_void_p(frames.data_ptr()), T, H, W, rows_p, int(resolution),
                                             MTV_IO_LOADER_WORKER, ctypes.c_void_p(out.data_ptr()), _stream(dev)), "mtv_io_prep_frames_ex")
    else:
        _lib.check(lib.mtv_io_prep_frames(dev.index or 0, ctypes.c_void_p(frames.data_ptr()), T, H, W, rows_p,
                                          int(resolution), ctypes.c_void_p(out.data_ptr()), _stream(dev)), "mtv_io_prep_frames")
    return out


def rasterize_landmarks(landmarks: torch.Tensor, WH: int, flip: bool = False) -> torch.Tensor:
    """Landmark clip ``[T, N, 3]`` (normalised) or ``[T, N, 2]`` (pixels at size ``WH``), fp32 or fp64 as stored ->
    key-point video fp32 ``[1, 3, T, 256, 256]`` in {-1, +1}."""
    _need_cuda(landmarks, "rasterize_landmarks")
    if landmarks.dtype not in (torch.float32, torch.float64) or landmarks.dim() != 3 or landmarks.shape[-1] not in (2, 3):
        raise ValueError("rasterize_landmarks: landmarks must be fp32 / fp64 [T, N, 2 or 3]")
    landmarks = landmarks.contiguous()
    T, N, dims = landmarks.shape
    dev = landmarks.device
    out = torch.empty((1, 3, T, CANVAS, CANVAS), dtype=torch.float32, device=dev)
    lib = _lib.load_library()
    _lib.check(lib.mtv_io_rasterize_landmarks(dev.index or 0, ctypes.c_void_p(landmarks.data_ptr()),
                                              1 if landmarks.dtype == torch.float64 else 0, T, N, dims, int(WH), 1 if flip else 0,
                                              ctypes.c_void_p(out.data_ptr()), _stream(dev)), "mtv_io_rasterize_landmarks")
    return out


def frames_out(decoded: torch.Tensor, batch_size: int, repeat: int = 16, want_frames: bool = True,
               want_reference: bool = True) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor], Optional[torch.Tensor]]:
    """``decode_from_sample`` output fp32 ``[(B T), 3, H, W]`` -> ``(frames_u8 [B, T, H, W, 3], last_u8 [B, H, W, 3],
    next_ref [B, 3, repeat, H, W])``: the video frames the script writes, the pixels of the last-frame PNG, and that PNG as
    the script reads it back for the next chunk (ready for ``extract``)."""
    _need_cuda(decoded, "frames_out")
    if decoded.dtype != torch.float32 or decoded.dim() != 4 or decoded.shape[1] != 3 or decoded.shape[0] % batch_size:
        raise ValueError("frames_out: decoded must be fp32 [(B T), 3, H, W]")
    decoded = decoded.contiguous()
    BT, _, H, W = decoded.shape
    B, T = int(batch_size), BT // int(batch_size)
    dev = decoded.device
    frames = torch.empty((B, T, H, W, 3), dtype=torch.uint8, device=dev) if want_frames else None
    last = torch.empty((B, H, W, 3), dtype=torch.uint8, device=dev) if want_reference else None
    ref = torch.empty((B, 3, repeat, H, W), dtype=torch.float32, device=dev) if want_reference else None
    lib = _lib.load_library()
    p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
    _lib.check(lib.mtv_io_frames_out(dev.index or 0, p(decoded), B, T, H, W, p(frames), p(last), p(ref), int(repeat), _stream(dev)),
               "mtv_io_frames_out")
    return frames, last, ref


class AsyncFrameWriter:
    """Writes what the sampling script writes per chunk — the GIF (``save_image_grid``, MToV/sample.py:56-76), the numbered PNG
    frames (``save_image_at_folder``, sample.py:79-106) and the last-frame reference PNGs (sample.py:385-396) — without stalling
    the GPU loop: each call enqueues a device-to-host copy into pinned memory on a side stream and returns; a worker thread waits
    for the copy and encodes with PIL.  File names, grid layout (samples side by side, ``grid_size = (k, 1)``) and encoder
    parameters are the reference's.  Pixels are the uint8 tensors ``frames_out`` produced, so the reference's ``normalize``
    step (``rint((img - 0) * 255 / 255)``) is the identity and is skipped.  ``close()`` (or leaving the ``with`` block) waits for
    every file; the first encoder error is re-raised there."""

    def __init__(self, max_pending: int = 8):
        import queue
        import threading

        self._q = queue.Queue(maxsize=max_pending)
        self._err = None
        self._stream = None
        self._t = threading.Thread(target=self._run, name="mtv-frame-writer", daemon=True)
        self._t.start()

    # ---- reference file layouts -------------------------------------------------------------------------------------------
    @staticmethod
    def grid(frames_u8: torch.Tensor, landmarks_clip: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``[B, T, H, W, 3]`` uint8 -> ``[T, H, B * W, 3]``: the reshape / transpose of save_image_grid for ``grid_size = (B, 1)``.
        ``landmarks_clip`` (``x_l`` of sample.py:324, fp32 ``[B, 3, T, H, W]`` in [-1, 1]) is put in front as the script does with
        ``--including_ldmk_video`` (sample.py:403-409: ``(x * 255 + 255) / 2``, then rint by the writer)."""
        cols = frames_u8
        if landmarks_clip is not None:
            lm = ((landmarks_clip.float() * 255 + 255) / 2).round().clamp(0, 255).to(torch.uint8)     # b c t h w
            cols = torch.cat([lm.permute(0, 2, 3, 4, 1).to(frames_u8.device), frames_u8], dim=0)
        B, T, H, W, C = cols.shape
        return cols.permute(1, 2, 0, 3, 4).reshape(T, H, B * W, C).contiguous()

    # ---- public calls (one per file group the script writes) --------------------------------------------------------------
    def save_gif(self, frames_u8: torch.Tensor, fname: str, landmarks_clip: Optional[torch.Tensor] = None) -> None:
        """sample.py:411-416: ``generated_{it}.gif`` is written as ``generated_gif_{it}.gif``, 100 ms per frame, looping."""
        fname = fname.replace("generated", "generated_gif")
        self._submit(self.grid(frames_u8, landmarks_clip), ("gif", fname))

    def save_frames(self, start_iter: int, frames_u8: torch.Tensor, folder: str, landmarks_clip: Optional[torch.Tensor] = None) -> None:
        """sample.py:418-424: one PNG per frame, named ``str(start_iter + i).zfill(4) + ".png"``."""
        self._submit(self.grid(frames_u8, landmarks_clip), ("frames", folder, int(start_iter)))

    def save_last_frames(self, last_u8: torch.Tensor, folder: str) -> None:
        """sample.py:385-396: ``{idx}.png`` per sample in ``references/{ldmk_end}``: the next chunk's reference frames."""
        self._submit(last_u8.contiguous(), ("last", folder))

    def close(self) -> None:
        if self._t is not None:
            self._q.put(None)
            self._t.join()
            self._t = None
        if self._err is not None:
            err, self._err = self._err, None
            raise err

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    # ---- plumbing ---------------------------------------------------------------------------------------------------------
    def _submit(self, t: torch.Tensor, job) -> None:
        if self._t is None:
            raise RuntimeError("AsyncFrameWriter is closed")
        if t.dtype != torch.uint8:
            raise ValueError("AsyncFrameWriter: expects the uint8 tensors frames_out returns")
        if t.is_cuda:
            if self._stream is None:
                self._stream = torch.cuda.Stream(device=t.device)
            host = torch.empty(t.shape, dtype=torch.uint8, pin_memory=True)
            self._stream.wait_stream(torch.cuda.current_stream(t.device))
            with torch.cuda.stream(self._stream):
                host.copy_(t, non_blocking=True)
                t.record_stream(self._stream)
                ev = torch.cuda.Event()
                ev.record(self._stream)
        else:
            host, ev = t.clone(), None
        self._q.put((host, ev, job))

    def _run(self) -> None:
        import os

        import PIL.Image

        while True:
            item = self._q.get()
            if item is None:
                return
            host, ev, job = item
            try:
                if ev is not None:
                    ev.synchronize()
                a = host.numpy()
                if job[0] == "gif":
                    imgs = [PIL.Image.fromarray(a[i], "RGB") for i in range(len(a))]
                    imgs[0].save(job[1], quality=95, save_all=True, append_images=imgs[1:], duration=100, loop=0)
                elif job[0] == "frames":
                    os.makedirs(job[1], exist_ok=True)
                    for i in range(len(a)):
                        PIL.Image.fromarray(a[i], "RGB").save(os.path.join(job[1], f"{job[2] + i}".zfill(4) + ".png"))
                else:
                    os.makedirs(job[1], exist_ok=True)
                    for i in range(len(a)):
                        PIL.Image.fromarray(a[i], "RGB").save(os.path.join(job[1], f"{i}.png"))
            except Exception as e:      # surfaced by close()
                if self._err is None:
                    self._err = e
