"""GPU parity of the chunk I/O kernels (moditalker_b200.chunkio -> libmtv_b200.so) — SURVEY §8(f)3.

Bit-exact against (1) the fixtures generated from the reference's own functions, (2) the numpy oracle on seeded inputs at the
sizes the shipped pipeline uses (16 frames, 634 x 634 sources, 256 x 256 model frames), and through size-independent properties
(idempotence of the PNG round trip, masking of exactly the rows numpy masks, disc count)."""
import ctypes
import glob
import os

import numpy as np
import pytest
import torch

from moditalker_b200 import _lib, chunkio
from oracle import chunkio_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLD = os.path.join(os.path.dirname(__file__), "golden")
# (the DataLoader-worker variant of the resize has its own file, tests/test_zz_loader_worker_gpu.py)
PREP = sorted(n for n in (os.path.basename(p)[8:-4] for p in glob.glob(os.path.join(GOLD, "chunkio_prep_*.npz"))) if "worker" not in n)
LM = sorted(os.path.basename(p)[8:-4] for p in glob.glob(os.path.join(GOLD, "chunkio_lm_*.npz")))


def load(name):
    return np.load(os.path.join(GOLD, f"chunkio_{name}.npz"))


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.mark.parametrize("name", PREP)
def test_prep_frames_matches_reference_fixture(name):
    d = load(name)
    fr = d["frames"]
    T, H = fr.shape[0], fr.shape[1]
    rows = [chunkio.lower_half_start(H, d["kpts"][t]) for t in range(T)] if int(d["masked"]) else None
    out = chunkio.prep_frames(dev(fr), rows, int(d["R"]))
    assert out.shape == (1,) + d["out"].shape and out.dtype == torch.float32
    assert np.array_equal(out[0].cpu().numpy(), d["out"])


@pytest.mark.parametrize("H,W,R,masked", [(634, 634, 256, False), (634, 634, 256, True), (256, 256, 256, True), (360, 640, 256, False),
                                           (726, 726, 128, True), (633, 641, 256, True), (301, 287, 256, False), (97, 97, 64, False),
                                           (97, 97, 68, True)])
def test_prep_frames_matches_oracle_at_pipeline_sizes(H, W, R, masked):
    # odd source sizes make the interpolation weights inexact in fp32, so the two rounding regimes of torch's kernel (outputs up
    # to 64 wide / wider) are really exercised; 634 / 726 / 360 -> 256 or 128 have exact weights
    rng = np.random.default_rng(H * 7 + W + R + int(masked))
    T = 16 if H * W > 100000 else 5
    fr = rng.integers(0, 256, size=(T, H, W, 3), dtype=np.uint8)
    rows = [int(r) for r in rng.integers(0, H + 1, size=T)] if masked else None
    if masked:
        rows[0], rows[1] = 0, H                      # everything masked / nothing masked
    got = chunkio.prep_frames(dev(fr), rows, R)[0].cpu().numpy()
    want = O.prep_frames(fr, rows, R)
    assert np.array_equal(got, want)
    if masked:
        assert (got[:, 0] == -1.0).all()             # a fully masked frame is black
        if H == R and W == R:                        # no resampling: exactly the rows numpy zeroes are black, the rest untouched
            for t in range(T):
                assert (got[:, t, rows[t]:] == -1.0).all()
                assert np.array_equal(got[:, t, :rows[t]], (fr[t, :rows[t]].astype(np.float32) / np.float32(127.5) - np.float32(1)).transpose(2, 0, 1))


@pytest.mark.parametrize("name", LM)
def test_rasterize_landmarks_matches_reference_fixture(name):
    d = load(name)
    out = chunkio.rasterize_landmarks(dev(d["lm"]), int(d["WH"]), bool(int(d["flip"])))
    T = d["lm"].shape[0]
    assert out.shape == (1, 3, T, 256, 256)
    o = out[0].cpu().numpy()
    assert set(np.unique(o)) <= {-1.0, 1.0}
    assert np.array_equal(o[0], o[1]) and np.array_equal(o[0], o[2])
    assert np.array_equal(np.packbits(o[0] > 0), d["canvas_bits"])


@pytest.mark.parametrize("dtype,dims,WH,flip", [(np.float32, 3, 634, False), (np.float64, 3, 634, True), (np.float64, 2, 726, False),
                                                (np.float32, 2, 256, True)])
def test_rasterize_landmarks_matches_oracle_at_pipeline_sizes(dtype, dims, WH, flip):
    rng = np.random.default_rng(dims * 1000 + WH + int(flip))
    T, N = 16, 478                                   # a 16-frame clip of a dense face mesh
    lm = rng.uniform(-1.1, 1.1, size=(T, N, 3)) if dims == 3 else rng.uniform(-20, WH + 20, size=(T, N, 2))
    lm = lm.astype(dtype)
    got = chunkio.rasterize_landmarks(dev(lm), WH, flip)[0].cpu().numpy()
    want = O.rasterize_landmarks(lm, WH, flip)
    assert np.array_equal(got, want)
    # one isolated on-canvas landmark paints exactly the 29 pixels of OpenCV's radius-3 disc
    one = np.zeros((1, 1, 3), dtype=dtype) if dims == 3 else np.full((1, 1, 2), WH / 2, dtype=dtype)
    assert int((chunkio.rasterize_landmarks(dev(one), WH, flip) > 0).sum().item()) == 3 * 29
    # no landmarks at all: a black canvas
    assert (chunkio.rasterize_landmarks(torch.zeros((2, 0, dims), dtype=torch.float32, device=DEV), WH) == -1).all()


def test_frames_out_matches_reference_fixture():
    d = load("frames_out")
    frames, last, ref = chunkio.frames_out(dev(d["dec"]), int(d["B"]), 16)
    assert np.array_equal(frames.cpu().numpy(), d["frames_u8"])
    assert np.array_equal(last.cpu().numpy(), d["last_u8"])
    assert np.array_equal(ref.cpu().numpy(), d["next_ref"])


@pytest.mark.parametrize("B,T,H,W", [(1, 16, 256, 256), (4, 16, 256, 256), (2, 16, 128, 128), (3, 5, 40, 52)])
def test_frames_out_matches_oracle_and_round_trips(B, T, H, W):
    rng = np.random.default_rng(B * 100 + T + H)
    dec = rng.uniform(-1.15, 1.15, size=(B * T, 3, H, W)).astype(np.float32)
    frames, last, ref = chunkio.frames_out(dev(dec), B, 16)
    f0, l0, r0 = O.frames_out(dec, B, 16)
    assert np.array_equal(frames.cpu().numpy(), f0)
    assert np.array_equal(last.cpu().numpy(), l0)
    assert np.array_equal(ref.cpu().numpy(), r0)
    # properties that hold at any size: all 16 reference frames are the same image; feeding the read-back reference through the
    # output stage again reproduces the PNG's pixels (the round trip is idempotent); truncation never exceeds rounding
    assert (ref == ref[:, :, :1]).all()
    again = ref[:, :, 0].contiguous()                                    # [B, 3, H, W] as a one-frame clip
    _, last2, _ = chunkio.frames_out(again, B, 1)
    assert torch.equal(last2, last)
    fl = frames[:, -1].to(torch.int16)
    assert ((last.to(torch.int16) - fl) >= 0).all() and ((last.to(torch.int16) - fl) <= 1).all()
    # optional outputs
    only_frames, none_last, none_ref = chunkio.frames_out(dev(dec), B, 16, want_reference=False)
    assert none_last is None and none_ref is None and torch.equal(only_frames, frames)


def test_results_do_not_depend_on_the_stream_or_on_repeats():
    rng = np.random.default_rng(5)
    fr = dev(rng.integers(0, 256, size=(16, 300, 280, 3), dtype=np.uint8))
    a = chunkio.prep_frames(fr, None, 256)
    s = torch.cuda.Stream(device=DEV)
    s.wait_stream(torch.cuda.current_stream(DEV))
    with torch.cuda.stream(s):
        b = chunkio.prep_frames(fr, None, 256)
    s.synchronize()
    assert torch.equal(a, b) and torch.equal(a, chunkio.prep_frames(fr, None, 256))


def test_async_writer_from_device_tensors(tmp_path):
    """frames_out -> AsyncFrameWriter: pinned copies on a side stream, files written by the worker thread hold exactly the
    device pixels (checked after close(), while the producing stream has long moved on)."""
    import PIL.Image

    rng = np.random.default_rng(9)
    B, T, H, W = 2, 16, 64, 64
    with chunkio.AsyncFrameWriter() as w:
        keep = []
        for it in range(3):
            dec = dev(rng.uniform(-1.1, 1.1, size=(B * T, 3, H, W)).astype(np.float32))
            frames, last, _ = chunkio.frames_out(dec, B, 16)
            w.save_gif(frames, str(tmp_path / f"generated_{it}.gif"))
            w.save_frames(16 * it, frames, str(tmp_path / "frames"))
            w.save_last_frames(last, str(tmp_path / "references" / str(16 * (it + 1))))
            keep.append((frames.cpu().numpy(), last.cpu().numpy()))
            del dec, frames, last                                        # the allocator may reuse the blocks: record_stream guards the copies
    assert len(list((tmp_path / "frames").iterdir())) == 48
    for it, (fr, la) in enumerate(keep):
        for t in (0, 7, 15):
            img = np.asarray(PIL.Image.open(tmp_path / "frames" / (f"{16 * it + t}".zfill(4) + ".png")))
            assert np.array_equal(img, np.concatenate([fr[0, t], fr[1, t]], axis=1))
        for i in range(B):
            assert np.array_equal(np.asarray(PIL.Image.open(tmp_path / "references" / str(16 * (it + 1)) / f"{i}.png")), la[i])
        assert PIL.Image.open(tmp_path / f"generated_gif_{it}.gif").n_frames == T


def test_c_abi_argument_errors():
    lib = _lib.load_library()
    buf = torch.zeros(1 << 16, dtype=torch.uint8, device=DEV)
    p = ctypes.c_void_p(buf.data_ptr())
    assert lib.mtv_io_prep_frames(0, None, 1, 8, 8, None, 8, p, None) != 0 and b"null" in lib.mtv_last_error()
    assert lib.mtv_io_prep_frames(0, p, 1, 8, 8, None, 6, p, None) != 0 and b"multiple of 4" in lib.mtv_last_error()
    assert lib.mtv_io_rasterize_landmarks(0, p, 0, 1, 4, 4, 256, 0, p, None) != 0 and b"dims" in lib.mtv_last_error()
    assert lib.mtv_io_frames_out(0, p, 1, 1, 8, 6, p, None, None, 1, None) != 0 and b"multiple of 4" in lib.mtv_last_error()
    assert lib.mtv_io_frames_out(0, p, 1, 1, 8, 8, None, None, p, 0, None) != 0 and b"Trep" in lib.mtv_last_error()
    with pytest.raises(ValueError):
        chunkio.prep_frames(torch.zeros(2, 8, 8, 3, device=DEV), None, 8)        # not uint8
    with pytest.raises(ValueError):
        chunkio.prep_frames(torch.zeros(2, 8, 8, 3, dtype=torch.uint8, device=DEV), [1], 8)
    with pytest.raises(ValueError):
        chunkio.frames_out(torch.zeros(5, 3, 8, 8, device=DEV), 2)
