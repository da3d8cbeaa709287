"""CPU tests of the chunk I/O oracle (oracle/chunkio_oracle.py) against the fixtures the REFERENCE's own functions produced
(oracle/make_golden_chunkio.py: EvalDataset._load_img_from_path / _crop_lower_half / _change_np_img_size,
data_utils.resize_crop, and the literal cv2 / PIL calls of MToV/sample.py:344-399), plus the host logic of
moditalker_b200.chunkio that needs no GPU.  Everything here is bit-exact."""
import glob
import os

import numpy as np
import pytest
import torch

from moditalker_b200 import chunkio
from oracle import chunkio_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")
PREP = sorted(os.path.basename(p)[8:-4] for p in glob.glob(os.path.join(GOLD, "chunkio_prep_*.npz")))
LM = sorted(os.path.basename(p)[8:-4] for p in glob.glob(os.path.join(GOLD, "chunkio_lm_*.npz")))


def load(name):
    return np.load(os.path.join(GOLD, f"chunkio_{name}.npz"))


def test_fixture_inventory():
    assert {"prep_plain", "prep_masked", "prep_identity", "prep_up", "prep_frac", "prep_wide", "prep_wide_down", "prep_worker"} <= set(PREP)
    assert {"lm_norm_f32", "lm_norm_f64", "lm_pixel_f64"} <= set(LM)


def test_to_tensor_times_255_is_the_identity_on_bytes():
    # dataloader_sample.py:144 computes ToTensor()(img) * 255 in fp32; the kernel loads the byte directly
    k = np.arange(256, dtype=np.uint8)
    assert np.array_equal(O.load_255(k), k.astype(np.float32))
    assert np.array_equal((torch.from_numpy(k).float().div(255) * 255).numpy(), k.astype(np.float32))


@pytest.mark.parametrize("name", PREP)
def test_prep_frames_matches_reference_fixture(name):
    d = load(name)
    fr = d["frames"]
    T, H = fr.shape[0], fr.shape[1]
    rows = [O.lower_half_start(H, d["kpts"][t]) for t in range(T)] if int(d["masked"]) else None
    out = O.prep_frames(fr, rows, int(d["R"]), loader_worker=bool(int(d["worker"])))
    assert out.dtype == np.float32 and out.shape == d["out"].shape
    assert np.array_equal(out, d["out"])
    if name == "prep_worker":                  # the two torch kernels really differ on this case: the flag is not a no-op
        assert not np.array_equal(O.prep_frames(fr, rows, int(d["R"]), loader_worker=False), d["out"])


def test_bilinear_matches_live_torch_cpu_over_random_sizes():
    """torch's CPU bilinear kernel rounds differently for outputs up to 64 pixels wide and for wider ones; the oracle reproduces
    both bit for bit.  Checked here against the installed torch over random (odd, non-dyadic) sizes on either side of the switch.
    The fixtures pin the behaviour of the build that generated them; if the installed torch no longer reproduces a fixture
    (another build / ISA), this live comparison is skipped rather than failed."""
    import torch.nn.functional as F

    d = load("prep_wide_down")
    fr = d["frames"]
    S = min(fr.shape[1], fr.shape[2])
    y0, x0 = (fr.shape[1] - S) // 2 if fr.shape[1] > fr.shape[2] else 0, (fr.shape[2] - S) // 2 if fr.shape[2] >= fr.shape[1] else 0
    v = torch.from_numpy(fr).permute(0, 3, 1, 2).float()[:, :, y0:y0 + S, x0:x0 + S]
    live = (F.interpolate(v, size=int(d["R"]), mode="bilinear", align_corners=False) / 127.5 - 1).permute(1, 0, 2, 3).numpy()
    if not np.array_equal(live, d["out"]):
        pytest.skip("installed torch rounds its CPU bilinear kernel differently from the build that generated the fixtures")
    rng = np.random.default_rng(7)
    for S, R in [(37, 20), (59, 60), (358, 64), (35, 65), (101, 68), (333, 128), (211, 256), (10, 236), (64, 64), (129, 4)]:
        img = rng.integers(0, 256, size=(2, S, S, 3)).astype(np.float32)
        ref = F.interpolate(torch.from_numpy(img).permute(0, 3, 1, 2), size=R, mode="bilinear", align_corners=False)
        got = O.bilinear_resize(img, R)
        assert np.array_equal(got, ref.permute(0, 2, 3, 1).numpy()), (S, R)


def test_bilinear_worker_form_matches_live_torch_cpu_with_one_thread():
    """Inside a DataLoader worker torch has one thread and resizes 3-channel input with its vectorized kernel at every size
    (`loader_worker=True`).  Same live comparison as above under torch.set_num_threads(1); skipped if the installed torch does
    not reproduce the worker fixture."""
    import torch.nn.functional as F

    n = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        d = load("prep_worker")
        fr = d["frames"]
        H, W = fr.shape[1:3]
        S = min(H, W)
        y0, x0 = ((H - W) // 2 if H > W else 0), ((W - H) // 2 if W >= H else 0)
        rows = np.array([O.lower_half_start(H, d["kpts"][t]) for t in range(fr.shape[0])])
        v = np.where(np.arange(H)[None, :, None, None] < rows[:, None, None, None], fr, 0).astype(np.float32)
        t = torch.from_numpy(v).permute(0, 3, 1, 2)[:, :, y0:y0 + S, x0:x0 + S]
        live = (F.interpolate(t, size=int(d["R"]), mode="bilinear", align_corners=False) / 127.5 - 1).permute(1, 0, 2, 3).numpy()
        if not np.array_equal(live, d["out"]):
            pytest.skip("installed torch rounds its single-thread CPU bilinear kernel differently from the build that generated the fixtures")
        rng = np.random.default_rng(8)
        for S, R in [(37, 20), (101, 68), (333, 128), (211, 256), (10, 236)]:
            img = rng.integers(0, 256, size=(2, S, S, 3)).astype(np.float32)
            ref = F.interpolate(torch.from_numpy(img).permute(0, 3, 1, 2), size=R, mode="bilinear", align_corners=False)
            assert np.array_equal(O.bilinear_resize(img, R, vectorized=True), ref.permute(0, 2, 3, 1).numpy()), (S, R)
    finally:
        torch.set_num_threads(n)


def test_lower_half_start_follows_numpy_slicing():
    H = 90
    for y, want in [(49.9, 49), (-22.5, 68), (-200.0, 0), (97.5, 90), (0.0, 0), (-0.5, 0)]:
        k = np.zeros((68, 2))
        k[33, 1] = y
        mask = np.ones((H, 4))
        mask[(k[33][1]).astype(int):, :] = 0.0            # the reference's statement (dataloader_sample.py:135)
        first_zero = int(np.argmax(mask[:, 0] == 0)) if (mask[:, 0] == 0).any() else H
        assert O.lower_half_start(H, k) == first_zero == want
        assert chunkio.lower_half_start(H, k) == want      # the product's host helper agrees with the oracle


@pytest.mark.parametrize("name", LM)
def test_rasterize_landmarks_matches_reference_fixture(name):
    d = load(name)
    out = O.rasterize_landmarks(d["lm"], int(d["WH"]), bool(int(d["flip"])))
    assert out.shape == (3, d["lm"].shape[0], 256, 256)
    assert set(np.unique(out)) <= {-1.0, 1.0}
    assert np.array_equal(out[0], out[1]) and np.array_equal(out[0], out[2])
    assert np.array_equal(np.packbits(out[0] > 0), d["canvas_bits"])


def test_rasterize_empty_and_offscreen():
    out = O.rasterize_landmarks(np.zeros((2, 0, 3), dtype=np.float32), 256)
    assert (out == -1).all()
    far = np.array([[[50.0, -40.0, 0.0], [-30.0, 2.0, 0.0]]], dtype=np.float64)       # normalised coordinates far outside
    assert (O.rasterize_landmarks(far, 634) == -1).all()


def test_frames_out_matches_reference_fixture():
    d = load("frames_out")
    frames, last, ref = O.frames_out(d["dec"], int(d["B"]), 16)
    assert np.array_equal(frames, d["frames_u8"])
    assert np.array_equal(last, d["last_u8"])
    assert np.array_equal(ref, d["next_ref"])
    # planted values: the clamp edges, and two exact half-way cases that round to even (127.5 -> 128, 254.5 -> 254)
    assert last[-1, 0, :8, 0].tolist() == [0, 255, 0, 1, 2, 128, 254, 0]
    assert frames[-1, -1, 0, :8, 0].tolist() == [0, 255, 0, 1, 2, 127, 254, 0]


def test_product_refuses_cpu_tensors_and_bad_shapes():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        chunkio.prep_frames(torch.zeros(2, 8, 8, 3, dtype=torch.uint8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        chunkio.rasterize_landmarks(torch.zeros(2, 5, 3), 256)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        chunkio.frames_out(torch.zeros(4, 3, 8, 8), 2)


def _reference_grid(img_nctHW, grid_size):
    """The reshape / transpose of save_image_grid and save_image_at_folder (MToV/sample.py:63-67, 86-90), restated."""
    gw, gh = grid_size
    _N, C, T, H, W = img_nctHW.shape
    img = img_nctHW.reshape(gh, gw, C, T, H, W).transpose(3, 0, 4, 1, 5, 2)
    return img.reshape(T, gh * H, gw * W, C)


def test_async_writer_files_match_the_reference_layout(tmp_path):
    import PIL.Image

    rng = np.random.default_rng(3)
    k, T, H, W = 3, 5, 8, 12
    frames = torch.from_numpy(rng.integers(0, 256, size=(k, T, H, W, 3), dtype=np.uint8))            # frames_out layout: b t h w c
    x_l = torch.from_numpy(rng.choice([-1.0, 1.0], size=(k, 3, T, H, W)).astype(np.float32))          # sample.py:324
    fakes = frames.permute(0, 4, 1, 2, 3).numpy()                                                      # sample.py:400: b c t h w
    want = _reference_grid(fakes, (k, 1))
    assert np.array_equal(chunkio.AsyncFrameWriter.grid(frames).numpy(), want)
    # --including_ldmk_video (sample.py:403-409): the key-point clip in front, grid one column wider, then the writer's rint
    both = np.concatenate([((x_l.numpy() * 255 + 255) / 2), fakes.astype(np.float32)])
    want_lm = np.rint(_reference_grid(both, (2 * k, 1))).clip(0, 255).astype(np.uint8)
    assert np.array_equal(chunkio.AsyncFrameWriter.grid(frames, x_l).numpy(), want_lm)

    last = frames[:, -1].contiguous()
    with chunkio.AsyncFrameWriter() as w:
        w.save_gif(frames, str(tmp_path / "gif" / "generated_7.gif").replace("/gif/", "/"))
        w.save_frames(32, frames, str(tmp_path / "frames"))
        w.save_last_frames(last, str(tmp_path / "references" / "48"))
    assert sorted(p.name for p in (tmp_path / "frames").iterdir()) == [f"{32 + i}".zfill(4) + ".png" for i in range(T)]
    for i in range(T):
        assert np.array_equal(np.asarray(PIL.Image.open(tmp_path / "frames" / (f"{32 + i}".zfill(4) + ".png"))), want[i])
    for i in range(k):
        assert np.array_equal(np.asarray(PIL.Image.open(tmp_path / "references" / "48" / f"{i}.png")), last[i].numpy())
    gif = PIL.Image.open(tmp_path / "generated_gif_7.gif")                                             # sample.py:75: "generated" -> "generated_gif"
    assert gif.n_frames == T and gif.info["duration"] == 100 and gif.info["loop"] == 0 and gif.size == (k * W, H)


def test_async_writer_reports_errors_and_refuses_floats(tmp_path):
    w = chunkio.AsyncFrameWriter()
    with pytest.raises(ValueError):
        w.save_last_frames(torch.zeros(1, 4, 4, 3), str(tmp_path))
    blocker = tmp_path / "not_a_dir"
    blocker.write_text("x")
    w.save_last_frames(torch.zeros(1, 4, 4, 3, dtype=torch.uint8), str(blocker))      # makedirs fails in the worker
    with pytest.raises(Exception):
        w.close()
    with pytest.raises(RuntimeError, match="closed"):
        w.save_last_frames(torch.zeros(1, 4, 4, 3, dtype=torch.uint8), str(tmp_path))
