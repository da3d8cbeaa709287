"""GPU parity of the DataLoader-worker variant of the frame resize (mtv_io_prep_frames_ex, flag MTV_IO_LOADER_WORKER).

The shipped script runs its loader in DataLoader workers, where torch has one thread and resizes 3-channel frames with its
"vectorized" CPU kernel at every output size; `chunkio.prep_frames(..., loader_worker=True)` reproduces that kernel bit for bit.
Checked against the fixture the reference's own resize_crop produced under torch.set_num_threads(1) and against the oracle at
odd sizes on both sides of the 64-pixel switch of the default form.  (This file sorts last on purpose: it covers an optional
variant, and a failure here must not hide the rest of the suite under `pytest -x`.)"""
import ctypes
import os

import numpy as np
import pytest
import torch

from moditalker_b200 import _lib, chunkio
from oracle import chunkio_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_worker_fixture():
    d = np.load(os.path.join(GOLD, "chunkio_prep_worker.npz"))
    fr = d["frames"]
    rows = [chunkio.lower_half_start(fr.shape[1], d["kpts"][t]) for t in range(fr.shape[0])]
    got = chunkio.prep_frames(torch.from_numpy(fr).to(DEV), rows, int(d["R"]), loader_worker=True)[0].cpu().numpy()
    assert np.array_equal(got, d["out"])
    default = chunkio.prep_frames(torch.from_numpy(fr).to(DEV), rows, int(d["R"]))[0].cpu().numpy()
    assert not np.array_equal(default, d["out"]) and np.abs(default - d["out"]).max() < 4e-7      # the other torch kernel: last-bit differences


@pytest.mark.parametrize("H,W,R,masked", [(633, 641, 256, True), (301, 287, 256, False), (97, 97, 64, False), (97, 97, 68, True), (634, 634, 256, False)])
def test_worker_form_matches_oracle(H, W, R, masked):
    rng = np.random.default_rng(H + W + R)
    T = 16 if H * W > 100000 else 5
    fr = rng.integers(0, 256, size=(T, H, W, 3), dtype=np.uint8)
    rows = [int(r) for r in rng.integers(0, H + 1, size=T)] if masked else None
    got = chunkio.prep_frames(torch.from_numpy(fr).to(DEV), rows, R, loader_worker=True)[0].cpu().numpy()
    assert np.array_equal(got, O.prep_frames(fr, rows, R, loader_worker=True))
    if (H, W) == (634, 634):       # exactly representable weights: both torch kernels, hence both forms, agree
        assert np.array_equal(got, chunkio.prep_frames(torch.from_numpy(fr).to(DEV), rows, R)[0].cpu().numpy())


def test_unknown_flag_bits_are_rejected():
    lib = _lib.load_library()
    buf = torch.zeros(1 << 12, dtype=torch.uint8, device=DEV)
    p = ctypes.c_void_p(buf.data_ptr())
    assert lib.mtv_io_prep_frames_ex(0, p, 1, 8, 8, None, 8, 2, p, None) != 0 and b"flag" in lib.mtv_last_error()
