import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

# Parity bar of BASELINE.json's north_star: <= 1e-3 relative (fp32).  Two metrics,
# both must hold (element-wise relative error is meaningless: outputs cross zero).
TOL_REL_L2 = 1e-3
TOL_MAX_ABS = 1e-3


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def rel_l2(a, b):
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def max_abs_rel(a, b):
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def assert_close(a, b, what, tol_l2=TOL_REL_L2, tol_max=TOL_MAX_ABS):
    assert tuple(a.shape) == tuple(b.shape), f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    assert bool(torch.isfinite(torch.as_tensor(a)).all()), f"{what}: non-finite values"
    e2, em = rel_l2(a, b), max_abs_rel(a, b)
    assert e2 <= tol_l2 and em <= tol_max, f"{what}: rel-L2 {e2:.3e} (tol {tol_l2:g}), max-abs/max-ref {em:.3e} (tol {tol_max:g})"
    return e2, em


def load_golden(name):
    p = os.path.join(GOLDEN, name + ".npz")
    if not os.path.exists(p):
        pytest.skip(f"fixture {name} missing")
    z = np.load(p, allow_pickle=False)
    return {k: z[k] for k in z.files}


def config_by_name(name):
    from moditalker_b200.arch import BASE_UNET_CONFIG, LONGVID_UNET_CONFIG, TINY_UNET_CONFIG
    return {"tiny": TINY_UNET_CONFIG, "base": BASE_UNET_CONFIG, "longvid": LONGVID_UNET_CONFIG}[str(name)]
