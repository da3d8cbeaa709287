"""Architecture walk, state-dict contract and the C-ABI surface (CPU)."""
import copy
import ctypes
import os
import re
import sys

import pytest
import torch

from conftest import ROOT
from moditalker_b200 import BASE_UNET_CONFIG, LONGVID_UNET_CONFIG, TINY_UNET_CONFIG, DiffusionWrapper, UNetModel, build_arch
from moditalker_b200.arch import param_shapes, plane_shapes, tokens_at
from moditalker_b200.synth import synth_inputs, synth_state_dict

REF = "/root/reference/MToV"


@pytest.mark.parametrize("cfg,nkeys,nres,nattn", [(BASE_UNET_CONFIG, 804, 28, 40), (LONGVID_UNET_CONFIG, 805, 28, 40),
                                                  (TINY_UNET_CONFIG, 292, 10, 15)])
def test_param_inventory(cfg, nkeys, nres, nattn):
    arch = build_arch(**cfg)
    shapes = param_shapes(arch)
    assert len(shapes) == nkeys                      # 804 keys for base.yaml
    assert len(arch.live_resblocks()) == nres       # 11 down + 2 mid + 15 up
    assert len(arch.live_attns()) == nattn          # 16 per-plane + 24 cross-plane
    names = [n for n, _ in shapes]
    assert len(set(names)) == len(names)


def test_module_state_dict_matches_walk():
    for cfg in (BASE_UNET_CONFIG, LONGVID_UNET_CONFIG, TINY_UNET_CONFIG):
        m = UNetModel(**cfg)
        got = [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
        assert got == param_shapes(build_arch(**cfg))


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")
def test_state_dict_matches_reference():
    sys.path.insert(0, REF)
    try:
        from models.ddpm.unet import UNetModel as RefUNet
    finally:
        sys.path.remove(REF)
    for cfg in (BASE_UNET_CONFIG, LONGVID_UNET_CONFIG):
        ref = [(k, tuple(v.shape)) for k, v in RefUNet(**cfg).state_dict().items()]
        assert ref == param_shapes(build_arch(**cfg))


def test_strict_load_and_deepcopy():
    w = DiffusionWrapper(UNetModel(**TINY_UNET_CONFIG))
    sd = synth_state_dict(TINY_UNET_CONFIG, 0, "diffusion_model.")
    assert set(sd) == set(w.state_dict())
    w.load_state_dict(sd, strict=True)
    bad = dict(sd); bad.pop("diffusion_model.out.2.bias")
    with pytest.raises(RuntimeError):
        w.load_state_dict(bad, strict=True)
    w2 = copy.deepcopy(w).eval()
    assert w2.diffusion_model._engine is not w.diffusion_model._engine
    for (ka, va), (kb, vb) in zip(w.state_dict().items(), w2.state_dict().items()):
        assert ka == kb and torch.equal(va, vb)
    assert w2.diffusion_model.in_channels == 4 and w2.diffusion_model.image_size == 32
    assert w2.diffusion_model.cond_model is False


def test_geometry():
    assert plane_shapes(0) == ((32, 32), (16, 32), (16, 32))
    assert [tokens_at(l) for l in range(4)] == [2048, 512, 128, 32]


def test_unsupported_options_raise():
    with pytest.raises(NotImplementedError):
        UNetModel(**dict(BASE_UNET_CONFIG, dims=3))
    with pytest.raises(NotImplementedError):
        UNetModel(**dict(BASE_UNET_CONFIG, use_spatial_transformer=True, context_dim=512))
    with pytest.raises(NotImplementedError):
        UNetModel(**dict(BASE_UNET_CONFIG, use_scale_shift_norm=False))


def test_forward_refuses_cpu_tensors():
    m = UNetModel(**TINY_UNET_CONFIG).eval()
    x, c, ic, t = synth_inputs(1)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(x, c, ic, t)


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    from moditalker_b200 import _lib
    monkeypatch.setenv("MTV_B200_LIB", str(tmp_path / "nope.so"))
    monkeypatch.setattr(_lib, "_lib", None)
    with pytest.raises(RuntimeError, match="no CPU / PyTorch fallback"):
        _lib.load_library()


def test_library_exports_every_declared_symbol():
    """The C-ABI library loads without a GPU and exports every function that
    include/mtv_b200.h declares (no compute calls here)."""
    from moditalker_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    hdr = open(os.path.join(ROOT, "include", "mtv_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(mtv_[a-z_0-9]+)\s*\(", hdr)))
    assert declared, "header parse failed"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for sym in declared:
        assert hasattr(lib, sym), f"libmtv_b200.so does not export {sym}"
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared
    lib.mtv_abi_version.restype = ctypes.c_int32
    assert lib.mtv_abi_version() == _lib.MTV_ABI_VERSION
