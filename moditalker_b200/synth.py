"""Deterministic synthetic weights and inputs, addressed by NAME.

No MToV diffusion checkpoint is published and the reference tree
is not present on the GPU box, so parity fixtures need weights that can be
re-created bit-identically anywhere from a (config, seed) pair, independent of
module construction order.  Every tensor is drawn from its own CPU generator
seeded by ``seed`` and a CRC of the tensor's state-dict key:

  * conv / linear weight, bias:  U(-1, 1) / sqrt(fan_in)   (== PyTorch's default
    kaiming_uniform(a=sqrt(5)) family; also applied to the reference's
    zero-initialised modules — ``out_layers.3``, ``proj_out``, ``out.2`` — since
    with those left at zero the whole forward is identically 0)
  * GroupNorm gamma: U(0.8, 1.2), beta: U(-0.2, 0.2)  (so affine terms matter)

Inputs follow SURVEY.md §8(d): x ~ N(0,1), cond / image_cond = tanh(N(0,1))
(autoencoder latents are post-tanh, models/autoencoder/autoencoder_vit.py:246-248).
"""
from __future__ import annotations

import zlib
from typing import Dict, Tuple

import torch

from .arch import UNetArch, build_arch, param_shapes


def _gen(seed: int, name: str) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((int(seed) * 1000003 + zlib.crc32(name.encode())) % (2**63 - 1))
    return g


def synth_tensor(name: str, shape: Tuple[int, ...], seed: int) -> torch.Tensor:
    g = _gen(seed, name)
    if name == "zeros":
        return torch.zeros(shape, dtype=torch.float32)
    is_norm = (len(shape) == 1) and (
        ".in_layers.0." in name or ".out_layers.0." in name or ".norm." in name or name.startswith("out.0.")
    )
    u = torch.rand(shape, generator=g, dtype=torch.float32)
    if is_norm:
        if name.endswith(".weight"):
            return 0.8 + 0.4 * u
        return -0.2 + 0.4 * u
    if name.endswith(".weight"):
        fan_in = 1
        for d in shape[1:]:
            fan_in *= d
    else:
        fan_in = _bias_fan_in(name, shape)
    return (2.0 * u - 1.0) / float(fan_in) ** 0.5


_FAN_CACHE: Dict[str, int] = {}


def _bias_fan_in(name: str, shape) -> int:
    # filled by synth_state_dict (a bias' fan_in is its weight's); fallback = len
    return _FAN_CACHE.get(name, shape[0])


def synth_state_dict(config: dict, seed: int = 0, prefix: str = "") -> Dict[str, torch.Tensor]:
    """State dict for ``UNetModel(**config)`` (or ``DiffusionWrapper`` with
    ``prefix='diffusion_model.'``).  Tensor values depend only on the
    un-prefixed key and the seed."""
    arch: UNetArch = build_arch(**config)
    shapes = param_shapes(arch, include_dead=True)
    by_name = dict(shapes)
    for n, s in shapes:
        if n.endswith(".bias"):
            w = by_name.get(n[:-5] + ".weight")
            if w is not None and len(w) > 1:
                f = 1
                for d in w[1:]:
                    f *= d
                _FAN_CACHE[n] = f
    return {prefix + n: synth_tensor(n, s, seed) for n, s in shapes}


def synth_inputs(batch: int, seed: int = 2, image_cond_len: int = 1024, t=500):
    """(x, cond, image_cond, t) for one UNet forward; sample b's tensors depend
    only on (seed, b), so a batch is the concatenation of its samples."""
    xs, cs, ics = [], [], []
    for b in range(batch):
        g = _gen(seed, f"input.{b}")
        xs.append(torch.randn((1, 4, 2048), generator=g))
        cs.append(torch.tanh(torch.randn((1, 8, 2048), generator=g)))
        ics.append(torch.tanh(torch.randn((1, 4, 2048), generator=g))[:, :, :image_cond_len])
    if isinstance(t, int):
        t = [t] * batch
    return (
        torch.cat(xs).contiguous(),
        torch.cat(cs).contiguous(),
        torch.cat(ics).contiguous(),
        torch.tensor(list(t), dtype=torch.long),
    )


def synth_noise(shape, seed: int, tag: str) -> torch.Tensor:
    return torch.randn(shape, generator=_gen(seed, tag))
