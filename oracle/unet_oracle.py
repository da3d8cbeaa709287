"""CPU ORACLE — test infrastructure only, never the product path.

A functional restatement (torch CPU ops over a flat state dict, no nn.Module)
of the MToV denoising hot path, each function citing the reference lines it
follows.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this file.

Parity pin: the reference ships no tests or golden vectors for this path
(SURVEY.md §4, §8c).  The oracle is therefore pinned against outputs of the
reference's OWN modules run in the build container (``oracle/make_golden.py``
imports /root/reference/MToV and writes ``tests/golden/*.npz``);
``tests/test_oracle_golden.py`` checks oracle == those fixtures.

All arithmetic below is what the reference delegates to PyTorch ATen
(conv2d / group_norm / softmax / einsum); running it in float64 gives the
tighter truth used for tolerance budgeting.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from moditalker_b200.arch import AttnSpec, ResSpec, Stage, UNetArch, build_arch


# --------------------------------------------------------------------------- primitives
def timestep_embedding(t: torch.Tensor, dim: int, dtype) -> torch.Tensor:
    """models/ddpm/diffusionmodules.py:108-128 — [cos(t f) | sin(t f)],
    f_i = exp(-ln(1e4) i / half); frequencies are built in fp32 like the
    reference, then promoted."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000) * torch.arange(0, half, dtype=torch.float32) / half).to(t.device)
    args = t[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    return emb.to(dtype)


def group_norm32(x, w, b):
    """diffusionmodules.py:156-173 — GroupNorm(32, C), eps 1e-5, affine."""
    return F.group_norm(x, 32, w, b, eps=1e-5)


def qkv_attention_legacy(qkv: torch.Tensor, n_heads: int) -> torch.Tensor:
    """unet.py:312-326 — heads split BEFORE q/k/v (channel layout
    [h0: q k v | h1: q k v | ...]), both q and k scaled by ch^-1/4, softmax in
    fp32 (or wider)."""
    bs, width, length = qkv.shape
    ch = width // (3 * n_heads)
    q, k, v = qkv.reshape(bs * n_heads, ch * 3, length).split(ch, dim=1)
    scale = 1 / math.sqrt(math.sqrt(ch))
    w = torch.einsum("bct,bcs->bts", q * scale, k * scale)
    w = torch.softmax(w, dim=-1)
    a = torch.einsum("bts,bcs->bct", w, v)
    return a.reshape(bs, -1, length)


class Oracle:
    def __init__(self, config: dict, state_dict: Dict[str, torch.Tensor], dtype=torch.float32, device="cpu"):
        """``device`` is "cpu" for the oracle proper; bench.py's gpu_eager_baseline fallback moves the same ATen ops to the
        GPU when the staged reference (oracle/_ref) is absent."""
        self.arch: UNetArch = build_arch(**config)
        strip = "diffusion_model."
        self.dtype = dtype
        self.device = torch.device(device)
        self.p = {(k[len(strip):] if k.startswith(strip) else k): v.detach().to(self.device, dtype)
                  for k, v in state_dict.items()}
        self.taps: Dict[str, torch.Tensor] = {}   # optional intermediate captures
        self.capture: Optional[set] = None

    # ----------------------------------------------------------------------- blocks
    def _w(self, name):
        return self.p[name + ".weight"], self.p[name + ".bias"]

    def res_block(self, r: ResSpec, x: torch.Tensor, emb: torch.Tensor) -> torch.Tensor:
        """unet.py:178-207 (use_scale_shift_norm=True, dropout p=0)."""
        p = r.name
        h = F.silu(group_norm32(x, *self._w(p + ".in_layers.0")))
        if r.updown == "up":      # unet.py:549-554 nearest x2 on both h and x
            h = F.interpolate(h, scale_factor=2, mode="nearest")
            x = F.interpolate(x, scale_factor=2, mode="nearest")
        elif r.updown == "down":  # unet.py:589-594 AvgPool2d(2, 2)
            h = F.avg_pool2d(h, 2, 2)
            x = F.avg_pool2d(x, 2, 2)
        h = F.conv2d(h, *self._w(p + ".in_layers.2"), padding=1)
        eo = F.linear(F.silu(emb), *self._w(p + ".emb_layers.1"))[:, :, None, None]
        scale, shift = torch.chunk(eo, 2, dim=1)
        h = group_norm32(h, *self._w(p + ".out_layers.0")) * (1 + scale) + shift
        h = F.conv2d(F.silu(h), *self._w(p + ".out_layers.3"), padding=1)
        if r.has_skip_conv:
            x = F.conv2d(x, *self._w(p + ".skip_connection"))
        return x + h

    def attn_block(self, a: AttnSpec, x: torch.Tensor) -> torch.Tensor:
        """unet.py:248-254 / 295-300: x + proj(attn(qkv(GN(x)))) over flattened tokens."""
        shp = x.shape
        x = x.reshape(shp[0], shp[1], -1)
        p = a.name
        qkv = F.conv1d(group_norm32(x, *self._w(p + ".norm")), *self._w(p + ".qkv"))
        h = qkv_attention_legacy(qkv, a.heads)
        h = F.conv1d(h, *self._w(p + ".proj_out"))
        return (x + h).reshape(shp)

    def _run_layers(self, st: Stage, planes: List[torch.Tensor], emb) -> List[torch.Tensor]:
        out = []
        for h in planes:   # same weights, three separate calls (unet.py:1032-1034)
            for l in st.layers:
                h = self.res_block(l, h, emb) if isinstance(l, ResSpec) else self.attn_block(l, h)
            out.append(h)
        return out

    def _joint(self, st: Stage, planes: List[torch.Tensor]) -> List[torch.Tensor]:
        """unet.py:1036-1049: flatten, concat xy|yt|xt, AttentionBlock1D, split back."""
        if st.joint is None:
            return planes
        shapes = [p.shape for p in planes]
        flat = torch.cat([p.reshape(p.shape[0], p.shape[1], -1) for p in planes], dim=-1)
        flat = self.attn_block(st.joint, flat)
        outs, o = [], 0
        for s in shapes:
            n = s[2] * s[3]
            outs.append(flat[:, :, o:o + n].reshape(s))
            o += n
        return outs

    def _tap(self, name, planes):
        if self.capture is not None and name in self.capture:
            self.taps[name] = torch.cat([p.reshape(p.shape[0], p.shape[1], -1) for p in planes], dim=-1).clone()

    # ----------------------------------------------------------------------- forward
    @torch.no_grad()
    def forward(self, x, cond, image_cond, t) -> torch.Tensor:
        """UNetModel.forward, unet.py:995-1117 (DiffusionWrapper.forward,
        unet.py:41-44, only forwards its arguments)."""
        A, dt = self.arch, self.dtype
        x, cond, image_cond = x.to(dt), cond.to(dt), image_cond.to(dt)
        B = x.shape[0]
        emb = timestep_embedding(t, A.model_channels, dt)
        emb = F.linear(emb, *self._w("time_embed.0"))
        emb = F.linear(F.silu(emb), *self._w("time_embed.2"))
        # unet.py:1022-1025: image_cond keeps its xy plane, yt/xt are zero-filled
        ic = torch.cat([image_cond[:, :, :1024], torch.zeros(B, cond.shape[1] // 2, 1024, dtype=dt, device=x.device)], dim=2)
        h = torch.cat([x, cond, ic], dim=1)
        planes = [
            h[:, :, 0:1024].reshape(B, -1, 32, 32),
            h[:, :, 1024:1536].reshape(B, -1, 16, 32),
            h[:, :, 1536:2048].reshape(B, -1, 16, 32),
        ]
        skips: List[List[torch.Tensor]] = []
        for i, st in enumerate(A.input_stages):
            if i == 0:
                planes = [F.conv2d(p, *self._w("input_blocks.0.0"), padding=1) for p in planes]
            else:
                planes = self._run_layers(st, planes, emb)
                planes = self._joint(st, planes)
            self._tap(f"in{i}", planes)
            skips.append(planes)
        planes = self._run_layers(A.middle, planes, emb)
        planes = self._joint(A.middle, planes)
        self._tap("mid", planes)
        for i, st in enumerate(A.output_stages):
            sk = skips.pop()
            planes = [torch.cat([p, s], dim=1) for p, s in zip(planes, sk)]  # unet.py:1080-1086
            planes = self._run_layers(st, planes, emb)
            planes = self._joint(st, planes)
            self._tap(f"out{i}", planes)
        outs = []
        for p in planes:  # unet.py:971-975, 1103-1109
            p = F.silu(group_norm32(p, *self._w("out.0")))
            p = F.conv2d(p, *self._w("out.2"), padding=1)
            outs.append(p.reshape(B, p.shape[1], -1))
        return torch.cat(outs, dim=-1)


# --------------------------------------------------------------------------- sampler
def schedule(timesteps=1000, linear_start=0.0015, linear_end=0.0195) -> Dict[str, torch.Tensor]:
    """losses/ddpm.py:79-81, 214-235 — 'linear' schedule (linear in sqrt(beta)),
    float64 numpy cumprod, every buffer rounded to fp32 from the float64 value."""
    import numpy as np
    betas = (torch.linspace(linear_start ** 0.5, linear_end ** 0.5, timesteps, dtype=torch.float64) ** 2).numpy()
    ac = np.cumprod(1.0 - betas, axis=0)
    f32 = lambda a: torch.tensor(a, dtype=torch.float32)
    return {
        "alphas_cumprod": f32(ac),
        "sqrt_alphas_cumprod": f32(np.sqrt(ac)),
        "sqrt_one_minus_alphas_cumprod": f32(np.sqrt(1.0 - ac)),
        "sqrt_recip_alphas_cumprod": f32(np.sqrt(1.0 / ac)),
        "sqrt_recipm1_alphas_cumprod": f32(np.sqrt(1.0 / ac - 1)),
    }


def ddim_time_pairs(total=1000, sampling=100):
    """losses/ddpm.py:372-376."""
    times = torch.linspace(-1, total - 1, steps=sampling + 1)
    times = list(reversed(times.int().tolist()))
    return list(zip(times[:-1], times[1:]))


def ddim_update(img, eps, noise, sch: Dict[str, torch.Tensor], time: int, time_next: int, eta: float = 1.0):
    """losses/ddpm.py:278-282, 346-351, 386-398 in fp32 tensor arithmetic."""
    x0 = sch["sqrt_recip_alphas_cumprod"][time] * img - sch["sqrt_recipm1_alphas_cumprod"][time] * eps
    x0 = x0.clamp_(-1.0, 1.0)
    if time_next < 0:
        return x0
    a, an = sch["alphas_cumprod"][time], sch["alphas_cumprod"][time_next]
    sigma = eta * ((1 - a / an) * (1 - an) / (1 - a)).sqrt()
    c = (1 - an - sigma ** 2).sqrt()
    return x0 * an.sqrt() + c * eps + sigma * noise


def ddim_sample(oracle: Oracle, cond, image_cond, noises: List[torch.Tensor], sampling: int,
                x_start=None, ratio: Optional[float] = None, eta: float = 1.0):
    """DDPM.ddim_sample (ddpm.py:363-404) when ``x_start`` is None, else
    ddim_sample_noised_start (ddpm.py:407-454).  ``noises[0]`` is the initial
    draw (img, or the q_sample noise); ``noises[1:]`` one per non-final step."""
    sch = schedule()
    pairs = ddim_time_pairs(1000, sampling)
    B = cond.shape[0]
    if x_start is None:
        img = noises[0].clone()
    else:
        t = int(1000 * ratio)   # q_sample, ddpm.py:422, 486-491
        img = sch["sqrt_alphas_cumprod"][t] * x_start + sch["sqrt_one_minus_alphas_cumprod"][t] * noises[0]
        pairs = pairs[int(len(pairs) * (1 - ratio)):]
    k = 1
    for time, time_next in pairs:
        tc = torch.full((B,), time, dtype=torch.long)
        eps = oracle.forward(img, cond, image_cond, tc).float()
        nz = None
        if time_next >= 0:
            nz = noises[k]
            k += 1
        img = ddim_update(img, eps, nz, sch, time, time_next, eta)
    return img
