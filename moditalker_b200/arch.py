"""Architecture walk of the MToV tri-plane UNet (pure Python, no torch).

One description of the network shared by the parameter-holding ``nn.Module``
mirror (``unet.py``), the CPU oracle (``oracle/unet_oracle.py``) and the tests.
The C++ plan builder in ``csrc/mtv_plan.cu`` walks the same structure from the
same integers; ``tests/test_arch.py`` checks the two agree on the weight-name
list.

Follows the constructor of the reference ``UNetModel``
(/root/reference/MToV/models/ddpm/unet.py:631-975):
  * stem conv 16 -> model_channels (unet.py:710-717; the 16 is hard-coded:
    x(4) + cond(8) + image_cond(4), unet.py:1025)
  * per level ``num_res_blocks`` ResBlocks (+ per-plane AttentionBlock when the
    level's downsample rate is in ``attention_resolutions``), each followed by a
    cross-plane AttentionBlock1D (unet.py:724-778), then a down ResBlock between
    levels (unet.py:780-812)
  * middle: Res, Attn, Res + mid_attn (unet.py:822-855)
  * decoder mirrors it with skip concatenation and up ResBlocks (unet.py:863-969)
  * dead ``output_bg_*`` copies of the decoder (unet.py:859-861, 879-968): they
    carry parameters in the state dict but ``forward`` never calls them.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple


@dataclass
class ResSpec:
    name: str            # state-dict prefix, e.g. "input_blocks.1.0"
    cin: int
    cout: int
    updown: str = "none"  # "none" | "up" | "down"

    @property
    def has_skip_conv(self) -> bool:
        return self.cin != self.cout


@dataclass
class AttnSpec:
    name: str            # e.g. "input_blocks.1.1" or "input_attns.1"
    channels: int
    heads: int
    joint: bool          # True: AttentionBlock1D over xy|yt|xt; False: per plane


@dataclass
class Stage:
    """One entry of input_blocks / output_blocks plus its cross-plane attention."""
    layers: List[object] = field(default_factory=list)   # ResSpec | AttnSpec(joint=False)
    joint: Optional[AttnSpec] = None
    skip_channels: int = 0      # decoder only: channels popped from the skip stack
    level_in: int = 0           # pyramid level (0 = 32x32) at the stage input
    level_out: int = 0


@dataclass
class UNetArch:
    model_channels: int
    in_channels: int
    out_channels: int
    num_heads: int
    time_embed_dim: int
    stem_in: int
    input_stages: List[Stage]
    middle: Stage
    output_stages: List[Stage]
    dead_bg: List[Stage]         # parameter-only mirror of output_stages
    head_channels: int
    cond_model: bool

    def live_resblocks(self) -> List[ResSpec]:
        out = []
        for st in self.input_stages + [self.middle] + self.output_stages:
            out += [l for l in st.layers if isinstance(l, ResSpec)]
        return out

    def live_attns(self) -> List[AttnSpec]:
        out = []
        for st in self.input_stages + [self.middle] + self.output_stages:
            out += [l for l in st.layers if isinstance(l, AttnSpec)]
            if st.joint is not None:
                out.append(st.joint)
        return out


def plane_shapes(level: int, res0: int = 32, t0: int = 16) -> Tuple[Tuple[int, int], ...]:
    """(h, w) of the xy, yt, xt planes at a pyramid level (unet.py:1027-1029,
    1047-1049: xy is res x res, yt and xt are t x res)."""
    res, t = res0 >> level, t0 >> level
    return ((res, res), (t, res), (t, res))


def tokens_at(level: int, res0: int = 32, t0: int = 16) -> int:
    return sum(h * w for h, w in plane_shapes(level, res0, t0))


def build_arch(
    image_size: int = 32,
    in_channels: int = 4,
    model_channels: int = 128,
    out_channels: int = 4,
    num_res_blocks: int = 2,
    attention_resolutions: Sequence[int] = (4, 2, 1),
    channel_mult: Sequence[int] = (1, 2, 4, 4),
    num_heads: int = 8,
    cond_model: bool = False,
    **_ignored,
) -> UNetArch:
    attention_resolutions = [int(a) for a in attention_resolutions]
    channel_mult = [int(m) for m in channel_mult]
    mc = int(model_channels)
    heads = int(num_heads)
    if heads <= 0:
        raise ValueError("num_heads must be set (num_head_channels mode is not on the MToV path)")

    stages: List[Stage] = [Stage(layers=[], joint=None, level_in=0, level_out=0)]  # stem
    skip_chans = [mc]
    ch, ds, level = mc, 1, 0
    idx = 1
    for li, mult in enumerate(channel_mult):
        for _ in range(int(num_res_blocks)):
            st = Stage(level_in=level, level_out=level)
            st.layers.append(ResSpec(f"input_blocks.{idx}.0", ch, mult * mc))
            ch = mult * mc
            if ds in attention_resolutions:
                st.layers.append(AttnSpec(f"input_blocks.{idx}.1", ch, heads, joint=False))
            st.joint = AttnSpec(f"input_attns.{idx}", ch, heads, joint=True)
            stages.append(st)
            skip_chans.append(ch)
            idx += 1
        if li != len(channel_mult) - 1:
            st = Stage(level_in=level, level_out=level + 1)
            st.layers.append(ResSpec(f"input_blocks.{idx}.0", ch, ch, updown="down"))
            st.joint = AttnSpec(f"input_attns.{idx}", ch, heads, joint=True)
            stages.append(st)
            skip_chans.append(ch)
            idx += 1
            ds *= 2
            level += 1

    middle = Stage(level_in=level, level_out=level)
    middle.layers = [
        ResSpec("middle_block.0", ch, ch),
        AttnSpec("middle_block.1", ch, heads, joint=False),
        ResSpec("middle_block.2", ch, ch),
    ]
    middle.joint = AttnSpec("mid_attn", ch, heads, joint=True)

    out_stages: List[Stage] = []
    bg_stages: List[Stage] = []
    oidx = 0
    for li, mult in list(enumerate(channel_mult))[::-1]:
        for i in range(int(num_res_blocks) + 1):
            ich = skip_chans.pop()
            st = Stage(skip_channels=ich, level_in=level, level_out=level)
            bg = Stage(skip_channels=ich, level_in=level, level_out=level)
            st.layers.append(ResSpec(f"output_blocks.{oidx}.0", ch + ich, mc * mult))
            bg.layers.append(ResSpec(f"output_bg_blocks.{oidx}.0", ch + ich, mc * mult))
            ch = mc * mult
            nxt = 1
            if ds in attention_resolutions:
                st.layers.append(AttnSpec(f"output_blocks.{oidx}.{nxt}", ch, heads, joint=False))
                nxt += 1
            if li and i == int(num_res_blocks):
                st.layers.append(ResSpec(f"output_blocks.{oidx}.{nxt}", ch, ch, updown="up"))
                # the bg copy never gets the per-plane attention (unet.py:879-889, 934-947)
                bg.layers.append(ResSpec(f"output_bg_blocks.{oidx}.1", ch, ch, updown="up"))
                ds //= 2
                level -= 1
                st.level_out = level
                bg.level_out = level
            st.joint = AttnSpec(f"output_attns.{oidx}", ch, heads, joint=True)
            bg.joint = AttnSpec(f"output_bg_attns.{oidx}", ch, heads, joint=True)
            out_stages.append(st)
            bg_stages.append(bg)
            oidx += 1

    return UNetArch(
        model_channels=mc,
        in_channels=int(in_channels),
        out_channels=int(out_channels),
        num_heads=heads,
        time_embed_dim=4 * mc,
        stem_in=16,
        input_stages=stages,
        middle=middle,
        output_stages=out_stages,
        dead_bg=bg_stages,
        head_channels=ch,
        cond_model=bool(cond_model),
    )


def param_shapes(arch: UNetArch, include_dead: bool = True) -> List[Tuple[str, Tuple[int, ...]]]:
    """(name, shape) of every tensor in ``UNetModel.state_dict()`` in the
    reference's registration order (time_embed, input_blocks, input_attns,
    middle_block, mid_attn, output_blocks, output_bg_blocks, output_attns,
    output_bg_attns, out — unet.py:701-975)."""
    out: List[Tuple[str, Tuple[int, ...]]] = []
    ted = arch.time_embed_dim

    def res(r: ResSpec):
        p = r.name
        out.append((f"{p}.in_layers.0.weight", (r.cin,)))
        out.append((f"{p}.in_layers.0.bias", (r.cin,)))
        out.append((f"{p}.in_layers.2.weight", (r.cout, r.cin, 3, 3)))
        out.append((f"{p}.in_layers.2.bias", (r.cout,)))
        out.append((f"{p}.emb_layers.1.weight", (2 * r.cout, ted)))
        out.append((f"{p}.emb_layers.1.bias", (2 * r.cout,)))
        out.append((f"{p}.out_layers.0.weight", (r.cout,)))
        out.append((f"{p}.out_layers.0.bias", (r.cout,)))
        out.append((f"{p}.out_layers.3.weight", (r.cout, r.cout, 3, 3)))
        out.append((f"{p}.out_layers.3.bias", (r.cout,)))
        if r.has_skip_conv:
            out.append((f"{p}.skip_connection.weight", (r.cout, r.cin, 1, 1)))
            out.append((f"{p}.skip_connection.bias", (r.cout,)))

    def attn(a: AttnSpec):
        p, c = a.name, a.channels
        out.append((f"{p}.norm.weight", (c,)))
        out.append((f"{p}.norm.bias", (c,)))
        out.append((f"{p}.qkv.weight", (3 * c, c, 1)))
        out.append((f"{p}.qkv.bias", (3 * c,)))
        out.append((f"{p}.proj_out.weight", (c, c, 1)))
        out.append((f"{p}.proj_out.bias", (c,)))

    def layers(st: Stage):
        for l in st.layers:
            res(l) if isinstance(l, ResSpec) else attn(l)

    if arch.cond_model:
        out.append(("zeros", (1, arch.in_channels, 2048)))
    mc = arch.model_channels
    out.append(("time_embed.0.weight", (ted, mc)))
    out.append(("time_embed.0.bias", (ted,)))
    out.append(("time_embed.2.weight", (ted, ted)))
    out.append(("time_embed.2.bias", (ted,)))
    out.append(("input_blocks.0.0.weight", (mc, arch.stem_in, 3, 3)))
    out.append(("input_blocks.0.0.bias", (mc,)))
    for st in arch.input_stages[1:]:
        layers(st)
    for st in arch.input_stages[1:]:
        attn(st.joint)
    layers(arch.middle)
    attn(arch.middle.joint)
    for st in arch.output_stages:
        layers(st)
    if include_dead:
        for st in arch.dead_bg:
            layers(st)
    for st in arch.output_stages:
        attn(st.joint)
    if include_dead:
        for st in arch.dead_bg:
            attn(st.joint)
    out.append(("out.0.weight", (arch.head_channels,)))
    out.append(("out.0.bias", (arch.head_channels,)))
    out.append(("out.2.weight", (arch.out_channels, mc, 3, 3)))
    out.append(("out.2.bias", (arch.out_channels,)))
    return out


BASE_UNET_CONFIG = dict(  # MToV/configs/latent-diffusion/base.yaml:28-39
    image_size=32, in_channels=4, out_channels=4, model_channels=128,
    attention_resolutions=[4, 2, 1], num_res_blocks=2, channel_mult=[1, 2, 4, 4],
    num_heads=8, use_scale_shift_norm=True, resblock_updown=True, cond_model=False,
)
LONGVID_UNET_CONFIG = dict(  # MToV/configs/latent-diffusion/base_longvid.yaml:27-38
    BASE_UNET_CONFIG, model_channels=256, cond_model=True,
)
# A reduced network for fast CPU tests (same code paths: up/down, skip conv,
# per-plane + joint attention, straddling GroupNorm groups in the decoder).
TINY_UNET_CONFIG = dict(
    BASE_UNET_CONFIG, model_channels=64, channel_mult=[1, 2], num_res_blocks=1,
    attention_resolutions=[2, 1], num_heads=4,
)
