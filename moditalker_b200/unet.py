"""Drop-in ``UNetModel`` / ``DiffusionWrapper`` for MoDiTalker's MToV stage.

Same constructor keywords, attributes, ``state_dict`` keys/shapes (804 keys for
base.yaml, including the never-called ``output_bg_*`` branch) and ``forward``
signature as /root/reference/MToV/models/ddpm/unet.py:34-61, 601-1117 — so
``sample.py:221-245`` runs unchanged with ``from moditalker_b200 import
UNetModel, DiffusionWrapper``.

The modules below only HOLD parameters (stock ``nn.Conv2d`` / ``nn.GroupNorm`` /
``nn.Linear`` objects, so ``.to()``, ``deepcopy``, ``load_state_dict(strict=True)``
behave as in the reference).  ``forward`` hands raw device pointers to
``libmtv_b200.so`` through the C ABI (``include/mtv_b200.h``); there is no PyTorch
or CPU implementation of the forward in this package.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import torch
import torch.nn as nn

from . import _lib
from .arch import AttnSpec, ResSpec, Stage, build_arch

__all__ = ["UNetModel", "DiffusionWrapper"]


class GroupNorm32(nn.GroupNorm):
    """Parameter holder for GroupNorm(32, C) (diffusionmodules.py:156-173)."""


def _zero(m: nn.Module) -> nn.Module:
    for p in m.parameters():   # the reference zero-initialises these (diffusionmodules.py:131-137)
        p.detach().zero_()
    return m


class ResBlockParams(nn.Module):
    """Weights of a ResBlock (unet.py:109-167); key layout in_layers.{0,2},
    emb_layers.1, out_layers.{0,3}, skip_connection."""

    def __init__(self, spec: ResSpec, emb_channels: int, dropout: float):
        super().__init__()
        self.channels, self.out_channels, self.updown = spec.cin, spec.cout, spec.updown
        self.in_layers = nn.Sequential(GroupNorm32(32, spec.cin), nn.SiLU(), nn.Conv2d(spec.cin, spec.cout, 3, padding=1))
        self.emb_layers = nn.Sequential(nn.SiLU(), nn.Linear(emb_channels, 2 * spec.cout))
        self.out_layers = nn.Sequential(
            GroupNorm32(32, spec.cout), nn.SiLU(), nn.Dropout(p=dropout),
            _zero(nn.Conv2d(spec.cout, spec.cout, 3, padding=1)),
        )
        self.skip_connection = nn.Conv2d(spec.cin, spec.cout, 1) if spec.has_skip_conv else nn.Identity()


class AttentionParams(nn.Module):
    """Weights of AttentionBlock / AttentionBlock1D (unet.py:217-242, 264-289)."""

    def __init__(self, spec: AttnSpec):
        super().__init__()
        self.channels, self.num_heads = spec.channels, spec.heads
        self.norm = GroupNorm32(32, spec.channels)
        self.qkv = nn.Conv1d(spec.channels, 3 * spec.channels, 1)
        self.proj_out = _zero(nn.Conv1d(spec.channels, spec.channels, 1))


class _EngineSlot:
    """Owns the native handle.  Never copied: deepcopy / pickle of the module give
    the copy an empty slot that lazily creates its own handle."""

    def __init__(self):
        self.handle: Optional[ctypes.c_void_p] = None
        self.device: Optional[torch.device] = None
        self.synced = False

    def __deepcopy__(self, memo):
        return _EngineSlot()

    def __reduce__(self):
        return (_EngineSlot, ())

    def release(self):
        if self.handle is not None:
            try:
                _lib.load_library().mtv_destroy(self.handle)
            except Exception:
                pass
            self.handle = None
            self.synced = False

    def __del__(self):
        self.release()


class UNetModel(nn.Module):
    """Tri-plane UNet epsilon-predictor.  Constructor mirrors unet.py:631-659; the
    options that the MToV configs never enable (3-D convs, class conditioning,
    spatial transformer, fp16 torso, num_head_channels, new attention order,
    non-resblock up/down, non-FiLM conditioning) raise ``NotImplementedError``."""

    def __init__(
        self,
        image_size,
        in_channels,
        model_channels,
        out_channels,
        num_res_blocks,
        attention_resolutions,
        dropout=0,
        channel_mult=(1, 2, 4, 8),
        conv_resample=True,
        dims=2,
        num_classes=None,
        use_checkpoint=False,
        use_fp16=False,
        num_heads=-1,
        num_head_channels=-1,
        num_heads_upsample=-1,
        use_scale_shift_norm=False,
        resblock_updown=False,
        use_new_attention_order=False,
        use_spatial_transformer=False,
        transformer_depth=1,
        context_dim=None,
        n_embed=None,
        legacy=True,
        cond_model=False,
    ):
        super().__init__()
        unsupported = {
            "dims != 2": dims != 2,
            "num_classes": num_classes is not None,
            "use_fp16": bool(use_fp16),
            "num_head_channels": num_head_channels != -1,
            "num_heads_upsample": num_heads_upsample not in (-1, num_heads),
            "use_scale_shift_norm=False": not use_scale_shift_norm,
            "resblock_updown=False": not resblock_updown,
            "use_new_attention_order": bool(use_new_attention_order),
            "use_spatial_transformer / context_dim": bool(use_spatial_transformer) or context_dim is not None,
            "n_embed": n_embed is not None,
            "legacy=False": not legacy,
            "num_heads unset": num_heads == -1,
            "dropout > 0": float(dropout) != 0.0,
        }
        bad = [k for k, v in unsupported.items() if v]
        if bad:
            raise NotImplementedError(
                "moditalker_b200.UNetModel covers the MToV configs (configs/latent-diffusion/*.yaml); "
                f"not on that path: {', '.join(bad)}"
            )
        attention_resolutions = [int(a) for a in attention_resolutions]   # accepts OmegaConf ListConfig
        channel_mult = [int(m) for m in channel_mult]

        self.image_size = image_size
        self.in_channels = in_channels
        self.model_channels = model_channels
        self.out_channels = out_channels
        self.num_res_blocks = num_res_blocks
        self.attention_resolutions = attention_resolutions
        self.dropout = dropout
        self.channel_mult = channel_mult
        self.conv_resample = conv_resample
        self.num_classes = num_classes
        self.use_checkpoint = use_checkpoint
        self.dtype = torch.float32
        self.num_heads = num_heads
        self.num_head_channels = num_head_channels
        self.num_heads_upsample = num_heads
        self.predict_codebook_ids = False
        self.cond_model = cond_model
        if cond_model:
            self.register_buffer("zeros", torch.zeros(1, self.in_channels, 2048))   # unet.py:697-698 (unused)

        arch = build_arch(
            image_size=image_size, in_channels=in_channels, model_channels=model_channels,
            out_channels=out_channels, num_res_blocks=num_res_blocks,
            attention_resolutions=attention_resolutions, channel_mult=channel_mult,
            num_heads=num_heads, cond_model=cond_model,
        )
        self._arch = arch
        ted = arch.time_embed_dim

        def holders(st: Stage):
            return [ResBlockParams(l, ted, dropout) if isinstance(l, ResSpec) else AttentionParams(l) for l in st.layers]

        self.time_embed = nn.Sequential(nn.Linear(model_channels, ted), nn.SiLU(), nn.Linear(ted, ted))
        self.input_blocks = nn.ModuleList([nn.Sequential(nn.Conv2d(arch.stem_in, model_channels, 3, padding=1))])
        self.input_attns = nn.ModuleList([nn.Identity()])
        for st in arch.input_stages[1:]:
            self.input_blocks.append(nn.Sequential(*holders(st)))
            self.input_attns.append(AttentionParams(st.joint))
        self.middle_block = nn.Sequential(*holders(arch.middle))
        self.mid_attn = AttentionParams(arch.middle.joint)
        self.output_blocks = nn.ModuleList([nn.Sequential(*holders(st)) for st in arch.output_stages])
        # parameter-only twin of the decoder: in the reference's state dict, never run (unet.py:859-861)
        self.output_bg_blocks = nn.ModuleList([nn.Sequential(*holders(st)) for st in arch.dead_bg])
        self.output_attns = nn.ModuleList([AttentionParams(st.joint) for st in arch.output_stages])
        self.output_bg_attns = nn.ModuleList([AttentionParams(st.joint) for st in arch.dead_bg])
        self.out = nn.Sequential(
            GroupNorm32(32, arch.head_channels), nn.SiLU(),
            _zero(nn.Conv2d(model_channels, out_channels, 3, padding=1)),
        )
        self._engine = _EngineSlot()
        # 0 = tcgen05 tensor-core kernels wherever a tile shape exists (default);
        # 1 = fp32 CUDA-core kernels everywhere (cross-check path used by the tests)
        self.kernel_path = int(os.environ.get("MTV_KERNEL_PATH", "0"))

    # ------------------------------------------------------------------ weight sync
    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self._engine.synced = False
        return r

    def load_state_dict(self, *a, **k):
        r = super().load_state_dict(*a, **k)
        self._engine.synced = False
        return r

    def _load_from_state_dict(self, *a, **k):   # reached when a PARENT module loads (DiffusionWrapper, DDPM)
        super()._load_from_state_dict(*a, **k)
        self._engine.synced = False

    def refresh_weights(self) -> None:
        """Re-upload parameters to the native engine (call after in-place edits
        such as an optimizer step)."""
        self._engine.synced = False

    def _ensure_engine(self, device: torch.device):
        lib = _lib.load_library()
        eng = self._engine
        if eng.handle is not None and eng.device != device:
            eng.release()
        if eng.handle is None:
            cfg = _lib.MtvConfig()
            cfg.abi_version = _lib.MTV_ABI_VERSION
            cfg.image_size = int(self.image_size)
            cfg.in_channels = int(self.in_channels)
            cfg.out_channels = int(self.out_channels)
            cfg.model_channels = int(self.model_channels)
            cfg.num_res_blocks = int(self.num_res_blocks)
            cfg.num_heads = int(self.num_heads)
            cfg.num_levels = len(self.channel_mult)
            for i, m in enumerate(self.channel_mult):
                cfg.channel_mult[i] = int(m)
                cfg.attn_at_level[i] = 1 if (1 << i) in self.attention_resolutions else 0
            cfg.device = device.index if device.index is not None else torch.cuda.current_device()
            cfg.kernel_path = int(self.kernel_path)
            h = ctypes.c_void_p()
            _lib.check(lib.mtv_create(ctypes.byref(cfg), ctypes.byref(h)), "mtv_create")
            eng.handle, eng.device, eng.synced = h, device, False
        if not eng.synced:
            stream = torch.cuda.current_stream(device).cuda_stream
            keep = []
            for name, tns in self.state_dict().items():
                if tns.device != device:
                    raise RuntimeError(f"parameter {name} is on {tns.device}, input is on {device}")
                src = tns.detach()
                if src.dtype != torch.float32 or not src.is_contiguous():
                    src = src.float().contiguous()
                keep.append(src)
                shape = (ctypes.c_int64 * src.dim())(*src.shape)
                used = ctypes.c_int32(0)
                _lib.check(
                    lib.mtv_load_weight(eng.handle, name.encode(), ctypes.c_void_p(src.data_ptr()), shape, src.dim(),
                                        ctypes.byref(used), ctypes.c_void_p(stream)),
                    "mtv_load_weight",
                )
            _lib.check(lib.mtv_weights_ready(eng.handle, None, None), "mtv_weights_ready")
            eng.synced = True
        return lib, eng.handle

    # ------------------------------------------------------------------ forward
    def forward(self, x, cond=None, image_cond=None, timesteps=None, context=None, y=None, **kwargs):
        """Epsilon prediction for tri-plane latents (unet.py:995-1117).

        x [B,4,2048], cond [B,8,2048], image_cond [B,4,>=1024], timesteps [B] ->
        [B,4,2048].  ``context`` is accepted and ignored exactly as in the reference
        (no cross-attention is built without ``use_spatial_transformer``)."""
        assert (y is not None) == (self.num_classes is not None), \
            "must specify y if and only if the model is class-conditional"
        if not x.is_cuda:
            raise RuntimeError("moditalker_b200.UNetModel runs on CUDA (sm_100a) tensors only; no CPU path exists")
        if self.training and torch.is_grad_enabled():
            raise NotImplementedError("training (backward) is the next scope item (SURVEY.md §8f); use eval()/no_grad()")
        dev = x.device
        B = x.shape[0]
        if x.dim() != 3 or x.shape[1] != self.in_channels or x.shape[2] != 2048:
            raise RuntimeError(f"x must be [B,{self.in_channels},2048], got {tuple(x.shape)}")
        if cond is None or cond.shape[0] != B or cond.shape[1] != 2 * self.in_channels or cond.shape[2] != 2048:
            raise RuntimeError(f"cond must be [B,{2 * self.in_channels},2048]")
        if image_cond is None or image_cond.shape[0] != B or image_cond.shape[1] != self.in_channels or image_cond.shape[2] < 1024:
            raise RuntimeError(f"image_cond must be [B,{self.in_channels},>=1024]")
        if timesteps is None or timesteps.shape[0] != B:
            raise RuntimeError("timesteps must be a 1-D batch of B steps")
        lib, handle = self._ensure_engine(dev)
        xf = x.detach().to(torch.float32).contiguous()
        cf = cond.detach().to(device=dev, dtype=torch.float32).contiguous()
        icf = image_cond.detach().to(device=dev, dtype=torch.float32).contiguous()
        tf = timesteps.detach().to(device=dev, dtype=torch.long).contiguous()
        out = torch.empty((B, self.out_channels, 2048), device=dev, dtype=torch.float32)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(
            lib.mtv_unet_forward(handle, ctypes.c_void_p(xf.data_ptr()), ctypes.c_void_p(cf.data_ptr()),
                                 ctypes.c_void_p(icf.data_ptr()), int(icf.shape[2]), ctypes.c_void_p(tf.data_ptr()),
                                 int(B), ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(stream)),
            "mtv_unet_forward",
        )
        return out.type(x.dtype)

    # ------------------------------------------------------------------ introspection (tests / bench)
    def native_handle(self, device=None):
        dev = torch.device(device) if device is not None else next(self.parameters()).device
        return self._ensure_engine(dev)

    def plan_info(self, batch: int):
        lib, h = self.native_handle()
        n, ws, wb = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int64(0)
        _lib.check(lib.mtv_plan_info(h, int(batch), ctypes.byref(n), ctypes.byref(ws), ctypes.byref(wb)), "mtv_plan_info")
        return {"launches": n.value, "workspace_bytes": ws.value, "weight_bytes": wb.value}

    def debug_read(self, tag: str, batch: int, channels: int, tokens: int) -> torch.Tensor:
        lib, h = self.native_handle()
        dev = next(self.parameters()).device
        dst = torch.empty((batch, channels, tokens), device=dev, dtype=torch.float32)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.mtv_debug_read(h, tag.encode(), ctypes.c_void_p(dst.data_ptr()), dst.numel(),
                                      ctypes.c_void_p(stream)), "mtv_debug_read")
        return dst

    def profile_forward(self, x, cond, image_cond, timesteps):
        """Per-launch CUDA-event timings of one (serialised) forward."""
        dev = x.device
        lib, h = self._ensure_engine(dev)
        B = x.shape[0]
        out = torch.empty((B, self.out_channels, 2048), device=dev, dtype=torch.float32)
        cap = 1024
        ents = (_lib.MtvKernelTime * cap)()
        n = ctypes.c_int32(0)
        stream = torch.cuda.current_stream(dev).cuda_stream
        xf, cf, icf = x.float().contiguous(), cond.float().contiguous(), image_cond.float().contiguous()
        tf = timesteps.long().contiguous()
        _lib.check(
            lib.mtv_profile_forward(h, ctypes.c_void_p(xf.data_ptr()), ctypes.c_void_p(cf.data_ptr()),
                                    ctypes.c_void_p(icf.data_ptr()), int(icf.shape[2]), ctypes.c_void_p(tf.data_ptr()),
                                    int(B), ctypes.c_void_p(out.data_ptr()), ents, cap, ctypes.byref(n),
                                    ctypes.c_void_p(stream)),
            "mtv_profile_forward",
        )
        rows = [(ents[i].name.decode(), ents[i].us, ents[i].flops, ents[i].bytes) for i in range(n.value)]
        return out, rows


class DiffusionWrapper(nn.Module):
    """unet.py:34-61.  Only ``conditioning_key=None`` is live in MoDiTalker; the other
    branches of the reference call ``UNetModel`` with argument lists it does not accept."""

    def __init__(self, model, conditioning_key=None):
        super().__init__()
        self.diffusion_model = model
        self.conditioning_key = conditioning_key
        assert self.conditioning_key in [None, "concat", "crossattn", "hybrid", "adm"]

    def forward(self, x, cond, image_cond, t, kpt_coord=None, c_concat: list = None, c_crossattn: list = None):
        if self.conditioning_key is None:
            # DDPM.model_predictions passes its `context` positionally into kpt_coord
            # (losses/ddpm.py:340); it is ignored there and here.
            return self.diffusion_model(x, cond, image_cond, t, context=c_crossattn)
        raise NotImplementedError(
            f"conditioning_key={self.conditioning_key!r} is unreachable in MoDiTalker (sample.py:222 passes None)"
        )
