"""Drop-in ``DDPM`` sampler for MoDiTalker's MToV stage.

Mirrors the part of /root/reference/MToV/losses/ddpm.py that ``sample.py:239-245,
377-384`` uses: the constructor, the schedule buffers (ddpm.py:195-253), ``sample``
(457-484), ``ddim_sample`` (363-404), ``ddim_sample_noised_start`` (407-454),
``q_sample`` (486-491), ``model_predictions`` (338-360) and
``predict_start_from_noise`` (278-282).  The per-step arithmetic runs in
``libmtv_b200.so`` (``mtv_ddim_step`` / ``mtv_q_sample``); the host computes the
step scalars with the same fp32 operations as the reference, and draws noise
with the same ``torch.randn`` / ``torch.randn_like`` calls in the same order so
the CUDA RNG stream is consumed identically.

Not mirrored (dead or out of scope): ``p_sample_loop`` /
``p_mean_variance`` (call the model with a signature it does not have),
``p_losses`` / ``forward`` (training, §8f).
"""
from __future__ import annotations

import ctypes
from collections import namedtuple
from typing import Callable, List, Optional

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .unet import DiffusionWrapper, UNetModel

ModelPrediction = namedtuple("ModelPrediction", ["pred_noise", "pred_x_start"])

__all__ = ["DDPM", "ModelPrediction"]


def _linear_betas(n_timestep: int, linear_start: float, linear_end: float) -> np.ndarray:
    # "linear" schedule of the reference is linear in sqrt(beta) (ddpm.py:80-81)
    return (torch.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=torch.float64) ** 2).numpy()


class DDPM(nn.Module):
    def __init__(
        self,
        model,
        timesteps=1000,
        beta_schedule="linear",
        loss_type="l2",
        ckpt_path=None,
        ignore_keys=(),
        load_only_unet=False,
        monitor="val/loss",
        use_ema=True,
        first_stage_key="image",
        image_size=256,
        channels=3,
        log_every_t=200,
        clip_denoised=True,
        linear_start=0.0015,
        linear_end=0.0195,
        cosine_s=8e-3,
        given_betas=None,
        original_elbo_weight=0.0,
        v_posterior=0.0,
        l_simple_weight=1.0,
        conditioning_key=None,
        parameterization="eps",
        use_positional_encodings=False,
        learn_logvar=False,
        logvar_init=0.0,
        sampling_timesteps=1000,
        ddim_sampling_eta=1.0,
        w=1.0,
        first_stage_model=None,
    ):
        super().__init__()
        if parameterization != "eps":
            raise NotImplementedError("MoDiTalker samples in eps-prediction mode (ddpm.py:145)")
        if beta_schedule != "linear" and given_betas is None:
            raise NotImplementedError("only the 'linear' beta schedule is on the MToV path")
        self.parameterization = parameterization
        self.clip_denoised = clip_denoised
        self.log_every_t = log_every_t
        self.first_stage_key = first_stage_key
        self.image_size = 2048          # ddpm.py:162: tri-plane token count, not the ctor argument
        self.channels = channels
        self.model = model
        self.use_ema = use_ema
        self.v_posterior = v_posterior
        self.original_elbo_weight = original_elbo_weight
        self.l_simple_weight = l_simple_weight
        self.loss_type = loss_type

        betas = np.asarray(given_betas, dtype=np.float64) if given_betas is not None else _linear_betas(
            timesteps, linear_start, linear_end)
        alphas = 1.0 - betas
        ac = np.cumprod(alphas, axis=0)
        ac_prev = np.append(1.0, ac[:-1])
        self.num_timesteps = int(betas.shape[0])
        self.linear_start, self.linear_end = linear_start, linear_end
        f32 = lambda a: torch.tensor(a, dtype=torch.float32)
        self.register_buffer("betas", f32(betas))
        self.register_buffer("alphas_cumprod", f32(ac))
        self.register_buffer("alphas_cumprod_prev", f32(ac_prev))
        self.register_buffer("sqrt_alphas_cumprod", f32(np.sqrt(ac)))
        self.register_buffer("sqrt_one_minus_alphas_cumprod", f32(np.sqrt(1.0 - ac)))
        self.register_buffer("log_one_minus_alphas_cumprod", f32(np.log(1.0 - ac)))
        self.register_buffer("sqrt_recip_alphas_cumprod", f32(np.sqrt(1.0 / ac)))
        self.register_buffer("sqrt_recipm1_alphas_cumprod", f32(np.sqrt(1.0 / ac - 1)))
        post_var = (1 - v_posterior) * betas * (1.0 - ac_prev) / (1.0 - ac) + v_posterior * betas
        self.register_buffer("posterior_variance", f32(post_var))
        self.register_buffer("posterior_log_variance_clipped", f32(np.log(np.maximum(post_var, 1e-20))))
        self.register_buffer("posterior_mean_coef1", f32(betas * np.sqrt(ac_prev) / (1.0 - ac)))
        self.register_buffer("posterior_mean_coef2", f32((1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac)))

        self.sampling_timesteps = sampling_timesteps if sampling_timesteps is not None else timesteps
        assert self.sampling_timesteps <= timesteps
        self.is_ddim_sampling = self.sampling_timesteps < timesteps
        self.ddim_sampling_eta = ddim_sampling_eta
        self.w = w
        self.first_stage_model = first_stage_model
        # test hook: noise_fn(kind, like_or_shape, device) -> tensor; default = torch.randn*
        self.noise_fn: Optional[Callable] = None
        # host copies of the schedule (the reference indexes the registered device buffers with Python ints; the fused step
        # takes host scalars).  Refreshed whenever a buffer has been replaced or edited (load_state_dict, .to(), in-place
        # writes): see _host_schedule().
        self._host = None
        self._host_key = None

    # ------------------------------------------------------------------ helpers
    def _unet(self) -> UNetModel:
        m = self.model
        if isinstance(m, DiffusionWrapper):
            m = m.diffusion_model
        if not isinstance(m, UNetModel):
            raise TypeError("moditalker_b200.DDPM wraps moditalker_b200.DiffusionWrapper / UNetModel")
        return m

    def _randn(self, shape, device):
        if self.noise_fn is not None:
            return self.noise_fn("randn", shape, device)
        return torch.randn(shape, device=device)

    def _randn_like(self, ref):
        if self.noise_fn is not None:
            return self.noise_fn("randn_like", tuple(ref.shape), ref.device)
        return torch.randn_like(ref)

    _HOST_KEYS = ("alphas_cumprod", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod",
                  "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod")

    def _host_schedule(self):
        """CPU copies of the live schedule buffers, re-taken when any of them changed identity or content version
        (``Tensor._version`` counts in-place edits; ``load_state_dict`` copies in place, ``.to()`` replaces the tensor)."""
        key = tuple((getattr(self, k).data_ptr(), getattr(self, k)._version, str(getattr(self, k).device)) for k in self._HOST_KEYS)
        if self._host is None or key != self._host_key:
            self._host = {k: getattr(self, k).detach().to("cpu", copy=True) for k in self._HOST_KEYS}
            self._host_key = key
        return self._host

    def time_pairs(self):
        """ddpm.py:372-376."""
        times = torch.linspace(-1, self.num_timesteps - 1, steps=self.sampling_timesteps + 1)
        times = list(reversed(times.int().tolist()))
        return list(zip(times[:-1], times[1:]))

    def step_scalars(self, time: int, time_next: int):
        """(sqrt_recip_ac, sqrt_recipm1_ac, sqrt(alpha_next), c, sigma) as Python floats
        holding exact fp32 values, computed with the reference's fp32 tensor ops
        (ddpm.py:280-281, 390-394)."""
        H = self._host_schedule()
        sr = float(H["sqrt_recip_alphas_cumprod"][time])
        srm1 = float(H["sqrt_recipm1_alphas_cumprod"][time])
        if time_next < 0:
            return sr, srm1, 0.0, 0.0, 0.0
        alpha, alpha_next = H["alphas_cumprod"][time], H["alphas_cumprod"][time_next]
        sigma = self.ddim_sampling_eta * ((1 - alpha / alpha_next) * (1 - alpha_next) / (1 - alpha)).sqrt()
        c = (1 - alpha_next - sigma ** 2).sqrt()
        return sr, srm1, float(alpha_next.sqrt()), float(c), float(sigma)

    # ------------------------------------------------------------------ reference API
    def predict_start_from_noise(self, x_t, t, noise):
        sr = self.sqrt_recip_alphas_cumprod.gather(-1, t).reshape(-1, *((1,) * (x_t.dim() - 1)))
        srm1 = self.sqrt_recipm1_alphas_cumprod.gather(-1, t).reshape(-1, *((1,) * (x_t.dim() - 1)))
        return sr * x_t - srm1 * noise

    def model_predictions(self, x, cond, image_cond, t, context=None, clip_x_start=False):
        pred_noise = self.model(x, cond, image_cond, t, context)
        x_start = self.predict_start_from_noise(x, t, pred_noise)
        if clip_x_start:
            x_start.clamp_(-1.0, 1.0)
        return ModelPrediction(pred_noise, x_start)

    def q_sample(self, x_start, t, noise=None):
        """ddpm.py:486-491 for a single-element ``t`` (how the sampler calls it,
        ddpm.py:422-429)."""
        if noise is None:
            noise = self._randn_like(x_start)
        if t.numel() != 1:
            a = self.sqrt_alphas_cumprod.gather(-1, t).reshape(-1, *((1,) * (x_start.dim() - 1)))
            b = self.sqrt_one_minus_alphas_cumprod.gather(-1, t).reshape(-1, *((1,) * (x_start.dim() - 1)))
            return a * x_start + b * noise
        ti = int(t.reshape(-1)[0])
        H = self._host_schedule()
        a = float(H["sqrt_alphas_cumprod"][ti])
        b = float(H["sqrt_one_minus_alphas_cumprod"][ti])
        lib, h = self._unet()._ensure_engine(x_start.device)
        xs = x_start.detach().float().contiguous()
        nz = noise.detach().float().contiguous()
        out = torch.empty_like(xs)
        stream = torch.cuda.current_stream(xs.device).cuda_stream
        _lib.check(lib.mtv_q_sample(h, ctypes.c_void_p(xs.data_ptr()), ctypes.c_void_p(nz.data_ptr()), xs.numel(),
                                    a, b, ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(stream)), "mtv_q_sample")
        return out

    def _loop(self, img, cond, image_cond, context, pairs):
        """The hot loop (ddpm.py:382-398 / 434-448): one UNet forward + one fused
        update kernel per step, no host synchronisation."""
        dev = img.device
        lib, h = self._unet()._ensure_engine(dev)
        B = img.shape[0]
        n = img.numel()
        t_cache = {}
        for time, time_next in pairs:
            tc = t_cache.get(time)
            if tc is None:
                tc = t_cache[time] = torch.full((B,), time, device=dev, dtype=torch.long)
            eps = self.model(img, cond, image_cond, tc, context)
            sr, srm1, san, c, sigma = self.step_scalars(time, time_next)
            last = time_next < 0
            noise = None if last else self._randn_like(img)
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(
                lib.mtv_ddim_step(h, ctypes.c_void_p(img.data_ptr()), ctypes.c_void_p(eps.data_ptr()),
                                  ctypes.c_void_p(noise.data_ptr()) if noise is not None else None, n,
                                  sr, srm1, san, c, sigma, 1 if last else 0, ctypes.c_void_p(stream)),
                "mtv_ddim_step",
            )
        return img

    @torch.no_grad()
    def ddim_sample(self, shape, cond, image_cond, context=None, clip_denoised=True):
        if not clip_denoised:
            raise NotImplementedError("the fused DDIM step clamps x_start like every call site of the reference")
        device = self.betas.device
        img = self._randn(shape, device).float().contiguous()
        return self._loop(img, cond, image_cond, context, self.time_pairs())

    @torch.no_grad()
    def ddim_sample_noised_start(self, shape, x_start, cond, image_cond, context=None, clip_denoised=True,
                                 ratio_=None, fixed_noise=False):
        if not clip_denoised:
            raise NotImplementedError("the fused DDIM step clamps x_start like every call site of the reference")
        pairs = self.time_pairs()
        t = torch.tensor([int(self.num_timesteps * ratio_)], device=x_start.device).long()
        if fixed_noise:
            torch.manual_seed(1004)          # ddpm.py:425
            noise = self._randn_like(x_start).contiguous()
        else:
            noise = self._randn_like(x_start)
        x_noisy = self.q_sample(x_start=x_start, t=t, noise=noise)
        pairs = pairs[int(len(pairs) * (1 - ratio_)):]
        return self._loop(x_noisy, cond, image_cond, context, pairs)

    @torch.no_grad()
    def sample(self, batch_size=16, cond=None, image_cond=None, context=None, return_intermediates=False,
               noised_start=None, first_stage_model=None, ratio_=None, fix_noise=False):
        """ddpm.py:457-484."""
        shape = (batch_size, self.channels, self.image_size)
        if not self.is_ddim_sampling:
            raise NotImplementedError(
                "sampling_timesteps == timesteps selects the reference's p_sample_loop, which calls the model "
                "with a signature it does not have (ddpm.py:299); MoDiTalker always samples with DDIM (eta=1)"
            )
        if noised_start is not None:
            return self.ddim_sample_noised_start(shape, noised_start, cond, image_cond, context, ratio_=ratio_,
                                                 fixed_noise=fix_noise)
        return self.ddim_sample(shape, cond, image_cond, context)
