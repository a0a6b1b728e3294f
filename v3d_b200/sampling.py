"""Sampling-side drop-ins for the sgm plugin surface (reference sgm/modules/diffusionmodules/):
`EDMDiscretization` (discretizer.py:18-39), `LinearPredictionGuider` (guiders.py:60-101),
`VScalingWithEDMcNoise` + `Denoiser` (denoiser_scaling.py:51-59, denoiser.py:12-39), `OpenAIWrapper`
(wrappers.py:9-34), `EulerEDMSampler` (sampling.py:24-133,214-218) and - SURVEY 8(f)-3, same denoiser, no new
tensor-core work - `HeunEDMSampler` (sampling.py:221-237), `VanillaCFG` (guiders.py:23-42) and
`CentralPredictionGuider` (guiders.py:104-146).

Call shapes are the reference's: `sampler(denoiser, x, cond=c, uc=uc) -> x`, `denoiser(network, input, sigma,
cond, **kw)`, `network(x, t, c, **kw)`; any denoiser closure works (scripts/pub/V3D_512.py:278-283).  The
per-step elementwise arithmetic runs in this package's own fp32 kernels (csrc/sampler.cu); the sigma schedule
lives on the host, so the loop has no device->host sync (the reference syncs once per step, sampling.py:118-122).
"""
from __future__ import annotations

import importlib
from typing import Dict, List, Optional, Tuple, Union

import torch
import torch.nn as nn

from . import ops


def get_obj_from_str(string: str):
    module, cls = string.rsplit(".", 1)
    return getattr(importlib.import_module(module, package=None), cls)


def instantiate_from_config(config):
    """Same contract as sgm.util.instantiate_from_config (sgm/util.py:170-187)."""
    if "target" not in config:
        raise KeyError("Expected key `target` to instantiate.")
    return get_obj_from_str(config["target"])(**config.get("params", dict()))


class EDMDiscretization:
    """Karras rho-schedule; `__call__` appends the final 0 (discretizer.py:18-39)."""

    def __init__(self, sigma_min: float = 0.002, sigma_max: float = 80.0, rho: float = 7.0):
        self.sigma_min, self.sigma_max, self.rho = sigma_min, sigma_max, rho

    def get_sigmas(self, n: int, device="cpu") -> torch.Tensor:
        # computed on the host in fp32 exactly like the reference's torch.linspace expression
        ramp = torch.linspace(0, 1, n, device="cpu")
        min_inv_rho = self.sigma_min ** (1 / self.rho)
        max_inv_rho = self.sigma_max ** (1 / self.rho)
        return ((max_inv_rho + ramp * (min_inv_rho - max_inv_rho)) ** self.rho).to(device)

    def __call__(self, n: int, do_append_zero: bool = True, device="cpu", flip: bool = False) -> torch.Tensor:
        sigmas = self.get_sigmas(n, device=device)
        if do_append_zero:
            sigmas = torch.cat([sigmas, sigmas.new_zeros([1])])
        return sigmas if not flip else torch.flip(sigmas, (0,))


class _FrameScaleGuider:
    """Shared machinery of the classifier-free guiders: batch order [uc; c], `x_u + scale_t (x_c - x_u)` with one
    scale per frame of a video (`self.scale`, shape [1, T]); the combine runs in v3d_cfg_combine."""

    additional_cond_keys: List[str] = []
    num_frames: int = 1

    def _init_keys(self, additional_cond_keys) -> None:
        if additional_cond_keys is None:
            additional_cond_keys = []
        if isinstance(additional_cond_keys, str):
            additional_cond_keys = [additional_cond_keys]
        self.additional_cond_keys = additional_cond_keys
        self._scale_dev: Optional[Tuple[torch.Tensor, torch.Tensor]] = None

    def _scale_on(self, device) -> torch.Tensor:
        # `self.scale` may be re-assigned by callers (scripts/pub/V3D_512.py:84-100); re-upload only when it changes
        # (re-assignment or an in-place edit such as `guider.scale[:, i] = ...`: the tensor's version counter moves)
        scale = self._scale_tensor()
        tag = (id(scale), scale._version)
        if self._scale_dev is None or self._scale_dev[0] != tag or self._scale_dev[1].device != device:
            self._scale_dev = (tag, scale.reshape(-1).float().to(device).contiguous())
        return self._scale_dev[1]

    def _scale_tensor(self) -> torch.Tensor:
        return self.scale

    def __call__(self, x: torch.Tensor, sigma: torch.Tensor) -> torch.Tensor:
        T = self.num_frames
        n2 = x.shape[0]
        assert n2 % (2 * T) == 0, "guider expects [uc; c] halves of whole videos"
        half = n2 // 2
        per = x[0].numel()
        x = x.float().contiguous()
        out = torch.empty((half,) + tuple(x.shape[1:]), device=x.device, dtype=torch.float32)
        ops.cfg_combine(x, self._scale_on(x.device), out, half // T, T, per)
        return out

    def prepare_inputs(self, x: torch.Tensor, s: torch.Tensor, c: dict, uc: dict):
        c_out = dict()
        for k in c:
            if k in ["vector", "crossattn", "concat"] + self.additional_cond_keys:
                c_out[k] = torch.cat((uc[k], c[k]), 0)
            else:
                if k == "rgb":
                    continue
                assert c[k] == uc[k]
                c_out[k] = c[k]
        return torch.cat([x] * 2), torch.cat([s] * 2), c_out


class LinearPredictionGuider(_FrameScaleGuider):
    """Per-frame CFG scale linspace(min, max, T); batch order [uc; c] (guiders.py:60-101)."""

    def __init__(self, max_scale: float, num_frames: int, min_scale: float = 1.0,
                 additional_cond_keys: Optional[Union[List[str], str]] = None):
        self.min_scale, self.max_scale, self.num_frames = min_scale, max_scale, num_frames
        self.scale = torch.linspace(min_scale, max_scale, num_frames, device="cpu").unsqueeze(0)
        self._init_keys(additional_cond_keys)


class CentralPredictionGuider(_FrameScaleGuider):
    """Triangular per-frame scale: linspace(min, 2 max, T) mirrored about the middle frame (guiders.py:104-146)."""

    def __init__(self, max_scale: float, num_frames: int, min_scale: float = 1.0,
                 additional_cond_keys: Optional[Union[List[str], str]] = None):
        self.min_scale, self.max_scale, self.num_frames = min_scale, max_scale, num_frames
        scale = torch.linspace(min_scale, 2 * max_scale, num_frames, device="cpu")
        scale[num_frames // 2:] = 2 * max_scale - scale[num_frames // 2:]
        self.scale = scale.unsqueeze(0)
        self._init_keys(additional_cond_keys)


class VanillaCFG(_FrameScaleGuider):
    """One scale for every sample: `x_u + scale (x_c - x_u)` (guiders.py:23-42)."""

    def __init__(self, scale: float):
        self.num_frames = 1
        self.scale_value = float(scale)
        self._init_keys(None)

    # the reference keeps a python float in `.scale` (guiders.py:24-25) and so does the attribute here; the [1, 1]
    # tensor the combine kernel reads is derived from it
    @property
    def scale(self) -> float:
        return float(self._scale_t.reshape(-1)[0])

    def _scale_tensor(self) -> torch.Tensor:
        return self._scale_t

    @scale.setter
    def scale(self, v):
        self._scale_t = v if isinstance(v, torch.Tensor) else torch.tensor([[float(v)]])

    @property
    def scale_value(self) -> float:
        return float(self._scale_t.reshape(-1)[0])

    @scale_value.setter
    def scale_value(self, v: float) -> None:
        self.scale = v


class VScalingWithEDMcNoise:
    """c_skip, c_out, c_in, c_noise of denoiser_scaling.py:51-59 (kept for API parity; the Denoiser below
    evaluates the same expressions inside its fused kernels)."""

    def __call__(self, sigma: torch.Tensor):
        c_skip = 1.0 / (sigma ** 2 + 1.0)
        c_out = -sigma / (sigma ** 2 + 1.0) ** 0.5
        c_in = 1.0 / (sigma ** 2 + 1.0) ** 0.5
        c_noise = 0.25 * sigma.log()
        return c_skip, c_out, c_in, c_noise


class Denoiser(nn.Module):
    """`network(input * c_in, c_noise, cond) * c_out + input * c_skip` (denoiser.py:23-39)."""

    def __init__(self, scaling_config: Dict):
        super().__init__()
        self.scaling = instantiate_from_config(scaling_config)
        if not isinstance(self.scaling, VScalingWithEDMcNoise):
            raise NotImplementedError("only VScalingWithEDMcNoise is implemented on the B200 path")

    def possibly_quantize_sigma(self, sigma: torch.Tensor) -> torch.Tensor:
        return sigma

    def possibly_quantize_c_noise(self, c_noise: torch.Tensor) -> torch.Tensor:
        return c_noise

    def forward(self, network: nn.Module, input: torch.Tensor, sigma: torch.Tensor, cond: Dict,
                **additional_model_inputs) -> torch.Tensor:
        sigma = self.possibly_quantize_sigma(sigma).float().contiguous()
        x = input.float().contiguous()
        n = x.shape[0]
        per = x[0].numel()
        assert sigma.numel() == n
        scaled = torch.empty_like(x)
        c_noise = torch.empty(n, device=x.device, dtype=torch.float32)
        ops.edm_scale_input(x, sigma, scaled, c_noise, n, per)
        c_noise = self.possibly_quantize_c_noise(c_noise.reshape(sigma.shape))
        net = network(scaled, c_noise, cond, **additional_model_inputs).float().contiguous()
        out = torch.empty_like(x)
        ops.edm_denoise_combine(net, x, sigma, out, n, per)
        return out


class IdentityWrapper(nn.Module):
    def __init__(self, diffusion_model, compile_model: bool = False):
        super().__init__()
        # compile_model is accepted and ignored: the network is already a fixed schedule of native kernels
        self.diffusion_model = diffusion_model

    def forward(self, *args, **kwargs):
        return self.diffusion_model(*args, **kwargs)


class OpenAIWrapper(IdentityWrapper):
    """wrappers.py:23-34: x = cat(x, c['concat']); context <- crossattn; y <- vector."""

    def forward(self, x: torch.Tensor, t: torch.Tensor, c: dict, **kwargs) -> torch.Tensor:
        if "concat" in c:
            x = torch.cat((x, c["concat"].type_as(x)), dim=1)
        return self.diffusion_model(x, timesteps=t, context=c.get("crossattn", None), y=c.get("vector", None),
                                    **kwargs)


DEFAULT_GUIDER = {"target": "v3d_b200.sampling.IdentityGuider"}


class IdentityGuider:
    def __call__(self, x, sigma):
        return x

    def prepare_inputs(self, x, s, c, uc):
        return x, s, {k: c[k] for k in c}


class EulerEDMSampler:
    """EDM Euler sampler (sampling.py:24-133,214-218). s_churn > 0 (noise injection) keeps the reference
    semantics via torch.randn_like; V3D_512 runs with s_churn = 0."""

    def __init__(self, discretization_config, num_steps: Optional[int] = None, guider_config=None,
                 verbose: bool = False, device: str = "cuda", s_churn: float = 0.0, s_tmin: float = 0.0,
                 s_tmax: float = float("inf"), s_noise: float = 1.0):
        self.num_steps = num_steps
        self.discretization = instantiate_from_config(discretization_config)
        self.guider = instantiate_from_config(guider_config if guider_config is not None else DEFAULT_GUIDER)
        self.verbose = verbose
        self.device = device
        self.s_churn, self.s_tmin, self.s_tmax, self.s_noise = s_churn, s_tmin, s_tmax, s_noise

    def prepare_sampling_loop(self, x, cond, uc=None, num_steps=None):
        sigmas = self.discretization(self.num_steps if num_steps is None else num_steps, device="cpu")
        uc = cond if uc is None else uc
        x *= float(torch.sqrt(1.0 + sigmas[0] ** 2.0))  # in place on the caller's noise, like sampling.py:50
        return x, sigmas, len(sigmas), cond, uc

    def denoise(self, x, denoiser, sigma, cond, uc):
        denoised = denoiser(*self.guider.prepare_inputs(x, sigma, cond, uc))
        return self.guider(denoised, sigma)

    def possible_correction_step(self, euler_step, x, d, dt, next_sigma, denoiser, cond, uc):
        return euler_step

    def sampler_step(self, sigma, next_sigma, denoiser, x, cond, uc=None, gamma: float = 0.0):
        sigma_hat = sigma if gamma == 0 else sigma * (gamma + 1.0)
        if gamma > 0:
            eps = torch.randn_like(x) * self.s_noise
            x = x + eps * (sigma_hat ** 2 - sigma ** 2).reshape(-1, *([1] * (x.ndim - 1))) ** 0.5
        denoised = self.denoise(x, denoiser, sigma_hat, cond, uc).float().contiguous()
        out = torch.empty_like(x)
        ops.euler_step(x, denoised, sigma_hat, next_sigma, out, x.shape[0], x[0].numel())
        return out

    def __call__(self, denoiser, x, cond, uc=None, num_steps=None):
        if not x.is_cuda:
            raise RuntimeError("v3d_b200.EulerEDMSampler needs CUDA tensors; there is no CPU fallback")
        return self._loop(denoiser, x, cond, uc, num_steps)

    def _loop(self, denoiser, x, cond, uc=None, num_steps=None):
        """EDMSampler.__call__ (sampling.py:112-133) below the CUDA check."""
        assert x.dtype == torch.float32, "sampler state is fp32 (denoiser.py:36-39)"
        x, sigmas, num_sigmas, cond, uc = self.prepare_sampling_loop(x, cond, uc, num_steps)
        n = x.shape[0]
        # per-step sigma vectors built once (one small H2D copy), indexed on the host
        sig_rows = sigmas.float().reshape(-1, 1).expand(num_sigmas, n).contiguous().to(x.device)
        host = [float(s) for s in sigmas]
        steps = range(num_sigmas - 1)
        if self.verbose:
            from tqdm import tqdm

            steps = tqdm(steps, total=num_sigmas - 1, desc=f"Sampling with {type(self).__name__}")
        x = x.contiguous()
        for i in steps:
            gamma = min(self.s_churn / (num_sigmas - 1), 2 ** 0.5 - 1) if self.s_tmin <= host[i] <= self.s_tmax else 0.0
            x = self.sampler_step(sig_rows[i], sig_rows[i + 1], denoiser, x, cond, uc, gamma)
        return x


class HeunEDMSampler(EulerEDMSampler):
    """Second-order EDM sampler (sampling.py:221-237): after the Euler proposal the denoiser is evaluated once more at
    sigma_next and the step is redone with the averaged slope; the last step (sigma_next = 0) stays first order, which
    also saves the network evaluation exactly as the reference does (`torch.sum(next_sigma) < 1e-14`)."""

    _next_sigma_host: Optional[float] = None  # set by __call__: lets the step skip the device->host sync

    def sampler_step(self, sigma, next_sigma, denoiser, x, cond, uc=None, gamma: float = 0.0):
        sigma_hat = sigma if gamma == 0 else sigma * (gamma + 1.0)
        if gamma > 0:
            eps = torch.randn_like(x) * self.s_noise
            x = x + eps * (sigma_hat ** 2 - sigma ** 2).reshape(-1, *([1] * (x.ndim - 1))) ** 0.5
        n, per = x.shape[0], x[0].numel()
        denoised = self.denoise(x, denoiser, sigma_hat, cond, uc).float().contiguous()
        x_euler = torch.empty_like(x)
        ops.euler_step(x, denoised, sigma_hat, next_sigma, x_euler, n, per)
        nsum = self._next_sigma_host if self._next_sigma_host is not None else float(next_sigma.sum())
        if nsum < 1e-14:
            return x_euler
        denoised2 = self.denoise(x_euler, denoiser, next_sigma, cond, uc).float().contiguous()
        out = torch.empty_like(x)
        ops.heun_step(x, denoised, x_euler, denoised2, sigma_hat.float().contiguous(),
                      next_sigma.float().contiguous(), out, n, per)
        return out

    def __call__(self, denoiser, x, cond, uc=None, num_steps=None):
        if not x.is_cuda:
            raise RuntimeError("v3d_b200.HeunEDMSampler needs CUDA tensors; there is no CPU fallback")
        return self._loop(denoiser, x, cond, uc, num_steps)

    def _loop(self, denoiser, x, cond, uc=None, num_steps=None):
        assert x.dtype == torch.float32, "sampler state is fp32 (denoiser.py:36-39)"
        x, sigmas, num_sigmas, cond, uc = self.prepare_sampling_loop(x, cond, uc, num_steps)
        n = x.shape[0]
        sig_rows = sigmas.float().reshape(-1, 1).expand(num_sigmas, n).contiguous().to(x.device)
        host = [float(s) for s in sigmas]
        x = x.contiguous()
        try:
            for i in range(num_sigmas - 1):
                gamma = (min(self.s_churn / (num_sigmas - 1), 2 ** 0.5 - 1)
                         if self.s_tmin <= host[i] <= self.s_tmax else 0.0)
                self._next_sigma_host = host[i + 1] * n
                x = self.sampler_step(sig_rows[i], sig_rows[i + 1], denoiser, x, cond, uc, gamma)
        finally:
            self._next_sigma_host = None
        return x

