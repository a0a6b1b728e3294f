"""fp32 CPU restatement of the EDM sampling loop of V3D_512 (TEST INFRASTRUCTURE).

EDMDiscretization (discretizer.py:18-39), LinearPredictionGuider (guiders.py:60-101), Denoiser with
VScalingWithEDMcNoise (denoiser.py:23-39, denoiser_scaling.py:51-59), EulerEDMSampler
(sampling.py:44-55,96-133,214-218; sampling_utils.py:34-35); HeunEDMSampler (sampling.py:221-237), VanillaCFG
(guiders.py:23-42), CentralPredictionGuider (guiders.py:104-146).
"""
from __future__ import annotations

from typing import Callable, Dict

import torch


def edm_sigmas(n: int, sigma_min: float = 0.002, sigma_max: float = 700.0, rho: float = 7.0) -> torch.Tensor:
    """EDMDiscretization.get_sigmas + Discretization.__call__ append_zero (discretizer.py:18-39)."""
    ramp = torch.linspace(0, 1, n)
    min_inv_rho = sigma_min ** (1 / rho)
    max_inv_rho = sigma_max ** (1 / rho)
    sigmas = (max_inv_rho + ramp * (min_inv_rho - max_inv_rho)) ** rho
    return torch.cat([sigmas, sigmas.new_zeros([1])])


def guider_scale(min_scale: float, max_scale: float, num_frames: int) -> torch.Tensor:
    return torch.linspace(min_scale, max_scale, num_frames).unsqueeze(0)  # guiders.py:71


def central_guider_scale(min_scale: float, max_scale: float, num_frames: int) -> torch.Tensor:
    """CentralPredictionGuider.__init__ (guiders.py:112-120): ramp to 2*max, second half mirrored."""
    scale = torch.linspace(min_scale, 2 * max_scale, num_frames)
    scale[num_frames // 2:] = 2 * max_scale - scale[num_frames // 2:]
    return scale.unsqueeze(0)


def vanilla_scale(scale: float) -> torch.Tensor:
    """VanillaCFG (guiders.py:23-31) as a one-frame scale row for guider_combine(…, num_frames=1)."""
    return torch.tensor([[float(scale)]])


def guider_prepare_inputs(x, s, c: Dict, uc: Dict):
    """guiders.py:88-101: [uc; c] batch order for vector / crossattn / concat."""
    c_out = {k: torch.cat((uc[k], c[k]), 0) for k in c if k in ("vector", "crossattn", "concat")}
    return torch.cat([x] * 2), torch.cat([s] * 2), c_out


def guider_combine(x, scale: torch.Tensor, num_frames: int):
    """guiders.py:78-86."""
    x_u, x_c = x.chunk(2)
    shp = x_u.shape
    x_u = x_u.reshape(-1, num_frames, *shp[1:])
    x_c = x_c.reshape(-1, num_frames, *shp[1:])
    sc = scale.reshape(1, num_frames, *([1] * (len(shp) - 1)))
    return (x_u + sc * (x_c - x_u)).reshape(shp)


def denoiser(network: Callable, x, sigma, cond: Dict, **kw):
    """Denoiser.forward (denoiser.py:23-39) with VScalingWithEDMcNoise."""
    sig = sigma.reshape(-1, *([1] * (x.ndim - 1)))
    c_skip = 1.0 / (sig ** 2 + 1.0)
    c_out = -sig / (sig ** 2 + 1.0) ** 0.5
    c_in = 1.0 / (sig ** 2 + 1.0) ** 0.5
    c_noise = (0.25 * sig.log()).reshape(sigma.shape)
    return network(x * c_in, c_noise, cond, **kw) * c_out + x * c_skip


def euler_edm_sample(denoise_fn: Callable, x, cond: Dict, uc: Dict, num_steps: int, scale: torch.Tensor,
                     num_frames: int, sigma_max: float = 700.0, trace=None):
    """EulerEDMSampler.__call__ with s_churn=0 (gamma=0): sampling.py:112-133.
    denoise_fn(x2, sigma2, c2) is the closure of scripts/pub/V3D_512.py:278-281."""
    sigmas = edm_sigmas(num_steps, sigma_max=sigma_max)
    x = x * torch.sqrt(1.0 + sigmas[0] ** 2.0)
    s_in = x.new_ones([x.shape[0]])
    for i in range(len(sigmas) - 1):
        sigma_hat = s_in * sigmas[i]
        next_sigma = s_in * sigmas[i + 1]
        den = denoise_fn(*guider_prepare_inputs(x, sigma_hat, cond, uc))
        den = guider_combine(den, scale, num_frames)
        d = (x - den) / sigma_hat.reshape(-1, *([1] * (x.ndim - 1)))
        dt = (next_sigma - sigma_hat).reshape(-1, *([1] * (x.ndim - 1)))
        x = x + dt * d
        if trace is not None:
            trace.append(x.clone())
    return x


def heun_edm_sample(denoise_fn: Callable, x, cond: Dict, uc: Dict, num_steps: int, scale: torch.Tensor,
                    num_frames: int, sigma_max: float = 700.0):
    """HeunEDMSampler.__call__ with s_churn=0: the Euler proposal, then (unless every next sigma is 0) a second
    denoiser evaluation at next_sigma and the step redone with the mean slope (sampling.py:96-133,221-237)."""
    sigmas = edm_sigmas(num_steps, sigma_max=sigma_max)
    x = x * torch.sqrt(1.0 + sigmas[0] ** 2.0)
    s_in = x.new_ones([x.shape[0]])
    bc = lambda v: v.reshape(-1, *([1] * (x.ndim - 1)))  # noqa: E731

    def denoise(xx, sig):
        return guider_combine(denoise_fn(*guider_prepare_inputs(xx, sig, cond, uc)), scale, num_frames)

    for i in range(len(sigmas) - 1):
        sigma_hat = s_in * sigmas[i]
        next_sigma = s_in * sigmas[i + 1]
        d = (x - denoise(x, sigma_hat)) / bc(sigma_hat)
        dt = bc(next_sigma - sigma_hat)
        x_euler = x + dt * d
        if torch.sum(next_sigma) < 1e-14:
            x = x_euler
            continue
        d_new = (x_euler - denoise(x_euler, next_sigma)) / bc(next_sigma)
        x = torch.where(bc(next_sigma) > 0.0, x + (d + d_new) / 2.0 * dt, x_euler)
    return x

