"""Denoise loop on the B200 UNet: classifier-free guidance + DDIM (eta = 0) or DPM-Solver++(2M), the inference call
pattern of the reference (`StableDiffusionPipeline.__call__` as used at train_text_to_image_control_lora.py:829-843 and
apps/gradio_canny2image.py:81-89; BASELINE config 3 = 50-step DDIM at batch 8 -> UNet batch 16; the reference swaps in
`DPMSolverMultistepScheduler` for validation / the apps: train_text_to_image_control_lora.py:817-823,
mix_lora_and_control_lora.py:80).

The control states do not depend on the timestep: `control_lora(guide)` runs once per image batch, every UNet evaluation
of the loop re-uses the injected states.  The guide must be tiled to the CFG batch 2B."""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from . import ops


def sd15_alphas_cumprod(num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012) -> List[float]:
    """SD-1.5 scheduler constants: scaled-linear betas (diffusers `scaled_linear`), cumulative product of 1 - beta."""
    import math

    out, acc = [], 1.0
    s0, s1 = math.sqrt(beta_start), math.sqrt(beta_end)
    for i in range(num_train_timesteps):
        b = (s0 + (s1 - s0) * i / (num_train_timesteps - 1)) ** 2
        acc *= 1.0 - b
        out.append(acc)
    return out


def ddim_timesteps(num_inference_steps: int, num_train_timesteps: int = 1000, steps_offset: int = 1) -> List[int]:
    ratio = num_train_timesteps // num_inference_steps
    return [i * ratio + steps_offset for i in range(num_inference_steps)][::-1]


def ddim_coeffs(t: int, num_inference_steps: int, ac: List[float], num_train_timesteps: int = 1000) -> Tuple[float, float]:
    """(alpha_prod_t, alpha_prod_prev) of diffusers' DDIMScheduler.step with set_alpha_to_one=False."""
    prev = t - num_train_timesteps // num_inference_steps
    return ac[t], (ac[prev] if prev >= 0 else ac[0])


@torch.no_grad()
def ddim_sample(unet, control_lora, guide: torch.Tensor, cond: torch.Tensor, uncond: torch.Tensor, num_inference_steps: int = 50,
                guidance_scale: float = 7.5, latents: Optional[torch.Tensor] = None, seed: int = 0) -> torch.Tensor:
    """guide [B,3,H,W] in [-1,1]; cond / uncond text states [B,77,D] bf16.  Returns the final latents [B,4,H/8,W/8] fp32."""
    B = guide.shape[0]
    dev = guide.device
    h, w = guide.shape[2] // 8, guide.shape[3] // 8
    if latents is None:
        g = torch.Generator(device="cpu").manual_seed(seed)
        latents = torch.randn(B, 4, h, w, generator=g).to(dev)
    latents = latents.float().contiguous().clone()
    control_lora(torch.cat([guide, guide], 0))                     # inject once, for [uncond | cond]
    ehs = torch.cat([uncond, cond], 0).to(torch.bfloat16).contiguous()
    ac = sd15_alphas_cumprod()
    for t in ddim_timesteps(num_inference_steps):
        x2 = torch.cat([latents, latents], 0)
        tt = torch.full((2 * B,), float(t), device=dev)
        eps2 = unet(x2, tt, ehs).sample
        a_t, a_prev = ddim_coeffs(t, num_inference_steps, ac)
        ops.cfg_ddim_step(eps2, latents, guidance_scale, a_t, a_prev)
    return latents


# ---------------------------------------------------------------------------------------------- DPM-Solver++ (2M)
def dpm_timesteps(num_inference_steps: int, num_train_timesteps: int = 1000) -> List[int]:
    """diffusers-0.13 DPMSolverMultistepScheduler.set_timesteps: linspace(0, T-1, N+1).round()[::-1][:-1]."""
    # reproduces numpy bit for bit: linspace = arange(N + 1) * step (step = (T-1)/N in double, last element = T-1 exactly),
    # np.round = Python round = half to even  (e.g. N = 30: 15 * 33.3 = 499.49999999999994 -> 499, not 500)
    step = (num_train_timesteps - 1) / num_inference_steps
    out = [int(round(i * step)) for i in range(num_inference_steps)] + [num_train_timesteps - 1]
    return out[::-1][:-1]


def dpmpp_2m_coeffs(i: int, timesteps: List[int], ac: List[float], lower_order_final: bool = True):
    """Scalars of step i of the multistep solver (epsilon prediction, midpoint, order 2; first step and - for < 15 steps -
    the last step are first order):  x0 = (x - sigma_s eps) / alpha_s;  x <- c_x x + c_m0 x0 + c_m1 x0_prev.
    Returns (alpha_s, sigma_s, c_x, c_m0, c_m1)."""
    import math

    def asl(t):
        a, s = math.sqrt(ac[t]), math.sqrt(1.0 - ac[t])
        return a, s, math.log(a) - math.log(s)

    n = len(timesteps)
    s0 = timesteps[i]
    t = 0 if i == n - 1 else timesteps[i + 1]
    a_t, sg_t, l_t = asl(t)
    a_s, sg_s, l_s = asl(s0)
    h = l_t - l_s
    c = a_t * (math.exp(-h) - 1.0)
    first = i == 0 or (i == n - 1 and lower_order_final and n < 15)
    if first:
        return a_s, sg_s, sg_t / sg_s, -c, 0.0
    _, _, l_s1 = asl(timesteps[i - 1])
    r0 = (l_s - l_s1) / h
    return a_s, sg_s, sg_t / sg_s, -c * (1.0 + 0.5 / r0), 0.5 * c / r0


@torch.no_grad()
def dpmpp_sample(unet, control_lora, guide: torch.Tensor, cond: torch.Tensor, uncond: torch.Tensor, num_inference_steps: int = 30,
                 guidance_scale: float = 7.5, latents: Optional[torch.Tensor] = None, seed: int = 0) -> torch.Tensor:
    """Same contract as ddim_sample with the DPM-Solver++(2M) update; one fused CFG + solver kernel per step."""
    B = guide.shape[0]
    dev = guide.device
    h, w = guide.shape[2] // 8, guide.shape[3] // 8
    if latents is None:
        g = torch.Generator(device="cpu").manual_seed(seed)
        latents = torch.randn(B, 4, h, w, generator=g).to(dev)
    latents = latents.float().contiguous().clone()
    x0_prev = torch.zeros_like(latents)
    control_lora(torch.cat([guide, guide], 0))                     # inject once, for [uncond | cond]
    ehs = torch.cat([uncond, cond], 0).to(torch.bfloat16).contiguous()
    ac = sd15_alphas_cumprod()
    ts = dpm_timesteps(num_inference_steps)
    for i, t in enumerate(ts):
        x2 = torch.cat([latents, latents], 0)
        tt = torch.full((2 * B,), float(t), device=dev)
        eps2 = unet(x2, tt, ehs).sample
        a_s, sg_s, c_x, c_m0, c_m1 = dpmpp_2m_coeffs(i, ts, ac)
        ops.cfg_dpmpp_step(eps2, latents, x0_prev, guidance_scale, a_s, sg_s, c_x, c_m0, c_m1)
    return latents
