"""ORACLE (tests only): CFG + DDIM (eta = 0) update restated from diffusers' DDIMScheduler.step / StableDiffusionPipeline
(the un-vendored dependency behind train_text_to_image_control_lora.py:829-843; scheduler config of SD-1.5:
scaled_linear betas 0.00085 -> 0.012, 1000 steps, steps_offset 1, set_alpha_to_one False, epsilon prediction, no clipping)."""
import torch


def alphas_cumprod(n=1000, b0=0.00085, b1=0.012):
    betas = torch.linspace(b0 ** 0.5, b1 ** 0.5, n, dtype=torch.float64) ** 2
    return torch.cumprod(1.0 - betas, 0)


def timesteps(num_inference_steps, n=1000, offset=1):
    ratio = n // num_inference_steps
    return (torch.arange(0, num_inference_steps) * ratio).flip(0) + offset


def cfg_ddim_step(eps_uncond, eps_cond, x, t, num_inference_steps, guidance, ac=None):
    ac = alphas_cumprod() if ac is None else ac
    eps = eps_uncond + guidance * (eps_cond - eps_uncond)
    prev = int(t) - 1000 // num_inference_steps
    a_t = ac[int(t)]
    a_p = ac[prev] if prev >= 0 else ac[0]
    x0 = (x - (1 - a_t).sqrt() * eps) / a_t.sqrt()
    return (a_p.sqrt() * x0 + (1 - a_p).sqrt() * eps).to(x.dtype)
