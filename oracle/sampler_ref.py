"""ORACLE (tests only): CFG + DDIM (eta = 0) update restated from diffusers' DDIMScheduler.step / StableDiffusionPipeline
(the un-vendored dependency behind train_text_to_image_control_lora.py:829-843; scheduler config of SD-1.5:
scaled_linear betas 0.00085 -> 0.012, 1000 steps, steps_offset 1, set_alpha_to_one False, epsilon prediction, no clipping),
and the DPM-Solver++(2M) multistep update of diffusers-0.13's DPMSolverMultistepScheduler with its defaults
(algorithm_type "dpmsolver++", solver_order 2, solver_type "midpoint", lower_order_final True, no thresholding) - the
scheduler the reference actually swaps in for validation / inference (train_text_to_image_control_lora.py:817-823,
mix_lora_and_control_lora.py:80, apps/gradio_canny2image.py)."""
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


# ---------------------------------------------------------------------------------------------- DPM-Solver++ (2M)
class DPMSolverPP2M:
    """Stateful restatement of DPMSolverMultistepScheduler (diffusers 0.13): set_timesteps + step for epsilon prediction."""

    def __init__(self, num_inference_steps, n=1000, b0=0.00085, b1=0.012, lower_order_final=True):
        import numpy as np

        ac = alphas_cumprod(n, b0, b1)
        self.alpha_t = ac.sqrt()
        self.sigma_t = (1 - ac).sqrt()
        self.lambda_t = self.alpha_t.log() - self.sigma_t.log()
        ts = np.linspace(0, n - 1, num_inference_steps + 1).round()[::-1][:-1].copy().astype(np.int64)
        self.timesteps = [int(v) for v in ts]
        self.model_outputs = [None, None]
        self.lower_order_nums = 0
        self.lower_order_final = lower_order_final

    def _first_order(self, m0, s, t, x):
        lt, ls = self.lambda_t[t], self.lambda_t[s]
        h = lt - ls
        return (self.sigma_t[t] / self.sigma_t[s]) * x - (self.alpha_t[t] * (torch.exp(-h) - 1.0)) * m0

    def _second_order(self, m0, m1, s1, s0, t, x):
        lt, ls0, ls1 = self.lambda_t[t], self.lambda_t[s0], self.lambda_t[s1]
        h, h0 = lt - ls0, ls0 - ls1
        r0 = h0 / h
        d0, d1 = m0, (1.0 / r0) * (m0 - m1)
        c = self.alpha_t[t] * (torch.exp(-h) - 1.0)
        return (self.sigma_t[t] / self.sigma_t[s0]) * x - c * d0 - 0.5 * c * d1

    def step(self, eps, timestep, x):
        """eps: the (already guidance-combined) noise prediction at `timestep`; x: current sample (fp64 math inside)."""
        i = self.timesteps.index(int(timestep))
        n = len(self.timesteps)
        prev = 0 if i == n - 1 else self.timesteps[i + 1]
        final = (i == n - 1) and self.lower_order_final and n < 15
        xd, ed = x.double(), eps.double()
        x0 = (xd - self.sigma_t[timestep] * ed) / self.alpha_t[timestep]
        self.model_outputs = [self.model_outputs[1], x0]
        if self.lower_order_nums < 1 or final:
            out = self._first_order(x0, int(timestep), prev, xd)
        else:
            out = self._second_order(x0, self.model_outputs[0], self.timesteps[i - 1], int(timestep), prev, xd)
        if self.lower_order_nums < 2:
            self.lower_order_nums += 1
        return out.to(x.dtype)


def cfg_combine(eps_uncond, eps_cond, guidance):
    return eps_uncond + guidance * (eps_cond - eps_uncond)


# ---------------------------------------------------------------------------------------------- training-step glue
def add_noise(x0, noise, t, ac=None):
    """diffusers DDPMScheduler.add_noise (train_text_to_image_control_lora.py:765): sqrt(ac[t]) x0 + sqrt(1 - ac[t]) noise."""
    ac = alphas_cumprod() if ac is None else ac
    shape = (-1,) + (1,) * (x0.dim() - 1)
    sa = ac[t.long()].sqrt().to(x0.dtype).view(shape)
    sb = (1 - ac[t.long()]).sqrt().to(x0.dtype).view(shape)
    return sa * x0 + sb * noise


def get_velocity(x0, noise, t, ac=None):
    """diffusers DDPMScheduler.get_velocity (train_text_to_image_control_lora.py:777): sqrt(ac) noise - sqrt(1 - ac) x0."""
    ac = alphas_cumprod() if ac is None else ac
    shape = (-1,) + (1,) * (x0.dim() - 1)
    sa = ac[t.long()].sqrt().to(x0.dtype).view(shape)
    sb = (1 - ac[t.long()]).sqrt().to(x0.dtype).view(shape)
    return sa * noise - sb * x0


def philox4x32_10(ctr, key):
    """Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11; Random123).
    ctr: uint32 array [..., 4], key: (k0, k1).  Returns uint32 [..., 4].  Pinned by Random123's known-answer vectors
    (tests/test_oracle.py)."""
    import numpy as np

    c = [np.asarray(ctr[..., i], dtype=np.uint64) for i in range(4)]
    k0, k1 = np.uint64(key[0]), np.uint64(key[1])
    M0, M1, MASK = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        k0 = (k0 + np.uint64(0x9E3779B9)) & MASK
        k1 = (k1 + np.uint64(0xBB67AE85)) & MASK
    return np.stack([v.astype(np.uint32) for v in c], -1)


def device_noise(seed, step, B, per_image, num_train_timesteps=1000):
    """Restatement of the counter layout of csrc/noise.cu: returns (noise float32 [B, per_image], timesteps int [B])."""
    import numpy as np

    key = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    hi = ((step >> 32) & 0xFFFFFFFF) << 8
    cb = np.zeros((B, 4), dtype=np.uint32)
    cb[:, 0] = np.arange(B, dtype=np.uint32)
    cb[:, 2] = np.uint32((1 + hi) & 0xFFFFFFFF)
    cb[:, 3] = np.uint32(step & 0xFFFFFFFF)
    rt = philox4x32_10(cb, key)[:, 0].astype(np.uint64)
    ts = ((rt * np.uint64(num_train_timesteps)) >> np.uint64(32)).astype(np.int64)
    groups = B * per_image // 4
    cg = np.zeros((groups, 4), dtype=np.uint32)
    gi = np.arange(groups, dtype=np.uint64)
    cg[:, 0] = (gi & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    cg[:, 1] = (gi >> np.uint64(32)).astype(np.uint32)
    cg[:, 2] = np.uint32(hi & 0xFFFFFFFF)
    cg[:, 3] = np.uint32(step & 0xFFFFFFFF)
    r = philox4x32_10(cg, key)
    u = ((r >> np.uint32(8)).astype(np.float32) + np.float32(1.0)) * np.float32(1.0 / 16777216.0)
    out = np.empty((groups, 4), dtype=np.float32)
    for h in range(2):
        rad = np.sqrt(np.float32(-2.0) * np.log(u[:, 2 * h]))
        ang = np.float32(2.0) * u[:, 2 * h + 1].astype(np.float64) * np.pi
        out[:, 2 * h] = (rad * np.cos(ang)).astype(np.float32)
        out[:, 2 * h + 1] = (rad * np.sin(ang)).astype(np.float32)
    return out.reshape(B, per_image), ts
