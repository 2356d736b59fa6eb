"""ORACLE (test infrastructure only — never imported by the product path).

CPU / fp32 restatement of diffusers-0.13 `AutoencoderKL` (SD-1.5 `vae/` config) for the two calls the reference makes
either side of the hot path:

  train_text_to_image_control_lora.py:753-754   latents = vae.encode(pixel_values).latent_dist.sample() * vae.config.scaling_factor
  StableDiffusionPipeline (…:829-843, apps/*)    image = vae.decode(latents / scaling_factor).sample

SD-1.5 config: in/out 3 channels, latent 4, block_out_channels (128, 256, 512, 512), layers_per_block 2, norm_num_groups 32,
act silu, scaling_factor 0.18215.  Restated pieces (diffusers `models/vae.py`, `unet_2d_blocks.py`, `resnet.py`,
`attention.py` of 0.13):
  Encoder  conv_in 3x3 -> 4 x DownEncoderBlock2D (2 ResnetBlock2D each, no time embedding, GroupNorm eps 1e-6;
           Downsample2D(padding=0): F.pad(x, (0,1,0,1)) + conv3x3 stride 2, on all but the last block) -> UNetMidBlock2D
           (resnet, AttentionBlock with ONE head of width 512, resnet) -> GroupNorm(32, eps 1e-6) -> SiLU -> conv_out 3x3 (-> 8)
  quant_conv 1x1 (8 -> 8); DiagonalGaussianDistribution: mean, logvar = chunk(2); logvar.clamp(-30, 20); sample = mean + exp(0.5 logvar) eps
  post_quant_conv 1x1 (4 -> 4)
  Decoder  conv_in 3x3 (4 -> 512) -> UNetMidBlock2D -> 4 x UpDecoderBlock2D (3 ResnetBlock2D each; Upsample2D = nearest x2 + conv3x3
           on all but the last) -> GroupNorm -> SiLU -> conv_out 3x3 (-> 3)
  AttentionBlock: h = GroupNorm(x); q,k,v = Linear(h) (with bias); softmax(q k^T / sqrt(C)) v; proj_attn; + x  (rescale factor 1)

PARITY STATUS: Encoder and Decoder are **pinned** to an independent published implementation of the same network - the CompVis
latent-diffusion / taming `Encoder` / `Decoder` (the code SD-1.5's VAE comes from; `transformers` ships it as JanusVQVAEEncoder /
JanusVQVAEDecoder): same weights -> same outputs to 1e-4 (tests/test_oracle.py).  diffusers itself is not installable here and no
SD-1.5 VAE weights are on disk, so the thin AutoencoderKL wrapper around them (quant_conv / post_quant_conv 1x1 convs, the
DiagonalGaussianDistribution clamp + sample, scaling_factor) is anchored by the parameter count of the SD-1.5 VAE (83 653 863) and
its state-dict key names only.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

SD15_VAE_CONFIG = dict(in_channels=3, out_channels=3, latent_channels=4, block_out_channels=(128, 256, 512, 512),
                       layers_per_block=2, norm_num_groups=32, scaling_factor=0.18215)


class ResnetBlock2D(nn.Module):
    """diffusers ResnetBlock2D with temb_channels=None, eps 1e-6, output_scale_factor 1."""

    def __init__(self, cin, cout, groups=32, eps=1e-6):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class AttentionBlock(nn.Module):
    """diffusers-0.13 AttentionBlock(channels, num_head_channels=None -> one head, norm_num_groups, eps 1e-6)."""

    def __init__(self, channels, groups=32, eps=1e-6):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, channels, eps=eps)
        self.query = nn.Linear(channels, channels)
        self.key = nn.Linear(channels, channels)
        self.value = nn.Linear(channels, channels)
        self.proj_attn = nn.Linear(channels, channels)
        self.channels = channels

    def forward(self, x):
        b, c, hh, ww = x.shape
        h = self.group_norm(x).view(b, c, hh * ww).transpose(1, 2)
        q, k, v = self.query(h), self.key(h), self.value(h)
        scale = 1.0 / math.sqrt(self.channels)
        p = torch.softmax(torch.baddbmm(torch.empty(b, hh * ww, hh * ww, dtype=q.dtype, device=q.device), q, k.transpose(1, 2),
                                        beta=0, alpha=scale).float(), dim=-1).to(q.dtype)
        h = self.proj_attn(torch.bmm(p, v))
        return h.transpose(1, 2).reshape(b, c, hh, ww) + x


class Downsample2D(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=0)

    def forward(self, x):
        return self.conv(F.pad(x, (0, 1, 0, 1)))


class Upsample2D(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class _Block(nn.Module):
    def __init__(self, cin, cout, layers, groups, down=False, up=False):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, groups) for i in range(layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(cout)]) if down else None
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if up else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class MidBlock(nn.Module):
    def __init__(self, c, groups):
        super().__init__()
        self.attentions = nn.ModuleList([AttentionBlock(c, groups)])
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c, groups), ResnetBlock2D(c, c, groups)])

    def forward(self, x):
        x = self.resnets[0](x)
        x = self.attentions[0](x)
        return self.resnets[1](x)


class Encoder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        ch, g, L = list(cfg["block_out_channels"]), cfg["norm_num_groups"], cfg["layers_per_block"]
        self.conv_in = nn.Conv2d(cfg["in_channels"], ch[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        out = ch[0]
        for i, c in enumerate(ch):
            inp, out = out, c
            self.down_blocks.append(_Block(inp, out, L, g, down=i != len(ch) - 1))
        self.mid_block = MidBlock(ch[-1], g)
        self.conv_norm_out = nn.GroupNorm(g, ch[-1], eps=1e-6)
        self.conv_out = nn.Conv2d(ch[-1], 2 * cfg["latent_channels"], 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for b in self.down_blocks:
            x = b(x)
        x = self.mid_block(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class Decoder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        ch, g, L = list(cfg["block_out_channels"]), cfg["norm_num_groups"], cfg["layers_per_block"]
        rev = list(reversed(ch))
        self.conv_in = nn.Conv2d(cfg["latent_channels"], ch[-1], 3, padding=1)
        self.mid_block = MidBlock(ch[-1], g)
        self.up_blocks = nn.ModuleList()
        out = rev[0]
        for i, c in enumerate(rev):
            inp, out = out, c
            self.up_blocks.append(_Block(inp, out, L + 1, g, up=i != len(ch) - 1))
        self.conv_norm_out = nn.GroupNorm(g, ch[0], eps=1e-6)
        self.conv_out = nn.Conv2d(ch[0], cfg["out_channels"], 3, padding=1)

    def forward(self, z):
        x = self.mid_block(self.conv_in(z))
        for b in self.up_blocks:
            x = b(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class AutoencoderKL(nn.Module):
    def __init__(self, **overrides):
        super().__init__()
        cfg = dict(SD15_VAE_CONFIG)
        cfg.update(overrides)
        self.config = cfg
        self.encoder = Encoder(cfg)
        self.decoder = Decoder(cfg)
        self.quant_conv = nn.Conv2d(2 * cfg["latent_channels"], 2 * cfg["latent_channels"], 1)
        self.post_quant_conv = nn.Conv2d(cfg["latent_channels"], cfg["latent_channels"], 1)

    def encode_moments(self, x):
        """(mean, logvar) of DiagonalGaussianDistribution: logvar clamped to [-30, 20]."""
        mean, logvar = self.quant_conv(self.encoder(x)).chunk(2, dim=1)
        return mean, logvar.clamp(-30.0, 20.0)

    def encode_sample(self, x, eps: Optional[torch.Tensor] = None):
        """`vae.encode(x).latent_dist.sample() * scaling_factor` with the caller's standard-normal eps (train_...:753-754)."""
        mean, logvar = self.encode_moments(x)
        eps = torch.randn_like(mean) if eps is None else eps
        return (mean + torch.exp(0.5 * logvar) * eps) * self.config["scaling_factor"]

    def decode(self, latents):
        """`vae.decode(latents / scaling_factor).sample`."""
        return self.decoder(self.post_quant_conv(latents / self.config["scaling_factor"]))


def init_synthetic_(model: nn.Module, seed: int = 0) -> nn.Module:
    """Seeded synthetic weights: W ~ N(0, 1/fan_in), biases 0.02 N, norm gamma 1 + 0.1 N, beta 0.1 N."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.dim() >= 2:
                p.copy_(torch.randn(p.shape, generator=g) / math.sqrt(p[0].numel()))
            elif "norm" in name and name.endswith("weight"):
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
            else:
                p.copy_((0.1 if "norm" in name else 0.02) * torch.randn(p.shape, generator=g))
    return model
