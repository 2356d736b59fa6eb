"""ORACLE (test infrastructure only — never imported by the product path).

CPU / fp32 restatement of the *un-vendored* dependency that holds the arithmetic of the ControlLoRA hot path:
`diffusers` 0.13/0.14 `UNet2DConditionModel` with the Stable-Diffusion-1.5 configuration, as called from
/root/reference/train_text_to_image_control_lora.py:407-409 (load), :487 (set_attn_processor) and :782
(`unet(noisy_latents, timesteps, encoder_hidden_states).sample`).  diffusers is pinned only loosely by the reference
(requirements.txt:4 = git main Feb-2023; `check_min_version("0.13.0.dev0")` train_...:49; the import path
`diffusers.models.cross_attention` at models.py:12 exists only in 0.12-0.14) and is NOT installed in this image, so
its published algorithm is restated here from the diffusers-0.13 sources' behaviour:

  * time embedding: sinusoidal (flip_sin_to_cos=True, freq_shift=0, 320 ch) -> Linear(320,1280) -> SiLU -> Linear
  * ResnetBlock2D: conv1(silu(GN(x))) + Linear(silu(temb)) -> conv2(silu(GN(.))) ; 1x1 shortcut iff Cin != Cout
  * Transformer2DModel: GN(eps 1e-6) -> conv1x1 -> [LN->attn1->+ ; LN->attn2(text)->+ ; LN->GEGLU FF->+] -> conv1x1 -> +res
  * CrossAttention: to_q/k/v without bias, to_out = [Linear(bias), Dropout(0)], softmax(q k^T / sqrt(d)) v, 8 heads
  * Downsample2D: conv3x3 stride 2 pad 1 ; Upsample2D: nearest x2 then conv3x3 pad 1
  * topology: 3x CrossAttnDownBlock2D + DownBlock2D, mid (res, attn, res), UpBlock2D + 3x CrossAttnUpBlock2D,
    GN -> SiLU -> conv_out.  Module / parameter names equal diffusers' state-dict keys.

PARITY STATUS: **parity unpinned as a whole** — the reference ships no tests / golden vectors and diffusers cannot
be imported here, so the assembled network could not be checked against diffusers' own outputs.  Its building blocks ARE pinned to
independent implementations of the same published blocks (tests/test_oracle.py): BasicTransformerBlock == torch.nn.TransformerDecoderLayer
(norm_first, GEGLU activation), ResnetBlock2D (time projection zeroed) / Upsample2D / Downsample2D(padding=0) == the CompVis LDM blocks
shipped in transformers, the timestep embedding == its closed form.  Unpinned: the block wiring (skip connections, resolution schedule),
the time-embedding injection, the Transformer2DModel wrapper.  The whole is anchored by
(a) the parameter-count sanity anchors of SURVEY.md §8c (859.08 M weights in conv/linear kernels), (b) the reference's
call sites, and (c) self-consistency tests in tests/test_oracle.py.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from types import SimpleNamespace
from typing import Dict, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F


SD15_CONFIG = dict(
    in_channels=4,
    out_channels=4,
    block_out_channels=(320, 640, 1280, 1280),
    down_block_types=("CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "DownBlock2D"),
    up_block_types=("UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D"),
    layers_per_block=2,
    attention_head_dim=8,  # diffusers quirk: for SD-1.5 this is the NUMBER of heads
    cross_attention_dim=768,
    norm_num_groups=32,
    norm_eps=1e-5,
    flip_sin_to_cos=True,
    freq_shift=0,
)


class LoRALinearLayer(nn.Module):
    """diffusers.models.cross_attention.LoRALinearLayer (imported at models.py:12).

    down ~ N(0, 1/rank), up = 0, no alpha; forward up-casts the input to the weight dtype and casts the result back."""

    def __init__(self, in_features: int, out_features: int, rank: int = 4):
        super().__init__()
        if rank > min(in_features, out_features):
            raise ValueError(f"LoRA rank {rank} must be less or equal than {min(in_features, out_features)}")
        self.down = nn.Linear(in_features, rank, bias=False)
        self.up = nn.Linear(rank, out_features, bias=False)
        nn.init.normal_(self.down.weight, std=1 / rank)
        nn.init.zeros_(self.up.weight)

    def forward(self, x):
        orig = x.dtype
        w = self.down.weight.dtype
        return self.up(self.down(x.to(w))).to(orig)


class PlainAttnProcessor:
    """diffusers CrossAttnProcessor (default, no LoRA)."""

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None):
        q = attn.head_to_batch_dim(attn.to_q(hidden_states))
        ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        k = attn.head_to_batch_dim(attn.to_k(ctx))
        v = attn.head_to_batch_dim(attn.to_v(ctx))
        p = attn.get_attention_scores(q, k, attention_mask)
        h = attn.batch_to_head_dim(torch.bmm(p, v))
        return attn.to_out[1](attn.to_out[0](h))


class CrossAttention(nn.Module):
    """diffusers.models.cross_attention.CrossAttention as used by SD-1.5 (no upcast, no group norm, no added kv)."""

    def __init__(self, query_dim: int, cross_attention_dim: Optional[int], heads: int, dim_head: int):
        super().__init__()
        inner = heads * dim_head
        self.heads = heads
        self.scale = dim_head**-0.5
        kv_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(kv_dim, inner, bias=False)
        self.to_v = nn.Linear(kv_dim, inner, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(0.0)])
        self.processor = PlainAttnProcessor()

    def set_processor(self, processor):
        # diffusers registers nn.Module processors as sub-modules (shared ownership with ControlLoRA.lora_layers)
        if isinstance(processor, nn.Module):
            self._modules.pop("processor", None)
            self.__dict__.pop("processor", None)
            self.add_module("processor", processor)
        else:
            self._modules.pop("processor", None)
            self.__dict__["processor"] = processor

    def prepare_attention_mask(self, attention_mask, target_length, batch_size=None):
        return None if attention_mask is None else attention_mask

    def head_to_batch_dim(self, t):
        b, n, c = t.shape
        h = self.heads
        return t.reshape(b, n, h, c // h).permute(0, 2, 1, 3).reshape(b * h, n, c // h)

    def batch_to_head_dim(self, t):
        bh, n, d = t.shape
        h = self.heads
        return t.reshape(bh // h, h, n, d).permute(0, 2, 1, 3).reshape(bh // h, n, d * h)

    def get_attention_scores(self, q, k, attention_mask=None):
        scores = torch.baddbmm(
            torch.empty(q.shape[0], q.shape[1], k.shape[1], dtype=q.dtype, device=q.device),
            q, k.transpose(-1, -2), beta=0, alpha=self.scale)
        if attention_mask is not None:
            scores = scores + attention_mask
        return scores.softmax(dim=-1)

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **kw):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask, **kw)


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        a, g = self.proj(x).chunk(2, dim=-1)
        return a * F.gelu(g)  # exact erf GELU


class FeedForward(nn.Module):
    def __init__(self, dim, mult=4):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * mult), nn.Dropout(0.0), nn.Linear(dim * mult, dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, heads, dim_head, cross_attention_dim):
        super().__init__()
        self.attn1 = CrossAttention(dim, None, heads, dim_head)
        self.ff = FeedForward(dim)
        self.attn2 = CrossAttention(dim, cross_attention_dim, heads, dim_head)
        self.norm1 = nn.LayerNorm(dim)
        self.norm2 = nn.LayerNorm(dim)
        self.norm3 = nn.LayerNorm(dim)

    def forward(self, h, encoder_hidden_states=None, **kw):
        h = self.attn1(self.norm1(h), encoder_hidden_states=None, **kw) + h
        h = self.attn2(self.norm2(h), encoder_hidden_states=encoder_hidden_states, **kw) + h
        h = self.ff(self.norm3(h)) + h
        return h


class Transformer2DModel(nn.Module):
    def __init__(self, heads, dim_head, in_channels, cross_attention_dim, norm_num_groups=32):
        super().__init__()
        inner = heads * dim_head
        self.norm = nn.GroupNorm(norm_num_groups, in_channels, eps=1e-6, affine=True)
        self.proj_in = nn.Conv2d(in_channels, inner, 1)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(inner, heads, dim_head, cross_attention_dim)])
        self.proj_out = nn.Conv2d(inner, in_channels, 1)

    def forward(self, x, encoder_hidden_states=None, **kw):
        b, c, hh, ww = x.shape
        res = x
        h = self.proj_in(self.norm(x))
        inner = h.shape[1]
        h = h.permute(0, 2, 3, 1).reshape(b, hh * ww, inner)
        for blk in self.transformer_blocks:
            h = blk(h, encoder_hidden_states=encoder_hidden_states, **kw)
        h = h.reshape(b, hh, ww, inner).permute(0, 3, 1, 2).contiguous()
        return self.proj_out(h) + res


class ResnetBlock2D(nn.Module):
    def __init__(self, in_channels, out_channels, temb_channels, groups=32, eps=1e-5):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, in_channels, eps=eps, affine=True)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels)
        self.norm2 = nn.GroupNorm(groups, out_channels, eps=eps, affine=True)
        self.dropout = nn.Dropout(0.0)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 1) if in_channels != out_channels else None

    def forward(self, x, temb):
        h = self.conv1(F.silu(self.norm1(x)))
        h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(self.dropout(F.silu(self.norm2(h))))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h  # output_scale_factor = 1


class Downsample2D(nn.Module):
    """diffusers Downsample2D(use_conv=True, name='op' -> attribute `conv`).  padding=1: plain stride-2 conv;
    padding=0 (hint encoder, models.py:591-598): F.pad(x, (0,1,0,1)) then stride-2 conv without padding."""

    def __init__(self, channels, out_channels=None, padding=1):
        super().__init__()
        self.padding = padding
        self.conv = nn.Conv2d(channels, out_channels or channels, 3, stride=2, padding=padding)

    def forward(self, x):
        if self.padding == 0:
            x = F.pad(x, (0, 1, 0, 1), mode="constant", value=0)
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class CrossAttnDownBlock2D(nn.Module):
    def __init__(self, cin, cout, temb, layers, heads, xdim, groups, eps, add_downsample):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, temb, groups, eps) for i in range(layers)])
        self.attentions = nn.ModuleList([Transformer2DModel(heads, cout // heads, cout, xdim, groups) for _ in range(layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(cout, cout, padding=1)]) if add_downsample else None

    def forward(self, h, temb, ehs, **kw):
        outs = []
        for r, a in zip(self.resnets, self.attentions):
            h = a(r(h, temb), encoder_hidden_states=ehs, **kw)
            outs.append(h)
        if self.downsamplers is not None:
            h = self.downsamplers[0](h)
            outs.append(h)
        return h, outs


class DownBlock2D(nn.Module):
    def __init__(self, cin, cout, temb, layers, groups, eps, add_downsample):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, temb, groups, eps) for i in range(layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(cout, cout, padding=1)]) if add_downsample else None

    def forward(self, h, temb, ehs=None, **kw):
        outs = []
        for r in self.resnets:
            h = r(h, temb)
            outs.append(h)
        if self.downsamplers is not None:
            h = self.downsamplers[0](h)
            outs.append(h)
        return h, outs


class UNetMidBlock2DCrossAttn(nn.Module):
    def __init__(self, ch, temb, heads, xdim, groups, eps):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(ch, ch, temb, groups, eps), ResnetBlock2D(ch, ch, temb, groups, eps)])
        self.attentions = nn.ModuleList([Transformer2DModel(heads, ch // heads, ch, xdim, groups)])

    def forward(self, h, temb, ehs, **kw):
        h = self.resnets[0](h, temb)
        h = self.attentions[0](h, encoder_hidden_states=ehs, **kw)
        return self.resnets[1](h, temb)


class UpBlockBase(nn.Module):
    def __init__(self, cin, cout, prev, temb, layers, groups, eps, add_upsample, heads=None, xdim=None):
        super().__init__()
        res = []
        for i in range(layers):
            skip = cin if i == layers - 1 else cout
            inp = prev if i == 0 else cout
            res.append(ResnetBlock2D(inp + skip, cout, temb, groups, eps))
        self.resnets = nn.ModuleList(res)
        if heads is not None:
            self.attentions = nn.ModuleList([Transformer2DModel(heads, cout // heads, cout, xdim, groups) for _ in range(layers)])
        else:
            self.attentions = None
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_upsample else None

    def forward(self, h, skips, temb, ehs=None, **kw):
        for i, r in enumerate(self.resnets):
            h = r(torch.cat([h, skips.pop()], dim=1), temb)
            if self.attentions is not None:
                h = self.attentions[i](h, encoder_hidden_states=ehs, **kw)
        if self.upsamplers is not None:
            h = self.upsamplers[0](h)
        return h


class TimestepEmbedding(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.linear_1 = nn.Linear(cin, cout)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(cout, cout)

    def forward(self, x):
        return self.linear_2(self.act(self.linear_1(x)))


def sinusoidal_embedding(timesteps: torch.Tensor, dim: int, flip_sin_to_cos=True, freq_shift=0.0) -> torch.Tensor:
    """diffusers get_timestep_embedding: fp32; [cos | sin] when flip_sin_to_cos."""
    half = dim // 2
    exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=timesteps.device)
    exponent = exponent / (half - freq_shift)
    emb = timesteps[:, None].float() * torch.exp(exponent)[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    return emb


class UNet2DConditionModel(nn.Module):
    """Restated diffusers UNet2DConditionModel (SD-1.5 options only).  `.config` mirrors the attributes the
    reference reads (train_...:411-425: block_out_channels, cross_attention_dim)."""

    def __init__(self, **overrides):
        super().__init__()
        cfg = dict(SD15_CONFIG)
        cfg.update(overrides)
        self.config = SimpleNamespace(**cfg)
        ch = list(cfg["block_out_channels"])
        layers = cfg["layers_per_block"]
        heads = cfg["attention_head_dim"]
        xdim = cfg["cross_attention_dim"]
        groups, eps = cfg["norm_num_groups"], cfg["norm_eps"]
        temb = ch[0] * 4
        self.conv_in = nn.Conv2d(cfg["in_channels"], ch[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(ch[0], temb)
        # diffusers creates both (empty) ModuleLists before the mid block, which fixes the registration order
        # down_blocks, up_blocks, mid_block and therefore the key order of `attn_processors`.
        self.down_blocks = nn.ModuleList()
        self.up_blocks = nn.ModuleList()
        out = ch[0]
        for i, typ in enumerate(cfg["down_block_types"]):
            inp, out = out, ch[i]
            last = i == len(ch) - 1
            if typ == "CrossAttnDownBlock2D":
                self.down_blocks.append(CrossAttnDownBlock2D(inp, out, temb, layers, heads, xdim, groups, eps, not last))
            else:
                self.down_blocks.append(DownBlock2D(inp, out, temb, layers, groups, eps, not last))
        self.mid_block = UNetMidBlock2DCrossAttn(ch[-1], temb, heads, xdim, groups, eps)
        rev = list(reversed(ch))
        out = rev[0]
        for i, typ in enumerate(cfg["up_block_types"]):
            prev, out = out, rev[i]
            inp = rev[min(i + 1, len(ch) - 1)]
            last = i == len(ch) - 1
            kw = dict(heads=heads, xdim=xdim) if typ == "CrossAttnUpBlock2D" else {}
            self.up_blocks.append(UpBlockBase(inp, out, prev, temb, layers + 1, groups, eps, not last, **kw))
        self.conv_norm_out = nn.GroupNorm(groups, ch[0], eps=eps)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(ch[0], cfg["out_channels"], 3, padding=1)

    # ---- processor plumbing (train_...:469-487) -------------------------------------------------------------
    def _attn_modules(self) -> "OrderedDict[str, CrossAttention]":
        out = OrderedDict()
        for name, m in self.named_modules():
            if isinstance(m, CrossAttention):
                out[f"{name}.processor"] = m
        return out

    @property
    def attn_processors(self) -> Dict[str, object]:
        return OrderedDict((k, m.processor) for k, m in self._attn_modules().items())

    def set_attn_processor(self, processors):
        mods = self._attn_modules()
        if isinstance(processors, dict):
            if len(processors) != len(mods):
                raise ValueError(f"A dict of processors was passed, but the number of processors {len(processors)} does "
                                 f"not match the number of attention layers: {len(mods)}.")
            for k, m in mods.items():
                m.set_processor(processors[k])
        else:
            for m in mods.values():
                m.set_processor(processors)

    # ---- forward ----------------------------------------------------------------------------------------------
    def forward(self, sample, timestep, encoder_hidden_states, cross_attention_kwargs=None):
        kw = cross_attention_kwargs or {}
        t = timestep
        if not torch.is_tensor(t):
            t = torch.tensor([t], dtype=torch.long, device=sample.device)
        if t.dim() == 0:
            t = t[None]
        t = t.expand(sample.shape[0])
        t_emb = sinusoidal_embedding(t, self.config.block_out_channels[0], self.config.flip_sin_to_cos,
                                     self.config.freq_shift).to(sample.dtype)
        emb = self.time_embedding(t_emb)
        h = self.conv_in(sample)
        skips = [h]
        for blk in self.down_blocks:
            h, outs = blk(h, emb, encoder_hidden_states, **kw)
            skips.extend(outs)
        h = self.mid_block(h, emb, encoder_hidden_states, **kw)
        for blk in self.up_blocks:
            h = blk(h, skips, emb, encoder_hidden_states, **kw)
        h = self.conv_out(self.conv_act(self.conv_norm_out(h)))
        return SimpleNamespace(sample=h)


def init_synthetic_(model: nn.Module, seed: int = 0) -> nn.Module:
    """Seeded synthetic frozen weights (no SD-1.5 checkpoint is reachable offline): W ~ N(0, 1/fan_in),
    biases 0.02 N, norm gamma = 1 + 0.1 N, beta = 0.1 N (SURVEY.md §8c "golden vectors we must mint")."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if "processor" in name:
                continue
            if p.dim() >= 2:
                fan_in = p[0].numel()
                p.copy_(torch.randn(p.shape, generator=g) / math.sqrt(fan_in))
            elif name.endswith("weight"):  # norm gammas
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
            else:
                is_norm = any(s in name for s in ("norm", "conv_norm_out"))
                p.copy_((0.1 if is_norm else 0.02) * torch.randn(p.shape, generator=g))
    return model
