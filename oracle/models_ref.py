"""ORACLE (test infrastructure only — never imported by the product path).

CPU / fp32 restatement of /root/reference/models.py: the LoRA attention processors (models.py:72-431) and the
ControlLoRA hint encoder (models.py:434-835), written against oracle/unet_ref.py's restated diffusers classes.
The arithmetic and its order follow the reference line by line (citations on each function); the code structure is
our own (one adapter-chain walk instead of three copies of the projection code).

PARITY STATUS: **parity unpinned** (the reference has no tests or golden vectors and cannot be imported here because
`diffusers` is absent — SURVEY.md §4, §8c).  Anchors: exact trainable-parameter counts of the shipped configs
(6 047 040 for fill50k/canny/pose, 6 048 576 post-add, 5 000 704 canny-v2, 19 810 304 danbooru-sketch) and the
state-dict key names of the published checkpoints, both checked in tests/test_oracle.py.
"""
from __future__ import annotations

import json
from dataclasses import dataclass
from pathlib import Path
from typing import List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .unet_ref import CrossAttention, Downsample2D, LoRALinearLayer


# ------------------------------------------------------------------------------------------------ processors
class LoRACrossAttnProcessor(nn.Module):
    """models.py:72-152."""

    def __init__(self, hidden_size, cross_attention_dim=None, rank=4, post_add=False, key_states_skipped=False,
                 value_states_skipped=False, output_states_skipped=False):
        super().__init__()
        self.hidden_size, self.cross_attention_dim, self.rank, self.post_add = hidden_size, cross_attention_dim, rank, post_add
        kv_in = hidden_size if post_add else (cross_attention_dim or hidden_size)  # models.py:92,95
        self.to_q_lora = LoRALinearLayer(hidden_size, hidden_size, rank)
        if not key_states_skipped:
            self.to_k_lora = LoRALinearLayer(kv_in, hidden_size, rank)
        if not value_states_skipped:
            self.to_v_lora = LoRALinearLayer(kv_in, hidden_size, rank)
        if not output_states_skipped:
            self.to_out_lora = LoRALinearLayer(hidden_size, hidden_size, rank)
        self.key_states_skipped = key_states_skipped
        self.value_states_skipped = value_states_skipped
        self.output_states_skipped = output_states_skipped

    # models.py:103-116 (including the reference's `to_q_lora` typo in skip_value_states' assert)
    def skip_key_states(self, is_skipped=True):
        if not is_skipped:
            assert hasattr(self, "to_k_lora")
        self.key_states_skipped = is_skipped

    def skip_value_states(self, is_skipped=True):
        if not is_skipped:
            assert hasattr(self, "to_q_lora")
        self.value_states_skipped = is_skipped

    def skip_output_states(self, is_skipped=True):
        if not is_skipped:
            assert hasattr(self, "to_out_lora")
        self.output_states_skipped = is_skipped

    def __call__(self, attn: CrossAttention, hidden_states, encoder_hidden_states=None, attention_mask=None, scale=1.0):
        # models.py:118-152
        b, n, _ = hidden_states.shape
        attention_mask = attn.prepare_attention_mask(attention_mask, n, b)
        q = attn.to_q(hidden_states)
        q = q + scale * self.to_q_lora(q if self.post_add else hidden_states)
        q = attn.head_to_batch_dim(q)
        ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        k = attn.to_k(ctx)
        if not self.key_states_skipped:
            k = k + scale * self.to_k_lora(k if self.post_add else ctx)
        v = attn.to_v(ctx)
        if not self.value_states_skipped:
            v = v + scale * self.to_v_lora(v if self.post_add else ctx)
        k, v = attn.head_to_batch_dim(k), attn.head_to_batch_dim(v)
        h = attn.batch_to_head_dim(torch.bmm(attn.get_attention_scores(q, k, attention_mask), v))
        out = attn.to_out[0](h)
        if not self.output_states_skipped:
            out = out + scale * self.to_out_lora(out if self.post_add else h)
        return attn.to_out[1](out)


def _reshape_control(proc, hidden_states):
    """models.py:202-206 / 337-341: cast, NCHW -> (B, HW, C) once, and cache the reshaped tensor on the processor."""
    cs = proc.control_states.to(hidden_states.dtype)
    if hidden_states.ndim == 3 and cs.ndim == 4:
        b, _, hh, ww = cs.shape
        cs = cs.permute(0, 2, 3, 1).reshape(b, hh * ww, -1)
        proc.control_states = cs
    return cs


def _control_term(proc, layer, hidden_states, scale):
    """models.py:207-220 / 342-355."""
    cs = _reshape_control(proc, hidden_states)
    inp = cs
    if proc.concat_hidden:
        b1, b2 = cs.shape[0], hidden_states.shape[0]
        if b1 != b2:  # repeat-interleave, models.py:209-212
            cs = cs[:, None].repeat(1, b2 // b1, *([1] * (cs.dim() - 1))).view(-1, *cs.shape[1:])
        inp = torch.cat([hidden_states, cs], -1)
    term = scale * layer(inp)
    return cs + term if proc.control_self_add else term


def _kv_chain(chain, me, base, ctx, scale, which):
    """k / v adapter walk, models.py:248-265 and 385-402.  Quirk kept: stacked (pre/post) adapters' VALUE deltas are
    not multiplied by `scale` (models.py:260,265,397,402)."""
    out = base
    for ad in chain:
        skipped = ad.key_states_skipped if which == "k" else ad.value_states_skipped
        if skipped:
            continue
        layer = ad.to_k_lora if which == "k" else ad.to_v_lora
        delta = layer(out if ad.post_add else ctx)
        out = out + (delta if (which == "v" and ad is not me) else scale * delta)
    return out


def _out_chain(chain, me, attn, h, scale):
    """to_out walk, models.py:275-282 / 419-426.  Quirk kept: the processor's own to_out_lora is applied even when its
    output_states_skipped flag is set (models.py:279, 423)."""
    out = attn.to_out[0](h)
    for ad in chain:
        if ad is not me and ad.output_states_skipped:
            continue
        out = out + scale * ad.to_out_lora(out if ad.post_add else h)
    return attn.to_out[1](out)


class ControlLoRACrossAttnProcessor(LoRACrossAttnProcessor):
    """models.py:155-287 ("v1"): control enters only the q path."""

    def __init__(self, hidden_size, cross_attention_dim=None, rank=4, control_rank=None, post_add=False,
                 concat_hidden=False, control_channels=None, control_self_add=True, key_states_skipped=False,
                 value_states_skipped=False, output_states_skipped=False, **kwargs):
        super().__init__(hidden_size, cross_attention_dim, rank, post_add=post_add,
                         key_states_skipped=key_states_skipped, value_states_skipped=value_states_skipped,
                         output_states_skipped=output_states_skipped)
        control_rank = rank if control_rank is None else control_rank
        control_channels = hidden_size if control_channels is None else control_channels
        self.concat_hidden = concat_hidden
        # models.py:180-182: control_channels was just made non-None, so this is always False.
        self.control_self_add = control_self_add if control_channels is None else False
        self.control_states: Optional[torch.Tensor] = None
        self.to_control = LoRALinearLayer(control_channels + (hidden_size if concat_hidden else 0), hidden_size, control_rank)
        self.pre_loras: List[LoRACrossAttnProcessor] = []
        self.post_loras: List[LoRACrossAttnProcessor] = []

    def inject_pre_lora(self, lora_layer):
        self.pre_loras.append(lora_layer)

    def inject_post_lora(self, lora_layer):
        self.post_loras.append(lora_layer)

    def inject_control_states(self, control_states):
        self.control_states = control_states

    def process_control_states(self, hidden_states, scale=1.0):
        return _control_term(self, self.to_control, hidden_states, scale)

    def __call__(self, attn: CrossAttention, hidden_states, encoder_hidden_states=None, attention_mask=None, scale=1.0):
        assert self.control_states is not None  # models.py:227
        b, n, _ = hidden_states.shape
        attention_mask = attn.prepare_attention_mask(attention_mask, n)
        chain = [*self.pre_loras, self, *self.post_loras]
        q = attn.to_q(hidden_states)
        for ad in chain:  # models.py:232-243
            lora_in = q if ad.post_add else hidden_states
            if isinstance(ad, ControlLoRACrossAttnProcessor):
                lora_in = lora_in + ad.process_control_states(hidden_states, scale)
            q = q + scale * ad.to_q_lora(lora_in)
        q = attn.head_to_batch_dim(q)
        ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        k = _kv_chain(chain, self, attn.to_k(ctx), ctx, scale, "k")
        v = _kv_chain(chain, self, attn.to_v(ctx), ctx, scale, "v")
        k, v = attn.head_to_batch_dim(k), attn.head_to_batch_dim(v)
        h = attn.batch_to_head_dim(torch.bmm(attn.get_attention_scores(q, k, attention_mask), v))
        return _out_chain(chain, self, attn, h, scale)


class ControlLoRACrossAttnProcessorV2(LoRACrossAttnProcessor):
    """models.py:292-431 ("V2"): control rewrites the hidden states before q (and, for self-attention, k/v) and the
    attention output before to_out; k/v LoRA are skipped by construction."""

    def __init__(self, hidden_size, cross_attention_dim=None, rank=4, control_rank=None, control_channels=None, **kwargs):
        super().__init__(hidden_size, cross_attention_dim, rank, post_add=False, key_states_skipped=True,
                         value_states_skipped=True, output_states_skipped=False)
        control_rank = rank if control_rank is None else control_rank
        control_channels = hidden_size if control_channels is None else control_channels
        self.concat_hidden = True
        self.control_self_add = False
        self.control_states: Optional[torch.Tensor] = None
        self.to_control = LoRALinearLayer(hidden_size + control_channels, hidden_size, control_rank)
        self.to_control_out = LoRALinearLayer(hidden_size + control_channels, hidden_size, control_rank)
        self.pre_loras: List[LoRACrossAttnProcessor] = []
        self.post_loras: List[LoRACrossAttnProcessor] = []

    def inject_pre_lora(self, lora_layer):
        self.pre_loras.append(lora_layer)

    def inject_post_lora(self, lora_layer):
        self.post_loras.append(lora_layer)

    def inject_control_states(self, control_states):
        self.control_states = control_states

    def process_control_states(self, hidden_states, scale=1.0, is_out=False):
        return _control_term(self, self.to_control_out if is_out else self.to_control, hidden_states, scale)

    def __call__(self, attn: CrossAttention, hidden_states, encoder_hidden_states=None, attention_mask=None, scale=1.0):
        assert self.control_states is not None  # models.py:362
        b, n, _ = hidden_states.shape
        attention_mask = attn.prepare_attention_mask(attention_mask, n)
        chain = [*self.pre_loras, self, *self.post_loras]
        for ad in chain:  # models.py:366-372
            if isinstance(ad, ControlLoRACrossAttnProcessorV2):
                hidden_states = hidden_states + ad.process_control_states(hidden_states, scale)
        q = attn.to_q(hidden_states)
        for ad in chain:  # models.py:374-380
            q = q + scale * ad.to_q_lora(q if ad.post_add else hidden_states)
        q = attn.head_to_batch_dim(q)
        # models.py:383: for self-attention the context is the *updated* hidden states
        ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        k = _kv_chain(chain, self, attn.to_k(ctx), ctx, scale, "k")
        v = _kv_chain(chain, self, attn.to_v(ctx), ctx, scale, "v")
        k, v = attn.head_to_batch_dim(k), attn.head_to_batch_dim(v)
        h = attn.batch_to_head_dim(torch.bmm(attn.get_attention_scores(q, k, attention_mask), v))
        for ad in chain:  # models.py:412-418
            if isinstance(ad, ControlLoRACrossAttnProcessorV2):
                h = h + ad.process_control_states(h, scale, is_out=True)
        return _out_chain(chain, self, attn, h, scale)


# ------------------------------------------------------------------------------------------------ hint encoder
class ConvBlock2D(nn.Module):
    """models.py:434-547 with the only options the ControlLoRA ctor uses (temb_channels=None, no up/down):
    SiLU(GN2(conv_k(SiLU(GN1(x))))), GroupNorm eps 1e-6, no residual."""

    def __init__(self, in_channels, out_channels, conv_kernel_size=3, groups=32, eps=1e-6):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, in_channels, eps=eps, affine=True)
        self.conv1 = nn.Conv2d(in_channels, out_channels, conv_kernel_size, stride=1, padding=conv_kernel_size // 2)
        self.norm2 = nn.GroupNorm(groups, out_channels, eps=eps, affine=True)

    def forward(self, x, temb=None):
        return F.silu(self.norm2(self.conv1(F.silu(self.norm1(x)))))


class SimpleDownEncoderBlock2D(nn.Module):
    """models.py:550-610."""

    def __init__(self, in_channels, out_channels, num_layers=1, eps=1e-6, groups=32, kernel_size=3,
                 add_downsample=True, downsample_padding=1):
        super().__init__()
        self.convnets = nn.ModuleList(
            [ConvBlock2D(in_channels if i == 0 else out_channels, out_channels, kernel_size, groups, eps) for i in range(num_layers)])
        cin = in_channels if num_layers == 0 else out_channels
        self.downsamplers = nn.ModuleList([Downsample2D(cin, out_channels, padding=downsample_padding)]) if add_downsample else None

    def forward(self, h):
        for c in self.convnets:
            h = c(h, temb=None)
        if self.downsamplers is not None:
            for d in self.downsamplers:
                h = d(h)
        return h


@dataclass
class ControlLoRAOutput:
    control_states: Tuple[torch.Tensor, ...]


_CONFIG_DEFAULTS = dict(
    in_channels=3,
    down_block_types=("SimpleDownEncoderBlock2D",) * 4,
    block_out_channels=(32, 64, 128, 256),
    layers_per_block=1,
    act_fn="silu",
    norm_num_groups=32,
    lora_pre_down_block_types=(None, "SimpleDownEncoderBlock2D", "SimpleDownEncoderBlock2D", "SimpleDownEncoderBlock2D"),
    lora_pre_down_layers_per_block=1,
    lora_pre_conv_skipped=False,
    lora_pre_conv_types=("SimpleDownEncoderBlock2D",) * 4,
    lora_pre_conv_layers_per_block=1,
    lora_pre_conv_layers_kernel_size=1,
    lora_block_in_channels=(256, 256, 256, 256),
    lora_block_out_channels=(320, 640, 1280, 1280),
    lora_cross_attention_dims=([None, 768] * 5, [None, 768] * 5, [None, 768] * 5, [None, 768]),
    lora_rank=4,
    lora_control_rank=None,
    lora_post_add=False,
    lora_concat_hidden=False,
    lora_control_channels=(None, None, None, None),
    lora_control_self_add=True,
    lora_key_states_skipped=False,
    lora_value_states_skipped=False,
    lora_output_states_skipped=False,
    lora_control_version=1,
)


class ControlLoRA(nn.Module):
    """models.py:618-835."""

    def __init__(self, **kwargs):
        super().__init__()
        cfg = dict(_CONFIG_DEFAULTS)
        unknown = set(kwargs) - set(cfg)
        if unknown:
            raise TypeError(f"unexpected ControlLoRA config keys: {sorted(unknown)}")
        cfg.update(kwargs)
        self.config = dict(cfg)
        c = cfg
        proc_cls = ControlLoRACrossAttnProcessorV2 if c["lora_control_version"] == 2 else ControlLoRACrossAttnProcessor
        assert c["lora_block_in_channels"][0] == c["block_out_channels"][-1]  # models.py:674
        control_channels = c["lora_control_channels"]
        self_add = c["lora_control_self_add"]
        if c["lora_pre_conv_skipped"]:  # models.py:676-678
            control_channels = c["lora_block_in_channels"]
            self_add = False
        groups = c["norm_num_groups"]
        boc = c["block_out_channels"]
        self.conv_in = nn.Conv2d(c["in_channels"], boc[0], 3, stride=1, padding=1)
        self.down_blocks = nn.ModuleList()
        self.pre_lora_layers = nn.ModuleList()
        self.lora_layers = nn.ModuleList()

        def pre_conv(i, cin):
            if c["lora_pre_conv_skipped"]:
                return nn.Identity()
            cout = c["lora_block_out_channels"][i] if control_channels[i] is None else control_channels[i]
            return SimpleDownEncoderBlock2D(cin, cout, c["lora_pre_conv_layers_per_block"], 1e-6, groups,
                                            c["lora_pre_conv_layers_kernel_size"], add_downsample=False, downsample_padding=0)

        def procs(i):
            return nn.ModuleList([
                proc_cls(c["lora_block_out_channels"][i], cross_attention_dim=xd, rank=c["lora_rank"],
                         control_rank=c["lora_control_rank"], post_add=c["lora_post_add"],
                         concat_hidden=c["lora_concat_hidden"], control_channels=control_channels[i],
                         control_self_add=self_add, key_states_skipped=c["lora_key_states_skipped"],
                         value_states_skipped=c["lora_value_states_skipped"],
                         output_states_skipped=c["lora_output_states_skipped"])
                for xd in c["lora_cross_attention_dims"][i]])

        # level 0: the whole 512^2 -> 64^2 pyramid (models.py:690-748)
        stack, out_ch = [], boc[0]
        for i in range(len(c["down_block_types"])):
            in_ch, out_ch = out_ch, boc[i]
            stack.append(SimpleDownEncoderBlock2D(in_ch, out_ch, c["layers_per_block"], 1e-6, groups, 3,
                                                  add_downsample=i != len(boc) - 1, downsample_padding=0))
        self.down_blocks.append(nn.Sequential(*stack))
        self.pre_lora_layers.append(pre_conv(0, c["lora_block_in_channels"][0]))
        self.lora_layers.append(procs(0))
        # levels 1..3 (models.py:750-808)
        out_ch = c["lora_block_in_channels"][0]
        for i in range(1, len(c["lora_pre_down_block_types"])):
            in_ch, out_ch = out_ch, c["lora_block_in_channels"][i]
            self.down_blocks.append(SimpleDownEncoderBlock2D(in_ch, out_ch, c["lora_pre_down_layers_per_block"], 1e-6,
                                                             groups, 3, add_downsample=True, downsample_padding=0))
            self.pre_lora_layers.append(pre_conv(i, out_ch))
            self.lora_layers.append(procs(i))

    @classmethod
    def from_config(cls, config):
        if isinstance(config, (str, Path)):
            p = Path(config)
            if p.is_dir():
                p = p / "config.json"
            config = json.loads(p.read_text())
        return cls(**{k: v for k, v in dict(config).items() if not k.startswith("_")})

    def forward(self, x, return_dict=True):
        # models.py:810-835
        orig_dtype = x.dtype
        h = self.conv_in(x.to(self.conv_in.weight.dtype))
        states = []
        for down, pre, procs in zip(self.down_blocks, self.pre_lora_layers, self.lora_layers):
            h = down(h)
            cs = pre(h)
            if isinstance(cs, tuple):
                cs = cs[0]
            cs = cs.to(orig_dtype)
            for proc in procs:
                proc.inject_control_states(cs)
            states.append(cs)
        if not return_dict:
            return tuple(states)
        return ControlLoRAOutput(control_states=tuple(states))


def wire_processors(unet, control_lora):
    """train_text_to_image_control_lora.py:469-487: pop processors from control_lora.lora_layers[control_id] in
    `unet.attn_processors` key order and install them on the UNet."""
    n_ch = len(unet.config.block_out_channels)
    ids = list(range(n_ch))
    pools = [list(l) for l in control_lora.lora_layers]
    procs = {}
    for name in unet.attn_processors.keys():
        if name.startswith("mid_block"):
            cid = ids[-1]
        elif name.startswith("up_blocks"):
            cid = list(reversed(ids))[int(name[len("up_blocks.")])]
        else:
            cid = ids[int(name[len("down_blocks.")])]
        if pools[cid]:
            procs[name] = pools[cid].pop(0)
    unet.set_attn_processor(procs)
    return procs


def randomize_lora_up_(control_lora: nn.Module, seed: int = 1, std: float = 0.02):
    """LoRA `up` weights are zero-initialised (diffusers), which would make every parity test vacuous: give them
    N(0, std) values (SURVEY.md 'five facts' #5)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in control_lora.named_parameters():
            if name.endswith("up.weight"):
                p.copy_(std * torch.randn(p.shape, generator=g))
