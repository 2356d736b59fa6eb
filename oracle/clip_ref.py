"""ORACLE (test infrastructure only — never imported by the product path).

CPU / fp32 restatement of `transformers.CLIPTextModel.forward` for the one call the reference makes in front of the hot path:

  train_text_to_image_control_lora.py:768     encoder_hidden_states = text_encoder(batch["input_ids"])[0]
  (text_encoder = CLIPTextModel.from_pretrained(..., subfolder="text_encoder"), :401-403; frozen, :431)

The arithmetic lives in the third-party `transformers` package (requirements.txt: `transformers>=4.25.1`), module
`models/clip/modeling_clip.py`: CLIPTextEmbeddings (token + learned position embedding), CLIPEncoderLayer (pre-LN:
x + attn(LN1(x)); x + mlp(LN2(x))), CLIPAttention (q/k/v/out Linear WITH bias, scale d^-0.5 applied to q, causal mask, softmax
in fp32), CLIPMLP (fc1 -> quick_gelu = x * sigmoid(1.702 x) -> fc2), final_layer_norm; output [0] = last_hidden_state.

PARITY STATUS: **pinned** — transformers 5.5 is importable in this image, and tests/test_oracle.py checks this restatement
against `transformers.CLIPTextModel` itself (random-init SD-1.5 text config, CPU fp32, max abs diff < 1e-4).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

SD15_TEXT_CONFIG = dict(vocab_size=49408, hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12,
                        max_position_embeddings=77, layer_norm_eps=1e-5, hidden_act="quick_gelu")


def clip_text_forward(sd: dict, input_ids: torch.Tensor, cfg: dict = SD15_TEXT_CONFIG) -> torch.Tensor:
    """last_hidden_state [B, T, hidden] of CLIPTextModel with state dict `sd` (transformers key names), computed in fp32."""
    pre = "text_model." if any(k.startswith("text_model.") for k in sd) else ""
    g = lambda k: sd[pre + k].float()
    B, T = input_ids.shape
    heads, Cw, eps = cfg["num_attention_heads"], cfg["hidden_size"], cfg["layer_norm_eps"]
    d = Cw // heads
    x = g("embeddings.token_embedding.weight")[input_ids] + g("embeddings.position_embedding.weight")[:T].unsqueeze(0)
    causal = torch.full((T, T), float("-inf")).triu(1)
    for i in range(cfg["num_hidden_layers"]):
        p = f"encoder.layers.{i}."
        h = F.layer_norm(x, (Cw,), g(p + "layer_norm1.weight"), g(p + "layer_norm1.bias"), eps)
        q = F.linear(h, g(p + "self_attn.q_proj.weight"), g(p + "self_attn.q_proj.bias")) * (d ** -0.5)
        k = F.linear(h, g(p + "self_attn.k_proj.weight"), g(p + "self_attn.k_proj.bias"))
        v = F.linear(h, g(p + "self_attn.v_proj.weight"), g(p + "self_attn.v_proj.bias"))
        sh = lambda t: t.view(B, T, heads, d).transpose(1, 2)
        s = sh(q) @ sh(k).transpose(-1, -2) + causal
        a = (torch.softmax(s, -1) @ sh(v)).transpose(1, 2).reshape(B, T, Cw)
        x = x + F.linear(a, g(p + "self_attn.out_proj.weight"), g(p + "self_attn.out_proj.bias"))
        h = F.layer_norm(x, (Cw,), g(p + "layer_norm2.weight"), g(p + "layer_norm2.bias"), eps)
        f = F.linear(h, g(p + "mlp.fc1.weight"), g(p + "mlp.fc1.bias"))
        f = f * torch.sigmoid(1.702 * f)
        x = x + F.linear(f, g(p + "mlp.fc2.weight"), g(p + "mlp.fc2.bias"))
    return F.layer_norm(x, (Cw,), g("final_layer_norm.weight"), g("final_layer_norm.bias"), eps)


def synthetic_state_dict(cfg: dict = SD15_TEXT_CONFIG, seed: int = 0) -> dict:
    """Seeded synthetic weights with transformers' key names: W ~ N(0, 1/fan_in), biases 0.02 N, LN gamma 1 + 0.1 N, beta 0.1 N,
    embeddings 0.02 N (CLIP's initializer range) scaled so activations stay O(1)."""
    gen = torch.Generator().manual_seed(seed)
    Cw, Fi = cfg["hidden_size"], cfg["intermediate_size"]
    r = lambda *s: torch.randn(*s, generator=gen)
    sd = {"text_model.embeddings.token_embedding.weight": r(cfg["vocab_size"], Cw) * 0.5,
          "text_model.embeddings.position_embedding.weight": r(cfg["max_position_embeddings"], Cw) * 0.5,
          "text_model.final_layer_norm.weight": 1 + 0.1 * r(Cw), "text_model.final_layer_norm.bias": 0.1 * r(Cw)}
    for i in range(cfg["num_hidden_layers"]):
        p = f"text_model.encoder.layers.{i}."
        for nm in ("q_proj", "k_proj", "v_proj", "out_proj"):
            sd[p + f"self_attn.{nm}.weight"] = r(Cw, Cw) / math.sqrt(Cw)
            sd[p + f"self_attn.{nm}.bias"] = 0.02 * r(Cw)
        sd[p + "mlp.fc1.weight"] = r(Fi, Cw) / math.sqrt(Cw); sd[p + "mlp.fc1.bias"] = 0.02 * r(Fi)
        sd[p + "mlp.fc2.weight"] = r(Cw, Fi) / math.sqrt(Fi); sd[p + "mlp.fc2.bias"] = 0.02 * r(Cw)
        for nm in ("layer_norm1", "layer_norm2"):
            sd[p + nm + ".weight"] = 1 + 0.1 * r(Cw); sd[p + nm + ".bias"] = 0.1 * r(Cw)
    return sd
