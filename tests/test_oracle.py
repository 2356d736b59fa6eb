"""The oracle (oracle/) against its anchors: parameter counts and state-dict keys of the published reference models
(SURVEY.md §8b/§8c), algebraic self-consistency, and the committed golden vectors."""
import json
import math
from pathlib import Path

import pytest
import torch

from oracle import models_ref as MR
from oracle import unet_ref as UR
from controllora_b200.configs import NAMED

ROOT = Path(__file__).resolve().parent.parent
TINY = dict(block_out_channels=(64, 128, 128, 128), layers_per_block=1, cross_attention_dim=64, attention_head_dim=8)
TINY_LORA = dict(lora_block_out_channels=(64, 128, 128, 128),
                 lora_cross_attention_dims=([None, 64] * 3, [None, 64] * 3, [None, 64] * 3, [None, 64]))

# README.md:7,17 ("~7M parameters / ~25M storage", "~5M / ~20M") -> exact counts, SURVEY.md §6
PARAM_COUNTS = {"base": 6047040, "fill50k": 6047040, "diffusiondb-canny": 6047040, "mpii-pose": 6047040,
                "post-add": 6048576, "diffusiondb-canny-v2": 5000704, "mpii-pose-v2": 5000704, "danbooru-sketch": 19810304}


@pytest.mark.parametrize("name", sorted(PARAM_COUNTS))
def test_controllora_param_counts(name):
    m = MR.ControlLoRA.from_config(NAMED[name])
    assert sum(p.numel() for p in m.parameters()) == PARAM_COUNTS[name]


def test_unet_param_count_matches_sd15():
    with torch.device("meta"):
        u = UR.UNet2DConditionModel()
    assert sum(p.numel() for p in u.parameters()) == 859_520_964           # SD-1.5 UNet
    assert sum(p.numel() for p in u.parameters() if p.dim() >= 2) == 859_077_120   # weights in conv/linear kernels
    keys = list(u.attn_processors.keys())
    assert len(keys) == 32
    # diffusers registration order: down_blocks, up_blocks, mid_block
    assert keys[0].startswith("down_blocks.0") and keys[12].startswith("up_blocks.1") and keys[-1].startswith("mid_block")


def test_state_dict_keys_of_published_checkpoints():
    sd = MR.ControlLoRA.from_config(NAMED["fill50k"]).state_dict()
    for k in ["conv_in.weight", "down_blocks.0.0.convnets.0.norm1.weight", "down_blocks.0.2.downsamplers.0.conv.bias",
              "down_blocks.3.downsamplers.0.conv.weight", "pre_lora_layers.2.convnets.0.conv1.weight",
              "lora_layers.0.9.to_control.up.weight", "lora_layers.3.1.to_out_lora.down.weight"]:
        assert k in sd, k
    assert "down_blocks.0.3.downsamplers.0.conv.weight" not in sd      # the last pyramid block has no downsampler
    v2 = MR.ControlLoRA.from_config(NAMED["diffusiondb-canny-v2"]).state_dict()
    assert "lora_layers.0.0.to_control_out.down.weight" in v2 and "lora_layers.0.0.to_k_lora.down.weight" not in v2
    assert not any(k.startswith("pre_lora_layers") for k in v2)


def _tiny_pair(**kw):
    torch.manual_seed(0)
    unet = UR.UNet2DConditionModel(**TINY)
    UR.init_synthetic_(unet, seed=1)
    cfg = dict(TINY_LORA)
    cfg.update(kw)
    cl = MR.ControlLoRA(**cfg)
    return unet, cl


def _inputs(B=2, HW=16):
    g = torch.Generator().manual_seed(5)
    return (torch.randn(B, 4, HW, HW, generator=g), torch.tensor([17, 801]), torch.randn(B, 77, 64, generator=g),
            torch.rand(B, 3, HW * 8, HW * 8, generator=g) * 2 - 1)


def test_zero_initialised_up_weights_make_lora_a_noop():
    """SURVEY 'five facts' #5: a freshly built ControlLoRA must not change the UNet output at all."""
    unet, cl = _tiny_pair()
    x, t, e, guide = _inputs()
    base = unet(x, t, e).sample
    MR.wire_processors(unet, cl)
    cl(guide)
    out = unet(x, t, e).sample
    assert torch.equal(base, out)
    MR.randomize_lora_up_(cl, seed=3, std=0.05)
    assert (unet(x, t, e).sample - base).abs().max() > 1e-4


def test_control_states_shapes_and_injection():
    unet, cl = _tiny_pair()
    x, t, e, guide = _inputs()
    MR.wire_processors(unet, cl)
    states = cl(guide).control_states
    assert [tuple(s.shape) for s in states] == [(2, 64, 16, 16), (2, 128, 8, 8), (2, 128, 4, 4), (2, 128, 2, 2)]
    for lvl, procs in enumerate(cl.lora_layers):
        for p in procs:
            assert p.control_states is states[lvl]
    # processors are wired level-wise: down_i and up_(3-i) share level i, mid is level 3 (train_...:469-487)
    names = list(unet.attn_processors.keys())
    assert unet.attn_processors["mid_block.attentions.0.transformer_blocks.0.attn1.processor"] is cl.lora_layers[3][0]
    assert unet.attn_processors[names[0]] is cl.lora_layers[0][0]


def test_v1_control_identity_used_by_the_fused_epilogue():
    """q-path algebra behind the CUDA path: Aq (h + s Bc Ac c) == Aq h + s (Aq Bc)(Ac c)."""
    torch.manual_seed(1)
    T, C, r = 50, 64, 4
    h, c = torch.randn(T, C), torch.randn(T, C)
    Aq, Bc, Ac = torch.randn(r, C), torch.randn(C, r), torch.randn(r, C)
    s = 0.7
    lhs = (h + s * (c @ Ac.t()) @ Bc.t()) @ Aq.t()
    rhs = h @ Aq.t() + s * (c @ Ac.t()) @ (Aq @ Bc).t()
    assert torch.allclose(lhs, rhs, atol=1e-4, rtol=1e-4)


def test_v2_rewrites_hidden_states_before_kv():
    """models.py:369-383: V2 adds the control term to the hidden states first, so for self-attention the keys/values
    are projected from the *updated* states (and the attention output is updated again before to_out, :415)."""
    torch.manual_seed(0)
    C, Cc, T = 64, 32, 10
    attn = UR.CrossAttention(C, None, heads=4, dim_head=16)
    p = MR.ControlLoRACrossAttnProcessorV2(C, None, rank=4, control_channels=Cc)
    for n, w in p.named_parameters():
        if n.endswith("up.weight"):
            torch.nn.init.normal_(w, std=0.2)
    h = torch.randn(2, T, C)
    c = torch.randn(2, T, Cc)
    p.inject_control_states(c)
    out = p(attn, h)
    hp = h + p.to_control(torch.cat([h, c], -1))
    q = attn.to_q(hp) + p.to_q_lora(hp)
    k, v = attn.to_k(hp), attn.to_v(hp)
    o = attn.batch_to_head_dim(torch.bmm(attn.get_attention_scores(attn.head_to_batch_dim(q), attn.head_to_batch_dim(k)),
                                         attn.head_to_batch_dim(v)))
    o = o + p.to_control_out(torch.cat([o, c], -1))
    want = attn.to_out[0](o) + p.to_out_lora(o)
    assert torch.allclose(out, want, atol=1e-5)
    k_old = attn.to_k(h)
    assert (k - k_old).abs().max() > 1e-2


def test_scale_kwarg_and_value_quirk():
    """`scale` multiplies every LoRA delta except stacked adapters' value deltas (models.py:260,265)."""
    unet, cl = _tiny_pair()
    MR.randomize_lora_up_(cl, seed=3, std=0.05)
    MR.wire_processors(unet, cl)
    x, t, e, guide = _inputs()
    cl(guide)
    a = unet(x, t, e, cross_attention_kwargs={"scale": 1.0}).sample
    b = unet(x, t, e, cross_attention_kwargs={"scale": 0.0}).sample
    torch.manual_seed(0)
    plain = UR.UNet2DConditionModel(**TINY)
    UR.init_synthetic_(plain, seed=1)
    assert torch.allclose(b, plain(x, t, e).sample, atol=1e-5)
    assert (a - b).abs().max() > 1e-4


GOLDEN = ROOT / "tests" / "golden" / "oracle_tiny.pt"


@pytest.mark.parametrize("variant", ["v1", "v2"])
def test_oracle_matches_committed_golden_vectors(variant):
    """Regression pin of the oracle itself (generated by tests/golden/make_golden.py)."""
    from tests.golden.make_golden import run_variant

    gold = torch.load(GOLDEN)[variant]
    now = run_variant(variant)
    for k in gold:
        assert torch.allclose(now[k], gold[k], atol=2e-5, rtol=2e-4), k


GOLDEN_EXTRA = ROOT / "tests" / "golden" / "oracle_extra.pt"


@pytest.mark.parametrize("variant", ["v1_post_add", "v1_concat"])
def test_oracle_variants_match_committed_golden_vectors(variant):
    """post_add (configs/post-add.json) and concat_hidden (configs/danbooru-sketch.json) flavours of the oracle."""
    from tests.golden.make_golden import run_variant

    gold = torch.load(GOLDEN_EXTRA)[variant]
    now = run_variant(variant)
    for k in gold:
        assert torch.allclose(now[k], gold[k], atol=2e-5, rtol=2e-4), k


def test_oracle_samplers_match_committed_golden_trajectories():
    """DDIM and DPM-Solver++(2M) restatements: 10- / 12-step trajectories and the 30-step timestep table."""
    from tests.golden.make_golden import run_samplers

    gold = torch.load(GOLDEN_EXTRA)["samplers"]
    now = run_samplers()
    assert torch.equal(now["dpm_timesteps_30"], gold["dpm_timesteps_30"])
    for k in ("ddim10", "dpmpp12"):
        assert torch.allclose(now[k], gold[k], atol=1e-5, rtol=1e-5), k


def test_philox_restatement_matches_random123_known_answers():
    """oracle/sampler_ref.philox4x32_10 (the checker of csrc/noise.cu) against the known-answer vectors shipped with
    Random123 (kat_vectors, philox4x32 10 rounds)."""
    import numpy as np
    from oracle import sampler_ref as SR

    kats = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
            ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
            ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kats:
        got = SR.philox4x32_10(np.array([ctr], dtype=np.uint32), key)[0]
        assert tuple(int(v) for v in got) == want


def test_device_noise_restatement_statistics_and_add_noise():
    import numpy as np
    import torch
    from oracle import sampler_ref as SR

    n, t = SR.device_noise(5, 3, 16, 4096)
    assert abs(float(n.mean())) < 2e-2 and abs(float(n.std()) - 1) < 2e-2 and t.min() >= 0 and t.max() < 1000
    n2, t2 = SR.device_noise(5, 4, 16, 4096)
    assert not np.array_equal(t, t2) and not np.array_equal(n, n2)
    x0 = torch.randn(16, 4, 32, 32)
    nz = torch.from_numpy(n).view(16, 4, 32, 32)
    tt = torch.from_numpy(t)
    ac = SR.alphas_cumprod()
    y = SR.add_noise(x0, nz, tt)
    b = 3
    assert torch.allclose(y[b], (ac[t[b]].sqrt() * x0[b] + (1 - ac[t[b]]).sqrt() * nz[b]).float(), atol=1e-6)
    v = SR.get_velocity(x0, nz, tt)
    assert torch.allclose(v[b], (ac[t[b]].sqrt() * nz[b] - (1 - ac[t[b]]).sqrt() * x0[b]).float(), atol=1e-6)


def test_vae_oracle_matches_sd15_anchors():
    """oracle/vae_ref.AutoencoderKL: exact parameter count of the SD-1.5 VAE (83 653 863) and the diffusers-0.13 key names the
    published vae/diffusion_pytorch_model.* files use; encode -> decode round-trips shapes."""
    import torch
    from oracle import vae_ref as VR

    m = VR.AutoencoderKL()
    assert sum(p.numel() for p in m.parameters()) == 83_653_863
    keys = set(m.state_dict().keys())
    for k in ("encoder.conv_in.weight", "encoder.down_blocks.0.resnets.1.conv2.bias", "encoder.down_blocks.2.downsamplers.0.conv.weight",
              "encoder.down_blocks.1.resnets.0.conv_shortcut.weight", "encoder.mid_block.attentions.0.query.weight",
              "encoder.mid_block.attentions.0.proj_attn.bias", "encoder.conv_norm_out.weight", "quant_conv.weight", "post_quant_conv.bias",
              "decoder.up_blocks.3.resnets.2.norm1.weight", "decoder.up_blocks.0.upsamplers.0.conv.weight", "decoder.conv_out.bias"):
        assert k in keys, k
    assert "encoder.down_blocks.3.downsamplers.0.conv.weight" not in keys and "decoder.up_blocks.3.upsamplers.0.conv.weight" not in keys
    t = VR.init_synthetic_(VR.AutoencoderKL(block_out_channels=(32, 64), layers_per_block=1), 0)
    with torch.no_grad():
        z = t.encode_sample(torch.rand(1, 3, 32, 32) * 2 - 1, torch.zeros(1, 4, 16, 16))
        assert z.shape == (1, 4, 16, 16) and t.decode(z).shape == (1, 3, 32, 32)
        mean, logvar = t.encode_moments(torch.zeros(1, 3, 32, 32))
        assert float(logvar.max()) <= 20.0 and float(logvar.min()) >= -30.0


def test_clip_oracle_matches_transformers():
    """oracle/clip_ref.clip_text_forward is pinned to the reference's own dependency: transformers.CLIPTextModel (the class
    train_text_to_image_control_lora.py:401-403 instantiates) with the same synthetic weights, CPU fp32; plus the parameter
    count of the SD-1.5 text encoder (123 060 480)."""
    import pytest
    import torch
    from oracle import clip_ref as CR

    tr = pytest.importorskip("transformers")
    cfg = dict(CR.SD15_TEXT_CONFIG)
    cfg.update(num_hidden_layers=2, vocab_size=1000)
    sd = CR.synthetic_state_dict(cfg, seed=0)
    hc = tr.CLIPTextConfig(vocab_size=cfg["vocab_size"], hidden_size=cfg["hidden_size"], intermediate_size=cfg["intermediate_size"],
                           num_hidden_layers=cfg["num_hidden_layers"], num_attention_heads=cfg["num_attention_heads"],
                           max_position_embeddings=cfg["max_position_embeddings"], hidden_act="quick_gelu", layer_norm_eps=1e-5)
    m = tr.CLIPTextModel(hc).eval()
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all("position_ids" in k for k in missing), (missing, unexpected)
    ids = torch.randint(0, cfg["vocab_size"], (2, 77), generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        want = m(ids)[0]
    got = CR.clip_text_forward(sd, ids, cfg)
    assert float((got - want).abs().max()) < 1e-4
    short = CR.clip_text_forward(sd, ids[:, :9], cfg)
    with torch.no_grad():
        assert float((short - m(ids[:, :9])[0]).abs().max()) < 1e-4
    full = CR.SD15_TEXT_CONFIG
    Cw, Fi, L = full["hidden_size"], full["intermediate_size"], full["num_hidden_layers"]
    n = (full["vocab_size"] + full["max_position_embeddings"]) * Cw + L * (4 * (Cw * Cw + Cw) + 2 * Cw * Fi + Fi + Cw + 4 * Cw) + 2 * Cw
    assert n == 123060480


def test_vae_oracle_encoder_decoder_match_the_ldm_implementation_in_transformers():
    """oracle/vae_ref.py's Encoder / Decoder are pinned to an independent published implementation of the same network: the
    CompVis latent-diffusion / taming `Encoder` and `Decoder` (the code SD-1.5's VAE was trained with and diffusers' AutoencoderKL
    re-implements), which `transformers` ships as JanusVQVAEEncoder / JanusVQVAEDecoder.  Same synthetic weights (key names mapped
    diffusers -> LDM: down_blocks.i.resnets.j -> down.i.block.j, conv_shortcut -> nin_shortcut, mid_block.resnets.0/1 -> mid.block_1/2,
    attention Linear [C, C] -> 1x1 conv), CPU fp32.  Janus adds attention blocks at the lowest resolution level, which SD's config
    (attn_resolutions = []) does not have: their output projection is zeroed, which makes them the identity (x + 0)."""
    import pytest
    import torch
    from oracle import vae_ref as VR

    pytest.importorskip("transformers")
    from transformers.models.janus.configuration_janus import JanusVQVAEConfig
    from transformers.models.janus.modeling_janus import JanusVQVAEDecoder, JanusVQVAEEncoder

    chans, layers = (32, 64, 64), 1
    cfg = JanusVQVAEConfig(base_channels=32, channel_multiplier=[c // 32 for c in chans], num_res_blocks=layers, in_channels=3,
                           out_channels=3, latent_channels=4, double_latent=True, dropout=0.0)
    torch.manual_seed(0)
    enc, dec = JanusVQVAEEncoder(cfg).eval(), JanusVQVAEDecoder(cfg).eval()
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for m in (enc, dec):
            for n, p in m.named_parameters():
                p.copy_(torch.randn(p.shape, generator=g) * (0.3 if p.dim() == 1 else 0.08) + (1.0 if ("norm" in n and n.endswith("weight")) else 0.0))
        for blk in (enc.down[-1], dec.up[0]):                      # level attention blocks -> identity
            for a in blk.attn:
                a.proj_out.weight.zero_()
                a.proj_out.bias.zero_()
    ovae = VR.AutoencoderKL(block_out_channels=chans, layers_per_block=layers)

    def res(dst, src):
        out = {f"{dst}.{k}.{w}": f"{src}.{k}.{w}" for k in ("norm1", "conv1", "norm2", "conv2") for w in ("weight", "bias")}
        out.update({f"{dst}.conv_shortcut.{w}": f"{src}.nin_shortcut.{w}" for w in ("weight", "bias")})
        return out

    def mid(dst, src):
        out = {**res(f"{dst}.resnets.0", f"{src}.block_1"), **res(f"{dst}.resnets.1", f"{src}.block_2")}
        for a, b in (("group_norm", "norm"), ("query", "q"), ("key", "k"), ("value", "v"), ("proj_attn", "proj_out")):
            out.update({f"{dst}.attentions.0.{a}.{w}": f"{src}.attn_1.{b}.{w}" for w in ("weight", "bias")})
        return out

    def load(omod, jmod, mapping):
        jsd, used = jmod.state_dict(), set()
        with torch.no_grad():
            for k, p in omod.state_dict().items():
                src = jsd[mapping[k]]
                used.add(mapping[k])
                p.copy_(src.reshape(p.shape))                      # [C, C, 1, 1] conv <-> [C, C] Linear
        rest = [k for k in jsd if k not in used]
        assert all(".attn." in k for k in rest), rest              # only the (identity) level attention blocks are unmapped

    emap = {f"conv_in.{w}": f"conv_in.{w}" for w in ("weight", "bias")}
    emap.update({f"conv_norm_out.{w}": f"norm_out.{w}" for w in ("weight", "bias")})
    emap.update({f"conv_out.{w}": f"conv_out.{w}" for w in ("weight", "bias")})
    dmap = dict(emap)
    for i in range(len(chans)):
        for j in range(layers):
            emap.update(res(f"down_blocks.{i}.resnets.{j}", f"down.{i}.block.{j}"))
        emap.update({f"down_blocks.{i}.downsamplers.0.conv.{w}": f"down.{i}.downsample.conv.{w}" for w in ("weight", "bias")})
        for j in range(layers + 1):
            dmap.update(res(f"up_blocks.{i}.resnets.{j}", f"up.{i}.block.{j}"))
        dmap.update({f"up_blocks.{i}.upsamplers.0.conv.{w}": f"up.{i}.upsample.conv.{w}" for w in ("weight", "bias")})
    emap.update(mid("mid_block", "mid"))
    dmap.update(mid("mid_block", "mid"))
    load(ovae.encoder, enc, emap)
    load(ovae.decoder, dec, dmap)
    x = torch.randn(2, 3, 32, 32, generator=g)
    z = torch.randn(2, 4, 8, 8, generator=g)
    with torch.no_grad():
        e_want, e_got = enc(x.clone()), ovae.encoder(x)
        d_want, d_got = dec(z.clone()), ovae.decoder(z)
    assert e_got.shape == e_want.shape == (2, 8, 8, 8) and d_got.shape == d_want.shape == (2, 3, 32, 32)
    assert float((e_got - e_want).abs().max()) < 1e-4 * float(e_want.abs().max())
    assert float((d_got - d_want).abs().max()) < 1e-4 * float(d_want.abs().max())


def test_unet_oracle_blocks_match_independent_implementations():
    """diffusers itself cannot run here, so the WHOLE-UNet restatement (oracle/unet_ref.py) stays unpinned; its building blocks are
    pinned to independent implementations of the same published blocks:
      BasicTransformerBlock (LN -> self-attn -> +, LN -> cross-attn -> +, LN -> GEGLU FF -> +)   == torch.nn.TransformerDecoderLayer(norm_first)
      ResnetBlock2D (time-embedding projection zeroed), Upsample2D, Downsample2D(padding=0)       == the LDM blocks shipped in transformers
      sinusoidal timestep embedding (flip_sin_to_cos, freq_shift 0)                                == the closed form, evaluated in float64."""
    import math

    import pytest
    import torch
    import torch.nn as nn
    import torch.nn.functional as F
    from oracle import unet_ref as UR

    torch.manual_seed(0)
    # ---- transformer block
    dim, heads, T, S, B = 64, 4, 24, 9, 2
    blk = UR.BasicTransformerBlock(dim, heads, dim // heads, dim)          # cross_attention_dim == dim: nn.MultiheadAttention's memory width
    with torch.no_grad():
        for p in blk.parameters():
            p.copy_(torch.randn(p.shape) * (0.2 if p.dim() > 1 else 0.1) + (1.0 if p.dim() == 1 and p.numel() == dim and p is not blk.ff.net[2].bias else 0.0))
    Fd = blk.ff.net[2].in_features

    def geglu_padded(p):                                                     # GEGLU as an `activation`: output padded back to 2F for linear2
        a, g = p.chunk(2, dim=-1)
        return torch.cat([a * F.gelu(g), torch.zeros_like(a)], -1)

    ref = nn.TransformerDecoderLayer(dim, heads, dim_feedforward=2 * Fd, dropout=0.0, activation=geglu_padded, batch_first=True, norm_first=True)
    with torch.no_grad():
        for mha, att in ((ref.self_attn, blk.attn1), (ref.multihead_attn, blk.attn2)):
            mha.in_proj_weight.copy_(torch.cat([att.to_q.weight, att.to_k.weight, att.to_v.weight], 0))
            mha.in_proj_bias.zero_()
            mha.out_proj.weight.copy_(att.to_out[0].weight)
            mha.out_proj.bias.copy_(att.to_out[0].bias)
        for a, b in ((ref.norm1, blk.norm1), (ref.norm2, blk.norm2), (ref.norm3, blk.norm3)):
            a.weight.copy_(b.weight); a.bias.copy_(b.bias)
        ref.linear1.weight.copy_(blk.ff.net[0].proj.weight); ref.linear1.bias.copy_(blk.ff.net[0].proj.bias)
        ref.linear2.weight.zero_()
        ref.linear2.weight[:, :Fd].copy_(blk.ff.net[2].weight); ref.linear2.bias.copy_(blk.ff.net[2].bias)
    x, mem = torch.randn(B, T, dim), torch.randn(B, S, dim)
    with torch.no_grad():
        want = ref.eval()(x, mem)
        got = blk(x, encoder_hidden_states=mem)
    assert float((got - want).abs().max()) < 2e-5 * float(want.abs().max())
    # ---- conv blocks (LDM implementation in transformers)
    pytest.importorskip("transformers")
    from transformers.models.janus.configuration_janus import JanusVQVAEConfig
    from transformers.models.janus.modeling_janus import JanusVQVAEConvDownsample, JanusVQVAEConvUpsample, JanusVQVAEResnetBlock

    cfg = JanusVQVAEConfig(dropout=0.0)
    for cin, cout in ((32, 64), (64, 64)):
        r = UR.ResnetBlock2D(cin, cout, temb_channels=16, groups=32, eps=1e-6)
        j = JanusVQVAEResnetBlock(cfg, cin, cout).eval()
        with torch.no_grad():
            for p in r.parameters():
                p.copy_(torch.randn(p.shape) * 0.1)
            r.time_emb_proj.weight.zero_(); r.time_emb_proj.bias.zero_()
            for a, b in (("norm1", "norm1"), ("conv1", "conv1"), ("norm2", "norm2"), ("conv2", "conv2")):
                getattr(j, b).weight.copy_(getattr(r, a).weight); getattr(j, b).bias.copy_(getattr(r, a).bias)
            if cin != cout:
                j.nin_shortcut.weight.copy_(r.conv_shortcut.weight); j.nin_shortcut.bias.copy_(r.conv_shortcut.bias)
            xi = torch.randn(2, cin, 12, 12)
            assert float((r(xi, torch.randn(2, 16)) - j(xi.clone())).abs().max()) < 1e-4
    up, jup = UR.Upsample2D(32), JanusVQVAEConvUpsample(32)
    dn, jdn = UR.Downsample2D(32, padding=0), JanusVQVAEConvDownsample(32)
    with torch.no_grad():
        jup.conv.load_state_dict(up.conv.state_dict()); jdn.conv.load_state_dict(dn.conv.state_dict())
        xi = torch.randn(2, 32, 10, 10)
        assert torch.allclose(up(xi), jup(xi), atol=1e-6) and torch.allclose(dn(xi), jdn(xi), atol=1e-6)
    # ---- timestep embedding: [cos | sin](t * 10000^(-i / half))
    t = torch.tensor([0.0, 1.0, 17.0, 999.0])
    emb = UR.sinusoidal_embedding(t, 320)
    i = torch.arange(160, dtype=torch.float64)
    ang = t.double()[:, None] * torch.exp(-math.log(10000.0) * i / 160)[None]
    assert torch.allclose(emb.double(), torch.cat([ang.cos(), ang.sin()], -1), atol=2e-4)       # fp32 argument reduction at t = 999


def test_scheduler_restatements_are_exact_on_the_point_mass_ode():
    """Property pin of oracle/sampler_ref.py (diffusers' DDIMScheduler / DPMSolverMultistepScheduler cannot be imported): for data
    concentrated on one point x* the optimal denoiser is eps(x, t) = (x - alpha_t x*) / sigma_t, the probability-flow ODE has the closed
    form x_t = alpha_t x* + (sigma_t / sigma_T)(x_T - alpha_T x*), and both deterministic DDIM and DPM-Solver++(2M) are EXACT on it for
    any number of steps (constant x0 prediction, constant noise direction).  A wrong first-order coefficient, timestep table, final-step rule or
    multistep history in the restatement breaks the identity (the second-order correction of 2M multiplies the CHANGE of the x0 prediction,
    which is zero here: its coefficient is covered by the committed golden trajectories only)."""
    import torch
    from oracle import sampler_ref as SR

    g = torch.Generator().manual_seed(0)
    xs, xT = torch.randn(2, 4, 8, 8, generator=g, dtype=torch.float64), torch.randn(2, 4, 8, 8, generator=g, dtype=torch.float64)
    ac = SR.alphas_cumprod()
    al, sg = ac.sqrt(), (1 - ac).sqrt()
    eps_star = lambda x, t: (x - al[t] * xs) / sg[t]
    for steps in (5, 20, 50):
        ts = [int(t) for t in SR.timesteps(steps)]
        x = xT.clone()
        for t in ts:
            e = eps_star(x, t)
            x = SR.cfg_ddim_step(e, e, x, t, steps, 7.5)                      # uncond == cond: the guidance term vanishes
        t_end = max(ts[-1] - 1000 // steps, 0)
        want = al[t_end] * xs + (sg[t_end] / sg[ts[0]]) * (xT - al[ts[0]] * xs)
        assert float((x - want).abs().max()) < 1e-9, ("ddim", steps)
    for steps in (4, 12, 30):                                                  # < 15: the last step is first order (lower_order_final)
        s = SR.DPMSolverPP2M(steps)
        x = xT.clone()
        for t in s.timesteps:
            x = s.step(eps_star(x, t), t, x)
        T = s.timesteps[0]
        want = al[0] * xs + (sg[0] / sg[T]) * (xT - al[T] * xs)
        assert float((x - want).abs().max()) < 1e-9, ("dpmpp", steps)
    # add_noise / get_velocity are the closed forms of the forward process (DDPMScheduler.add_noise / get_velocity)
    t = torch.tensor([3, 700])
    n = torch.randn(2, 4, 8, 8, generator=g, dtype=torch.float64)
    a, b = al[t].view(2, 1, 1, 1), sg[t].view(2, 1, 1, 1)
    assert torch.allclose(SR.add_noise(xs, n, t).double(), a * xs + b * n, atol=1e-6)
    assert torch.allclose(SR.get_velocity(xs, n, t).double(), a * n - b * xs, atol=1e-6)
