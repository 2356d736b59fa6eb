"""Generates tests/golden/oracle_tiny.pt: golden vectors of the fp32 oracle on the tiny SD-style configuration
(noise prediction, loss, a digest of every ControlLoRA gradient).  The reference ships no golden vectors and cannot be
imported (diffusers is absent), so these pin the *oracle*; the CUDA path is compared with the oracle on the same inputs.

    python -m tests.golden.make_golden
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from oracle import models_ref as MR  # noqa: E402
from oracle import unet_ref as UR  # noqa: E402

TINY = dict(block_out_channels=(64, 128, 128, 128), layers_per_block=1, cross_attention_dim=64, attention_head_dim=8)
TINY_LORA = dict(lora_block_out_channels=(64, 128, 128, 128),
                 lora_cross_attention_dims=([None, 64] * 3, [None, 64] * 3, [None, 64] * 3, [None, 64]))


def run_variant(variant: str):
    torch.manual_seed(0)
    torch.set_num_threads(1)       # deterministic reduction order
    unet = UR.UNet2DConditionModel(**TINY)
    UR.init_synthetic_(unet, seed=1)
    unet.requires_grad_(False)
    kw = dict(TINY_LORA)
    if variant == "v2":
        kw.update(lora_control_version=2, lora_pre_conv_skipped=True, lora_key_states_skipped=True, lora_value_states_skipped=True)
    if variant == "v1_post_add":        # configs/post-add.json flavour
        kw.update(lora_post_add=True)
    if variant == "v1_concat":          # configs/danbooru-sketch.json flavour (control MLP over [h ; c])
        kw.update(lora_concat_hidden=True, lora_control_rank=32, lora_pre_conv_skipped=True, lora_control_self_add=False)
    cl = MR.ControlLoRA(**kw)
    MR.randomize_lora_up_(cl, seed=3, std=0.05)
    MR.wire_processors(unet, cl)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 4, 16, 16, generator=g)
    t = torch.tensor([17, 801])
    e = torch.randn(2, 77, 64, generator=g)
    guide = torch.rand(2, 3, 128, 128, generator=g) * 2 - 1
    tgt = torch.randn(2, 4, 16, 16, generator=g)
    states = cl(guide).control_states
    pred = unet(x, t, e).sample
    loss = torch.nn.functional.mse_loss(pred, tgt)
    loss.backward()
    out = {"pred": pred.detach().clone(), "loss": loss.detach().clone(),
           "state0": states[0].detach()[:, :8, :4, :4].clone()}
    gn = []
    for n, p in cl.named_parameters():
        gn.append(p.grad.double().norm().float() if p.grad is not None else torch.zeros(()))
    out["grad_norms"] = torch.stack(gn)
    return out


def run_samplers():
    """Trajectories of the two scheduler restatements (oracle/sampler_ref.py) on a fixed synthetic eps model."""
    from oracle import sampler_ref as SR

    g = torch.Generator().manual_seed(0)
    w = torch.randn(4, 4, generator=g) * 0.3
    x0 = torch.randn(2, 4, 8, 8, generator=g)
    eps = lambda v, t: torch.einsum("oc,bchw->bohw", w, v) * (0.5 + t / 2000.0)
    out = {}
    x = x0.clone()
    for t in SR.timesteps(10):
        x = SR.cfg_ddim_step(eps(x, int(t)) * 0.9, eps(x, int(t)), x, t, 10, 7.5)
    out["ddim10"] = x.clone()
    sched = SR.DPMSolverPP2M(12)
    x = x0.clone()
    for t in sched.timesteps:
        x = sched.step(SR.cfg_combine(eps(x, t) * 0.9, eps(x, t), 7.5), t, x)
    out["dpmpp12"] = x.clone()
    out["dpm_timesteps_30"] = torch.tensor(SR.DPMSolverPP2M(30).timesteps)
    return out


if __name__ == "__main__":
    res = {v: run_variant(v) for v in ("v1", "v2")}
    path = Path(__file__).resolve().parent / "oracle_tiny.pt"
    torch.save(res, path)
    print("wrote", path, {k: float(v["loss"]) for k, v in res.items()})
    extra = {v: run_variant(v) for v in ("v1_post_add", "v1_concat")}
    extra["samplers"] = run_samplers()
    path2 = Path(__file__).resolve().parent / "oracle_extra.pt"
    torch.save(extra, path2)
    print("wrote", path2, {k: float(v["loss"]) for k, v in extra.items() if "loss" in v})
