"""Host-side logic of the drop-in surface (no GPU): class surface / configs / state-dict compatibility with the
oracle's restatement of /root/reference/models.py, processor wiring, error conventions, checkpoint round trip."""
import json

import pytest
import torch

import controllora_b200 as cb
from controllora_b200.configs import NAMED, wire_processors
from oracle import models_ref as MR


@pytest.mark.parametrize("name", sorted(NAMED))
def test_state_dict_keys_and_shapes_equal_the_reference_layout(name):
    ours = cb.ControlLoRA.from_config(NAMED[name]).state_dict()
    ref = MR.ControlLoRA.from_config(NAMED[name]).state_dict()
    assert list(ours.keys()) == list(ref.keys())
    for k in ours:
        assert ours[k].shape == ref[k].shape, k


def test_from_config_accepts_path_dict_and_ignores_private_keys(tmp_path):
    cfg = {"_class_name": "ControlLoRA", "_diffusers_version": "0.13.0.dev0", "lora_rank": 4, "lora_control_version": 2,
           "lora_pre_conv_skipped": True}
    p = tmp_path / "c.json"
    p.write_text(json.dumps(cfg))
    a = cb.ControlLoRA.from_config(str(p))
    b = cb.ControlLoRA.from_config(cfg)
    assert type(a.lora_layers[0][0]).__name__ == "ControlLoRACrossAttnProcessorV2"
    assert list(a.state_dict()) == list(b.state_dict())
    with pytest.raises(TypeError):
        cb.ControlLoRA.from_config({"not_a_key": 1})


def test_save_and_load_pretrained_roundtrip(tmp_path):
    m = cb.ControlLoRA.from_config(NAMED["diffusiondb-canny-v2"])
    with torch.no_grad():
        for p in m.parameters():
            p.add_(torch.randn_like(p) * 0.01)
    m.save_pretrained(tmp_path / "bin")
    m.save_pretrained(tmp_path / "st", safe_serialization=True)
    for sub in ("bin", "st"):
        r = cb.ControlLoRA.from_pretrained(tmp_path, subfolder=sub)
        for (k1, v1), (k2, v2) in zip(m.state_dict().items(), r.state_dict().items()):
            assert k1 == k2 and torch.equal(v1, v2)
    saved = json.loads((tmp_path / "bin" / "config.json").read_text())
    assert saved["_class_name"] == "ControlLoRA" and saved["lora_control_version"] == 2


def test_lora_linear_layer_init_and_rank_check():
    l = cb.LoRALinearLayer(64, 32, 4)
    assert l.down.weight.shape == (4, 64) and l.up.weight.shape == (32, 4)
    assert float(l.up.weight.abs().max()) == 0.0           # diffusers: up = 0
    assert 0.1 < float(l.down.weight.std()) < 0.4           # ~ N(0, 1/rank)
    with pytest.raises(ValueError):
        cb.LoRALinearLayer(2, 64, 4)


def test_processor_flags_and_skip_helpers():
    p = cb.LoRACrossAttnProcessor(64, 32, rank=4, key_states_skipped=True)
    assert not hasattr(p, "to_k_lora") and hasattr(p, "to_v_lora")
    assert p.to_v_lora.down.weight.shape == (4, 32)
    with pytest.raises(AssertionError):
        p.skip_key_states(False)
    p.skip_output_states(True)
    assert p.output_states_skipped
    v1 = cb.ControlLoRACrossAttnProcessor(64, None, control_self_add=True)
    assert v1.control_self_add is False                      # models.py:180-182 quirk
    assert v1.to_control.down.weight.shape == (4, 64)
    cat = cb.ControlLoRACrossAttnProcessor(64, None, concat_hidden=True, control_channels=256, control_rank=8)
    assert cat.to_control.down.weight.shape == (8, 320)
    v2 = cb.ControlLoRACrossAttnProcessorV2(64, 32, control_channels=256)
    assert v2.key_states_skipped and v2.value_states_skipped and v2.to_control_out.down.weight.shape == (4, 320)
    lo = cb.LoRACrossAttnProcessor(64)
    v1.inject_pre_lora(lo)
    assert v1.pre_loras == [lo] and "pre_loras" not in dict(v1.named_modules())   # plain lists, not sub-modules


def test_processors_have_no_cpu_fallback(monkeypatch):
    """A processor is callable the way diffusers calls it (controllora_b200/eager_attn.py) - on CUDA.  CPU tensors are refused:
    there is no PyTorch implementation of the arithmetic in the package."""
    from oracle import unet_ref as UR

    monkeypatch.delenv("CLB_DRYRUN", raising=False)
    p = cb.LoRACrossAttnProcessor(64)
    attn = UR.CrossAttention(64, None, 8, 8)
    with pytest.raises(RuntimeError, match="runs only on CUDA"):
        p(attn, torch.zeros(1, 4, 64))
    with pytest.raises(RuntimeError, match="runs only on CUDA"):
        cb.LoRALinearLayer(64, 64, 4)(torch.zeros(2, 64))
    with pytest.raises(NotImplementedError):
        p(attn, torch.zeros(1, 4, 64), attention_mask=torch.zeros(1, 4, 4))


def test_unet_wrapper_refuses_cpu_inputs(monkeypatch):
    from controllora_b200.unet_module import UNet2DConditionModel

    monkeypatch.delenv("CLB_DRYRUN", raising=False)
    with pytest.raises(RuntimeError):
        UNet2DConditionModel._prep_inputs(torch.zeros(1, 4, 8, 8), torch.tensor([1]), torch.zeros(1, 77, 64))


TINY = dict(block_out_channels=(64, 128, 128, 128), layers_per_block=1, cross_attention_dim=64, attention_head_dim=8)
TINY_LORA = dict(lora_block_out_channels=(64, 128, 128, 128),
                 lora_cross_attention_dims=([None, 64] * 3, [None, 64] * 3, [None, 64] * 3, [None, 64]))


def _tiny_unet():
    from controllora_b200.unet import synthetic_state_dict

    return cb.UNet2DConditionModel.from_state_dict(synthetic_state_dict(TINY, 0), "cpu", TINY)


def test_attn_processor_names_and_wiring_match_the_oracle():
    from oracle import unet_ref as UR

    mu = _tiny_unet()
    ou = UR.UNet2DConditionModel(**TINY)
    assert list(mu.attn_processors.keys()) == list(ou.attn_processors.keys())
    mcl = cb.ControlLoRA(**TINY_LORA)
    ocl = MR.ControlLoRA(**TINY_LORA)
    mp = wire_processors(mu, mcl)
    op = MR.wire_processors(ou, ocl)
    idx = lambda cl, p: [(i, j) for i, l in enumerate(cl.lora_layers) for j, q in enumerate(l) if q is p][0]
    assert {k: idx(mcl, v) for k, v in mp.items()} == {k: idx(ocl, v) for k, v in op.items()}
    with pytest.raises(ValueError):
        mu.set_attn_processor({"only.one": mcl.lora_layers[0][0]})
    # processors become sub-modules of the UNet wrapper as well (shared ownership, like diffusers)
    assert any(p is mcl.lora_layers[0][0].to_q_lora.down.weight for p in mu.parameters())
    assert mu.config.cross_attention_dim == 64 and tuple(mu.config.block_out_channels) == (64, 128, 128, 128)


def test_synthetic_state_dict_has_diffusers_keys_and_sd15_size():
    from controllora_b200.unet import SD15_CONFIG, synthetic_state_dict
    from oracle import unet_ref as UR

    with torch.device("meta"):
        ref = UR.UNet2DConditionModel(**TINY)
    sd = synthetic_state_dict(TINY, 0)
    assert set(sd.keys()) == set(ref.state_dict().keys())
    for k, v in ref.state_dict().items():
        assert tuple(sd[k].shape) == tuple(v.shape), k


def test_wirings_beyond_the_fused_path_select_the_general_chain_path():
    """The one-launch path covers the shipped configs; every other wiring models.py:118-431 allows is routed, per attention
    layer, to the general chain path (lora_generic.py) instead of being refused (numerics: tests/check_variants.py)."""
    from controllora_b200.lora_runtime import LoraRuntime
    from controllora_b200.unet_module import GradStore

    def kinds(rt):
        return {lp.kind for lp in rt.layers.values() if lp.proc is not None}

    mu = _tiny_unet()
    mcl = cb.ControlLoRA(lora_concat_hidden=True, **TINY_LORA)          # configs/danbooru-sketch.json flavour ...
    procs = wire_processors(mu, mcl)
    assert kinds(LoraRuntime(mu.weights, torch.device("cpu"), GradStore().get)) == {"v1cat"}      # ... fused on its own,
    for p in procs.values():                                            # general path with another adapter stacked on it
        p.inject_pre_lora(cb.LoRACrossAttnProcessor(p.hidden_size, p.cross_attention_dim, rank=4))
    rt = LoraRuntime(mu.weights, torch.device("cpu"), GradStore().get)
    assert kinds(rt) == {"generic"} and all("concat_hidden" in lp.generic_reason for lp in rt.layers.values() if lp.proc is not None)
    mu = _tiny_unet()
    mcl = cb.ControlLoRA(lora_post_add=True, **TINY_LORA)
    procs = wire_processors(mu, mcl)
    for p in procs.values():
        p.inject_pre_lora(cb.LoRACrossAttnProcessor(p.hidden_size, p.cross_attention_dim, rank=4))
    assert kinds(LoraRuntime(mu.weights, torch.device("cpu"), GradStore().get)) == {"generic"}
    # the standard stacked wiring (mix_lora_and_control_lora.py: one rank-4 pre-LoRA) stays on the fused path
    mu = _tiny_unet()
    procs = wire_processors(mu, cb.ControlLoRA(**TINY_LORA))
    for p in procs.values():
        p.inject_pre_lora(cb.LoRACrossAttnProcessor(p.hidden_size, p.cross_attention_dim, rank=4))
    assert kinds(LoraRuntime(mu.weights, torch.device("cpu"), GradStore().get)) == {"v1"}


def test_post_add_config_builds_a_runtime_plan():
    """lora_post_add=True (configs/post-add.json): every slot is flagged, and the text k / v adapters take C inputs."""
    from controllora_b200.lora_runtime import LoraRuntime
    from controllora_b200.unet_module import GradStore

    mu = _tiny_unet()
    mcl = cb.ControlLoRA(lora_post_add=True, **TINY_LORA)
    procs = wire_processors(mu, mcl)
    rt = LoraRuntime(mu.weights, torch.device("cpu"), GradStore().get)
    n = 0
    for lp in rt.layers.values():
        if getattr(lp, "proc", None) is None:
            continue
        n += 1
        assert lp.post_add and all(sl.post_add for sl in (lp.q, lp.k, lp.v, lp.out))
        assert lp.k.K == lp.k.N and lp.v.K == lp.v.N
    assert n == len(procs) > 0


@pytest.mark.parametrize("steps", [8, 30, 50])
def test_dpmpp_2m_coefficients_reproduce_the_oracle_scheduler(steps):
    """sampler.dpmpp_2m_coeffs (the five scalars the fused CFG + DPM-Solver++ kernel consumes) against the stateful
    restatement of diffusers' DPMSolverMultistepScheduler: same timesteps, same trajectory on a synthetic eps model."""
    from controllora_b200.sampler import dpm_timesteps, dpmpp_2m_coeffs, sd15_alphas_cumprod
    from oracle import sampler_ref as SR

    sched = SR.DPMSolverPP2M(steps)
    ts = dpm_timesteps(steps)
    assert ts == sched.timesteps
    ac = sd15_alphas_cumprod()
    g = torch.Generator().manual_seed(0)
    x_ref = torch.randn(2, 4, 8, 8, generator=g, dtype=torch.float64)
    x = x_ref.clone()
    x0_prev = torch.zeros_like(x)
    w = torch.randn(4, 4, generator=g, dtype=torch.float64) * 0.3
    for i, t in enumerate(ts):
        eps_fn = lambda v: torch.einsum("oc,bchw->bohw", w, v) * (0.5 + t / 2000.0)      # any deterministic "model"
        x_ref = sched.step(eps_fn(x_ref), t, x_ref)
        a_s, sg_s, c_x, c_m0, c_m1 = dpmpp_2m_coeffs(i, ts, ac)
        x0 = (x - sg_s * eps_fn(x)) / a_s
        x, x0_prev = c_x * x + c_m0 * x0 + c_m1 * x0_prev, x0
        assert torch.allclose(x, x_ref, rtol=1e-9, atol=1e-9), (i, t)


def test_unet_from_pretrained_reads_a_diffusers_directory(tmp_path):
    """config.json + diffusion_pytorch_model.safetensors under <root>/unet, the layout of train_...:407-409."""
    import json
    from safetensors.torch import save_file
    from controllora_b200.unet import synthetic_state_dict

    sd = synthetic_state_dict(TINY, 3)
    root = tmp_path / "sd" / "unet"
    root.mkdir(parents=True)
    save_file({k: v.contiguous() for k, v in sd.items()}, str(root / "diffusion_pytorch_model.safetensors"))
    cfg = {k: (list(v) if isinstance(v, tuple) else v) for k, v in TINY.items()}
    cfg.update(_class_name="UNet2DConditionModel", sample_size=64)
    (root / "config.json").write_text(json.dumps(cfg))
    mu = cb.UNet2DConditionModel.from_pretrained(str(tmp_path / "sd"), subfolder="unet", device="cpu")
    assert tuple(mu.config.block_out_channels) == tuple(TINY["block_out_channels"])
    assert list(mu.attn_processors.keys()) == list(_tiny_unet().attn_processors.keys())
    with pytest.raises(FileNotFoundError):
        cb.UNet2DConditionModel.from_pretrained(str(tmp_path / "nope"), subfolder="unet", device="cpu")


def test_trainer_checkpoint_roundtrip(tmp_path):
    """checkpoint-N directories (reference: accelerator.save_state / resume from the highest N): parameters, AdamW moments
    and the step counter survive a save -> fresh Trainer -> load."""
    from controllora_b200.trainer import Trainer

    def make(seed):
        torch.manual_seed(seed)
        mu = _tiny_unet()
        mcl = cb.ControlLoRA(**TINY_LORA)
        wire_processors(mu, mcl)
        return Trainer(mu, mcl, lr=1e-4), mcl

    tr, cl = make(0)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        tr.flat_p[:tr.numel].copy_(torch.randn(tr.numel, generator=g))
        tr.flat_m[:tr.numel].copy_(torch.randn(tr.numel, generator=g))
        tr.flat_v[:tr.numel].copy_(torch.rand(tr.numel, generator=g))
    tr.step_idx = 37
    tr.rng_counter.fill_(9)                       # 9 noise draws so far (device Philox step counter)
    assert Trainer.latest_checkpoint(tmp_path) is None
    tr.save_checkpoint(tmp_path, global_step=30)
    p37 = tr.save_checkpoint(tmp_path)
    expected_next = torch.rand(3)                 # what the host RNG yields right after the save
    assert Trainer.latest_checkpoint(tmp_path) == p37 and p37.endswith("checkpoint-37")
    import os
    assert {"config.json", "diffusion_pytorch_model.safetensors", "optimizer.bin", "scheduler.bin", "random_states_0.pkl"} <= set(os.listdir(p37))
    tr2, cl2 = make(5)
    assert not torch.equal(tr2.flat_p, tr.flat_p)
    assert tr2.load_checkpoint(p37) == 37 and tr2.step_idx == 37
    assert int(tr2.rng_counter) == 9 and tr2.noise_seed == tr.noise_seed        # RNG state resumes (train_...:805-809 save_state)
    assert torch.equal(torch.rand(3), expected_next)
    assert torch.equal(tr2.flat_p[:tr.numel], tr.flat_p[:tr.numel])
    assert torch.equal(tr2.flat_m, tr.flat_m) and torch.equal(tr2.flat_v, tr.flat_v)
    for (n1, a), (n2, b) in zip(cl.named_parameters(), cl2.named_parameters()):
        assert n1 == n2 and torch.equal(a, b)
    # the ControlLoRA part of the checkpoint is a plain reference-format model directory
    cl3 = cb.ControlLoRA.from_pretrained(p37)
    assert all(torch.equal(a, b.cpu()) for a, b in zip(cl3.state_dict().values(), cl.state_dict().values()))


def test_trainer_checkpoint_keeps_arena_only_parameters_and_checks_names(tmp_path):
    """Stacked pre_loras are trained through the arena but are not part of control_lora's state dict (models.py:189-190):
    the checkpoint stores them by arena name and refuses to load into a differently wired Trainer."""
    from controllora_b200.trainer import Trainer

    def make(stacked):
        torch.manual_seed(0)
        mu = _tiny_unet()
        mcl = cb.ControlLoRA(**TINY_LORA)
        procs = wire_processors(mu, mcl)
        extra = []
        if stacked:
            for name, p in procs.items():
                pre = cb.LoRACrossAttnProcessor(p.hidden_size, p.cross_attention_dim, rank=4)
                p.inject_pre_lora(pre)
                extra.append(pre)
        return Trainer(mu, mcl, lr=1e-4), extra

    tr, extra = make(True)
    assert tr.numel > sum(p.numel() for p in tr.cl.parameters())
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        tr.flat_p[:tr.numel].copy_(torch.randn(tr.numel, generator=g))
    path = tr.save_checkpoint(tmp_path, global_step=3)
    tr2, extra2 = make(True)
    tr2.load_checkpoint(path)
    assert torch.equal(tr2.flat_p[:tr.numel], tr.flat_p[:tr.numel])
    assert all(torch.equal(a, b) for ea, eb in zip(extra, extra2) for a, b in zip(ea.parameters(), eb.parameters()))
    tr3, _ = make(False)
    with pytest.raises(ValueError):
        tr3.load_checkpoint(path)


def _tiny_trainer(seed):
    from controllora_b200.trainer import Trainer

    torch.manual_seed(seed)
    mu = _tiny_unet()
    mcl = cb.ControlLoRA(**TINY_LORA)
    wire_processors(mu, mcl)
    return Trainer(mu, mcl, lr=3e-4, betas=(0.9, 0.99), weight_decay=2e-2, eps=1e-7), mcl


def test_checkpoint_is_loadable_by_the_reference_stack(tmp_path):
    """`accelerator.load_state(checkpoint-N)` (train_text_to_image_control_lora.py:733) does, for the reference's objects:
    control_lora.load_state_dict(torch.load(pytorch_model.bin)), optimizer.load_state_dict(torch.load(optimizer.bin)),
    lr_scheduler.load_state_dict(torch.load(scheduler.bin)) and restores random_states_0.pkl.  The files save_checkpoint() writes
    must go through exactly those calls with the reference's classes (oracle restatement of ControlLoRA, torch.optim.AdamW,
    LambdaLR) and reproduce the arena's state."""
    import pickle

    tr, cl = _tiny_trainer(0)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        tr.flat_p[:tr.numel].copy_(torch.randn(tr.numel, generator=g))
        tr.flat_m[:tr.numel].copy_(torch.randn(tr.numel, generator=g))
        tr.flat_v[:tr.numel].copy_(torch.rand(tr.numel, generator=g))
    tr.step_idx = 11
    path = tr.save_checkpoint(tmp_path)
    ref = MR.ControlLoRA(**TINY_LORA)
    ref.load_state_dict(torch.load(f"{path}/pytorch_model.bin"))                       # strict: same keys, same shapes
    for (n1, a), (n2, b) in zip(ref.named_parameters(), cl.named_parameters()):
        assert n1 == n2 and torch.equal(a, b)
    opt = torch.optim.AdamW(ref.parameters(), lr=1.0)                                   # train_...:512-518
    opt.load_state_dict(torch.load(f"{path}/optimizer.bin", weights_only=False))
    grp = opt.param_groups[0]
    assert grp["lr"] == 3e-4 and tuple(grp["betas"]) == (0.9, 0.99) and grp["weight_decay"] == 2e-2 and grp["eps"] == 1e-7
    off = 0
    for p in ref.parameters():
        st = opt.state[p]
        k = p.numel()
        assert float(st["step"]) == 11
        assert torch.equal(st["exp_avg"].flatten(), tr.flat_m[off:off + k]) and torch.equal(st["exp_avg_sq"].flatten(), tr.flat_v[off:off + k])
        off += k
    assert off == tr.numel
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda _: 1.0)                       # diffusers get_scheduler("constant")
    sched.load_state_dict(torch.load(f"{path}/scheduler.bin", weights_only=False))
    assert sched.last_epoch == 11 and sched.get_last_lr() == [3e-4]
    with open(f"{path}/random_states_0.pkl", "rb") as f:
        rs = pickle.load(f)
    assert {"random_state", "numpy_random_seed", "torch_manual_seed", "torch_cuda_manual_seed"} <= set(rs)      # accelerate's keys
    # and one optimizer step of the reference stack from this state is what the arena formulas give
    grads = [torch.randn(p.shape, generator=g) for p in ref.parameters()]
    for p, gr in zip(ref.parameters(), grads):
        p.grad = gr.clone()
    before = torch.cat([p.detach().flatten() for p in ref.parameters()])
    opt.step()
    after = torch.cat([p.detach().flatten() for p in ref.parameters()])
    gflat = torch.cat([x.flatten() for x in grads])
    m = 0.9 * tr.flat_m[:tr.numel] + 0.1 * gflat
    v = 0.99 * tr.flat_v[:tr.numel] + 0.01 * gflat * gflat
    want = before * (1 - 3e-4 * 2e-2) - (3e-4 / (1 - 0.9 ** 12)) * m / (v.sqrt() / (1 - 0.99 ** 12) ** 0.5 + 1e-7)
    assert torch.allclose(after, want, rtol=1e-5, atol=1e-7)


def test_checkpoint_written_by_the_reference_stack_resumes_here(tmp_path):
    """The reverse direction: a `checkpoint-N` as `accelerator.save_state` leaves it (model / optimizer / scheduler pickles of the
    reference's objects + accelerate's random_states file) restores the arena, the moments and the step counter."""
    import pickle
    import random

    import numpy as np

    ref = MR.ControlLoRA(**TINY_LORA)
    opt = torch.optim.AdamW(ref.parameters(), lr=2e-4, betas=(0.9, 0.999), weight_decay=1e-2, eps=1e-8)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda _: 1.0)
    g = torch.Generator().manual_seed(4)
    for _ in range(3):
        for p in ref.parameters():
            p.grad = torch.randn(p.shape, generator=g)
        opt.step()
        sched.step()
    d = tmp_path / "checkpoint-3"
    d.mkdir()
    torch.save(ref.state_dict(), d / "pytorch_model.bin")
    torch.save(opt.state_dict(), d / "optimizer.bin")
    torch.save(sched.state_dict(), d / "scheduler.bin")
    with open(d / "random_states_0.pkl", "wb") as f:
        pickle.dump({"random_state": random.getstate(), "numpy_random_seed": np.random.get_state(),
                     "torch_manual_seed": torch.get_rng_state(), "torch_cuda_manual_seed": None}, f)
    expected_next = torch.rand(2)
    tr, cl = _tiny_trainer(9)
    seed_before = tr.noise_seed
    assert tr.load_checkpoint(str(d)) == 3 and tr.step_idx == 3 and int(tr.step_dev) == 3
    assert torch.equal(torch.rand(2), expected_next) and tr.noise_seed == seed_before
    assert tr.lr == 2e-4 and tr.wd == 1e-2
    off = 0
    for (n1, a), (n2, b) in zip(ref.named_parameters(), cl.named_parameters()):
        k = a.numel()
        assert n1 == n2 and torch.equal(a.detach(), b.detach())
        assert torch.equal(opt.state[a]["exp_avg"].flatten(), tr.flat_m[off:off + k])
        assert torch.equal(opt.state[a]["exp_avg_sq"].flatten(), tr.flat_v[off:off + k])
        off += k
    # a mismatched wiring is refused, not silently mis-assigned
    opt2 = opt.state_dict()
    opt2["param_groups"][0]["params"] = opt2["param_groups"][0]["params"][:-1]
    torch.save(opt2, d / "optimizer.bin")
    with pytest.raises(ValueError):
        tr.load_checkpoint(str(d))


@pytest.mark.parametrize("name", ["constant", "constant_with_warmup", "linear", "cosine", "cosine_with_restarts"])
def test_lr_schedules_equal_transformers_get_scheduler(name):
    """`get_scheduler(args.lr_scheduler, optimizer, num_warmup_steps, num_training_steps)` (train_...:675-681; diffusers' optimization.py is
    the transformers file): the multiplier the Trainer applies to its k-th optimizer step equals the learning rate LambdaLR holds there."""
    tr_mod = pytest.importorskip("transformers")
    from controllora_b200.trainer import lr_lambda_for

    warm, total, base = 7, 50, 3e-4
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.AdamW([p], lr=base)
    sched = tr_mod.get_scheduler(name, opt, num_warmup_steps=warm, num_training_steps=total)
    lam = lr_lambda_for(name, warm, total)
    for k in range(1, 61):                       # k-th optimizer step
        ours = base if lam is None else base * lam(k - 1)
        assert abs(ours - opt.param_groups[0]["lr"]) <= 1e-12 * base + 1e-18, (name, k)
        opt.step()
        sched.step()
    from controllora_b200.trainer import Trainer
    with pytest.raises(NotImplementedError):
        tr, _ = _tiny_trainer(0)
        Trainer(tr.unet, tr.cl, cuda_graph=True, lr_scheduler="cosine", max_train_steps=10)
