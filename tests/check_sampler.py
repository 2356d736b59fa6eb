"""Denoise loops and the device-side step glue against the oracle (tiny SD-style UNet):
  ddim / dpmpp    `ddim_sample` / `dpmpp_sample` (CFG, UNet batch 2B, control injected once; train_...:824-843, mix_lora_and_control_lora.py:
                  153-164) vs the oracle UNet driven by the restated schedulers (oracle/sampler_ref.py)
  step_glue       `Trainer.step_from_latents` (train_...:757-796): the draw of step k equals the Philox restatement for counter k, the
                  loss equals the oracle's on the same noisy latents / target, fresh draws every step
usage: python tests/check_sampler.py [ddim|dpmpp|step_glue]        (CLB_EMU=1: host-logic mode on the CPU)"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from tests._device import DEV, sync  # noqa: E402
from tests import check_unet  # noqa: E402


def loop_case(which):
    import torch
    from controllora_b200.sampler import ddim_sample, dpmpp_sample
    from oracle import sampler_ref as SR

    ounet, munet, ocl, mcl = check_unet.build_pair("v1" if which == "ddim" else "v2")
    g = torch.Generator().manual_seed(9 if which == "ddim" else 11)
    B, HW = 2, 16
    guide = (torch.rand(B, 3, HW * 8, HW * 8, generator=g) * 2 - 1).to(torch.bfloat16).float()
    cond = torch.randn(B, 77, 64, generator=g).to(torch.bfloat16).float()
    unc = torch.randn(B, 77, 64, generator=g).to(torch.bfloat16).float()
    lat0 = torch.randn(B, 4, HW, HW, generator=g)
    steps = 3 if which == "ddim" else 4
    with torch.no_grad():
        ocl(torch.cat([guide, guide], 0))
        x = lat0.clone()
        if which == "ddim":
            for t in SR.timesteps(steps):
                eps = ounet(torch.cat([x, x], 0), torch.full((2 * B,), int(t)), torch.cat([unc, cond], 0)).sample
                x = SR.cfg_ddim_step(eps[:B], eps[B:], x, t, steps, 7.5)
        else:
            sched = SR.DPMSolverPP2M(steps)
            for t in sched.timesteps:
                eps = ounet(torch.cat([x, x], 0), torch.full((2 * B,), int(t)), torch.cat([unc, cond], 0)).sample
                x = sched.step(SR.cfg_combine(eps[:B], eps[B:], 7.5), t, x)
    fn = ddim_sample if which == "ddim" else dpmpp_sample
    out = fn(munet, mcl, guide.to(DEV), cond.to(DEV).to(torch.bfloat16), unc.to(DEV).to(torch.bfloat16), num_inference_steps=steps,
             guidance_scale=7.5, latents=lat0.to(DEV))
    sync()
    err = float((out.cpu() - x).norm() / x.norm())
    print(f"[sampler {which}] {steps}-step latent rel err {err:.3e}")
    ok = err < (6e-2 if which == "ddim" else 8e-2)
    print("CASE_OK" if ok else "CASE_FAIL")
    return ok


def step_glue_case():
    import numpy as np
    import torch
    from controllora_b200.trainer import Trainer
    from oracle import sampler_ref as SR

    ounet, munet, ocl, mcl = check_unet.build_pair("v2")
    B, HW = 2, 16
    g = torch.Generator().manual_seed(3)
    lat = torch.randn(B, 4, HW, HW, generator=g)
    ehs = torch.randn(B, 77, 64, generator=g).to(torch.bfloat16).float()
    guide = (torch.rand(B, 3, HW * 8, HW * 8, generator=g) * 2 - 1).to(torch.bfloat16).float()
    tr = Trainer(munet, mcl, lr=0.0, noise_seed=1234)      # lr 0: the step runs, the adapters stay where the oracle's are
    ok, prev = True, None
    for k in range(3):
        loss = tr.step_from_latents(lat.to(DEV), ehs.to(DEV).to(torch.bfloat16), guide.to(DEV))
        sync()
        noisy, target, ts = [t.detach().cpu() for t in tr.last_noise_draw]
        n_ref, t_ref = SR.device_noise(tr._rank_seed, k, B, 4 * HW * HW)
        n_ref, t_t = torch.from_numpy(n_ref).view(B, 4, HW, HW), torch.from_numpy(t_ref)
        same_draw = np.array_equal(ts.numpy().astype(np.int64), t_ref) and torch.allclose(noisy, SR.add_noise(lat, n_ref, t_t), atol=2e-5, rtol=1e-5) \
            and torch.allclose(target, n_ref, atol=2e-5, rtol=1e-5)
        fresh = prev is None or not torch.equal(prev, noisy)
        prev = noisy
        with torch.no_grad():
            ocl(guide)
            lo = torch.nn.functional.mse_loss(ounet(noisy, ts.long(), ehs).sample, target)
        e = abs(float(loss) - float(lo)) / abs(float(lo))
        print(f"[step_glue] step {k}: timesteps {ts.tolist()} draw==restatement {same_draw} fresh {fresh} loss ours={float(loss):.6f} oracle={float(lo):.6f} rel={e:.2e}")
        ok = ok and same_draw and fresh and e < 2e-3 and int(tr.rng_counter) == k + 1
    print("CASE_OK" if ok else "CASE_FAIL")
    return ok


def pixels_case():
    """Trainer.step_from_pixels: the whole loop body train_text_to_image_control_lora.py:751-796 from pixels / token ids / guide - VAE
    encode + latent sample, CLIP text tower, device noise + timesteps + add_noise, control injection, UNet, MSE - against the oracle
    chain (vae_ref -> clip_ref -> sampler_ref.add_noise -> models_ref + unet_ref) on the same weights, the same latent-sampling noise
    and the noise / timesteps the device drew."""
    import torch
    import controllora_b200 as cb
    from controllora_b200.trainer import Trainer
    from oracle import clip_ref as CR
    from oracle import sampler_ref as SR
    from oracle import vae_ref as VR

    ounet, munet, ocl, mcl = check_unet.build_pair("v2")
    vcfg = dict(block_out_channels=(32, 64, 64), layers_per_block=1)             # 3 levels: 64x64 pixels -> 16x16 latents
    ovae = VR.AutoencoderKL(**vcfg)
    VR.init_synthetic_(ovae, seed=4)
    mvae = cb.AutoencoderKL.from_state_dict({k: v.detach().clone() for k, v in ovae.state_dict().items()}, DEV, vcfg)
    ccfg = dict(CR.SD15_TEXT_CONFIG)
    ccfg.update(num_hidden_layers=2, vocab_size=1000, hidden_size=64, intermediate_size=256, num_attention_heads=4)
    csd = CR.synthetic_state_dict(ccfg, seed=0)
    mclip = cb.CLIPTextModel.from_state_dict(csd, DEV, ccfg)
    B = 2
    g = torch.Generator().manual_seed(13)
    pix = (torch.rand(B, 3, 64, 64, generator=g) * 2 - 1)
    ids = torch.randint(0, ccfg["vocab_size"], (B, 77), generator=g)
    guide = (torch.rand(B, 3, 128, 128, generator=g) * 2 - 1).to(torch.bfloat16).float()
    eps = torch.randn(B, 4, 16, 16, generator=g)
    tr = Trainer(munet, mcl, lr=0.0, noise_seed=99)
    loss = tr.step_from_pixels(mvae, mclip, pix.to(DEV), ids.to(DEV), guide.to(DEV), latent_noise=eps)
    sync()
    noisy_m, target, ts = [t.detach().cpu() for t in tr.last_noise_draw]
    with torch.no_grad():
        lat_o = ovae.encode_sample(pix, eps)
        ehs_o = CR.clip_text_forward(csd, ids, ccfg)
        noisy_o = SR.add_noise(lat_o, target, ts.long())
        ocl(guide)
        lo = torch.nn.functional.mse_loss(ounet(noisy_o, ts.long(), ehs_o).sample, target)
    e_lat = float((tr.last_latents.cpu() - lat_o).norm() / lat_o.norm())
    e_noisy = float((noisy_m - noisy_o).norm() / noisy_o.norm())
    e_loss = abs(float(loss) - float(lo)) / abs(float(lo))
    print(f"[pixels] latents rel={e_lat:.3e} noisy latents rel={e_noisy:.3e} loss ours={float(loss):.6f} oracle={float(lo):.6f} rel={e_loss:.2e} timesteps {ts.tolist()}")
    ok = e_lat < 2e-2 and e_noisy < 2e-2 and e_loss < 1e-2
    print("CASE_OK" if ok else "CASE_FAIL")
    return ok


def generate_case():
    """sampler.generate: prompt ids -> CLIP text states -> 3 CFG + DDIM steps -> VAE decode -> [0, 1] images, against the oracle chain
    (clip_ref -> models_ref + unet_ref driven by sampler_ref -> vae_ref.decode)."""
    import torch
    import controllora_b200 as cb
    from controllora_b200.sampler import generate
    from oracle import clip_ref as CR
    from oracle import sampler_ref as SR
    from oracle import vae_ref as VR

    ounet, munet, ocl, mcl = check_unet.build_pair("v1")
    vcfg = dict(block_out_channels=(32, 64, 64, 64), layers_per_block=1)          # 4 levels: 16x16 latents -> 128x128 pixels
    ovae = VR.AutoencoderKL(**vcfg)
    VR.init_synthetic_(ovae, seed=4)
    mvae = cb.AutoencoderKL.from_state_dict({k: v.detach().clone() for k, v in ovae.state_dict().items()}, DEV, vcfg)
    ccfg = dict(CR.SD15_TEXT_CONFIG)
    ccfg.update(num_hidden_layers=2, vocab_size=1000, hidden_size=64, intermediate_size=256, num_attention_heads=4)
    csd = CR.synthetic_state_dict(ccfg, seed=0)
    mclip = cb.CLIPTextModel.from_state_dict(csd, DEV, ccfg)
    B, HW, steps = 2, 16, 3
    g = torch.Generator().manual_seed(17)
    guide = (torch.rand(B, 3, HW * 8, HW * 8, generator=g) * 2 - 1).to(torch.bfloat16).float()
    ids = torch.randint(0, ccfg["vocab_size"], (B, 77), generator=g)
    neg = torch.randint(0, ccfg["vocab_size"], (B, 77), generator=g)
    lat0 = torch.randn(B, 4, HW, HW, generator=g)
    with torch.no_grad():
        cond, unc = CR.clip_text_forward(csd, ids, ccfg), CR.clip_text_forward(csd, neg, ccfg)
        ocl(torch.cat([guide, guide], 0))
        x = lat0.clone()
        for t in SR.timesteps(steps):
            eps = ounet(torch.cat([x, x], 0), torch.full((2 * B,), int(t)), torch.cat([unc, cond], 0)).sample
            x = SR.cfg_ddim_step(eps[:B], eps[B:], x, t, steps, 7.5)
        want = (ovae.decode(x) / 2 + 0.5).clamp(0, 1)
    got = generate(munet, mcl, mvae, mclip, guide.to(DEV), ids.to(DEV), neg.to(DEV), num_inference_steps=steps, guidance_scale=7.5,
                   scheduler="ddim", latents=lat0.to(DEV))
    sync()
    err = float((got.cpu() - want).norm() / want.norm())
    rng_ok = float(got.min()) >= 0.0 and float(got.max()) <= 1.0 and tuple(got.shape) == (B, 3, HW * 8, HW * 8)
    print(f"[generate] {steps}-step DDIM + decode: image rel err {err:.3e}, range/shape ok {rng_ok}")
    ok = err < 5e-2 and rng_ok
    print("CASE_OK" if ok else "CASE_FAIL")
    return ok


def graphed_program_case():
    """GraphedSampler's per-step program (timestep-invariant products prepared once, text k / v cache, timestep and solver
    coefficients read from device tables through a device step counter) run WITHOUT capturing it, against the plain Python loops: the
    same latents, bit for bit on the CPU (host-logic mode), to rounding on the GPU.  Capture + replay itself is a GPU test
    (tests/test_gpu_parity.py::test_graphed_sampler_matches_eager_loop)."""
    import torch
    from controllora_b200.sampler import GraphedSampler, ddim_sample, dpmpp_sample

    ok = True
    for sched, variant, fn in (("ddim", "v1", ddim_sample), ("dpmpp", "v2", dpmpp_sample)):
        _, munet, _, mcl = check_unet.build_pair(variant)
        g = torch.Generator().manual_seed(9)
        B, HW = 2, 16
        guide = (torch.rand(B, 3, HW * 8, HW * 8, generator=g) * 2 - 1).to(torch.bfloat16).float().to(DEV)
        cond = torch.randn(B, 77, 64, generator=g).to(torch.bfloat16).to(DEV)
        unc = torch.randn(B, 77, 64, generator=g).to(torch.bfloat16).to(DEV)
        lat0 = torch.randn(B, 4, HW, HW, generator=g).to(DEV)
        gs = GraphedSampler(munet, mcl, B, HW * 8, HW * 8, scheduler=sched, num_inference_steps=4, guidance_scale=7.5)
        a = gs(guide, cond, unc, latents=lat0, use_graph=False).clone()
        b = fn(munet, mcl, guide, cond, unc, num_inference_steps=4, guidance_scale=7.5, latents=lat0)
        sync()
        err = float((a - b).norm() / b.norm())
        print(f"[graphed program {sched}] prepared-context step program vs plain loop: rel {err:.3e}")
        ok = ok and err <= (0.0 if DEV == "cpu" else 2e-2)
    print("CASE_OK" if ok else "CASE_FAIL")
    return ok


CASES = {"graphed_program": graphed_program_case, "generate": generate_case, "step_from_pixels": pixels_case, "sampler_ddim": lambda: loop_case("ddim"), "sampler_dpmpp": lambda: loop_case("dpmpp"), "step_glue": step_glue_case}

if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    bad = [n for n in names if not CASES[n if n in CASES else "sampler_" + n]()]
    print("SUMMARY", "all ok" if not bad else f"FAILED {bad}")
    sys.exit(1 if bad else 0)
