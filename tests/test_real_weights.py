"""Published weights.  Neither an SD-1.5 checkpoint nor a `HighCWu/ControlLoRA` snapshot exists on the build or the GPU machines (no
network), so every other test runs on seeded synthetic weights of the exact shapes / key names.  These tests run when the files are
provided:
    CLB_SD15_DIR=/path/to/stable-diffusion-v1-5          (diffusers layout: unet/, vae/, text_encoder/)
    CLB_CONTROLLORA_DIR=/path/to/HighCWu-ControlLoRA      (subfolders sd-*-model-control-lora with config.json + weights)
and otherwise SKIP with that reason - they never pass vacuously."""
import os
from pathlib import Path

import pytest
import torch

SD15 = os.environ.get("CLB_SD15_DIR")
CLORA = os.environ.get("CLB_CONTROLLORA_DIR")
need_sd15 = pytest.mark.skipif(not (SD15 and Path(SD15, "unet").is_dir()), reason="no SD-1.5 checkpoint on this machine (set CLB_SD15_DIR)")
need_clora = pytest.mark.skipif(not (CLORA and Path(CLORA).is_dir()), reason="no HighCWu/ControlLoRA snapshot on this machine (set CLB_CONTROLLORA_DIR)")


@need_sd15
def test_published_sd15_unet_vae_text_encoder_load_with_exact_parameter_counts():
    """train_text_to_image_control_lora.py:401-409: the three `from_pretrained(..., subfolder=...)` calls on the real files."""
    import controllora_b200 as cb
    from safetensors.torch import load_file

    def n_params(sub):
        d = Path(SD15, sub)
        f = next((p for p in (d / "diffusion_pytorch_model.safetensors", d / "model.safetensors") if p.exists()), None)
        sd = load_file(str(f)) if f else torch.load(next(d.glob("*.bin")), map_location="cpu")
        return sum(v.numel() for k, v in sd.items() if "position_ids" not in k)

    assert n_params("unet") == 859_520_964 and n_params("vae") == 83_653_863 and n_params("text_encoder") == 123_060_480
    unet = cb.UNet2DConditionModel.from_pretrained(SD15, subfolder="unet", device="cpu")      # layout conversion only: no kernel runs
    assert len(unet.attn_processors) == 32 and unet.config.cross_attention_dim == 768


@need_clora
def test_published_controllora_snapshots_load_into_the_drop_in_class():
    """apps/gradio_canny2image.py:38, mix_lora_and_control_lora.py:84-88: `ControlLoRA.from_pretrained("HighCWu/ControlLoRA", subfolder=...)`."""
    import controllora_b200 as cb

    subs = sorted(p.name for p in Path(CLORA).iterdir() if (p / "config.json").exists())
    assert subs, "no <subfolder>/config.json under CLB_CONTROLLORA_DIR"
    for sub in subs:
        m = cb.ControlLoRA.from_pretrained(CLORA, subfolder=sub)          # strict load_state_dict: every published key must exist
        assert sum(p.numel() for p in m.parameters()) in (6_047_040, 6_048_576, 5_000_704, 19_810_304), sub


@pytest.mark.gpu
@need_sd15
@need_clora
def test_published_weights_run_one_denoise_step_on_the_cuda_path():
    import controllora_b200 as cb
    from controllora_b200.configs import wire_processors

    unet = cb.UNet2DConditionModel.from_pretrained(SD15, subfolder="unet", device="cuda")
    sub = sorted(p.name for p in Path(CLORA).iterdir() if (p / "config.json").exists())[0]
    cl = cb.ControlLoRA.from_pretrained(CLORA, subfolder=sub).cuda()
    wire_processors(unet, cl)
    g = torch.Generator().manual_seed(0)
    with torch.no_grad():
        cl((torch.rand(2, 3, 512, 512, generator=g) * 2 - 1).cuda())
        pred = unet(torch.randn(2, 4, 64, 64, generator=g).cuda(), torch.tensor([500, 500]).cuda(),
                    torch.randn(2, 77, 768, generator=g).cuda().to(torch.bfloat16)).sample
    assert pred.shape == (2, 4, 64, 64) and bool(torch.isfinite(pred).all()) and 0.1 < float(pred.std()) < 10
