"""GPU parity of the AutoencoderKL forward paths (encode moments / sample, decode) against the fp32 CPU oracle
(oracle/vae_ref.py) on a small config and - `full` - on the SD-1.5 VAE shapes at 256x256.
usage: python tests/check_vae.py [tiny|full]"""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from tests._device import DEV, sync  # noqa: E402


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def run(which="tiny"):
    import torch
    from oracle import vae_ref as VR
    import controllora_b200 as cb

    torch.backends.cuda.matmul.allow_tf32 = False
    if which == "tiny":
        cfg = dict(block_out_channels=(32, 64, 64), layers_per_block=1)
        n, size = 2, 64
    else:
        cfg = {}
        n, size = 1, 256
    ovae = VR.AutoencoderKL(**cfg)
    VR.init_synthetic_(ovae, seed=4)
    with torch.no_grad():
        for p in ovae.parameters():
            p.copy_(p.to(torch.bfloat16).float())          # the reference runs the frozen VAE in weight_dtype = bf16
    mvae = cb.AutoencoderKL.from_state_dict({k: v.detach().clone() for k, v in ovae.state_dict().items()}, DEV, cfg)
    g = torch.Generator().manual_seed(1)
    x = (torch.rand(n, 3, size, size, generator=g) * 2 - 1).to(torch.bfloat16).float()
    lat_hw = size // (2 ** (len(ovae.config["block_out_channels"]) - 1))
    eps = torch.randn(n, 4, lat_hw, lat_hw, generator=g)
    t0 = time.time()
    with torch.no_grad():
        mean_o, logvar_o = ovae.encode_moments(x)
        z_o = (mean_o + torch.exp(0.5 * logvar_o) * eps) * ovae.config["scaling_factor"]
        img_o = ovae.decode(z_o)
    t_or = time.time() - t0
    dist = mvae.encode(x.to(DEV)).latent_dist
    z_m = (dist.mean + dist.std * eps.to(DEV)) * mvae.config.scaling_factor
    img_m = mvae.decode(z_o.to(DEV) / mvae.config.scaling_factor).sample       # decode the ORACLE's latents: isolates the decoder
    sync()
    e = {"mean": rel(dist.mean, mean_o), "logvar": rel(dist.logvar, logvar_o), "latents": rel(z_m, z_o), "decode": rel(img_m, img_o)}
    print(f"[vae {which}] n={n} {size}x{size}: oracle {t_or:.1f}s  " + "  ".join(f"{k} rel={v:.3e}" for k, v in e.items()))
    ok = e["mean"] < 2e-2 and e["logvar"] < 2e-2 and e["latents"] < 2e-2 and e["decode"] < 2e-2
    print("CASE_OK" if ok else "CASE_FAIL")
    return ok


if __name__ == "__main__":
    sys.exit(0 if run(sys.argv[1] if len(sys.argv) > 1 else "tiny") else 1)
