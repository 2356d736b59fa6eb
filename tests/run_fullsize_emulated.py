"""One fused train step at the REAL SD-1.5 shapes (320/640/1280/1280 channels, 64x64 latents, 77x768 text states, 512x512 guide, batch
1) in host-logic mode: every C-ABI call the host program issues goes through the real ops.py wrapper and the real launcher's argument
validation (shape / alignment / shared-memory limits at C = 1280, K = 768 ...) before its torch restatement runs (tests/emu_ops.py).
usage: python tests/run_fullsize_emulated.py <config> <generic 0|1> <stacked 0|1> [...]      prints `FULLSIZE <config> ... OK`"""
import os
import sys
import time
from pathlib import Path

os.environ["CLB_EMU"] = "1"
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from tests import _device  # noqa: E402,F401  (installs the host-logic mode)


def run(name, generic, stacked):
    import torch
    import controllora_b200 as cb
    from controllora_b200.configs import NAMED, wire_processors
    from controllora_b200.trainer import Trainer

    if generic:
        os.environ["CLB_GENERIC_CHAIN"] = "1"
    else:
        os.environ.pop("CLB_GENERIC_CHAIN", None)
    t0 = time.time()
    unet = cb.UNet2DConditionModel.synthetic("cpu", None, seed=0)
    cl = cb.ControlLoRA.from_config(NAMED[name])
    g = torch.Generator().manual_seed(3)
    with torch.no_grad():
        for n, p in cl.named_parameters():
            if n.endswith("up.weight"):
                p.copy_(0.02 * torch.randn(p.shape, generator=g))
    procs = wire_processors(unet, cl)
    if stacked:
        for p in procs.values():
            p.inject_pre_lora(cb.LoRACrossAttnProcessor(p.hidden_size, p.cross_attention_dim, rank=4))
    tr = Trainer(unet, cl, lr=1e-4)
    x = torch.randn(1, 4, 64, 64, generator=g)
    e = torch.randn(1, 77, 768, generator=g).to(torch.bfloat16)
    guide = torch.rand(1, 3, 512, 512, generator=g) * 2 - 1
    loss = tr.step(x, torch.tensor([500.0]), e, guide, torch.randn(1, 4, 64, 64, generator=g))
    kinds = {}
    for lp in unet._get_runtime().layers.values():
        kinds[lp.kind] = kinds.get(lp.kind, 0) + 1
    ok = bool(torch.isfinite(loss).all()) and bool(torch.isfinite(tr.flat_p).all()) and float(tr.flat_m.abs().sum()) > 0
    print(f"FULLSIZE {name} generic={int(generic)} stacked={int(stacked)} layers={kinds} loss={float(loss):.5f} launches={tr.launches_per_step} "
          f"{time.time() - t0:.0f}s {'OK' if ok else 'FAIL'}", flush=True)
    return ok


if __name__ == "__main__":
    import torch

    torch.set_num_threads(min(8, os.cpu_count() or 1))
    a = sys.argv[1:]
    bad = 0
    for i in range(0, len(a), 3):
        bad += 0 if run(a[i], a[i + 1] == "1", a[i + 2] == "1") else 1
    sys.exit(1 if bad else 0)
