"""Full-size parity of the drop-in path on the BASELINE configs, and the precision contract.

For a named ControlLoRA config (`diffusiondb-canny` = BASELINE C2, v1 processors + pre-convs; `diffusiondb-canny-v2` =
BASELINE C4, V2 processors) on the REAL SD-1.5 shapes (block_out_channels 320/640/1280/1280, 64x64 latents, 77x768 text
states, 512x512 guide) three implementations run the same training-step front half -- hint encoder, UNet, MSE, backward --
on the same seeded inputs and weights:

  (o) the fp32 oracle on the CPU (oracle/unet_ref.py + oracle/models_ref.py)                        = the reference value
  (a) ours: `control_lora(guide)`, `unet(...).sample`, `loss.backward()` through the drop-in classes = the product path
  (b) the same oracle code in eager PyTorch on the GPU the way the reference runs it: frozen UNet in bf16, trainable
      ControlLoRA in fp32, `torch.autocast(bf16)` (train_text_to_image_control_lora.py:437-447)       = the reference's own
      precision

and the test asserts the SURVEY.md section 7 "precision contract": err(a vs o) <= MARGIN * err(b vs o) for the noise
prediction, the loss and the concatenated gradient of all trainable parameters, plus absolute caps.  north_star's
"1e-3 relative" is below the bf16 floor of ANY bf16 pipeline (arm (b) shows where that floor is on these shapes).

usage: python tests/check_fullsize.py [diffusiondb-canny|diffusiondb-canny-v2] [B]
"""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

# err(ours) <= MARGIN * err(eager bf16): both are bf16 pipelines with different rounding points, so they are compared
# directly (measured round 2: ours 0.31-0.78 of the eager-bf16 error, profiles/r02_parity_fullsize.log); the absolute caps are the tolerance contract written into the test.
MARGIN = 1.0
CAP_PRED = 1.5e-2        # relative L2 of the noise prediction vs the fp32 oracle
CAP_GRAD_ALL = 5e-2      # relative L2 of the concatenated gradient of every trainable parameter
CAP_LOSS = 2e-3


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def zero_grad_biases(cl):
    """Names of conv biases whose true gradient is exactly zero: the conv feeds a GroupNorm that has one channel per group
    (32 channels / 32 groups in the first pyramid level of models.py:690-748), which removes any per-channel constant."""
    groups = cl.config["norm_num_groups"] if isinstance(cl.config, dict) else cl.config.norm_num_groups
    names = set()
    for n, p in cl.named_parameters():
        if n.endswith("bias") and p.dim() == 1 and p.numel() == groups and ("conv1" in n or "downsamplers" in n or n == "conv_in.bias"):
            names.add(n)
    return names


def run(config_name="diffusiondb-canny-v2", B=1, log=None):
    import torch
    from oracle import models_ref as MR
    from oracle import unet_ref as UR
    import controllora_b200 as cb
    from controllora_b200.configs import NAMED
    import bench

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dev = "cuda"
    out = {"config": config_name, "batch": B}

    # ---------------- models: one set of weights for the three arms
    ounet = UR.UNet2DConditionModel()
    UR.init_synthetic_(ounet, seed=1)
    with torch.no_grad():
        for p in ounet.parameters():
            p.copy_(p.to(torch.bfloat16).float())       # weight_dtype = bf16 for the frozen network in every arm
    ounet.requires_grad_(False)
    ocl = MR.ControlLoRA.from_config(NAMED[config_name])
    MR.randomize_lora_up_(ocl, seed=3, std=0.02)
    g = torch.Generator().manual_seed(2)
    with torch.no_grad():
        for n, p in ocl.named_parameters():
            if "norm" in n:
                p.add_(0.1 * torch.randn(p.shape, generator=g))
    import copy
    bunet_cpu, bcl_cpu = copy.deepcopy(ounet), copy.deepcopy(ocl)      # arm (b)'s twins, copied before any forward caches state
    MR.wire_processors(ounet, ocl)
    x, t, e, guide, tgt = bench.synth_inputs(torch, B)
    x = x.to(torch.bfloat16).float()
    e = e.to(torch.bfloat16).float()

    def arm_oracle(unet, cl, device, autocast):
        xs, ts, es, gs, tg = (v.to(device) for v in (x, t, e, guide, tgt))
        cl.zero_grad(set_to_none=True)
        if autocast:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                cl(gs)
                pred = unet(xs.to(torch.bfloat16), ts.long(), es.to(torch.bfloat16)).sample
                loss = torch.nn.functional.mse_loss(pred.float(), tg)
        else:
            cl(gs)
            pred = unet(xs, ts.long(), es).sample
            loss = torch.nn.functional.mse_loss(pred, tg)
        loss.backward()
        grads = {n: p.grad.detach().float().cpu().clone() for n, p in cl.named_parameters() if p.grad is not None}
        return pred.detach().float().cpu(), float(loss), grads

    # ---------------- (a) ours, through the drop-in classes
    munet = cb.UNet2DConditionModel.from_state_dict({k: v.detach().clone() for k, v in ounet.state_dict().items()
                                                     if "processor" not in k}, dev)
    mcl = cb.ControlLoRA.from_config(NAMED[config_name])
    mcl.load_state_dict(ocl.state_dict())
    mcl.to(dev)
    from controllora_b200.configs import wire_processors
    wire_processors(munet, mcl)
    t0 = time.time()
    mcl(guide.to(dev))
    pm = munet(x.to(dev), t.to(dev), e.to(dev).to(torch.bfloat16)).sample
    lm = torch.nn.functional.mse_loss(pm.float(), tgt.to(dev))
    lm.backward()
    torch.cuda.synchronize()
    out["ours_s"] = time.time() - t0
    pred_a, loss_a = pm.detach().float().cpu(), float(lm)
    grads_a = {n: p.grad.detach().float().cpu().clone() for n, p in mcl.named_parameters() if p.grad is not None}
    del munet, mcl, pm, lm
    torch.cuda.empty_cache()

    # ---------------- (o) fp32 oracle on the host
    t0 = time.time()
    pred_o, loss_o, grads_o = arm_oracle(ounet, ocl, "cpu", False)
    out["oracle_cpu_s"] = time.time() - t0

    # ---------------- (b) the reference's own precision: eager PyTorch, bf16 UNet, fp32 adapters under bf16 autocast
    bunet = bunet_cpu.to(dev).to(torch.bfloat16)
    bcl = bcl_cpu.to(dev)
    MR.wire_processors(bunet, bcl)
    pred_b, loss_b, grads_b = arm_oracle(bunet, bcl, dev, True)
    del bunet, bcl
    torch.cuda.empty_cache()

    # ---------------- compare
    zero_b = zero_grad_biases(ocl)
    names = [n for n in grads_o if n not in zero_b]
    missing = [n for n in names if n not in grads_a]
    assert not missing, f"ours produced no gradient for {missing[:5]}"
    cat = lambda gr: torch.cat([gr[n].flatten() for n in names])
    go, ga, gb = cat(grads_o), cat(grads_a), cat(grads_b)
    res = {
        "pred": (rel(pred_a, pred_o), rel(pred_b, pred_o)),
        "loss": (abs(loss_a - loss_o) / abs(loss_o), abs(loss_b - loss_o) / abs(loss_o)),
        "grad_all": (rel(ga, go), rel(gb, go)),
    }
    groups = {"hint_encoder": lambda n: not n.startswith("lora_layers"),
              "lora_q_out": lambda n: n.startswith("lora_layers") and ("to_q_lora" in n or "to_out_lora" in n or "to_k_lora" in n or "to_v_lora" in n),
              "control": lambda n: n.startswith("lora_layers") and "to_control" in n}
    for gname, pred_fn in groups.items():
        sel = [n for n in names if pred_fn(n)]
        if not sel:
            continue
        c = lambda gr: torch.cat([gr[n].flatten() for n in sel])
        res["grad_" + gname] = (rel(c(grads_a), c(grads_o)), rel(c(grads_b), c(grads_o)))
    per = sorted(((rel(grads_a[n], grads_o[n]), rel(grads_b[n], grads_o[n]), float(grads_o[n].norm()), n) for n in names), reverse=True)
    # provably-zero-gradient biases: ours must stay at noise level next to the layer's weight gradient
    zero_rows = []
    for n in sorted(zero_b):
        wn = n[:-4] + "weight"
        wnorm = float(grads_o[wn].norm()) if wn in grads_o else 1.0
        zero_rows.append((n, float(grads_o[n].norm()) / wnorm, float(grads_a[n].norm()) / wnorm, float(grads_b[n].norm()) / wnorm))

    def emit(s):
        print(s, flush=True)
        if log is not None:
            log.write(s + "\n")

    emit(f"== full-size parity: {config_name}, B={B}, SD-1.5 shapes (64x64 latents, 512x512 guide); "
         f"oracle fp32 CPU {out['oracle_cpu_s']:.1f}s, ours {out['ours_s']:.2f}s (first call, incl. weight layout)")
    emit(f"   loss: oracle {loss_o:.6f}  ours {loss_a:.6f}  eager-bf16 {loss_b:.6f}")
    emit(f"   {'quantity':<18} {'ours vs fp32':>14} {'eager-bf16 vs fp32':>20}   ratio")
    for k, (ea, eb) in res.items():
        emit(f"   {k:<18} {ea:14.3e} {eb:20.3e}   {ea / max(eb, 1e-30):5.2f}")
    emit(f"   worst single tensors (ours | eager-bf16 | |g| | name), {len(names)} tensors compared:")
    for ea, eb, nrm, n in per[:8]:
        emit(f"     {ea:10.3e} {eb:10.3e} {nrm:10.3e} {n}")
    for n, zo, za, zb in zero_rows:
        emit(f"   zero-gradient bias {n}: |g|/|g_weight| oracle {zo:.1e} ours {za:.1e} eager-bf16 {zb:.1e}")
    ok = True
    for k in ("pred", "grad_all"):
        ea, eb = res[k]
        ok = ok and ea <= MARGIN * eb
    ok = ok and res["pred"][0] <= CAP_PRED and res["grad_all"][0] <= CAP_GRAD_ALL and res["loss"][0] <= CAP_LOSS
    ok = ok and all(za <= 5e-2 for _, _, za, _ in zero_rows)
    out.update({k: {"ours": v[0], "eager_bf16": v[1]} for k, v in res.items()})
    out["worst_tensor"] = {"name": per[0][3], "ours": per[0][0], "eager_bf16": per[0][1]}
    out["ok"] = bool(ok)
    emit("   JSON " + json.dumps(out))
    emit("CASE_OK" if ok else "CASE_FAIL")
    return ok


if __name__ == "__main__":
    name = sys.argv[1] if len(sys.argv) > 1 else "diffusiondb-canny-v2"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    logp = ROOT / "gpurun_out"
    logp.mkdir(exist_ok=True)
    with open(logp / f"parity_fullsize_{name}_b{B}.log", "w") as f:
        sys.exit(0 if run(name, B, f) else 1)
