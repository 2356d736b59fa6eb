"""Parity of the GENERAL adapter chains (controllora_b200/lora_generic.py) against the fp32 oracle: the wirings of
/root/reference/models.py:118-431 that the fused one-launch path cannot express - `post_add` inside stacked chains, ranks above 8,
control ranks above 4, a second ControlLoRA stacked as pre / post LoRA, `concat_hidden` control combined with stacking.
Tiny SD-style UNet (tests/check_unet.py), noise prediction + every adapter / control gradient.

usage: python tests/check_variants.py [case ...]       (CLB_EMU=1: host-logic mode on the CPU, tests/_device.py)"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from tests._device import DEV, sync  # noqa: E402
from tests.check_unet import TINY, TINY_LORA, rel  # noqa: E402

# name -> (ControlLoRA kwargs, stacking recipe)
#   recipe entries: ("plain", where, rank, post_add)  a LoRACrossAttnProcessor injected as pre / post LoRA of every processor
#                   ("control", where, kwargs)        the processors of a SECOND ControlLoRA (own control states) injected likewise
SPECS = {
    # post_add inside stacked chains (models.py:232-243: adapter i reads the base projection plus all earlier deltas)
    "v1_pre_post_add": (dict(), [("plain", "pre", 4, True)]),
    "v1_post_add_main_stacked": (dict(lora_post_add=True), [("plain", "pre", 4, False), ("plain", "post", 4, True)]),
    "v2_post_post_add": (dict(lora_control_version=2, lora_pre_conv_skipped=True), [("plain", "post", 4, True)]),
    # ranks: one adapter above 8, a chain summing above 8, control rank above 4
    "v1_rank16": (dict(lora_rank=16, lora_control_rank=4), []),
    "v1_rank8_stacked8": (dict(lora_rank=8), [("plain", "pre", 8, False)]),
    "v1_control_rank12": (dict(lora_control_rank=12), []),
    "v2_control_rank8": (dict(lora_control_version=2, lora_pre_conv_skipped=True, lora_control_rank=8), []),
    "post_add_rank8": (dict(lora_post_add=True, lora_rank=8, lora_control_rank=4), []),
    # a second ControlLoRA stacked on the first (multi-control): v1 on v1 (models.py:234-236), V2 on V2 (models.py:366-372, 412-418)
    "v1_on_v1": (dict(), [("control", "pre", dict())]),
    "v2_on_v2": (dict(lora_control_version=2, lora_pre_conv_skipped=True),
                 [("control", "post", dict(lora_control_version=2, lora_pre_conv_skipped=True))]),
    # concat_hidden control (dense MLP, control rank 32) combined with stacking / post_add; and the small-rank (hi/lo block) form
    "v1_concat_stacked": (dict(lora_concat_hidden=True, lora_control_rank=32, lora_pre_conv_skipped=True, lora_control_self_add=False),
                          [("plain", "pre", 4, False)]),
    "v1_concat_post_add": (dict(lora_concat_hidden=True, lora_control_rank=32, lora_pre_conv_skipped=True, lora_control_self_add=False,
                                lora_post_add=True), []),
    "v1_concat_rank8": (dict(lora_concat_hidden=True, lora_control_rank=8, lora_rank=8, lora_pre_conv_skipped=True,
                             lora_control_self_add=False), []),
}


def _assign(unet, control_lora):
    """name -> processor, the pop order of train_text_to_image_control_lora.py:469-487 (oracle.models_ref.wire_processors) without
    installing anything on the UNet."""
    n = len(unet.config.block_out_channels)
    pools = [list(l) for l in control_lora.lora_layers]
    out = {}
    for name in unet.attn_processors.keys():
        if name.startswith("mid_block"):
            cid = n - 1
        elif name.startswith("up_blocks"):
            cid = n - 1 - int(name[len("up_blocks.")])
        else:
            cid = int(name[len("down_blocks.")])
        if pools[cid]:
            out[name] = pools[cid].pop(0)
    return out


def _control_channels(kw):
    ch = TINY["block_out_channels"]
    return [256] * 4 if (kw.get("lora_control_version") == 2 or kw.get("lora_concat_hidden")) else list(ch)


def run(case, B=2, HW=16):
    import torch
    from oracle import models_ref as MR
    from oracle import unet_ref as UR
    import controllora_b200 as cb

    torch.backends.cuda.matmul.allow_tf32 = False
    kw_main, recipe = SPECS[case]
    torch.manual_seed(0)
    ounet = UR.UNet2DConditionModel(**TINY)
    UR.init_synthetic_(ounet, seed=1)
    with torch.no_grad():
        for p in ounet.parameters():
            p.copy_(p.to(torch.bfloat16).float())
    munet = cb.UNet2DConditionModel.from_state_dict({k: v.detach().clone() for k, v in ounet.state_dict().items()}, DEV, TINY)

    def make_cl(kw, seed):
        k = dict(TINY_LORA)
        k.update(kw)
        o = MR.ControlLoRA(**k)
        MR.randomize_lora_up_(o, seed=seed, std=0.05)
        m = cb.ControlLoRA(**k)
        m.load_state_dict(o.state_dict())
        return o, m.to(DEV), k

    ocl, mcl, kmain = make_cl(kw_main, 3)
    oprocs = MR.wire_processors(ounet, ocl)
    mprocs = MR.wire_processors(munet, mcl)
    g = torch.Generator().manual_seed(5)
    controls = []          # (oracle tensor, ours tensor)

    def inject(ocl_, mcl_, kw):
        cc = _control_channels(kw)
        for lvl in range(4):
            s = HW >> lvl
            c = (0.5 * torch.randn(B, cc[lvl], s, s, generator=g)).to(torch.bfloat16).float()
            co, cm = c.clone().requires_grad_(True), c.clone().to(DEV).requires_grad_(True)
            controls.append((co, cm))
            for p in ocl_.lora_layers[lvl]:
                p.inject_control_states(co)
            for p in mcl_.lora_layers[lvl]:
                p.inject_control_states(cm)

    inject(ocl, mcl, kmain)
    extras = []            # (oracle module, ours module) pairs whose gradients are compared besides the main processors'
    gp = torch.Generator().manual_seed(11)
    for i, item in enumerate(recipe):
        if item[0] == "plain":
            _, where, rank, post_add = item
            for name in oprocs:
                C, xd = oprocs[name].hidden_size, oprocs[name].cross_attention_dim
                op = MR.LoRACrossAttnProcessor(C, xd, rank=rank, post_add=post_add)
                with torch.no_grad():
                    for n_, p_ in op.named_parameters():
                        if n_.endswith("up.weight"):
                            p_.copy_(0.05 * torch.randn(p_.shape, generator=gp))
                mp = cb.LoRACrossAttnProcessor(C, xd, rank=rank, post_add=post_add).to(DEV)
                mp.load_state_dict(op.state_dict())
                getattr(oprocs[name], f"inject_{where}_lora")(op)
                getattr(mprocs[name], f"inject_{where}_lora")(mp)
                extras.append((op, mp))
        else:
            _, where, kw2 = item
            ocl2, mcl2, k2 = make_cl(kw2, 23 + i)
            inject(ocl2, mcl2, k2)
            o2, m2 = _assign(ounet, ocl2), _assign(munet, mcl2)
            for name in oprocs:
                getattr(oprocs[name], f"inject_{where}_lora")(o2[name])
                getattr(mprocs[name], f"inject_{where}_lora")(m2[name])
                extras.append((o2[name], m2[name]))
    sample = torch.randn(B, 4, HW, HW, generator=g).to(torch.bfloat16).float()
    t = torch.tensor([17, 801])
    ehs = torch.randn(B, 77, TINY["cross_attention_dim"], generator=g).to(torch.bfloat16).float()
    target = torch.randn(B, 4, HW, HW, generator=g)
    po = ounet(sample, t, ehs).sample
    lo = torch.nn.functional.mse_loss(po, target)
    lo.backward()
    pm = munet(sample.to(DEV), t.to(DEV), ehs.to(DEV).to(torch.bfloat16)).sample
    lm = torch.nn.functional.mse_loss(pm, target.to(DEV))
    lm.backward()
    sync()
    rt = munet._get_runtime()
    kinds = {}
    for lp in rt.layers.values():
        kinds[lp.kind] = kinds.get(lp.kind, 0) + 1
    e_pred = rel(pm, po)
    print(f"[{case}] layers {kinds}; noise-pred rel={e_pred:.3e}  loss oracle={float(lo):.6f} ours={float(lm):.6f}")
    rows = []
    pairs = [(ounet.attn_processors[n], munet.attn_processors[n], n) for n in ounet.attn_processors] + [(o, m, "stacked") for o, m in extras]
    go, gm = [], []
    for o_, m_, tag in pairs:
        for (n1, p1), (n2, p2) in zip(o_.named_parameters(), m_.named_parameters()):
            assert n1 == n2
            if p1.grad is None:
                continue
            assert p2.grad is not None, (tag, n2)
            rows.append((rel(p2.grad, p1.grad), tag, n1, float(p1.grad.norm())))
            go.append(p1.grad.flatten())
            gm.append(p2.grad.flatten().cpu())
    rows.sort(reverse=True)
    for e, tag, n1, nrm in rows[:5]:
        print(f"    grad rel={e:.3e} |g|={nrm:.3e} {tag} {n1}")
    worst = rows[0][0]
    e_all = rel(torch.cat(gm), torch.cat(go))
    e_ctrl = max(rel(cm.grad, co.grad) for co, cm in controls)
    print(f"[{case}] all adapter grads (concatenated, {len(rows)} tensors) rel={e_all:.3e}  worst tensor rel={worst:.3e}  worst d control rel={e_ctrl:.3e}")
    ok = kinds.get("generic", 0) > 0 and e_pred < 2e-2 and e_all < 5e-2 and worst < 1e-1 and e_ctrl < 8e-2
    print("CASE_OK" if ok else "CASE_FAIL")
    return ok


def run_forced(variant):
    """A standard wiring (tests/check_unet.py) with EVERY layer forced through the general chain path: the same oracle numbers
    must come out of both implementations."""
    import os

    from tests import check_unet

    os.environ["CLB_GENERIC_CHAIN"] = "1"
    try:
        return check_unet.run(variant)
    finally:
        os.environ.pop("CLB_GENERIC_CHAIN", None)


FORCED = ["plain", "v1", "v2", "v1_stacked@0.5", "v1_post_add", "v1_concat"]
CASES = {"variant_" + k: (lambda k=k: run(k)) for k in SPECS}
CASES.update({"generic_" + v: (lambda v=v: run_forced(v)) for v in FORCED})
CASE_NAMES = list(CASES)


SCALES = (1.0, 0.6)


def run_two_forwards(variant="v1_stacked"):
    """Two UNet forwards (different inputs, the second one with a different `scale`) BEFORE one backward - gradient accumulation /
    several UNet calls per loss (train_text_to_image_control_lora.py:751 `accelerator.accumulate`): the per-forward LoRA products must
    live with their forward, not on the shared runtime."""
    import torch
    from tests import check_unet

    ounet, munet, ocl, mcl = check_unet.build_pair(variant)
    g = torch.Generator().manual_seed(21)
    B, HW = 2, 16
    cc = [256] * 4 if variant in ("v2", "v1_concat") else list(check_unet.TINY["block_out_channels"])
    ctrl = []
    for lvl in range(4):
        c = (0.5 * torch.randn(B, cc[lvl], HW >> lvl, HW >> lvl, generator=g)).to(torch.bfloat16).float()
        co, cm = c.clone().requires_grad_(True), c.clone().to(DEV).requires_grad_(True)
        ctrl.append((co, cm))
        for p in ocl.lora_layers[lvl]:
            p.inject_control_states(co)
        for p in mcl.lora_layers[lvl]:
            p.inject_control_states(cm)
    lo = lm = 0.0
    preds = []
    for k, scale in enumerate(SCALES):
        x = torch.randn(B, 4, HW, HW, generator=g).to(torch.bfloat16).float()
        t = torch.randint(0, 1000, (B,), generator=g)
        e = torch.randn(B, 77, 64, generator=g).to(torch.bfloat16).float()
        tgt = torch.randn(B, 4, HW, HW, generator=g)
        cak = {"scale": scale}
        po = ounet(x, t, e, cross_attention_kwargs=cak).sample
        pm = munet(x.to(DEV), t.to(DEV), e.to(DEV).to(torch.bfloat16), cross_attention_kwargs=cak).sample
        lo = lo + torch.nn.functional.mse_loss(po, tgt)
        lm = lm + torch.nn.functional.mse_loss(pm, tgt.to(DEV))
        preds.append(rel(pm, po))
    lo.backward()
    lm.backward()
    sync()
    go, gm = [], []
    for n in ounet.attn_processors:
        for o_, m_ in [(ounet.attn_processors[n], munet.attn_processors[n])]:
            for (n1, p1), (n2, p2) in zip(o_.named_parameters(), m_.named_parameters()):
                if p1.grad is not None:
                    go.append(p1.grad.flatten())
                    gm.append(p2.grad.flatten().cpu())
    e_all = rel(torch.cat(gm), torch.cat(go))
    e_ctrl = max(rel(cm.grad, co.grad) for co, cm in ctrl)
    print(f"[two forwards {variant}] preds rel={preds[0]:.3e} / {preds[1]:.3e}; all adapter grads rel={e_all:.3e}; worst d control rel={e_ctrl:.3e}")
    ok = max(preds) < 2e-2 and e_all < 5e-2 and e_ctrl < 8e-2
    print("CASE_OK" if ok else "CASE_FAIL")
    return ok


CASES["two_forwards_v1_stacked"] = lambda: run_two_forwards("v1_stacked")
CASES["two_forwards_v2"] = lambda: run_two_forwards("v2")
CASE_NAMES = list(CASES)


def run_cfg_broadcast(variant):
    """Classifier-free guidance the way the reference's pipelines run it: `control_lora(guide)` on batch 1, UNet batch 2
    (apps/gradio_*2image.py:75-89, mix_lora_and_control_lora.py:163-164) - the injected control states are broadcast / repeated
    along the batch (models.py:209-212, 236-238, 344-347)."""
    import torch
    from tests import check_unet

    ounet, munet, ocl, mcl = check_unet.build_pair(variant)
    g = torch.Generator().manual_seed(31)
    HW = 16
    cc = [256] * 4 if variant in ("v2", "v1_concat") else list(check_unet.TINY["block_out_channels"])
    for lvl in range(4):
        c = (0.5 * torch.randn(1, cc[lvl], HW >> lvl, HW >> lvl, generator=g)).to(torch.bfloat16).float()
        for p in ocl.lora_layers[lvl]:
            p.inject_control_states(c.clone())
        for p in mcl.lora_layers[lvl]:
            p.inject_control_states(c.clone().to(DEV))
    x = torch.randn(1, 4, HW, HW, generator=g).to(torch.bfloat16).float().repeat(2, 1, 1, 1)
    t = torch.tensor([400, 400])
    e = torch.randn(2, 77, 64, generator=g).to(torch.bfloat16).float()
    with torch.no_grad():
        po = ounet(x, t, e).sample
        pm = munet(x.to(DEV), t.to(DEV), e.to(DEV).to(torch.bfloat16)).sample
    sync()
    err = rel(pm, po)
    print(f"[cfg broadcast {variant}] control batch 1, UNet batch 2: noise-pred rel={err:.3e}")
    ok = err < 2e-2
    print("CASE_OK" if ok else "CASE_FAIL")
    return ok


CASES["cfg_broadcast_v1"] = lambda: run_cfg_broadcast("v1")
CASES["cfg_broadcast_v2"] = lambda: run_cfg_broadcast("v2")
CASE_NAMES = list(CASES)


if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    bad = [n for n in names if not CASES[n if n in CASES else "variant_" + n]()]
    print("SUMMARY", "all ok" if not bad else f"FAILED {bad}")
    sys.exit(1 if bad else 0)
