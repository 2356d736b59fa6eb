"""Parity of the processors CALLED THE WAY DIFFUSERS CALLS THEM - `processor(attn, hidden_states, encoder_hidden_states, None, scale)`
on a stand-alone attention module (/root/reference/models.py:118-152, 222-287, 357-431) - against the oracle's processors on the
oracle's CrossAttention restatement: output, d hidden_states, every adapter / control parameter gradient, d control states.
Also `LoRALinearLayer.forward`.     usage: python tests/check_eager.py [case ...]     (CLB_EMU=1: host-logic mode on the CPU)"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from tests._device import DEV, sync  # noqa: E402
from tests.check_unet import rel  # noqa: E402

C, XD, HEADS, B, N, CC = 128, 64, 8, 2, 64, 256


def _pair(kind, cross, **kw):
    import torch
    from oracle import models_ref as MR
    import controllora_b200 as cb

    xd = XD if cross else None
    cls = {"plain": "LoRACrossAttnProcessor", "v1": "ControlLoRACrossAttnProcessor", "v2": "ControlLoRACrossAttnProcessorV2"}[kind]
    if kind != "plain":
        kw.setdefault("control_channels", CC)
    o = getattr(MR, cls)(C, xd, **kw)
    g = torch.Generator().manual_seed(3)
    with torch.no_grad():
        for n_, p_ in o.named_parameters():
            if n_.endswith("up.weight"):
                p_.copy_(0.05 * torch.randn(p_.shape, generator=g))
    m = getattr(cb, cls)(C, xd, **kw)
    m.load_state_dict(o.state_dict())
    return o, m.to(DEV)


def run(case):
    import torch
    from oracle import unet_ref as UR

    kind, cross, scale, stacked = {
        "plain_self": ("plain", False, 1.0, False), "plain_cross": ("plain", True, 0.7, False),
        "v1_self": ("v1", False, 1.0, False), "v1_cross_stacked": ("v1", True, 0.5, True),
        "v2_self": ("v2", False, 1.0, False), "v2_cross": ("v2", True, 1.0, False),
    }[case]
    torch.manual_seed(0)
    attn = UR.CrossAttention(C, XD if cross else None, HEADS, C // HEADS)
    with torch.no_grad():
        for p in attn.parameters():
            p.copy_((p * 2).to(torch.bfloat16).float())
    import copy

    attn_m = copy.deepcopy(attn).to(DEV)
    o, m = _pair(kind, cross)
    extras = []
    if stacked:
        o2, m2 = _pair("plain", cross, rank=4)
        o.inject_pre_lora(o2)
        m.inject_pre_lora(m2)
        extras.append((o2, m2))
    g = torch.Generator().manual_seed(5)
    hs = torch.randn(B, N, C, generator=g).to(torch.bfloat16).float()
    ehs = torch.randn(B, 77, XD, generator=g).to(torch.bfloat16).float() if cross else None
    w = torch.randn(B, N, C, generator=g)
    hs_o, hs_m = hs.clone().requires_grad_(True), hs.clone().to(DEV).requires_grad_(True)
    ctrl = None
    if kind != "plain":
        c = (0.5 * torch.randn(B, CC, 8, 8, generator=g)).to(torch.bfloat16).float()        # NCHW like ControlLoRA.forward's states
        co, cm = c.clone().requires_grad_(True), c.clone().to(DEV).requires_grad_(True)
        o.inject_control_states(co)
        m.inject_control_states(cm)
        ctrl = (co, cm)
    yo = o(attn, hs_o, ehs, None, scale)
    (yo * w).sum().backward()
    ym = m(attn_m, hs_m, None if ehs is None else ehs.to(DEV), None, scale)
    assert ym.dtype == hs_m.dtype and ym.shape == yo.shape
    (ym * w.to(DEV)).sum().backward()
    sync()
    e_out, e_hs = rel(ym, yo), rel(hs_m.grad, hs_o.grad)
    rows = []
    for o_, m_ in [(o, m)] + extras:
        for (n1, p1), (n2, p2) in zip(o_.named_parameters(), m_.named_parameters()):
            assert n1 == n2
            if p1.grad is None:
                continue
            assert p2.grad is not None, n2
            rows.append((rel(p2.grad, p1.grad), n1))
    rows.sort(reverse=True)
    e_c = rel(ctrl[1].grad, ctrl[0].grad) if ctrl else 0.0
    print(f"[eager {case}] out rel={e_out:.3e}  d hidden rel={e_hs:.3e}  worst param grad rel={rows[0][0]:.3e} ({rows[0][1]}; {len(rows)} tensors)"
          f"  d control rel={e_c:.3e}")
    ok = e_out < 1e-2 and e_hs < 2e-2 and rows[0][0] < 3e-2 and e_c < 3e-2
    print("CASE_OK" if ok else "CASE_FAIL")
    return ok


def run_lora_linear():
    import torch
    from oracle import models_ref as MR
    import controllora_b200 as cb

    torch.manual_seed(1)
    o = MR.LoRALinearLayer(128, 256, 12)
    with torch.no_grad():
        o.up.weight.copy_(0.05 * torch.randn(o.up.weight.shape))
    m = cb.LoRALinearLayer(128, 256, 12)
    m.load_state_dict(o.state_dict())
    m.to(DEV)
    x = torch.randn(3, 40, 128).to(torch.bfloat16).float()
    w = torch.randn(3, 40, 256)
    xo, xm = x.clone().requires_grad_(True), x.clone().to(DEV).requires_grad_(True)
    (o(xo) * w).sum().backward()
    ym = m(xm)
    (ym * w.to(DEV)).sum().backward()
    sync()
    errs = [rel(ym, o(xo)), rel(xm.grad, xo.grad), rel(m.down.weight.grad, o.down.weight.grad), rel(m.up.weight.grad, o.up.weight.grad)]
    print("[eager lora_linear] y / dx / d down / d up rel = " + " ".join(f"{e:.3e}" for e in errs))
    ok = max(errs) < 1e-2
    print("CASE_OK" if ok else "CASE_FAIL")
    return ok


CASES = {"eager_" + k: (lambda k=k: run(k)) for k in ("plain_self", "plain_cross", "v1_self", "v1_cross_stacked", "v2_self", "v2_cross")}
CASES["eager_lora_linear"] = run_lora_linear
CASE_NAMES = list(CASES)

if __name__ == "__main__":
    names = sys.argv[1:] or CASE_NAMES
    bad = [n for n in names if not CASES[n]()]
    print("SUMMARY", "all ok" if not bad else f"FAILED {bad}")
    sys.exit(1 if bad else 0)
