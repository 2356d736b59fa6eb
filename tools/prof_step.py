"""Full-size (SD-1.5 dims) UNet train-step timing harness: python tools/prof_step.py [--batch 8] [--variant v1] [--iters 3]

Prints ms per UNet fwd+bwd (CUDA events) and, with --breakdown, per-op-family GPU time measured by bracketing the
ops.* wrappers with CUDA events (serialising; for a quick where-does-the-time-go picture)."""
import argparse
import sys
import time
from collections import defaultdict
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import torch

import controllora_b200 as cb
from controllora_b200 import ops
from controllora_b200.engine import Tape, Var


def build(variant, B, dev="cuda"):
    t0 = time.time()
    unet = cb.UNet2DConditionModel.synthetic(dev)
    kw = {}
    if variant == "v2":
        kw = dict(lora_control_version=2, lora_pre_conv_skipped=True, lora_key_states_skipped=True, lora_value_states_skipped=True)
    cl = cb.ControlLoRA(**kw).to(dev)
    g = torch.Generator().manual_seed(3)
    with torch.no_grad():
        for n, p in cl.named_parameters():
            if n.endswith("up.weight"):
                p.copy_((0.02 * torch.randn(p.shape, generator=g)).to(dev))
    # wiring as train_text_to_image_control_lora.py:469-487
    pools = [list(l) for l in cl.lora_layers]
    procs = {}
    for name in unet.attn_processors.keys():
        if name.startswith("mid_block"):
            cid = 3
        elif name.startswith("up_blocks"):
            cid = 3 - int(name[len("up_blocks.")])
        else:
            cid = int(name[len("down_blocks.")])
        procs[name] = pools[cid].pop(0)
    unet.set_attn_processor(procs)
    ch = [256] * 4 if variant == "v2" else [320, 640, 1280, 1280]
    ctrl = []
    for lvl in range(4):
        s = 64 >> lvl
        c = (0.5 * torch.randn(B, s, s, ch[lvl], generator=g)).to(dev).to(torch.bfloat16)
        c = c.permute(0, 3, 1, 2).requires_grad_(True)  # NCHW view of channels-last memory
        for p in cl.lora_layers[lvl]:
            p.inject_control_states(c)
        ctrl.append(c)
    print(f"build {time.time()-t0:.1f}s", flush=True)
    return unet, cl, ctrl


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--variant", default="v1")
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--breakdown", action="store_true")
    ap.add_argument("--fwd-only", action="store_true")
    ap.add_argument("--trainer", action="store_true", help="time the full fused Trainer.step (hint encoder + UNet + optimizer)")
    a = ap.parse_args()
    B = a.batch
    unet, cl, ctrl = build(a.variant, B)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, 4, 64, 64, generator=g).cuda()
    t = torch.randint(0, 1000, (B,), generator=g).cuda().float()
    e = torch.randn(B, 77, 768, generator=g).cuda().to(torch.bfloat16)
    tgt = torch.randn(B, 4, 64, 64, generator=g).cuda()

    if a.trainer:
        from controllora_b200.trainer import Trainer
        tr = Trainer(unet, cl, lr=1e-4)
        guide = ((torch.rand(B, 1, 512, 512, generator=g) < 0.08).float() * 2 - 1).expand(B, 3, 512, 512).contiguous().cuda()

    def step():
        if a.trainer:
            return tr.step(x, t, e, guide, tgt)
        control, _ = unet.collect_control(need_grad=not a.fwd_only)
        tape = None if a.fwd_only else Tape()
        pred, ctx, rt = unet.run_engine(x, t, e, control, tape)
        loss, dpred = ops.mse_loss(pred.data, tgt)
        if not a.fwd_only:
            unet.grad_store.zero()
            pred.grad = dpred
            tape.backward()
        return loss

    times = defaultdict(float)
    counts = defaultdict(int)
    if a.breakdown:
        import functools

        def wrap(name, fn):
            @functools.wraps(fn)
            def w(*args, **kw):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                key = name
                if name == "gemm":
                    A_, B_ = args[0], args[1]
                    conv = kw.get("conv_stride", 0)
                    M = A_.shape[0] if not conv else A_.shape[0] * A_.shape[1] * A_.shape[2] // (conv * conv)
                    key = f"gemm{'_conv' if conv else ''}{'_lora' if kw.get('lora_up') is not None else ''} M={M} N={B_.shape[0]} K={B_.shape[1]}"
                e0.record()
                r = fn(*args, **kw)
                e1.record()
                pend.append((key, e0, e1))     # keep no tensor references: the allocator must be able to recycle memory
                return r
            return w

        pend = []
        for name in ["gemm", "attention_fwd", "attention_bwd", "groupnorm_fwd", "groupnorm_bwd", "layernorm_fwd", "layernorm_bwd",
                     "geglu_fwd", "geglu_bwd", "add", "upsample2x_fwd", "upsample2x_bwd", "zero_insert2x", "concat_channels",
                     "slice_channels", "skinny_atb", "rowdot", "rowmat", "skinny_small", "small_matmul", "hilo_combine",
                     "rank_update", "conv_in", "conv_out", "conv_out_bwd", "small_linear", "mse_loss", "conv_wgrad", "colsum",
                     "conv_weight_prep", "conv_in_wgrad", "sumsq", "adamw", "f32_to_bf16", "nchw_to_nhwc"]:
            setattr(ops, name, wrap(name, getattr(ops, name)))

    for i in range(2):
        loss = step()
    torch.cuda.synchronize()
    print("warm loss", float(loss), "mem GB", torch.cuda.max_memory_allocated() / 2**30, flush=True)
    if a.breakdown:
        pend.clear()
    n0 = cb._lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    for i in range(a.iters):
        step()
    e1.record()
    torch.cuda.synchronize()
    wall = (time.time() - t0) / a.iters * 1e3
    ms = e0.elapsed_time(e1) / a.iters
    n1 = cb._lib.launch_count()
    print(f"variant={a.variant} B={B} fwd_only={a.fwd_only}: {ms:.2f} ms/step (GPU events), wall {wall:.2f} ms, "
          f"{(n1-n0)//a.iters} launches/step, {B/ms*1e3:.1f} img/s", flush=True)
    if a.breakdown:
        for key, s0, s1 in pend:
            times[key] += s0.elapsed_time(s1) / a.iters
            counts[key] += 1
        fam = defaultdict(float)
        for k, v in times.items():
            fam[k.split(" ")[0]] += v
        print("---- per family (ms/step, serialised event timing)")
        for k, v in sorted(fam.items(), key=lambda kv: -kv[1]):
            print(f"  {k:24s} {v:8.3f}")
        print("---- top GEMM shapes")
        for k, v in sorted(times.items(), key=lambda kv: -kv[1])[:40]:
            print(f"  {v:8.3f} ms  x{counts[k]//a.iters:3d}  {k}")


if __name__ == "__main__":
    main()
