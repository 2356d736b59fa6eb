"""GPU parity of `controllora_b200.CLIPTextModel` (train_text_to_image_control_lora.py:768 `text_encoder(input_ids)[0]`) against the
fp32 CPU oracle oracle/clip_ref.py (itself pinned to transformers.CLIPTextModel in tests/test_oracle.py), next to the SAME oracle
math run the way the reference runs the frozen encoder: bf16 weights and activations in eager PyTorch on the GPU.
usage: python tests/check_clip.py [tiny|full]"""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from tests._device import DEV  # noqa: E402

CAP = 3e-2          # relative L2 of last_hidden_state vs the fp32 oracle (12 bf16 transformer layers)
MARGIN = 1.25       # ours <= MARGIN * eager-bf16 error (both are bf16 pipelines with different rounding points)


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def run(which="tiny"):
    import torch
    from oracle import clip_ref as CR
    import controllora_b200 as cb

    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = dict(CR.SD15_TEXT_CONFIG)
    if which == "tiny":
        cfg.update(num_hidden_layers=2, vocab_size=1000)
        B, T = 2, 77
    else:
        B, T = 8, 77
    sd = CR.synthetic_state_dict(cfg, seed=5)
    sd = {k: v.to(torch.bfloat16).float() for k, v in sd.items()}            # weight_dtype = bf16 in every arm
    g = torch.Generator().manual_seed(2)
    ids = torch.randint(0, cfg["vocab_size"], (B, T), generator=g)
    ids[:, 0] = cfg["vocab_size"] - 2                                        # BOS / padding-like repeated ids as in real prompts
    ids[:, 20:] = cfg["vocab_size"] - 1
    t0 = time.time()
    ref = CR.clip_text_forward(sd, ids, cfg)
    t_or = time.time() - t0
    model = cb.CLIPTextModel.from_state_dict(sd, DEV, cfg)
    out = model(ids.to(DEV))
    y = out[0]
    assert y.shape == (B, T, cfg["hidden_size"]) and y.dtype == torch.bfloat16 and out.last_hidden_state is y
    # short prompt (T < 77) goes through the same kernels
    y_short = model(ids[:, :16].to(DEV))[0]
    ref_short = CR.clip_text_forward(sd, ids[:, :16], cfg)
    # the reference's own precision: the same math in eager PyTorch, bf16 weights + activations on the GPU
    sd16 = {k: v.to(DEV).to(torch.bfloat16) for k, v in sd.items()}

    import torch.nn.functional as F
    x = sd16["text_model.embeddings.token_embedding.weight"][ids.to(DEV)] + sd16["text_model.embeddings.position_embedding.weight"][:T]
    Cw, heads = cfg["hidden_size"], cfg["num_attention_heads"]
    d = Cw // heads
    causal = torch.full((T, T), float("-inf"), device=DEV, dtype=torch.bfloat16).triu(1)
    for i in range(cfg["num_hidden_layers"]):
        p = f"text_model.encoder.layers.{i}."
        w = lambda k: sd16[p + k]
        h = F.layer_norm(x, (Cw,), w("layer_norm1.weight"), w("layer_norm1.bias"), 1e-5)
        q = F.linear(h, w("self_attn.q_proj.weight"), w("self_attn.q_proj.bias")) * (d ** -0.5)
        k = F.linear(h, w("self_attn.k_proj.weight"), w("self_attn.k_proj.bias"))
        v = F.linear(h, w("self_attn.v_proj.weight"), w("self_attn.v_proj.bias"))
        sh = lambda t: t.view(B, T, heads, d).transpose(1, 2)
        s = sh(q) @ sh(k).transpose(-1, -2) + causal
        a = (torch.softmax(s.float(), -1).to(torch.bfloat16) @ sh(v)).transpose(1, 2).reshape(B, T, Cw)
        x = x + F.linear(a, w("self_attn.out_proj.weight"), w("self_attn.out_proj.bias"))
        h = F.layer_norm(x, (Cw,), w("layer_norm2.weight"), w("layer_norm2.bias"), 1e-5)
        f = F.linear(h, w("mlp.fc1.weight"), w("mlp.fc1.bias"))
        f = f * torch.sigmoid(1.702 * f)
        x = x + F.linear(f, w("mlp.fc2.weight"), w("mlp.fc2.bias"))
    eager = F.layer_norm(x, (Cw,), sd16["text_model.final_layer_norm.weight"], sd16["text_model.final_layer_norm.bias"], 1e-5)
    e_ours, e_eager, e_short = rel(y, ref), rel(eager, ref), rel(y_short, ref_short)
    print(f"[clip {which}] B={B} T={T} layers={cfg['num_hidden_layers']}: ours vs fp32 oracle {e_ours:.3e} | eager bf16 vs fp32 oracle {e_eager:.3e} | "
          f"T=16 prompt {e_short:.3e} | oracle CPU {t_or:.1f}s", flush=True)
    ok = e_ours <= CAP and e_short <= CAP and e_ours <= MARGIN * e_eager and bool(torch.isfinite(y.float()).all())
    print("CASE_OK" if ok else "CASE_FAIL")
    return ok


if __name__ == "__main__":
    sys.exit(0 if run(sys.argv[1] if len(sys.argv) > 1 else "tiny") else 1)
