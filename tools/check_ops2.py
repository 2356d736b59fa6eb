"""GPU bring-up checks, part 2: attention backward, LoRA side-path kernels, optimizer.  usage as check_ops.py."""
import subprocess
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

CASES = {}


def case(fn):
    CASES[fn.__name__] = fn
    return fn


def _setup():
    import torch

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(0)
    return torch


def _rel(a, b):
    return float((a.detach().float() - b.detach().float()).norm() / (b.detach().float().norm() + 1e-30))


def _attn_bwd_case(B, H, Nq, Nk, d, tol=1e-2):
    torch = _setup()
    from controllora_b200 import ops

    q = torch.randn(B, Nq, H * d, device="cuda").to(torch.bfloat16)
    k = torch.randn(B, Nk, H * d, device="cuda").to(torch.bfloat16)
    v = torch.randn(B, Nk, H * d, device="cuda").to(torch.bfloat16)
    scale = d**-0.5
    o, lse = ops.attention_fwd(q, k, v, H, scale)
    d_o = torch.randn(B, Nq, H * d, device="cuda").to(torch.bfloat16)
    dq, dk, dv = ops.attention_bwd(q, k, v, o, d_o, lse, H, scale)
    torch.cuda.synchronize()
    qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
    qh = qf.view(B, Nq, H, d).transpose(1, 2)
    kh = kf.view(B, Nk, H, d).transpose(1, 2)
    vh = vf.view(B, Nk, H, d).transpose(1, 2)
    p = ((qh @ kh.transpose(-1, -2)) * scale).softmax(-1)
    oref = (p @ vh).transpose(1, 2).reshape(B, Nq, H * d)
    oref.backward(d_o.float())
    e = (_rel(dq, qf.grad), _rel(dk, kf.grad), _rel(dv, vf.grad))
    print(f"attn bwd B={B} H={H} Nq={Nq} Nk={Nk} d={d}: dq rel={e[0]:.3e} dk rel={e[1]:.3e} dv rel={e[2]:.3e}")
    assert max(e) < tol, e


@case
def attn_bwd_one_block():
    _attn_bwd_case(1, 1, 128, 128, 64)


@case
def attn_bwd_d40():
    _attn_bwd_case(2, 8, 1024, 1024, 40)
    _attn_bwd_case(2, 8, 512, 77, 40)


@case
def attn_bwd_d80_d160():
    _attn_bwd_case(2, 8, 1024, 1024, 80)
    _attn_bwd_case(2, 8, 256, 256, 160)
    _attn_bwd_case(2, 8, 64, 64, 160)
    _attn_bwd_case(2, 8, 256, 77, 160)


@case
def attn_bwd_small_d():
    _attn_bwd_case(2, 4, 256, 256, 16)
    _attn_bwd_case(1, 2, 200, 300, 32)


@case
def attn_perf2():
    torch = _setup()
    from controllora_b200 import ops

    for (B, H, N, Nk, d) in [(8, 8, 4096, 4096, 40), (8, 8, 4096, 77, 40), (8, 8, 1024, 1024, 80), (8, 8, 256, 256, 160)]:
        q = torch.randn(B, N, H * d, device="cuda").to(torch.bfloat16)
        k = torch.randn(B, Nk, H * d, device="cuda").to(torch.bfloat16)
        v = torch.randn(B, Nk, H * d, device="cuda").to(torch.bfloat16)
        d_o = torch.randn(B, N, H * d, device="cuda").to(torch.bfloat16)
        o, lse = ops.attention_fwd(q, k, v, H, d**-0.5)
        dq, dk, dv = ops.attention_bwd(q, k, v, o, d_o, lse, H, d**-0.5)
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        for _ in range(5):
            ops.attention_fwd(q, k, v, H, d**-0.5, out=o)
        e1.record()
        for _ in range(5):
            ops.attention_bwd(q, k, v, o, d_o, lse, H, d**-0.5, dq=dq, dk=dk, dv=dv)
        e2.record()
        torch.cuda.synchronize()
        msf, msb = e0.elapsed_time(e1) / 5, e1.elapsed_time(e2) / 5
        fl = 4.0 * B * H * N * Nk * d
        print(f"attn perf B={B} N={N} Nk={Nk} d={d}: fwd {msf*1e3:.1f} us {fl/msf/1e9:.1f} TFLOP/s | bwd {msb*1e3:.1f} us {2.5*fl/msb/1e9:.1f} TFLOP/s(5-matmul count)")


@case
def lora_kernels():
    torch = _setup()
    from controllora_b200 import ops

    dev = "cuda"
    r, K, N, M = 4, 320, 640, 3000
    down = torch.randn(r, K, device=dev) / r
    up = torch.randn(N, r, device=dev) * 0.3
    ext = torch.zeros(16, K, device=dev, dtype=torch.bfloat16)
    tab = torch.zeros(N, 4, device=dev)
    ext_t = torch.zeros(16, N, device=dev, dtype=torch.bfloat16)
    tab_t = torch.zeros(K, 4, device=dev)
    plan = ops.PackPlan(dev)
    plan.add_ext(down, ext)
    plan.add_table(up, tab)
    plan.add_ext(up, ext_t, transposed=True)          # rows j = up[:, j]
    plan.add_table(down, tab_t, transposed=True)      # tab_t[k, j] = down[j, k]
    plan.run()
    torch.cuda.synchronize()
    ref = ops.split_bf16_ext(down, K)
    assert torch.equal(ext, ref)
    assert torch.equal(tab, up)
    assert torch.equal(ext_t, ops.split_bf16_ext(up.t().contiguous(), N))
    assert torch.equal(tab_t, down.t().contiguous())
    # skinny reductions
    a = torch.randn(M, 4, device=dev)
    b = torch.randn(M, K, device=dev).to(torch.bfloat16)
    out = torch.zeros(r, K, device=dev)
    ops.skinny_atb(a, r, b, out, K, 1, 0.5)
    outT = torch.zeros(K, r, device=dev)
    ops.skinny_atb(a, r, b, outT, 1, r, 0.5)
    refo = 0.5 * a.t() @ b.float()
    print(f"skinny_atb rel={_rel(out, refo):.3e} transposed rel={_rel(outT, refo.t()):.3e}")
    assert _rel(out, refo) < 1e-5 and _rel(outT, refo.t()) < 1e-5
    bw = torch.randn(M, 1280, device=dev).to(torch.bfloat16)
    a8 = torch.randn(M, 8, device=dev)
    o8 = torch.zeros(8, 1280, device=dev)
    ops.skinny_atb(a8, 8, bw, o8, 1280, 1, 1.0)
    assert _rel(o8, a8.t() @ bw.float()) < 1e-5
    # rowdot
    u = torch.randn(N, 4, device=dev)
    dy = torch.randn(M, N, device=dev).to(torch.bfloat16)
    e = ops.rowdot(dy, u)
    print(f"rowdot rel={_rel(e, dy.float() @ u):.3e}")
    assert _rel(e, dy.float() @ u) < 1e-5
    # rowmat fp32 + bf16 hi/lo
    w = torch.randn(4, 4, device=dev)
    o = torch.empty(M, 4, device=dev)
    ops.rowmat(a, w, 4, 1, 4, 4, 0.7, o, 4)
    assert _rel(o, 0.7 * a @ w.t()) < 1e-5
    ob = torch.zeros(M, 128, device=dev, dtype=torch.bfloat16)
    ops.rowmat(a, w, 4, 1, 4, 4, 0.7, ob, 128, out_mode=1, col_off=8, lo_off=64)
    rec = ob[:, 8:12].float() + ob[:, 72:76].float()
    assert _rel(rec, 0.7 * a @ w.t()) < 1e-4
    # skinny_small
    b4 = torch.randn(M, 4, device=dev)
    ss = torch.zeros(4, 4, device=dev)
    ops.skinny_small(a, 4, b4, 4, ss, 2.0)
    assert _rel(ss, 2.0 * a.t() @ b4) < 1e-5
    # small_matmul: out[K, 4] = down^T[K, r] @ w[r, 4]
    sm = torch.zeros(K, 4, device=dev)
    ops.small_matmul(down, 1, K, w, 4, 1, sm, 4, 1, K, r, 4, alpha=1.5)
    assert _rel(sm, 1.5 * down.t() @ w) < 1e-5
    print("lora kernels ok")


@case
def optimizer():
    torch = _setup()
    from controllora_b200 import ops

    n = 100003
    p = torch.randn(n, device="cuda")
    g = torch.randn(n, device="cuda") * 3
    p_ref = p.clone().requires_grad_(True)
    opt = torch.optim.AdamW([p_ref], lr=1e-4, betas=(0.9, 0.999), weight_decay=1e-2, eps=1e-8)
    m = torch.zeros(n, device="cuda")
    v = torch.zeros(n, device="cuda")
    for step in range(1, 4):
        gi = g * step
        p_ref.grad = gi.clone()
        torch.nn.utils.clip_grad_norm_([p_ref], 1.0)
        opt.step()
        gg = gi.clone()
        nrm = torch.zeros(1, device="cuda")
        ops.sumsq(gg, nrm)
        ops.adamw(p, gg, m, v, 1e-4, 0.9, 0.999, 1e-8, 1e-2, step, gnorm_sq=nrm, max_norm=1.0)
        torch.cuda.synchronize()
        assert float(gg.abs().max()) == 0.0
        err = float((p - p_ref.detach()).abs().max())
        print(f"adamw step {step}: max abs diff {err:.3e}")
        assert err < 1e-6


def main():
    if len(sys.argv) > 1:
        CASES[sys.argv[1]]()
        print("CASE_OK")
        return
    results = {}
    for name in CASES:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, __file__, name], capture_output=True, text=True, timeout=150)
            ok = r.returncode == 0 and "CASE_OK" in r.stdout
            out = r.stdout + r.stderr
        except subprocess.TimeoutExpired as e:
            ok, out = False, f"TIMEOUT {e}"
        results[name] = ok
        print(f"=== {name}: {'PASS' if ok else 'FAIL'} ({time.time()-t0:.1f}s)")
        tail = out.strip().splitlines()
        for line in (tail if ok else tail[-25:]):
            print("    " + line)
        sys.stdout.flush()
    print("SUMMARY", results)
    sys.exit(0 if all(results.values()) else 1)


if __name__ == "__main__":
    main()
