"""Tiny driver for ncu captures of cl_gemm:
    python tools/ncu_gemm.py M N K [lora] [iters]
    python tools/ncu_gemm.py conv n H W C N [iters]        (3x3 stride-1 implicit GEMM on an NHWC image)"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from controllora_b200 import ops

if sys.argv[1] == "conv":
    n, H, W, Cc, N = (int(v) for v in sys.argv[2:7])
    iters = int(sys.argv[7]) if len(sys.argv) > 7 else 8
    x = torch.randn(n, H, W, Cc, device="cuda").to(torch.bfloat16)
    w = (torch.randn(N, 9 * Cc, device="cuda") / (9 * Cc) ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    out = torch.empty(n, H, W, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(iters):
        ops.gemm(x, w, conv_stride=1, pad_lo=1, bias=bias, out=out)
    torch.cuda.synchronize()
    print("done")
    sys.exit(0)

M, N, K = (int(v) for v in sys.argv[1:4])
lora = len(sys.argv) > 4 and sys.argv[4] == "lora"
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 8
a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
b = (torch.randn(N, K, device="cuda") / K**0.5).to(torch.bfloat16)
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
kw = {}
if lora:
    down = torch.randn(4, K, device="cuda") / 4
    up = torch.randn(N, 4, device="cuda") * 0.1
    kw = dict(ext=ops.split_bf16_ext(down, K), lora_up=up, lora_scale=1.0, t_out=torch.empty(M, 4, device="cuda"))
for _ in range(iters):
    ops.gemm(a, b, out=out, **kw)
torch.cuda.synchronize()
print("done")
