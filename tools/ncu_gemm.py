"""Tiny driver for ncu captures of cl_gemm: python tools/ncu_gemm.py M N K [lora] [iters]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from controllora_b200 import ops

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
