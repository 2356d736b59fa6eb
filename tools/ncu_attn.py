"""Driver for ncu captures of the attention kernels: python tools/ncu_attn.py [B H N Nk d]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from controllora_b200 import ops

B, H, N, Nk, d = (int(v) for v in (sys.argv[1:6] if len(sys.argv) > 5 else (8, 8, 4096, 4096, 40)))
q = torch.randn(B, N, H * d, device="cuda").to(torch.bfloat16)
k = torch.randn(B, Nk, H * d, device="cuda").to(torch.bfloat16)
v = torch.randn(B, Nk, H * d, device="cuda").to(torch.bfloat16)
do = torch.randn(B, N, H * d, device="cuda").to(torch.bfloat16)
for _ in range(2):
    o, lse = ops.attention_fwd(q, k, v, H, d ** -0.5)
    ops.attention_bwd(q, k, v, o, do, lse, H, d ** -0.5)
torch.cuda.synchronize()
print("done")
