"""Driver for ncu captures of the GroupNorm kernels: python tools/ncu_norm.py [n HW C]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from controllora_b200 import ops

n, HW, C = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (8, 4096, 320)))
x = torch.randn(n, HW, C, device="cuda").to(torch.bfloat16)
dy = torch.randn(n, HW, C, device="cuda").to(torch.bfloat16)
g = torch.randn(C, device="cuda")
b = torch.randn(C, device="cuda")
for _ in range(3):
    y, st = ops.groupnorm_fwd(x, g, b, 32, 1e-5, True)
    ops.groupnorm_bwd(x, dy, g, b, st, 32, True)
torch.cuda.synchronize()
print("done")
