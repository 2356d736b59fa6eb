"""Debug: per-SM event timeline of gemm_tc_kernel (needs a CLB_TIMELINE build):
   CLB_EXTRA_NVCC=-DCLB_TIMELINE python -m controllora_b200.build -f ; python tools/gemm_timeline.py M N K [block_n]"""
import ctypes as C
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from controllora_b200 import _lib, ops

M, N, K = (int(v) for v in sys.argv[1:4])
bn = int(sys.argv[4]) if len(sys.argv) > 4 and sys.argv[4].isdigit() else 0
lora = "lora" in sys.argv[4:]
res = "res" in sys.argv[4:]
a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
b = (torch.randn(N, K, device="cuda") / K**0.5).to(torch.bfloat16)
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
kw = {}
if lora:
    down = torch.randn(4, K, device="cuda") / 4
    up = torch.randn(N, 4, device="cuda") * 0.1
    kw = dict(ext=ops.split_bf16_ext(down, K), lora_up=up, lora_scale=1.0, t_out=torch.empty(M, 4, device="cuda"))
if res:
    kw["residual"] = torch.randn(M, N, device="cuda").to(torch.bfloat16)
    kw["bias"] = torch.randn(N, device="cuda")
for _ in range(3):
    ops.gemm(a, b, out=out, block_n=bn, **kw)
torch.cuda.synchronize()
lib = _lib.lib()
lib.cl_debug_timeline(None, None, 1)
ops.gemm(a, b, out=out, block_n=bn, **kw)
buf = (C.c_ulonglong * (160 * 256 * 2))()
cnt = (C.c_uint * 160)()
lib.cl_debug_timeline(buf, cnt, 0)
names = {1: "entry", 10: "prod:tile", 20: "mma:wait_acc", 21: "mma:acc_free", 22: "mma:stage0", 23: "mma:issued", 30: "epi:wait",
         31: "epi:ready", 32: "epi:done"}
for sm in (0, 1, 77):
    n = min(cnt[sm], 256)
    ev = sorted((buf[(sm * 256 + i) * 2], buf[(sm * 256 + i) * 2 + 1]) for i in range(n))
    if not ev:
        continue
    t0 = ev[0][0]
    print(f"--- SM {sm}: {n} events")
    for t, tag in ev:
        print(f"  {t - t0:8d}  {names.get(tag, tag)}")
