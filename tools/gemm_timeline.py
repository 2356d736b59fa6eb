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
names = {1: "entry", 10: "prod:tile", 11: "prod:slot", 12: "prod:issued", 24: "mma:kb_ready", 25: "mma:kb_issued", 20: "mma:wait_acc", 21: "mma:acc_free", 22: "mma:stage0", 23: "mma:issued", 30: "epi:wait",
         31: "epi:ready", 32: "epi:done", 33: "epi:gran_regs", 34: "epi:lora_done", 35: "epi:staged", 36: "epi:p2_loaded", 37: "epi:stores_issued",
         38: "epi:t_ready", 39: "epi:acc2_ready"}
for sm in (0, 77):
    n = min(cnt[sm], 256)
    ev = sorted((buf[(sm * 256 + i) * 2], buf[(sm * 256 + i) * 2 + 1]) for i in range(n) if buf[(sm * 256 + i) * 2 + 1] != 0)
    if not ev:
        continue
    t0 = ev[0][0]
    print(f"--- SM {sm}: {n} events")
    for t, tag in ev:
        print(f"  {t - t0:8d}  {names.get(tag, tag)}")

# summary: tensor-pipe time per tcgen05.mma dispatch (K = 16) on SM 0, from the MMA thread's stage0 -> issued intervals
import os
reps = int(os.environ.get("CLB_TL_MMA_REPS", "0"))
n = min(cnt[0], 256)
ev = sorted((buf[i * 2], buf[i * 2 + 1]) for i in range(n) if buf[i * 2 + 1] in (22, 23))
pairs = [(b[0] - a[0]) for a, b in zip(ev, ev[1:]) if a[1] == 22 and b[1] == 23]
nkb = (K + 63) // 64
if pairs:
    per = [d / (nkb * 4 * (reps + 1)) for d in pairs]
    print(f"SUMMARY M={M} N={N} K={K} bn={bn} reps={reps}: tiles={len(pairs)} clk/dispatch min={min(per):.1f} avg={sum(per)/len(per):.1f}  (k-blocks/tile {nkb})")
