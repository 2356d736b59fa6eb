"""Debug (CLB_TIMELINE build): globaltimer entry / exit of back-to-back gemm_tc_kernel launches -> in-kernel CTA lifetime, launch
skew and the idle gap between consecutive dependent launches.   python tools/gemm_gaps.py M N K [lora]"""
import ctypes as C
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from controllora_b200 import _lib, ops

M, N, K = (int(v) for v in sys.argv[1:4])
lora = "lora" in sys.argv[4:]
a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
b = (torch.randn(N, K, device="cuda") / K**0.5).to(torch.bfloat16)
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
kw = {}
if lora:
    down = torch.randn(4, K, device="cuda") / 4
    up = torch.randn(N, 4, device="cuda") * 0.1
    kw = dict(ext=ops.split_bf16_ext(down, K), lora_up=up, lora_scale=1.0, t_out=torch.empty(M, 4, device="cuda"))
lib = _lib.lib()
for _ in range(70):          # warm-up; leaves the launch ordinal at a known phase (mod 64 = 6)
    ops.gemm(a, b, out=out, **kw)
torch.cuda.synchronize()
lib.cl_debug_timeline(None, None, 1)
n = 12
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n):
    ops.gemm(a, b, out=out, **kw)
e1.record()
buf = (C.c_ulonglong * (64 * 4))()
lib.cl_debug_gtimes(buf)
rows = [(buf[i * 4], buf[i * 4 + 1], buf[i * 4 + 2], buf[i * 4 + 3]) for i in range(64) if buf[i * 4 + 3] != 0]
rows.sort()
t0 = rows[0][0]
print(f"shape {M}x{N}x{K} lora={lora}: {len(rows)} launches, event time per launch {e0.elapsed_time(e1) / n * 1e3:.2f} us")
prev_exit = None
for fe, le, fx, lx in rows:
    gap = (fe - prev_exit) / 1e3 if prev_exit is not None else float('nan')
    print(f"  first entry {(fe - t0) / 1e3:8.2f} us | entry skew {(le - fe) / 1e3:5.2f} | first exit +{(fx - fe) / 1e3:6.2f} | last exit +{(lx - fe) / 1e3:6.2f} | gap from previous last exit {gap:5.2f} us")
    prev_exit = lx
