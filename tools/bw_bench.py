"""HBM-roofline microbenchmark of the memory-bound kernels at the SD-1.5 training shapes.

Every op is timed with CUDA events over a rotation of operand sets whose total footprint exceeds the 126 MB L2 (so the
numbers are DRAM numbers, not L2-hit numbers), and reported as achieved GB/s over the ALGORITHMIC bytes of one call
(each tensor read / written once) next to the measured HBM peak of MEASURED_PEAKS.json.

usage: python tools/bw_bench.py [--iters 20]
"""
import argparse
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import torch

from controllora_b200 import ops

BF16 = torch.bfloat16


def hbm_peak():
    p = Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json"
    try:
        d = json.loads(p.read_text())
        for k in ("hbm_gbs_sustained", "hbm_gbs", "hbm_gbps"):
            if k in d:
                return float(d[k])
    except Exception:
        pass
    return 6500.0


def timed(fn, sets, iters):
    for s in sets:
        fn(*s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 0
    for _ in range(iters):
        for s in sets:
            fn(*s)
            n += 1
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n      # us per call


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    a = ap.parse_args()
    dev = "cuda"
    peak = hbm_peak()
    g = torch.Generator(device=dev).manual_seed(0)
    rows = []

    def rnd(*shape, dtype=BF16):
        return torch.randn(*shape, device=dev, generator=g).to(dtype)

    def nsets(bytes_per_set):
        return max(2, int(400e6 // bytes_per_set) + 1)

    def report(name, us, nbytes):
        gbps = nbytes / us / 1e3
        rows.append((name, us, nbytes / 1e6, gbps, gbps / peak))
        print(f"{name:44s} {us:9.1f} us  {nbytes/1e6:8.1f} MB  {gbps:8.0f} GB/s  {100*gbps/peak:5.1f} % of HBM peak", flush=True)

    shapes = [(8, 4096, 320), (8, 1024, 640), (8, 256, 1280), (8, 4096, 960)]
    for (n, HW, Cc) in shapes:
        M = n * HW
        tb = M * Cc * 2
        gam, bet = torch.ones(Cc, device=dev), torch.zeros(Cc, device=dev)
        # ---- GroupNorm
        K = nsets(3 * tb)
        xs = [rnd(n, HW, Cc) for _ in range(K)]
        dys = [rnd(n, HW, Cc) for _ in range(K)]
        outs = [torch.empty_like(x) for x in xs]
        stats = [ops.groupnorm_fwd(x, gam, bet, 32, 1e-5, True)[1] for x in xs]
        us = timed(lambda x, o: ops.groupnorm_fwd(x, gam, bet, 32, 1e-5, True, out=o), list(zip(xs, outs)), a.iters)
        report(f"groupnorm_fwd+silu  [{M},{Cc}]", us, 2 * tb)           # x read (the second read may hit L2), y written
        dg, db = torch.zeros(Cc, device=dev), torch.zeros(Cc, device=dev)
        us = timed(lambda x, dy, st, o: ops.groupnorm_bwd(x, dy, gam, bet, st, 32, True, dx=o, dgamma=dg, dbeta=db),
                   list(zip(xs, dys, stats, outs)), a.iters)
        report(f"groupnorm_bwd+silu  [{M},{Cc}]", us, 3 * tb)
        # ---- LayerNorm
        lst = [ops.layernorm_fwd(x.view(M, Cc), gam, bet)[1] for x in xs]
        us = timed(lambda x: ops.layernorm_fwd(x.view(M, Cc), gam, bet), [(x,) for x in xs], a.iters)
        report(f"layernorm_fwd       [{M},{Cc}]", us, 2 * tb)
        us = timed(lambda x, dy, st, o: ops.layernorm_bwd(x.view(M, Cc), dy.view(M, Cc), gam, st, dx=o.view(M, Cc)),
                   list(zip(xs, dys, lst, outs)), a.iters)
        report(f"layernorm_bwd       [{M},{Cc}]", us, 3 * tb)
        # ---- residual add
        us = timed(lambda x, dy, o: ops.add(x, dy, out=o), list(zip(xs, dys, outs)), a.iters)
        report(f"add                 [{M},{Cc}]", us, 3 * tb)
        # ---- rank-r update (V2 control injection) and its backward reductions
        ts = [torch.randn(M, 4, device=dev, generator=g) for _ in range(K)]
        tab = torch.randn(Cc, 4, device=dev, generator=g)
        us = timed(lambda x, t, o: ops.rank_update(x.view(M, Cc), t, tab, 0.5, out=o.view(M, Cc)), list(zip(xs, ts, outs)), a.iters)
        report(f"rank_update r=4     [{M},{Cc}]", us, 2 * tb + M * 16)
        us = timed(lambda x: ops.rowdot(x.view(M, Cc), tab), [(x,) for x in xs], a.iters)
        report(f"rowdot r=4          [{M},{Cc}]", us, tb + M * 16)
        # ---- LoRA weight-gradient reductions, 16 problems per launch (as the training step batches them)
        if Cc <= 1280:
            acc = [torch.zeros(4, Cc, device=dev) for _ in range(16)]
            pool = list(zip(xs, ts))

            def skinny_batch(off):
                for i in range(16):
                    x, t = pool[(off + i) % len(pool)]
                    ops.SKINNY.add(t, 4, x.view(M, Cc), acc[i], Cc, 1, 1.0)
                ops.SKINNY.flush()

            us = timed(skinny_batch, [(i,) for i in range(0, len(pool), 4)], max(2, a.iters // 4))
            report(f"skinny_atb_batch x16 [{M},{Cc}]", us, 16 * (tb + M * 16))
        # ---- column sum (bias gradients)
        cs = torch.zeros(Cc, device=dev)
        us = timed(lambda x: ops.colsum(x.view(M, Cc), cs), [(x,) for x in xs], a.iters)
        report(f"colsum              [{M},{Cc}]", us, tb)
        del xs, dys, outs, stats, lst, ts
        torch.cuda.empty_cache()

    # ---- GEGLU at the two large FFN shapes
    for (M, F) in [(32768, 1280), (8192, 2560)]:
        K = nsets(M * 2 * F * 2 * 2)
        ps = [rnd(M, 2 * F) for _ in range(K)]
        ds = [rnd(M, F) for _ in range(K)]
        us = timed(lambda p: ops.geglu_fwd(p), [(p,) for p in ps], a.iters)
        report(f"geglu_fwd           [{M},{2*F}]", us, M * 3 * F * 2)
        us = timed(lambda p, d: ops.geglu_bwd(p, d), list(zip(ps, ds)), a.iters)
        report(f"geglu_bwd           [{M},{2*F}]", us, M * 5 * F * 2)
        del ps, ds
        torch.cuda.empty_cache()
    print(json.dumps({"hbm_peak_gbps": peak, "rows": [dict(zip(("op", "us", "MB", "GBps", "frac"), r)) for r in rows]}))


if __name__ == "__main__":
    main()
