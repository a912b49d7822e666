"""Quick timing of the fused scoring kernel (K3): python tools/perf_score.py [n] [d] [kind]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bodywork_mlops_demo_b200 as b2
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 128
kind = sys.argv[3] if len(sys.argv) > 3 else "f32"
ctx = b2.Context(0)
X, y = ctx.synth(n, d, kind=kind)
coef = np.full(d, 0.5)
out = ctx.empty((n,), "f32")          # preallocated: the timed region is the kernel, not a cudaMalloc
for want_yhat in (False, True):
    best = 1e9
    for _ in range(6):
        ctx.sync(); ctx.timer_start()
        yh, st = ctx.score(X, coef, 1.0, y=y, want_yhat=want_yhat, out=out if want_yhat else None)
        ms = ctx.timer_stop(); best = min(best, ms)
    bpr = d * (4 if kind == "f32" else 2) + 4 + (4 if want_yhat else 0)
    print(f"score {kind} n={n} d={d} yhat={want_yhat}: {best:.3f} ms  {n/best/1e6:.2f} G rows/s  {n*bpr/best/1e6:.0f} GB/s = {n*bpr/best/1e6/6575.1:.3f} of HBM peak  (rows {st[5]:.0f})")
