"""Default fit() on pageable (numpy) rows vs pinned rows: python tools/perf_pageable.py [n] [d]   (B2_COPY_THREADS=k)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bodywork_mlops_demo_b200 as b2
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 128
ctx = b2.Context(0)
rng = np.random.default_rng(1)
X = rng.random((n, d), dtype=np.float32) * 100.0
y = (1.0 + 0.5 * X.sum(axis=1) + 10.0 * rng.standard_normal(n, dtype=np.float32)).astype(np.float32)
est = b2.B200LinearRegression(ctx=ctx)
est.fit(X, y)
best = 1e9
for _ in range(3):
    t0 = time.perf_counter(); est.fit(X, y); best = min(best, time.perf_counter() - t0)
print(f"threads={os.environ.get('B2_COPY_THREADS','default')} pageable {n} x {d}: {best*1e3:.1f} ms  {n/best/1e6:.1f} M rows/s  {n*(4*d+4)/best/1e9:.1f} GB/s  coef0 {est.coef_[0]:.5f}")
