"""B200LinearRegression().fit on float64 host rows (the scikit-learn habit): python tools/perf_f64_fit.py [n] [d]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bodywork_mlops_demo_b200 as b2
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 128
rng = np.random.default_rng(1)
X = rng.random((n, d)) * 100.0
y = 1.0 + 0.5 * X.sum(axis=1) + 10.0 * rng.standard_normal(n)
est = b2.B200LinearRegression()
est.fit(X, y)
best = 1e9
for _ in range(3):
    t0 = time.perf_counter(); est.fit(X, y); best = min(best, time.perf_counter() - t0)
t0 = time.perf_counter(); Xf = np.ascontiguousarray(X, dtype=np.float32); t_np = time.perf_counter() - t0
print(f"fit(float64 host {n} x {d}): {best*1e3:.1f} ms = {n/best/1e6:.1f} M rows/s ({n*d*8/best/1e9:.1f} GB/s of float64 read); "
      f"numpy astype(float32) alone: {t_np*1e3:.0f} ms; coef0 {est.coef_[0]:.5f}")
