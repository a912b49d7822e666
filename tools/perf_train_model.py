"""train_model(DataFrame) end to end, the stage's own entry point: python tools/perf_train_model.py [n] [d]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pandas as pd
import bodywork_mlops_demo_b200 as b2
from bodywork_mlops_demo_b200 import stage_1_train_model as s1
n = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 128
rng = np.random.default_rng(0)
X = rng.random((n, d)) * 100.0
y = 1.0 + 0.5 * X.sum(axis=1) + 10.0 * rng.standard_normal(n)
df = pd.DataFrame({"date": "2021-01-01", "y": y, **{f"X{j}": X[:, j] for j in range(d)}})
s1.train_model(df)
best = 1e9
for _ in range(3):
    t0 = time.perf_counter(); model, metrics = s1.train_model(df); best = min(best, time.perf_counter() - t0)
t0 = time.perf_counter(); cols = [df[c].to_numpy() for c in s1.feature_columns(df)]; t_cols = time.perf_counter() - t0
t0 = time.perf_counter(); Xd = s1.default_context().upload_columns(cols); t_up = time.perf_counter() - t0; Xd.free()
t0 = time.perf_counter(); s1.split_mask(n); t_mask = time.perf_counter() - t0
print(f"train_model(DataFrame {n} x {d} float64): {best*1e3:.1f} ms = {n/best/1e6:.2f} M rows/s  "
      f"(column views {t_cols*1e3:.1f} ms, upload_columns {t_up*1e3:.1f} ms = {n*d*8/t_up/1e9:.1f} GB/s of float64 read, "
      f"split mask {t_mask*1e3:.1f} ms); r2 {float(metrics['r_squared'][0]):.4f}")
from sklearn.linear_model import LinearRegression
from sklearn.model_selection import train_test_split
t0 = time.perf_counter()
Xs = df[s1.feature_columns(df)].values; ys = df["y"].values
Xtr, Xte, ytr, yte = train_test_split(Xs, ys, test_size=0.2, random_state=42)
m = LinearRegression(fit_intercept=True).fit(Xtr, ytr); m.predict(Xte)
t_ref = time.perf_counter() - t0
print(f"reference train_model arithmetic (sklearn, float64, all host threads): {t_ref*1e3:.0f} ms = {n/t_ref/1e6:.3f} M rows/s")
