"""Measure the BASELINE.json configs that are not bench.py's headline line (development / evidence tool).

    python tools/config_runs.py [c3] [c4] [c5] [shapes] [refshape] [realmask]  -> gpurun_out/r02_configs.json (one record per config)

  c3      configs[2] on ONE GPU: 100 M x 128 fp32 resident in HBM (51.6 GB) -- the north-star target point
  c4      configs[3]: D = 32 rows streamed from pinned host DRAM through the C-ABI (B2_MEM_HOST); 200 M rows
          (26 GB pinned) stand in for 1 B (132 GB): the path is PCIe-bound, the rate does not depend on N
  c5      configs[4]: 30-day concept-drift replay, D = 1 reference-faithful tranches and a 1 M x 128 variant
  refshape  the reference's one-feature shape at 1 B resident rows: masked fit + hold-out metrics (narrow kernels)
  shapes  device-resident Gram-kernel rate for D in {1, 8, 16, 32, 64, 128} x {f32, bf16}
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bodywork_mlops_demo_b200 as b2  # noqa: E402

PEAK = 6575.1
out = {}


def gram_rate(ctx, X, y, n, d, kind, reps=6):
    ctx.set_kernel(b2.KERNEL_AUTO)      # D <= 16: CUDA-core streaming kernel; wider: tcgen05
    best_k, best_t = 1e9, 1e9
    for _ in range(reps):
        ctx.gram_reset(d)
        ctx.timer_start()
        ctx.gram_accumulate(X, y)
        coef, b0 = ctx.solve()
        t = ctx.timer_stop()
        k, _n = ctx.last_kernel_ms()
        best_k, best_t = min(best_k, k), min(best_t, t)
    bpr = d * (4 if kind == "f32" else 2) + 4
    return {"n": n, "d": d, "x": kind, "gram_kernel_ms": best_k, "fit_ms": best_t,
            "gram_rows_per_s": n / best_k * 1e3, "fit_rows_per_s": n / best_t * 1e3,
            "hbm_gbs": n * bpr / best_k / 1e6, "frac_of_measured_peak": n * bpr / best_k / 1e6 / PEAK,
            "coef_head": [float(c) for c in coef[:2]], "intercept": float(b0)}


def c3(ctx):
    n, d = 100_000_000, 128
    X, y = ctx.synth(n, d, seed=1234)
    ctx.sync()
    r = gram_rate(ctx, X, y, n, d, "f32", reps=5)
    X.free(); y.free()
    out["config3_100Mx128_f32_1gpu"] = r
    print("c3", r, flush=True)


def c4(ctx):
    n, d = 200_000_000, 32
    Xp, yp = ctx.pinned((n, d), np.float32), ctx.pinned((n,), np.float32)
    blk = 25_000_000
    for lo in range(0, n, blk):             # fill the pinned rows from the device generator
        Xd, yd = ctx.synth(blk, d, seed=77, row_offset=lo)
        b2.native._check(b2.native.load().b2_copy_d2h(ctx._h, Xp.ptr + lo * d * 4, Xd.ptr, Xd.nbytes), "d2h")
        b2.native._check(b2.native.load().b2_copy_d2h(ctx._h, yp.ptr + lo * 4, yd.ptr, yd.nbytes), "d2h")
        Xd.free(); yd.free()
    ctx.set_kernel(b2.KERNEL_AUTO)
    est = b2.B200LinearRegression(ctx=ctx)
    est.fit(Xp.array[:1_000_000], yp.array[:1_000_000], with_spectrum=False)     # warm-up: staging ring
    best = 1e9
    for _ in range(2):
        t0 = time.perf_counter()
        est.fit(Xp.array, yp.array, with_spectrum=False)
        best = min(best, time.perf_counter() - t0)
    gbs = n * (d * 4 + 4) / best / 1e9
    # the full 1 B rows of configs[3]: five passes over the 26 GB pinned ring, one statistic (no reset in between)
    passes = 5
    ctx.gram_reset(d)
    t0 = time.perf_counter()
    for _ in range(passes):
        ctx.gram_accumulate(Xp.array, yp.array)
    coef_1b, b0_1b = ctx.solve()
    t_1b = time.perf_counter() - t0
    n_1b = int(round(ctx.gram_export()[d, d]))
    r = {"n": n, "d": d, "x": "f32 in pinned host memory", "seconds": best, "rows_per_s": n / best,
         "h2d_gb_per_s": gbs, "note": "PCIe Gen5 x16 bound (~63 GB/s nominal)",
         "coef_head": [float(c) for c in est.coef_[:2]],
         "one_billion_rows": {"rows_accumulated": n_1b, "seconds": t_1b, "rows_per_s": n_1b / t_1b,
                              "h2d_gb_per_s": n_1b * (d * 4 + 4) / t_1b / 1e9,
                              "how": "5 passes over a 200 M-row (26.4 GB) pinned ring into one statistic, then one solve",
                              "coef_head": [float(c) for c in coef_1b[:2]]}}
    m = 20_000_000
    Xn, yn = np.array(Xp.array[:m]), np.array(yp.array[:m])                  # ordinary (pageable) numpy rows
    est.fit(Xn, yn, with_spectrum=False)
    t0 = time.perf_counter()
    est.fit(Xn, yn, with_spectrum=False)
    tp = time.perf_counter() - t0
    r["pageable_rows"] = {"n": m, "seconds": tp, "rows_per_s": m / tp, "h2d_gb_per_s": m * (d * 4 + 4) / tp / 1e9,
                          "how": "numpy rows (not page-locked): pinned bounce ring filled by 8 host threads"}
    Xp.free(); yp.free()
    out["config4_host_streamed_d32"] = r
    print("c4", r, flush=True)


def c5(ctx):
    from bodywork_mlops_demo_b200 import incremental
    res = {}
    for tag, n, d, days in (("reference_tranches_1440x1", 1440, 1, 30), ("scaled_1Mx128", 1_000_000, 128, 10)):
        tranches = []
        for day in range(days):
            alpha = 1.0 + 0.5 * np.sin(2.0 * np.pi * 6.0 * day / 364.0)       # stage_3...:33,38 intercept drift
            if d == 1:                                   # stage_3...:28-43 on the device: alpha(day), y >= 0 filter
                Xd, yd, kept = ctx.synth_tranche(n, day + 1, seed=900 + day)
                tranches.append((Xd.to_host()[:kept].copy(), yd.to_host()[:kept].copy()))
                Xd.free(); yd.free()
            else:
                Xd, yd = ctx.synth(n, d, seed=900 + day, alpha=alpha)
                tranches.append((Xd.to_host(), yd.to_host()))
                Xd.free(); yd.free()
        incremental.replay(tranches[:2], d, ctx=ctx)                      # warm-up
        days_res = incremental.replay(tranches, d, mode="incremental", ctx=ctx)
        secs = [r.seconds for r in days_res[1:]]
        res[tag] = {"days": days, "rows_per_day": n, "d": d, "median_day_seconds": float(np.median(secs)),
                    "rows_per_s_per_day": n / float(np.median(secs)),
                    "last_day_test_r2": days_res[-1].test_r2, "last_day_test_mape": days_res[-1].test_mape,
                    "coef_head_last": [float(c) for c in days_res[-1].coef[:2]],
                    "note": "day = score tranche t with model(t-1) + fold tranche t into S + re-solve; host rows "
                            "(H2D included)"}
    out["config5_replay"] = res
    print("c5", res, flush=True)


def refshape(ctx):
    """The reference's own shape (one feature, stage_1_train_model.py:95) at 1 B resident rows: train_model's two passes --
    masked fit on the 80 % train rows, hold-out metrics on the other 20 % -- through the narrow kernels."""
    n, d = 1_000_000_000, 1
    X, y = ctx.synth(n, d, seed=4242)
    mask = ctx.empty((n,), "u8")
    blk = 100_000_000
    pattern = (np.arange(blk, dtype=np.int64) % 5 != 0).astype(np.uint8)          # 80 / 20 membership
    lib = b2.native.load()
    for lo in range(0, n, blk):
        b2.native._check(lib.b2_copy_h2d(ctx._h, mask.ptr + lo, pattern.ctypes.data, blk), "h2d")
    ctx.set_kernel(b2.KERNEL_AUTO)
    est = b2.B200LinearRegression(ctx=ctx)
    best_fit, best_score = 1e9, 1e9
    for _ in range(4):
        ctx.sync(); ctx.timer_start()
        est.fit(X, y, row_mask=mask, mask_keep=1, with_spectrum=False)
        best_fit = min(best_fit, ctx.timer_stop())
        ctx.sync(); ctx.timer_start()
        _, stats = ctx.score(X, est.coef_, float(est.intercept_), y=y, row_mask=mask, mask_keep=0, want_yhat=False)
        best_score = min(best_score, ctx.timer_stop())
    from bodywork_mlops_demo_b200 import stage_1_train_model as s1
    mape, r2, mx = s1.metrics_from_stats(stats)
    r = {"n": n, "d": d, "x": "f32 resident", "fit_ms": best_fit, "score_ms": best_score,
         "fit_rows_per_s": n / best_fit * 1e3, "score_rows_per_s": n / best_score * 1e3,
         "fit_hbm_gbs": n * 9 / best_fit / 1e6, "score_hbm_gbs": n * 9 / best_score / 1e6,
         "bytes_per_row": "4 (x) + 4 (y) + 1 (mask) per pass", "coef": float(est.coef_[0]),
         "intercept": float(est.intercept_), "test_rows": float(stats[5]), "MAPE": mape, "r_squared": r2,
         "max_residual": mx}
    X.free(); y.free(); mask.free()
    out["reference_shape_1Bx1_train_model"] = r
    print("refshape", r, flush=True)


def realmask(ctx):
    """The reference's shape with the REAL train_test_split(random_state=42) membership at 200 M rows: the host-side
    MT19937 + Fisher-Yates shuffle (b2_split_mask) next to the two GPU passes it feeds."""
    from bodywork_mlops_demo_b200 import stage_1_train_model as s1
    n, d = 200_000_000, 1
    X, y = ctx.synth(n, d, seed=4242)
    ctx.sync()
    t0 = time.perf_counter()
    job = s1.split_mask_async(n)
    mask_h = job.result()
    t_mask = time.perf_counter() - t0
    t0 = time.perf_counter()
    mask = ctx.to_device(mask_h)
    t_h2d = time.perf_counter() - t0
    ctx.set_kernel(b2.KERNEL_AUTO)
    est = b2.B200LinearRegression(ctx=ctx)
    best_fit, best_score = 1e9, 1e9
    for _ in range(3):
        ctx.sync(); ctx.timer_start()
        est.fit(X, y, row_mask=mask, mask_keep=1, with_spectrum=False)
        best_fit = min(best_fit, ctx.timer_stop())
        ctx.sync(); ctx.timer_start()
        _, stats = ctx.score(X, est.coef_, float(est.intercept_), y=y, row_mask=mask, mask_keep=0, want_yhat=False)
        best_score = min(best_score, ctx.timer_stop())
    r = {"n": n, "d": d, "split_mask_host_seconds": t_mask, "mask_h2d_seconds": t_h2d, "fit_ms": best_fit,
         "score_ms": best_score, "train_rows": float(mask_h.sum()), "test_rows": float(stats[5]),
         "host_ns_per_row": t_mask / n * 1e9,
         "note": "the split is sequential host work by definition (bit-exact numpy legacy generator); the two GPU "
                 "passes over the same rows take milliseconds"}
    X.free(); y.free(); mask.free()
    out["reference_shape_200Mx1_real_split"] = r
    print("realmask", r, flush=True)


def shapes(ctx):
    res = []
    for d in (1, 8, 16, 32, 64, 128):
        for kind in ("f32", "bf16"):
            n = 200_000_000 if d == 1 else 40_000_000 if d <= 32 else 10_000_000
            X, y = ctx.synth(n, d, seed=5, kind=kind)
            ctx.sync()
            r = gram_rate(ctx, X, y, n, d, kind)
            res.append(r)
            print("shape", r, flush=True)
            X.free(); y.free()
    out["gram_kernel_shapes"] = res


if __name__ == "__main__":
    which = sys.argv[1:] or ["c3", "c4", "c5", "shapes", "refshape"]
    ctx = b2.Context(0)
    for w in which:
        {"c3": c3, "c4": c4, "c5": c5, "shapes": shapes, "refshape": refshape, "realmask": realmask}[w](ctx)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    path = os.path.join(ROOT, "gpurun_out", "r02_configs.json")
    prev = json.load(open(path)) if os.path.exists(path) else {}
    prev.update(out)
    json.dump(prev, open(path, "w"), indent=1)
    print("wrote", path)
