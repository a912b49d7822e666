"""GPU bring-up checks (development tool; run on the B200 box through gpurun).

    python tests/tools/gpu_check.py [stage ...]      stages: simt tc perf solve score host bf16

Each stage runs in its own subprocess with a timeout so that a trapped kernel (sticky CUDA error)
cannot take the later stages down.  Results are printed and appended to gpurun_out/gpu_check.log.
"""
from __future__ import annotations

import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def _ctx():
    import bodywork_mlops_demo_b200 as b2
    return b2, b2.Context(0)


def _rel(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def stage_simt():
    from oracle import ols_oracle as orc
    b2, ctx = _ctx()
    print("device:", ctx.info())
    ctx.set_kernel(b2.KERNEL_SIMT)
    for n, d in ((1000, 1), (5000, 8), (4096, 128), (3001, 37)):
        X, y = orc.generate_dataset(n, d, seed=n + d, dtype=np.float32)
        ctx.gram_reset(d)
        ctx.gram_accumulate(ctx.to_device(X), ctx.to_device(y))
        S = ctx.gram_export()
        So = orc.gram_stats(X, y)
        print(f"simt n={n} d={d}: rel err vs oracle S = {_rel(S, So):.3e}")


def stage_tc():
    from oracle import ols_oracle as orc
    b2, ctx = _ctx()
    for n, d, drain in ((4096, 128, 8192), (64, 128, 8192), (100_000, 128, 8192), (100_003, 128, 1024),
                        (50_000, 32, 8192), (20_000, 8, 8192), (300_000, 64, 4096)):
        X, y = orc.generate_dataset(n, d, seed=n + d, dtype=np.float32)
        Xd, yd = ctx.to_device(X), ctx.to_device(y)
        ctx.set_drain_rows(drain)
        ctx.set_kernel(b2.KERNEL_TCGEN05)
        ctx.gram_reset(d)
        ctx.gram_accumulate(Xd, yd)
        S = ctx.gram_export()
        So = orc.gram_stats(X, y)
        dS = np.abs(S - So)
        if n <= 2 * d:
            print(f"tc n={n} d={d}: S rel err {_rel(S, So):.3e} (n <= 2d: no solve)")
            continue
        fo = orc.fit_from_stats(So)
        coef, b0 = ctx.solve()
        i, j = np.unravel_index(np.argmax(dS), dS.shape)
        print(f"tc n={n} d={d} drain={drain}: S rel err {_rel(S, So):.3e} (worst [{i},{j}] got {S[i, j]:.6e} "
              f"want {So[i, j]:.6e}) | coef linf {np.max(np.abs(coef - fo['coef'])):.3e} "
              f"intercept err {abs(b0 - fo['intercept']):.3e}")
        if _rel(S, So) > 1e-3:
            np.set_printoptions(linewidth=200, precision=4)
            print("  S[:4,:4] got\n", S[:4, :4], "\n  want\n", So[:4, :4])
            print("  last rows got\n", S[-2:, :6], "\n  want\n", So[-2:, :6])
            print("  ratio diag:", (np.diag(S) / np.diag(So))[:8])
        Xd.free(); yd.free()


def stage_perf():
    b2, ctx = _ctx()
    d = 128
    for kind in ("f32", "bf16"):
        for n in (10_000_000, 40_000_000):
            Xd, yd = ctx.synth(n, d, kind=kind)
            ctx.sync()
            ctx.set_kernel(b2.KERNEL_TCGEN05)
            times = []
            for it in range(6):
                ctx.gram_reset(d)
                ctx.timer_start()
                ctx.gram_accumulate(Xd, yd)
                ms = ctx.timer_stop()
                kms, nl = ctx.last_kernel_ms()
                times.append((ms, kms))
            ms, kms = min(times[1:])
            bpr = d * (4 if kind == "f32" else 2) + 4
            print(f"perf {kind} n={n}: accumulate {ms:.3f} ms (gram kernel {kms:.3f} ms) -> "
                  f"{n / ms / 1e6:.2f} G rows/s, kernel {n * bpr / kms / 1e6:.1f} GB/s "
                  f"= {n * bpr / kms / 1e6 / 6575.1:.3f} of measured HBM peak")
            coef, b0 = ctx.solve()
            print("   coef[:4]", coef[:4], "intercept", b0)
            Xd.free(); yd.free()


def stage_solve():
    from oracle import ols_oracle as orc
    b2, ctx = _ctx()
    for n, d in ((2000, 1), (5000, 8), (20000, 128), (3000, 33)):
        X, y = orc.generate_dataset(n, d, seed=5 * n + d)
        S = orc.gram_stats(X, y)
        ctx.gram_import(S)
        for alpha in (0.0, 10.0):
            coef, b0 = ctx.solve(alpha=alpha)
            fo = orc.fit_from_stats(S, alpha=alpha)
            print(f"solve n={n} d={d} alpha={alpha}: coef linf {np.max(np.abs(coef - fo['coef'])):.3e} "
                  f"intercept {abs(b0 - fo['intercept']):.3e}")
        t0 = time.time()
        c2, b2_, sing, rank = ctx.solve_spectral()
        fo = orc.fit_from_stats(S)
        print(f"spectral d={d}: {1e3 * (time.time() - t0):.2f} ms coef linf {np.max(np.abs(c2 - fo['coef'])):.3e} "
              f"sing rel {_rel(sing, fo['singular']):.3e} rank {rank} vs {fo['rank']}")
    # rank deficient
    X, y = orc.generate_dataset(500, 4, seed=31)
    X = np.concatenate([X, X[:, :1], np.full((500, 1), 7.0)], axis=1)
    ctx.gram_import(orc.gram_stats(X, y))
    try:
        ctx.solve()
        print("rank-deficient: cholesky did NOT flag singular")
    except np.linalg.LinAlgError as e:
        print("rank-deficient: cholesky flagged:", str(e)[:60])
    c2, b2_, sing, rank = ctx.solve_spectral()
    fo = orc.fit_lstsq(X, y)
    print(f"rank-deficient spectral: coef linf {np.max(np.abs(c2 - fo['coef'])):.3e} rank {rank} vs {fo['rank']}")


def stage_score():
    from oracle import ols_oracle as orc
    b2, ctx = _ctx()
    for n, d in ((10_000, 128), (777, 5), (100_000, 32)):
        X, y = orc.generate_dataset(n, d, seed=n, dtype=np.float32)
        coef = np.linspace(0.3, 0.7, d)
        p = orc.predict(X, coef, 1.5)
        mask = (np.arange(n) % 5 == 0).astype(np.uint8)
        yhat, stats = ctx.score(ctx.to_device(X), coef, 1.5, y=ctx.to_device(y), row_mask=ctx.to_device(mask),
                                mask_keep=1)
        so = orc.score_stats(y[mask == 1], p[mask == 1])
        yh = yhat.to_host()
        print(f"score n={n} d={d}: yhat max err {np.max(np.abs(yh[mask == 1] - p[mask == 1])):.3e} "
              f"stats rel {np.max(np.abs(stats - so) / np.maximum(np.abs(so), 1e-300)):.3e}")


def stage_host():
    from oracle import ols_oracle as orc
    b2, ctx = _ctx()
    n, d = 700_000, 128
    X, y = orc.generate_dataset(n, d, seed=9, dtype=np.float32)
    ctx.gram_reset(d)
    t0 = time.time()
    ctx.gram_accumulate(X, y)
    coef, b0 = ctx.solve()
    dt = time.time() - t0
    fo = orc.fit_from_stats(orc.gram_stats(X, y))
    print(f"host-streamed n={n}: {dt * 1e3:.1f} ms coef linf {np.max(np.abs(coef - fo['coef'])):.3e}")


def stage_bf16():
    from oracle import ols_oracle as orc
    b2, ctx = _ctx()
    n, d = 200_000, 128
    X, y = orc.generate_dataset(n, d, seed=19, dtype=np.float32)
    Xb = b2.native.to_bf16_bits(X)
    Xr = b2.native.from_bf16_bits(Xb)
    ctx.gram_reset(d)
    ctx.gram_accumulate(ctx.to_device(Xb, "bf16"), ctx.to_device(y))
    coef, b0 = ctx.solve()
    fo = orc.fit_from_stats(orc.gram_stats(Xr, y))
    print(f"bf16 storage n={n}: coef linf vs fit on the same bf16 rows {np.max(np.abs(coef - fo['coef'])):.3e}")


STAGES = {"simt": stage_simt, "tc": stage_tc, "perf": stage_perf, "solve": stage_solve, "score": stage_score,
          "host": stage_host, "bf16": stage_bf16}

if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "--one":
        STAGES[sys.argv[2]]()
        sys.exit(0)
    names = sys.argv[1:] or list(STAGES)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "gpu_check.log"), "a") as logf:
        for nm in names:
            t0 = time.time()
            try:
                p = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", nm], capture_output=True,
                                   text=True, timeout=300)
                out = p.stdout + ("\n[stderr]\n" + p.stderr[-3000:] if p.returncode != 0 else "")
                rc = p.returncode
            except subprocess.TimeoutExpired as e:
                out, rc = f"TIMEOUT\n{(e.stdout or b'')[-2000:]}", -999
            msg = f"===== stage {nm}: rc={rc} ({time.time() - t0:.1f}s)\n{out}\n"
            print(msg, flush=True)
            logf.write(msg)
