"""Tiny drivers for ncu / dev timing (development tool):  python tools/prof_target.py <what> [args]
  fit n d kind prec      : 3 fused fits (b2_fit) of n x d synthetic rows (kind f32|bf16, prec split|bf16)
  solve d                : 5 Cholesky (LDL^T) solves, 3 eigenvalue solves, 1 Jacobi solve of a d-feature statistic, timed
  score n d              : 3 scoring passes (metrics only) over n x d rows
  drain                  : drain-interval sweep of the tensor-core kernel at 12.5 M x 128: kernel ms and coefficient error
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bodywork_mlops_demo_b200 as b2

what = sys.argv[1]
ctx = b2.Context(0)
if what == "fit":
    n, d, kind, prec = int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5]
    X, y = ctx.synth(n, d, kind=kind)
    ctx.set_kernel(b2.KERNEL_TCGEN05)
    ctx.set_precision(b2.PRECISION_BF16 if prec == "bf16" else b2.PRECISION_SPLIT)
    for _ in range(3):
        coef, b0 = ctx.fit(X, y)
    print(coef[:3], b0)
elif what == "solve":
    d = int(sys.argv[2])
    X, y = ctx.synth(200_000, d)
    ctx.set_kernel(b2.KERNEL_SIMT if d % 4 else b2.KERNEL_AUTO)
    ctx.gram_reset(d); ctx.gram_accumulate(X, y); ctx.sync()
    for name, fn, reps in (("cholesky", lambda: ctx.solve(), 5), ("eigvals", lambda: ctx.solve_eigvals(), 3),
                           ("spectral_jacobi", lambda: ctx.solve_spectral(), 1)):
        best = 1e9
        for _ in range(reps):
            ctx.sync(); ctx.timer_start(); r = fn(); best = min(best, ctx.timer_stop())
        print(f"d={d} {name}: {best * 1e3:.1f} us (CUDA events around the call incl. the result fetch)")
elif what == "score":
    n, d = int(sys.argv[2]), int(sys.argv[3])
    X, y = ctx.synth(n, d)
    coef = np.full(d, 0.5)
    for _ in range(3):
        ctx.sync(); ctx.timer_start()
        _, st = ctx.score(X, coef, 1.0, y=y, want_yhat=False)
        ms = ctx.timer_stop()
    print(f"score n={n} d={d}: {ms:.3f} ms {n / ms / 1e6:.1f} G rows/s  {n * (4 * d + 4) / ms / 1e6 / 6575.1:.3f} of HBM peak")
elif what == "drain":
    n, d = 12_500_000, 128
    X, y = ctx.synth(n, d)
    ctx.set_kernel(b2.KERNEL_SIMT)
    ctx.gram_reset(d); ctx.gram_accumulate(X, y); c_ex, b_ex = ctx.solve()
    ctx.set_kernel(b2.KERNEL_TCGEN05)
    for drain in (8192, 4096, 2048, 1024, 512):
        ctx.set_drain_rows(drain)
        for _ in range(3):
            ctx.fit(X, y)
        ctx.last_kernel_ms()
        ctx.sync(); ctx.timer_start()
        for _ in range(10):
            coef, b0 = ctx.fit(X, y)
        ms = ctx.timer_stop() / 10
        kms, kl = ctx.last_kernel_ms()
        print(f"drain_rows={drain}: fit {ms:.4f} ms, gram kernel {kms / kl:.4f} ms ({n * 516 / (kms / kl) / 1e6 / 6575.1:.3f} of peak), "
              f"coef linf vs exact {np.max(np.abs(coef - c_ex)):.2e}, intercept err {abs(b0 - b_ex):.2e}", flush=True)
ctx.close()
