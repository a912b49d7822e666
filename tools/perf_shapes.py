"""Gram-kernel rate through the AUTO dispatch for a list of feature counts (development / evidence tool):
    python tools/perf_shapes.py [d ...]   -> one JSON line per (d, storage)"""
from __future__ import annotations

import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bodywork_mlops_demo_b200 as b2  # noqa: E402

PEAK = 6575.1


def main():
    dims = [int(a) for a in sys.argv[1:]] or [20, 24, 28, 32, 36, 40, 48, 56, 64, 72, 96, 100, 128]
    ctx = b2.Context(0)
    out = []
    for d in dims:
        for kind in ("f32", "bf16"):
            es = 4 if kind == "f32" else 2
            if (d * es) % 16:
                continue
            bpr = d * es + 4
            n = int(2.0e9 // bpr)
            X, y = ctx.synth(n, d, seed=5, kind=kind)
            ctx.sync()
            best = 1e9
            for _ in range(5):
                ctx.gram_reset(d)
                ctx.gram_accumulate(X, y)
                ctx.sync()
                k, _n = ctx.last_kernel_ms()
                best = min(best, k)
            coef, b0 = ctx.solve()
            rec = {"d": d, "x": kind, "n": n, "bytes_per_row": bpr, "kernel_ms": best, "g_rows_per_s": n / best / 1e6,
                   "frac_of_measured_peak": n * bpr / best / 1e6 / PEAK, "coef0": float(coef[0]), "intercept": float(b0)}
            X.free(); y.free()
            print(json.dumps(rec), flush=True)
            out.append(rec)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r01_shapes.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
