"""Gram-kernel rate for narrow rows (development / evidence tool):  python tools/perf_narrow.py [d ...]"""
from __future__ import annotations

import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bodywork_mlops_demo_b200 as b2  # noqa: E402

PEAK = 6575.1


def main():
    dims = [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8, 12, 16]
    ctx = b2.Context(0)
    out = []
    for d in dims:
        for kind in ("f32", "bf16"):
            es = 4 if kind == "f32" else 2
            bpr = d * es + 4
            n = min(400_000_000, int(2.4e9 // bpr))
            X, y = ctx.synth(n, d, seed=5, kind=kind)
            ctx.sync()
            rec = {"d": d, "x": kind, "n": n, "bytes_per_row": bpr}
            for name, kernel in (("narrow", b2.KERNEL_NARROW), ("tcgen05", b2.KERNEL_TCGEN05)):
                if kernel == b2.KERNEL_TCGEN05 and (d * es) % 16:
                    continue
                ctx.set_kernel(kernel)
                best = 1e9
                for _ in range(5):
                    ctx.gram_reset(d)
                    ctx.gram_accumulate(X, y)
                    ctx.sync()
                    k, _n = ctx.last_kernel_ms()
                    best = min(best, k)
                coef, b0 = ctx.solve()
                rec[name] = {"kernel_ms": best, "g_rows_per_s": n / best / 1e6, "gb_per_s": n * bpr / best / 1e6,
                             "frac_of_measured_peak": n * bpr / best / 1e6 / PEAK, "coef0": float(coef[0]),
                             "intercept": float(b0)}
            ctx.set_kernel(b2.KERNEL_AUTO)
            X.free(); y.free()
            print(json.dumps(rec), flush=True)
            out.append(rec)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r01_narrow.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
