"""Quick timing of the Gram kernel on synthetic rows: python tools/perf_quick.py [n] [d] [kind]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bodywork_mlops_demo_b200 as b2
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 128
kind = sys.argv[3] if len(sys.argv) > 3 else "f32"
ctx = b2.Context(0)
X, y = ctx.synth(n, d, kind=kind)
ctx.set_kernel(b2.KERNEL_TCGEN05)
if os.environ.get("B2_PRECISION") == "bf16":
    ctx.set_precision(b2.PRECISION_BF16)
best = 1e9
for _ in range(8):
    ctx.gram_reset(d); ctx.gram_accumulate(X, y); ctx.sync()
    ms, _n = ctx.last_kernel_ms(); best = min(best, ms)
bpr = d * (4 if kind == "f32" else 2) + 4
print(f"{os.environ.get('TAG','')} n={n} d={d} {kind}: gram {best:.3f} ms  {n/best/1e6:.2f} G rows/s  {n*bpr/best/1e6/6575.1:.3f} of HBM peak; " + ("" if os.environ.get('B2_TC_DEBUG','0')!='0' else f"coef0 {ctx.solve()[0][0]:.6f}"))
