"""Tiny driver for ncu: one Gram accumulation of n x d synthetic rows (development tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bodywork_mlops_demo_b200 as b2
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 128
kind = sys.argv[3] if len(sys.argv) > 3 else "f32"
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
ctx = b2.Context(0)
X, y = ctx.synth(n, d, kind=kind)
kernel = {"auto": b2.KERNEL_AUTO, "tc": b2.KERNEL_TCGEN05, "narrow": b2.KERNEL_NARROW, "simt": b2.KERNEL_SIMT}[
    sys.argv[5] if len(sys.argv) > 5 else ("narrow" if d <= 16 else "tc")]
ctx.set_kernel(kernel)
for _ in range(reps):
    ctx.gram_reset(d)
    ctx.gram_accumulate(X, y)
ctx.sync()
print(ctx.solve()[0][:3])
