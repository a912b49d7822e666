"""Small driver for compute-sanitizer (memcheck / racecheck): every kernel once on small inputs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bodywork_mlops_demo_b200 as b2
from oracle import ols_oracle as orc
which = sys.argv[1] if len(sys.argv) > 1 else "all"
ctx = b2.Context(0)
X, y = orc.generate_dataset(4096 + 37, 128, seed=1, dtype=np.float32)
mask = (np.arange(len(y)) % 5 != 0).astype(np.uint8)
Xd, yd, md = ctx.to_device(X), ctx.to_device(y), ctx.to_device(mask)
if which in ("all", "simt"):
    ctx.set_kernel(b2.KERNEL_SIMT); ctx.gram_reset(128); ctx.gram_accumulate(Xd, yd, md, 1)
    print("simt n", ctx.gram_export()[128, 128])
if which in ("all", "tc"):
    ctx.set_kernel(b2.KERNEL_TCGEN05); ctx.set_drain_rows(1024); ctx.gram_reset(128); ctx.gram_accumulate(Xd, yd, md, 1)
    print("tc n", ctx.gram_export()[128, 128])
if which in ("all", "solve"):
    ctx.gram_import(orc.gram_stats(X, y))
    c, b = ctx.solve(); c2, b2_, s, r = ctx.solve_spectral()
    print("solve", c[:2], r)
    X3, y3 = orc.generate_dataset(500, 33, seed=2)
    ctx.gram_import(orc.gram_stats(X3, y3)); print("solve33", ctx.solve()[0][:2], ctx.solve_spectral()[3])
if which in ("all", "score"):
    yh, st = ctx.score(Xd, np.full(128, 0.5), 1.0, y=yd, row_mask=md, mask_keep=0)
    print("score", st[5])
    Xs, ys = ctx.synth(1000, 128, seed=3); print("synth", float(Xs.to_host().mean()))
print("done")
