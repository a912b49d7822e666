#!/bin/bash
# same-box A/B of the fp32 headline kernel: tools/bin/libb2gram_base.so (before) vs the in-tree library (after)
mkdir -p gpurun_out
(
for r in 1 2; do
  B2_LIB_PATH=tools/bin/libb2gram_base.so TAG=base timeout 120 python tools/perf_quick.py 12500000 128 f32
  B2_NO_REBUILD=1 TAG=new timeout 120 python tools/perf_quick.py 12500000 128 f32
done
B2_LIB_PATH=tools/bin/libb2gram_base.so TAG=base timeout 120 python tools/perf_quick.py 40000000 32 f32
B2_NO_REBUILD=1 TAG=new timeout 120 python tools/perf_quick.py 40000000 32 f32
B2_LIB_PATH=tools/bin/libb2gram_base.so TAG=base B2_PRECISION=bf16 timeout 120 python tools/perf_quick.py 12500000 128 f32
B2_NO_REBUILD=1 TAG=new B2_PRECISION=bf16 timeout 120 python tools/perf_quick.py 12500000 128 f32
nvidia-smi --query-gpu=clocks.sm,power.draw,temperature.gpu --format=csv,noheader
) 2>&1 | tee gpurun_out/ab_f32.txt
