#!/bin/bash
# the bf16-storage D = 128 kernel: parity tests, then the kernel rate next to the generic kernel on the same box
mkdir -p gpurun_out; O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_b16.py -m gpu -x -q --timeout 300 -p no:cacheprovider 2>&1 | tail -15
for P in split bf16; do
  TAG="new-$P" B2_PRECISION=$P timeout 120 python tools/perf_quick.py 10000000 128 bf16
  TAG="old-$P" B2_TC_B16_GENERIC=1 B2_PRECISION=$P timeout 120 python tools/perf_quick.py 10000000 128 bf16
done 2>&1 | tee $O/b16_perf.txt
