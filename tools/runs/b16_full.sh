#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
for P in split bf16; do
  TAG="new-$P" B2_PRECISION=$P timeout 120 python tools/perf_quick.py 10000000 128 bf16
done 2>&1 | tee $O/b16_perf2.txt
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 600 -p no:cacheprovider 2>&1 | tail -6
