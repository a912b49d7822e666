#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_b16.py -m gpu -x -q --timeout 300 -p no:cacheprovider 2>&1 | tail -4
for P in split bf16; do
  TAG="new-$P" B2_PRECISION=$P timeout 120 python tools/perf_quick.py 10000000 128 bf16
done 2>&1 | tee $O/b16_perf.txt
TAG="old-split" B2_TC_B16_GENERIC=1 B2_PRECISION=split timeout 120 python tools/perf_quick.py 10000000 128 bf16
for T in 8 16 24 32; do B2_COPY_THREADS=$T timeout 200 python tools/perf_pageable.py 4000000 128; done 2>&1 | tee $O/pageable_threads.txt
