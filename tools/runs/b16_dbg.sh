#!/bin/bash
mkdir -p gpurun_out
(
TAG="old-split" B2_TC_B16_GENERIC=1 B2_PRECISION=split timeout 120 python tools/perf_quick.py 10000000 128 bf16
for D in 0 4 8 12 3 7 15; do
  TAG="split-dbg$D" B2_B16_DBG=$D B2_PRECISION=split B2_TC_DEBUG=1 timeout 120 python tools/perf_quick.py 10000000 128 bf16
done
for D in 0 4 12; do
TAG="single-dbg$D" B2_B16_DBG=$D B2_PRECISION=bf16 B2_TC_DEBUG=1 timeout 120 python tools/perf_quick.py 10000000 128 bf16
done
) 2>&1 | tee gpurun_out/b16_dbg5.txt
