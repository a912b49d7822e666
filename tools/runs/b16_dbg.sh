#!/bin/bash
mkdir -p gpurun_out
(
for D in 0 8 16 24 31; do
  TAG="split-dbg$D" B2_B16_DBG=$D B2_PRECISION=split B2_TC_DEBUG=1 timeout 120 python tools/perf_quick.py 10000000 128 bf16
done
) 2>&1 | tee gpurun_out/b16_dbg9.txt
