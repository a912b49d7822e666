#!/bin/bash
for rep in 1 2; do
for n in 10000000; do
  (cd tools/bin/r01 && B2_NO_REBUILD=1 TAG=r01 timeout 120 python tools/perf_quick.py $n 128 f32)
  B2_NO_REBUILD=1 TAG=now timeout 120 python tools/perf_quick.py $n 128 f32
  B2_LIB_PATH=tools/bin/libb2gram_fusedoff.so TAG=fusedoff timeout 120 python tools/perf_quick.py $n 128 f32
done; done
