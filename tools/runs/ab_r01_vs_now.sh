#!/bin/bash
# Same-box A/B of the round-1 Gram kernel against the current one.  The round-1 tree is rebuilt into tools/bin/r01 first:
#   git worktree add /tmp/r01 e57a35a && (cd /tmp/r01/bodywork-mlops-demo_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a \
#     -O3 -std=c++17 -Xcompiler -fPIC -shared -o ../libb2gram.so *.cu -ldl) && mkdir -p tools/bin/r01 && \
#   cp -r /tmp/r01/{bodywork-mlops-demo_b200,bodywork_mlops_demo_b200,include,tools} tools/bin/r01/
# same-box A/B of the round-1 Gram kernel vs the current one (stand-alone accumulate path), 10 M and 12.5 M x 128 fp32
for n in 10000000 12500000; do
  (cd tools/bin/r01 && B2_NO_REBUILD=1 TAG=r01 timeout 120 python tools/perf_quick.py $n 128 f32)
  B2_NO_REBUILD=1 TAG=now timeout 120 python tools/perf_quick.py $n 128 f32
done
(cd tools/bin/r01 && B2_NO_REBUILD=1 TAG=r01 timeout 120 python tools/perf_quick.py 10000000 128 bf16)
B2_NO_REBUILD=1 TAG=now timeout 120 python tools/perf_quick.py 10000000 128 bf16
nvidia-smi --query-gpu=name,clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.active --format=csv
