#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
for d in 128 32; do B2_LIB_PATH=tools/bin/libb2gram_dev.so B2_SOLVE_TIMING=1 timeout 120 python tools/prof_target.py solve $d 2>&1 | grep -E "eigvals" | tail -2; done
echo "== score"; timeout 120 python tools/prof_target.py score 400000000 1; timeout 120 python tools/prof_target.py score 200000000 2; timeout 120 python tools/prof_target.py score 100000000 4; timeout 120 python tools/prof_target.py score 100000000 8; timeout 120 python tools/prof_target.py score 50000000 16
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider 2>&1 | tail -6
