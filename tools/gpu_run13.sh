#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
for d in 128 100 32; do B2_LIB_PATH=tools/bin/libb2gram_dev.so B2_SOLVE_TIMING=1 timeout 120 python tools/prof_target.py solve $d 2>&1 | grep -E "eigvals" | tail -2; done
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity.py -m gpu -q --timeout 300 -p no:cacheprovider 2>&1 | tail -6
timeout 300 ncu --set full --clock-control none --import-source on -k regex:score_narrow_kernel -s 1 -c 1 -f -o $O/r02_score_narrow_400Mx1_f32_v2 python tools/prof_target.py score 400000000 1 > $O/ncu6.log 2>&1; tail -1 $O/ncu6.log
