#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:gram_b16_split -s 2 -c 1 -f -o $O/r02_gram_b16_10Mx128_split python tools/prof_target.py fit 10000000 128 bf16 split > $O/ncu_b16a.log 2>&1; tail -1 $O/ncu_b16a.log
