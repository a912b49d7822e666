#!/bin/bash
# ncu --set full captures of the bf16-storage D = 128 kernel, both operand modes
mkdir -p gpurun_out; O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:gram_b16_ -s 2 -c 1 -f -o $O/r02_gram_b16_10Mx128_split python tools/prof_target.py fit 10000000 128 bf16 split > $O/ncu_b16a.log 2>&1; tail -1 $O/ncu_b16a.log
timeout 300 $NCU -k regex:gram_b16_ -s 2 -c 1 -f -o $O/r02_gram_b16_10Mx128_single python tools/prof_target.py fit 10000000 128 bf16 bf16 > $O/ncu_b16b.log 2>&1; tail -1 $O/ncu_b16b.log
ls -la $O/*.ncu-rep | awk '{print $5, $9}'
