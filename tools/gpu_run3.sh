#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
echo "== solve phases"; B2_LIB_PATH=tools/bin/libb2gram_dev.so B2_SOLVE_TIMING=1 timeout 120 python tools/prof_target.py solve 128 2>&1 | grep -v "^\[b2_solve\]" | tail -3; B2_LIB_PATH=tools/bin/libb2gram_dev.so B2_SOLVE_TIMING=1 timeout 120 python tools/prof_target.py solve 128 2>&1 | grep "^\[b2_solve\]" | tail -1
timeout 300 $NCU -k regex:gram_tc_kernel -s 2 -c 1 -f -o $O/r02_gram_fused_12M5x128_f32 python tools/prof_target.py fit 12500000 128 f32 split > $O/ncu1.log 2>&1; tail -2 $O/ncu1.log
timeout 300 $NCU -k regex:gram_tc_kernel -s 2 -c 1 -f -o $O/r02_gram_fused_10Mx128_bf16_split python tools/prof_target.py fit 10000000 128 bf16 split > $O/ncu2.log 2>&1; tail -2 $O/ncu2.log
timeout 300 $NCU -k regex:gram_tc_kernel -s 2 -c 1 -f -o $O/r02_gram_fused_10Mx128_bf16_single python tools/prof_target.py fit 10000000 128 bf16 bf16 > $O/ncu3.log 2>&1; tail -2 $O/ncu3.log
timeout 300 $NCU -k regex:solve_cholesky_kernel -s 2 -c 1 -f -o $O/r02_solve_ldlt_d128 python tools/prof_target.py solve 128 > $O/ncu4.log 2>&1; tail -2 $O/ncu4.log
timeout 300 $NCU -k regex:solve_eigvals_kernel -s 1 -c 1 -f -o $O/r02_solve_eigvals_d128 python tools/prof_target.py solve 128 > $O/ncu5.log 2>&1; tail -2 $O/ncu5.log
timeout 300 $NCU -k regex:score_narrow_kernel -s 1 -c 1 -f -o $O/r02_score_narrow_400Mx1_f32 python tools/prof_target.py score 400000000 1 > $O/ncu6.log 2>&1; tail -2 $O/ncu6.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 24 --csv --log-file $O/r02_launches_bench_n1.csv python bench.py --steps 4 --warmup 3 --no-extras --no-e2e --no-cpu-baseline > $O/ncu7.log 2>&1; tail -c 300 $O/ncu7.log
timeout 120 python tools/prof_target.py score 400000000 1
ls -la $O/*.ncu-rep
