#!/bin/bash
# gpurun -- "bash tools/runs/evidence_run.sh"
# evidence run: ncu --set full captures of the final kernels (1 GPU), launch list of the bench step, bench line
mkdir -p gpurun_out; O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:gram_tc_kernel -s 2 -c 1 -f -o $O/r02_gram_tc_12M5x128_f32 python tools/prof_target.py fit 12500000 128 f32 split > $O/ncu1.log 2>&1; tail -1 $O/ncu1.log
timeout 300 $NCU -k regex:tc_finalize_kernel -s 2 -c 1 -f -o $O/r02_tc_finalize python tools/prof_target.py fit 12500000 128 f32 split > $O/ncu1b.log 2>&1; tail -1 $O/ncu1b.log
timeout 300 $NCU -k regex:gram_tc_kernel -s 2 -c 1 -f -o $O/r02_gram_tc_10Mx128_bf16_split python tools/prof_target.py fit 10000000 128 bf16 split > $O/ncu2.log 2>&1; tail -1 $O/ncu2.log
timeout 300 $NCU -k regex:gram_tc_kernel -s 2 -c 1 -f -o $O/r02_gram_tc_10Mx128_bf16_single python tools/prof_target.py fit 10000000 128 bf16 bf16 > $O/ncu3.log 2>&1; tail -1 $O/ncu3.log
timeout 300 $NCU -k regex:solve_cholesky_kernel -s 2 -c 1 -f -o $O/r02_solve_ldlt_d128 python tools/prof_target.py solve 128 > $O/ncu4.log 2>&1; tail -1 $O/ncu4.log
timeout 300 $NCU -k regex:solve_eigvals_kernel -s 1 -c 1 -f -o $O/r02_solve_eigvals_d128 python tools/prof_target.py solve 128 > $O/ncu5.log 2>&1; tail -1 $O/ncu5.log
timeout 300 $NCU -k regex:score_narrow_kernel -s 1 -c 1 -f -o $O/r02_score_narrow_400Mx1_f32 python tools/prof_target.py score 400000000 1 > $O/ncu6.log 2>&1; tail -1 $O/ncu6.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 16 --csv --log-file $O/r02_launches_bench_n1.csv python bench.py --steps 4 --warmup 3 --no-extras --no-e2e --no-cpu-baseline > $O/ncu7.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_reference.json 2>/dev/null
python tools/config_runs.py c4 c5 refshape realmask shapes > $O/r02_configs.log 2>&1; tail -3 $O/r02_configs.log
ls -la $O/*.ncu-rep | awk '{print $5, $9}'
