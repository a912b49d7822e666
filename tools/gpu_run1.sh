#!/bin/bash
# round-2 bring-up: full GPU suite, fp64 latency microbenchmark, 1-GPU bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --maxfail=25 --timeout 300 -p no:cacheprovider > gpurun_out/r2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
tail -5 gpurun_out/r2_pytest.log
timeout 60 tools/bin/ubench_fp64 > gpurun_out/r2_ubench_fp64.txt 2>&1
cat gpurun_out/r2_ubench_fp64.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
echo "bench rc=$?"
tail -c 1500 gpurun_out/r2_bench_n1.err
head -c 600 gpurun_out/r2_bench_n1.json
