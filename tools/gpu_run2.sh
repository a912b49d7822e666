#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 60 tools/bin/ubench_fp64 > $O/r2_ubench_fp64.txt 2>&1; cat $O/r2_ubench_fp64.txt
echo "== solve phases (dev build)"; B2_LIB_PATH=tools/bin/libb2gram_dev.so B2_SOLVE_TIMING=1 timeout 120 python tools/prof_target.py solve 128 2>&1 | tail -12 | tee $O/r2_solve_phases.txt
timeout 120 python tools/prof_target.py solve 128 2>&1 | tail -3 | tee -a $O/r2_solve_phases.txt
timeout 120 python tools/prof_target.py solve 32 2>&1 | tail -3 | tee -a $O/r2_solve_phases.txt
echo "== drain sweep"; timeout 300 python tools/prof_target.py drain 2>&1 | tee $O/r2_drain_sweep.txt
echo "== A/B fused vs unfused"
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-e2e --no-cpu-baseline > $O/r2_ab_fused.json 2>$O/r2_ab.err
B2_NO_FUSED=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-e2e --no-cpu-baseline > $O/r2_ab_unfused.json 2>>$O/r2_ab.err
python - <<'PY'
import json
for t in ("fused","unfused"):
    d=json.load(open(f"gpurun_out/r2_ab_{t}.json"))
    print(t, "ms/step %.4f kernel %.4f frac %.3f tail_us %.1f launches %d coef_linf %.2e" % (d["ms_per_step"], d["roofline"]["kernel_ms_avg"], d["roofline"]["frac"], d["roofline"]["step_tail_us"], d["gpu_launches"], d["parity"]["coef_linf"]))
PY
echo "== pytest fused + parity subset"; timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity.py -m gpu -q --timeout 300 -p no:cacheprovider -x 2>&1 | tail -4
