#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
bash tools/gpu_run4.sh 2>&1 | grep -E "^r01|^now" | tee $O/r02_ab_r01_vs_now.txt
echo "== solve phases"; B2_LIB_PATH=tools/bin/libb2gram_dev.so B2_SOLVE_TIMING=1 timeout 120 python tools/prof_target.py solve 128 2>&1 | tail -4 | tee $O/r02_solve_phases.txt
B2_LIB_PATH=tools/bin/libb2gram_dev.so B2_SOLVE_TIMING=1 timeout 120 python tools/prof_target.py solve 100 2>&1 | tail -4 | tee -a $O/r02_solve_phases.txt
timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity.py -m gpu -q --timeout 300 -p no:cacheprovider 2>&1 | tail -8
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-e2e --no-cpu-baseline > $O/r2_ab_fused.json 2>$O/r2_ab.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_ab_fused.json"))
print("fused ms/step %.4f kernel %.4f frac %.3f tail_us %.1f launches %d coef_linf %.2e" % (d["ms_per_step"], d["roofline"]["kernel_ms_avg"], d["roofline"]["frac"], d["roofline"]["step_tail_us"], d["gpu_launches"], d["parity"]["coef_linf"]))
PY
