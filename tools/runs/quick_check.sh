#!/bin/bash
# tests of the Gram / fit paths + a short bench line + the launch list (after a change to the shift / finalize / solve kernels)
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity.py -m gpu -x -q --timeout 300 -p no:cacheprovider 2>&1 | tail -4
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-e2e --no-cpu-baseline > $O/quick.json 2>$O/quick.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/quick.json"))
print("fit ms/step %.4f kernel %.4f frac %.3f tail_us %.1f launches %d coef_linf %.2e" % (d["ms_per_step"], d["roofline"]["kernel_ms_avg"], d["roofline"]["frac"], d["roofline"]["step_tail_us"], d["gpu_launches"], d["parity"]["coef_linf"]))
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 8 --csv --log-file $O/quick_launches.csv python bench.py --steps 4 --warmup 3 --no-extras --no-e2e --no-cpu-baseline > $O/ncu7.log 2>&1
grep -E "tc_|gram_tc|solve" $O/quick_launches.csv | awk -F'","' '{split($5,a,"("); print a[1], $NF}' | head -8
