#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
(cd tools/bin/r01 && B2_NO_REBUILD=1 TAG=r01 timeout 120 python tools/perf_quick.py 10000000 128 f32)
B2_NO_REBUILD=1 TAG=now timeout 120 python tools/perf_quick.py 10000000 128 f32
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider 2>&1 | tail -12
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-e2e --no-cpu-baseline > $O/r2_ab_fused.json 2>$O/r2_ab.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_ab_fused.json"))
print("fit ms/step %.4f kernel %.4f frac %.3f tail_us %.1f launches %d coef_linf %.2e" % (d["ms_per_step"], d["roofline"]["kernel_ms_avg"], d["roofline"]["frac"], d["roofline"]["step_tail_us"], d["gpu_launches"], d["parity"]["coef_linf"]))
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 16 --csv --log-file $O/r02_launches_bench_n1.csv python bench.py --steps 4 --warmup 3 --no-extras --no-e2e --no-cpu-baseline > $O/ncu7.log 2>&1
grep -E "tc_|gram_tc|solve" $O/r02_launches_bench_n1.csv | awk -F'","' '{print $5, $NF}' | head -12
