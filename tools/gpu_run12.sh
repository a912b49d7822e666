#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
for d in 128 32; do B2_LIB_PATH=tools/bin/libb2gram_dev.so B2_SOLVE_TIMING=1 timeout 120 python tools/prof_target.py solve $d 2>&1 | grep -E "eigvals|cholesky" | tail -3; done
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider 2>&1 | tail -6
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_n1.json").read().strip().splitlines()[-1])
print("N=1 value %.3e ms/step %.4f kernel %.4f frac %.3f tail_us %.1f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_avg"], d["roofline"]["frac"], d["roofline"]["step_tail_us"]))
print("e2e", json.dumps(d["e2e"])[:1200]); print("north", json.dumps(d["north_star"])[:700])
PY
