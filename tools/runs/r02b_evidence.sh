#!/bin/bash
# second evidence run of round 2: sanitizer (incl. the b16 kernel), bench line N = 1, launch list
mkdir -p gpurun_out; O=gpurun_out
bash tools/runs/sanitizer.sh > /dev/null 2>&1; grep -c "ERROR SUMMARY: 0 errors" $O/r02_compute_sanitizer.txt; grep -E "==|SUMMARY" $O/r02_compute_sanitizer.txt | tail -24
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 16 --csv --log-file $O/r02_launches_bench_n1.csv python bench.py --steps 4 --warmup 3 --no-extras --no-e2e --no-cpu-baseline > $O/ncu7.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_n1.json").read().strip().splitlines()[-1])
print("N=1 value %.3e ms/step %.4f kernel %.4f frac %.3f tail_us %.1f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_avg"], d["roofline"]["frac"], d["roofline"]["step_tail_us"]))
for k,v in d["config1_10Mx128"].items():
    if isinstance(v, dict): print(k, "%.3f ms  frac %.3f  coef_linf %.2e" % (v["gram_kernel_ms"], v["frac_of_hbm_peak"], v["coef_linf_vs_exact"]))
print("north_star", json.dumps(d["north_star"])[:600])
print("e2e", json.dumps(d["e2e"])[:500]); print("clocks", d.get("clocks"))
PY
