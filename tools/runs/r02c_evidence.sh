#!/bin/bash
# third evidence run of round 2 (final bf16-storage kernels): parity tests, ncu captures, sanitizer, bench line, launch list
mkdir -p gpurun_out; O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_b16.py -m gpu -q --timeout 300 -p no:cacheprovider 2>&1 | tail -4
bash tools/runs/b16_ncu.sh 2>&1 | tail -3
bash tools/runs/sanitizer.sh > /dev/null 2>&1; echo "sanitizer clean runs: $(grep -c -E 'ERROR SUMMARY: 0 errors|0 hazards displayed' $O/r02_compute_sanitizer.txt) of 8"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 16 --csv --log-file $O/r02_launches_bench_n1.csv python bench.py --steps 4 --warmup 3 --no-extras --no-e2e --no-cpu-baseline > $O/ncu7.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_n1.json").read().strip().splitlines()[-1])
print("N=1 value %.3e ms/step %.4f kernel %.4f frac %.3f tail_us %.1f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_avg"], d["roofline"]["frac"], d["roofline"]["step_tail_us"]))
for k,v in d["config1_10Mx128"].items():
    if isinstance(v, dict): print(k, "%.3f ms  frac %.3f  coef_linf %.2e" % (v["gram_kernel_ms"], v["frac_of_hbm_peak"], v["coef_linf_vs_exact"]))
ns=d["north_star"]; print("north_star kernel %.3f ms frac %.3f coef_linf %.2e" % (ns["gram_kernel_ms"], ns["gram_kernel_frac_of_hbm_peak"], ns["coef_linf_vs_exact"]))
print("e2e %.4g rows/s" % d["e2e"]["value"]); print("clocks", d.get("clocks"))
PY
