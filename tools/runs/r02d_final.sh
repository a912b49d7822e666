#!/bin/bash
# final refresh of the single-GPU evidence on the last code of round 2: full parity suite, smoke(), bench lines, launch list
mkdir -p gpurun_out; O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 16 --csv --log-file $O/r02_launches_bench_n1.csv python bench.py --steps 4 --warmup 3 --no-extras --no-e2e --no-cpu-baseline > $O/ncu7.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_reference.json 2>/dev/null
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_n1.json").read().strip().splitlines()[-1])
print("N=1 value %.3e ms/step %.4f kernel %.4f frac %.3f tail_us %.1f launches %d" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_avg"], d["roofline"]["frac"], d["roofline"]["step_tail_us"], d["gpu_launches"]))
for k,v in d["config1_10Mx128"].items():
    if isinstance(v, dict): print(k, "%.3f ms  frac %.3f  coef_linf %.2e" % (v["gram_kernel_ms"], v["frac_of_hbm_peak"], v["coef_linf_vs_exact"]))
ns=d["north_star"]; print("north_star kernel %.3f ms frac %.3f whole fit %.3f coef_linf %.2e" % (ns["gram_kernel_ms"], ns["gram_kernel_frac_of_hbm_peak"], ns["whole_fit_frac_of_hbm_peak"], ns["coef_linf_vs_exact"]))
e=d["e2e"]; print("e2e pinned %.4g pageable %.4g float64 %.4g train_model %.4g rows/s" % (e["value"], e["pageable_rows_per_s"], e["float64_rows_per_s"], e["train_model_rows_per_s"])); print("clocks", d.get("clocks"))
r=json.loads(open("gpurun_out/r02_bench_reference.json").read().strip().splitlines()[-1]); print("reference %.4g rows/s" % r["value"])
PY
