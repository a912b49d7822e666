#!/bin/bash
# final single-GPU evidence run of round 2: full parity suite, ncu captures (fp32 headline kernel + the two bf16-storage
# kernels), sanitizer, launch list, bench line
mkdir -p gpurun_out; O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider 2>&1 | tail -4
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:gram_tc_kernel -s 2 -c 1 -f -o $O/r02_gram_tc_12M5x128_f32 python tools/prof_target.py fit 12500000 128 f32 split > $O/ncu1.log 2>&1; tail -1 $O/ncu1.log
bash tools/runs/b16_ncu.sh 2>&1 | tail -3
bash tools/runs/sanitizer.sh > /dev/null 2>&1; echo "sanitizer clean runs: $(grep -c -E 'ERROR SUMMARY: 0 errors|0 hazards displayed' $O/r02_compute_sanitizer.txt) of 8"
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
print("parity", d["parity"]["coef_linf"], "e2e %.4g rows/s" % d["e2e"]["value"]); print("clocks", d.get("clocks"))
r=json.loads(open("gpurun_out/r02_bench_reference.json").read().strip().splitlines()[-1]); print("reference %.4g rows/s" % r["value"])
PY
