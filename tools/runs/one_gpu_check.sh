#!/bin/bash
# what the driver runs at round end on one GPU: the GPU suite, smoke(), the bench line
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 -p no:cacheprovider 2>&1 | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err; echo "bench rc=$?"; tail -c 400 $O/r02_bench_n1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_n1.json").read().strip().splitlines()[-1])
print("N=1 value %.4e ms/step %.4f kernel %.4f frac %.3f tail_us %.1f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_avg"], d["roofline"]["frac"], d["roofline"]["step_tail_us"]))
n=d["north_star"]; print("north_star kernel %.3f ms frac %.3f fit %.3e rows/s coef_linf %.2e" % (n["gram_kernel_ms"], n["gram_kernel_frac_of_hbm_peak"], n["fit_rows_per_s"], n["coef_linf_vs_exact"]), n["clocks"], n["per_fit_host_ms"])
print(d["default_fit_resident"])
PY
