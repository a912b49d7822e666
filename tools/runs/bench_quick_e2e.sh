mkdir -p gpurun_out
timeout 900 python bench.py --steps 6 --warmup 3 > gpurun_out/bq.json 2> gpurun_out/bq.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bq.json").read().strip().splitlines()[-1]); e=d["e2e"]
print({k:e[k] for k in ("value","pageable_rows_per_s","float64_rows_per_s","float64_fit_ms","train_model_rows_per_s")})
PY
