#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r02_bench_n2.json 2> $O/r02_bench_n2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_n2.json").read().strip().splitlines()[-1])
print("N=2 value %.3e ms/step %.4f kernel %.4f tail_us %.1f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_avg"], d["roofline"]["step_tail_us"]))
s=d["strong_100M"]; print("strong", s["ms_per_fit"], s["gram_kernel_ms"], s["per_fit_host_ms"]); print(json.dumps(d["e2e"])[:300]); print(json.dumps(d["oracle_pin"]))
PY
