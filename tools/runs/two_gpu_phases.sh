#!/bin/bash
# solve-kernel phase cycles (build = peer wait + gather + normal equations) at N = 2 vs N = 1, development library
mkdir -p gpurun_out; O=gpurun_out
export B2_LIB_PATH=tools/bin/libb2gram_dev.so B2_SOLVE_TIMING=1
timeout 300 python bench.py --gpus 1 --steps 8 --warmup 3 --no-extras --no-e2e --no-cpu-baseline > $O/ph_n1.json 2> $O/ph_n1.err
grep "b2_solve" $O/ph_n1.err | tail -4
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 8 --warmup 3 --no-extras --no-e2e --no-cpu-baseline > $O/ph_n2.json 2> $O/ph_n2.err
grep "b2_solve" $O/ph_n2.err | tail -8
python - <<'PY'
import json
for f in ("ph_n1","ph_n2"):
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1]); print(f, d["ms_per_step"], d["roofline"]["step_tail_us"])
PY
