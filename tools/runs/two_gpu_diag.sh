#!/bin/bash
# 8-GPU diagnostic: the headline loop only, peer-memory exchange and NCCL alternating, per-rank kernel times
mkdir -p gpurun_out; O=gpurun_out
for r in 1; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2952$r bench.py --gpus 2 --steps 30 --warmup 5 --no-extras --no-e2e --no-cpu-baseline > $O/diag_n8_p2p_$r.json 2> $O/diag.err
B2_NO_P2P=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$r bench.py --gpus 2 --steps 30 --warmup 5 --no-extras --no-e2e --no-cpu-baseline > $O/diag_n8_nccl_$r.json 2> $O/diag.err
done
python - <<'PY'
import json
for f in ("p2p_1","nccl_1"):
    d=json.loads(open(f"gpurun_out/diag_n8_{f}.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print(f, "step %.4f ms tail %.1f us vs slowest kernel %.1f us" % (d["ms_per_step"], r["step_tail_us"], r["step_tail_us_vs_slowest_kernel"]), "kernels", r["kernel_ms_by_rank"])
PY
