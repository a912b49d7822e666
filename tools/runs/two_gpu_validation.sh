#!/bin/bash
# 2-GPU validation: two-rank test (NCCL + IPC peer exchange), bench at N=2 (full line incl. strong_100M, score_shard)
mkdir -p gpurun_out; O=gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_multi_gpu.py tests/test_gpu_fused.py -m gpu -q --timeout 300 -p no:cacheprovider 2>&1 | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r02_bench_n2.json 2> $O/r02_bench_n2.err
echo "bench n2 rc=$?"; tail -c 1200 $O/r02_bench_n2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_n2.json").read().strip().splitlines()[-1])
print("N=2 value %.3e ms/step %.4f kernel %.4f frac %.3f tail_us %.1f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_avg"], d["roofline"]["frac"], d["roofline"]["step_tail_us"]))
for k in ("parity","exchange","score_shard","strong_100M","e2e"): print(k, json.dumps(d[k])[:900])
PY
B2_NO_P2P=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-extras --no-e2e > $O/r02_bench_n2_nccl.json 2> $O/r02_bench_n2_nccl.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_n2_nccl.json").read().strip().splitlines()[-1])
print("N=2 NCCL value %.3e ms/step %.4f tail_us %.1f" % (d["value"], d["ms_per_step"], d["roofline"]["step_tail_us"]), d["exchange"], d["parity"]["coef_linf"], d["parity"].get("bit_identical_across_ranks"))
PY
