#!/bin/bash
# gpurun --gpus 8 -- "bash tools/runs/eight_gpu_validation.sh": bench at N = 8 and N = 4 (full lines), NCCL flavour at N = 8
mkdir -p gpurun_out; O=gpurun_out
nvidia-smi -L | wc -l
for n in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 > $O/r02_bench_n$n.json 2> $O/r02_bench_n$n.err
echo "bench n$n rc=$?"; tail -c 600 $O/r02_bench_n$n.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_n$n.json").read().strip().splitlines()[-1])
print("N=$n value %.4e ms/step %.4f kernel %.4f frac %.3f tail_us %.1f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_avg"], d["roofline"]["frac"], d["roofline"]["step_tail_us"]))
for k in ("parity","exchange","score_shard","strong_100M"): print(k, json.dumps(d[k])[:700])
print("e2e", d["e2e"]["value"])
PY
done
B2_NO_P2P=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 20 --warmup 5 --no-extras --no-e2e --no-cpu-baseline > $O/r02_bench_n8_nccl.json 2> $O/r02_bench_n8_nccl.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_n8_nccl.json").read().strip().splitlines()[-1])
print("N=8 NCCL value %.4e ms/step %.4f tail_us %.1f" % (d["value"], d["ms_per_step"], d["roofline"]["step_tail_us"]), d["exchange"]["exchange_used"], d["parity"]["coef_linf"], d["parity"].get("bit_identical_across_ranks"))
PY
