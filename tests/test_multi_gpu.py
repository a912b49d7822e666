"""Two-rank GPU test of the exchange step (skipped on a single-GPU box): NCCL all-reduce and the one-shot
peer-memory all-reduce must both turn per-rank partial statistics into the statistic of the whole dataset."""
import json
import os
import socket
import subprocess
import sys
import textwrap

import pytest

import bodywork_mlops_demo_b200 as b2

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = textwrap.dedent("""
    import os, sys, json
    import numpy as np
    import torch.distributed as dist
    sys.path.insert(0, os.environ["B2_ROOT"])
    import bodywork_mlops_demo_b200 as b2
    from bodywork_mlops_demo_b200 import sharding
    from oracle import ols_oracle as orc
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", init_method="env://", rank=rank, world_size=world)
    ctx = b2.Context(rank)
    uid = [b2.Context.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(world, rank, uid[0])
    n, d = 200_003, 128
    X, y = orc.generate_dataset(n, d, seed=31, dtype=np.float32)
    lo, hi = sharding.shard_bounds(n, world, rank)
    Xd, yd = ctx.to_device(X[lo:hi]), ctx.to_device(y[lo:hi])
    out = {}
    for mode in ("nccl", "p2p"):
        if mode == "p2p":
            handles = [None] * world
            dist.all_gather_object(handles, ctx.comm_p2p_export())
            ctx.comm_p2p_attach(world, rank, handles)
        for rep in range(3):                       # several exchanges: epoch / parity handling
            ctx.gram_reset(d); ctx.gram_accumulate(Xd, yd); ctx.gram_allreduce()
            S = ctx.gram_export()
        coef, b0 = ctx.solve()
        out[mode] = S
        out[mode + "_coef"] = coef
    full = orc.gram_stats(X, y)
    fo = orc.fit_from_stats(full)
    res = {"rank": rank,
           "nccl_rel": float(np.max(np.abs(out["nccl"] - full)) / np.max(np.abs(full))),
           "p2p_rel": float(np.max(np.abs(out["p2p"] - full)) / np.max(np.abs(full))),
           "n": float(out["p2p"][d, d]),
           "coef_err": float(np.max(np.abs(out["p2p_coef"] - fo["coef"]))),
           "p2p_vs_nccl": float(np.max(np.abs(out["p2p"] - out["nccl"])) / np.max(np.abs(full)))}
    gathered = [None] * world
    dist.all_gather_object(gathered, out["p2p"].tobytes())
    res["bit_identical_across_ranks"] = all(g == gathered[0] for g in gathered)
    if rank == 0:
        print(json.dumps(res))
    dist.barrier(); ctx.close(); dist.destroy_process_group()
""")


def test_two_rank_allreduce_nccl_and_peer_memory(tmp_path):
    if b2.native.device_count() < 2:
        pytest.skip("needs two GPUs on one box")
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   B2_ROOT=ROOT)
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=300) for p in procs]
    for p, (so, se) in zip(procs, outs):
        assert p.returncode == 0, se[-3000:]
    res = json.loads(outs[0][0].strip().splitlines()[-1])
    assert res["n"] == 200_003
    assert res["nccl_rel"] < 2e-6 and res["p2p_rel"] < 2e-6 and res["p2p_vs_nccl"] < 1e-12
    assert res["coef_err"] < 2e-5
    assert res["bit_identical_across_ranks"]
