"""Multi-process host logic on CPU (gloo, world_size 2): row sharding + the single exchange step.

The per-shard statistic is computed by the oracle here (no GPU in this tier of tests); what is under test is
the shard arithmetic, the all-reduce protocol bench.py / the stage use, and the score-stat combination rule."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from bodywork_mlops_demo_b200 import sharding
from oracle import ols_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n,world", [(0, 1), (1, 2), (63, 2), (64, 2), (65, 3), (100_003, 8), (12_500_000 * 8, 8)])
def test_shard_bounds_partition(n, world):
    shards = sharding.all_shards(n, world)
    assert shards[0][0] == 0 and shards[-1][1] == n
    for (lo, hi), (lo2, _) in zip(shards, shards[1:]):
        assert lo <= hi == lo2
    for lo, hi in shards[:-1]:
        assert lo % sharding.TILE_ROWS == 0 and hi % sharding.TILE_ROWS == 0 or hi == n
    sizes = [hi - lo for lo, hi in shards]
    assert max(sizes) - min(sizes) <= sharding.TILE_ROWS or n < sharding.TILE_ROWS * world
    with pytest.raises(ValueError):
        sharding.shard_bounds(10, 2, 2)


def test_combine_score_stats_equals_global():
    rng = np.random.RandomState(0)
    y = rng.normal(50, 20, 1000); p = y + rng.normal(0, 3, 1000)
    parts = [orc.score_stats(y[lo:hi], p[lo:hi]) for lo, hi in sharding.all_shards(1000, 3, align=1)]
    np.testing.assert_allclose(sharding.combine_score_stats(parts), orc.score_stats(y, p), rtol=1e-12)


_WORKER = textwrap.dedent("""
    import os, sys, json
    import numpy as np
    import torch, torch.distributed as dist
    sys.path.insert(0, os.environ["B2_ROOT"])
    from bodywork_mlops_demo_b200 import sharding
    from oracle import ols_oracle as orc
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", init_method="env://", rank=rank, world_size=world)
    n, d = 50_003, 16
    X, y = orc.generate_dataset(n, d, seed=11)               # every rank draws the same global dataset
    lo, hi = sharding.shard_bounds(n, world, rank)
    S = torch.from_numpy(orc.gram_stats(X[lo:hi], y[lo:hi]))  # this rank's partial statistic
    dist.all_reduce(S, op=dist.ReduceOp.SUM)                  # the one exchange step (NCCL on the GPU box)
    fit = orc.fit_from_stats(S.numpy())
    # scoring: five sums + one max
    p = orc.predict(X[lo:hi], fit["coef"], fit["intercept"])
    st = torch.from_numpy(orc.score_stats(y[lo:hi], p))
    mx = st[[4, 9]].clone(); st[4] = 0.0; st[9] = 0.0       # eight sums + two maxima (b2_score_allreduce)
    dist.all_reduce(st, op=dist.ReduceOp.SUM); dist.all_reduce(mx, op=dist.ReduceOp.MAX); st[4] = mx[0]; st[9] = mx[1]
    # max-over-ranks timing rule used by bench.py
    t = torch.tensor([1.0 + rank], dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        full = orc.fit_from_stats(orc.gram_stats(X, y))
        ref = orc.score_stats(y, orc.predict(X, full["coef"], full["intercept"]))
        print(json.dumps({"coef_err": float(np.max(np.abs(fit["coef"] - full["coef"]))),
                          "n": float(S[d, d]), "stats_rel": float(np.max(np.abs(st.numpy() - ref) / np.abs(ref))),
                          "tmax": float(t.item())}))
    dist.barrier(); dist.destroy_process_group()
""")


def test_two_rank_gloo_allreduce_of_partial_statistics(tmp_path):
    pytest.importorskip("torch")
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   B2_ROOT=ROOT, OMP_NUM_THREADS="2")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=240) for p in procs]
    for p, (so, se) in zip(procs, outs):
        assert p.returncode == 0, se[-2000:]
    import json
    res = json.loads(outs[0][0].strip().splitlines()[-1])
    assert res["n"] == 50_003 and res["coef_err"] < 1e-11 and res["stats_rel"] < 1e-10 and res["tmax"] == 2.0
