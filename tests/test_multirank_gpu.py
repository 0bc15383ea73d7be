"""GPU, two ranks on ONE device over gloo: the N>1 host logic of the engine that does not need two GPUs.

* ``evaluate_candidates(group=)`` (tuning/pipeline_tuner.py:213-239 over ranks): candidates dealt round-robin, every
  rank ends with the same full cost list, equal to the single-process evaluation;
* the sharded ``MPPI(group=)`` host path with ``exchange='nccl'``-style gather replaced by gloo is NOT covered here
  (NCCL cannot run two ranks on one device); its data path is covered on one device by
  test_fused_peer_exchange_two_shards_one_device / test_sharded_partials_merge_equals_single_handle and on real
  multi-GPU hardware by bench.py's parity figure and scripts/multi_gpu_check.py (logs under profiles/).
* the rank-0 broadcast of the nominal action sequence and of host-drawn noise (round-1 advisor finding: ranks with
  diverging NumPy streams rolled out around different sequences).
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _candidates():
    from autompc_b200 import MPPI, B200MLP
    from autompc_b200.mlp import MLPWeights
    from autompc_b200.plugin import QuadCost, Task, ThresholdCost
    from autompc_b200.problems import cartpole_problem
    from tests.helpers import GOLDEN
    z = np.load(os.path.join(GOLDEN, "cartpole_mlp.npz"))
    system, task, w, x0 = cartpole_problem(MLPWeights.from_npz(z))
    model = B200MLP(system, w)
    rng = np.random.default_rng(5)
    ctls = []
    np.random.seed(3)
    for i in range(5):
        t = Task(system)
        t.set_ctrl_bound("u", -20.0, 20.0)
        g = np.exp(rng.uniform(np.log(1e-2), np.log(1e2), size=9))
        t.set_cost(QuadCost(system, np.diag(g[:4]), np.diag(g[8:9]), np.diag(g[4:8]), goal=np.zeros(4)))
        ctls.append(MPPI(system, t, model, horizon=int(rng.integers(5, 16)), num_path=int(rng.integers(100, 400)),
                         sigma=float(rng.uniform(0.2, 1.5)), lmda=float(rng.uniform(0.3, 1.5)), seed=i, precision="fp32"))
    score = ThresholdCost(system, np.zeros(4), [0, 3], 0.2)
    return ctls, model, x0, score


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from autompc_b200 import evaluate_candidates
        ctls, model, x0, score = _candidates()
        costs, mine = evaluate_candidates(ctls, x0, 12, model, group=dist.group.WORLD, cost=score)
        # rank-0 broadcast: every rank seeds its NumPy stream differently, the nominal sequence must still agree
        from autompc_b200 import MPPI
        np.random.seed(100 + rank)
        c = MPPI(ctls[0].system, ctls[0].task, model, horizon=6, num_path=64, precision="fp32", noise="numpy")
        c.world, c.rank, c.group = world, rank, dist.group.WORLD          # host-side sharding logic only
        a = c._from_rank0(np.random.normal(size=(6, 1)))
        e = c._from_rank0(np.random.normal(size=(3, 2)))
        out[rank] = dict(costs=list(costs), mine=sorted(mine.keys()), act=a.tolist(), eps=e.tolist())
        for k in ctls + [c]:
            k.close()
    finally:
        dist.destroy_process_group()


def test_evaluate_candidates_over_two_ranks_equals_one_process():
    from autompc_b200 import evaluate_candidates
    world = 2
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        res = [dict(out[r]) for r in range(world)]
    assert res[0]["mine"] == [0, 2, 4] and res[1]["mine"] == [1, 3]      # dealt round-robin
    assert res[0]["costs"] == res[1]["costs"]                            # every rank holds the full list
    assert res[0]["act"] == res[1]["act"] and res[0]["eps"] == res[1]["eps"]   # rank 0's draws everywhere
    ctls, model, x0, score = _candidates()
    costs, _ = evaluate_candidates(ctls, x0, 12, model, cost=score)
    assert list(costs) == res[0]["costs"]                                # Philox noise: identical closed loops
    assert all(0 <= c <= 13 for c in costs)
    for k in ctls:
        k.close()
