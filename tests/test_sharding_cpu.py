"""CPU, world_size 2 over gloo: the N>1 host logic and the exchange protocol.

Two processes each own a contiguous shard of the samples (``autompc_b200.mppi.shard_of``), run the float64
oracle rollouts on their shard only, all-gather one (2 + H*nu)-number record and merge it; every rank must end
with the same action sequence as the unsharded oracle (mppi.py:110-118).  No GPU, no compute call into the
CUDA library."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.mppi_oracle import MPPIOracle, QuadCostParams
from oracle.shard_merge import merge_records, shard_record
from tests.helpers import synthetic_mlp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _problem():
    p = synthetic_mlp(5, 2, [16, 16], seed=4)
    cost = QuadCostParams(np.eye(5), 0.1 * np.eye(2), 5 * np.eye(5))
    return p, cost, np.array([-1.0, -2.0]), np.array([1.0, 2.0])


K, H, LMDA = 101, 6, 0.7          # odd K: ragged shards


def _worker(rank, world, port, out):
    from autompc_b200.mppi import shard_of
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p, cost, umin, umax = _problem()
        k_local, k_off = shard_of(K, world, rank)
        geo = torch.tensor([k_local, k_off], dtype=torch.int64)
        geos = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(geos, geo)
        # every rank draws the same global noise and the same initial action sequence (seeded), then slices
        np.random.seed(7)
        full = MPPIOracle(p, cost, umin, umax, horizon=H, num_path=K, lmda=LMDA)
        eps = full.sample_eps()
        x0 = np.linspace(-1, 1, 5)
        mine = MPPIOracle(p, cost, umin, umax, horizon=H, num_path=k_local, lmda=LMDA, draw_init=False)
        mine.act_sequence = full.act_sequence.copy()
        costs, eps_c = mine.do_rollouts(x0, eps[:, k_off:k_off + k_local].copy())
        costs = costs - mine.term_const            # the reference's common terminal scalar cancels in the softmax
        rec = torch.from_numpy(shard_record(costs, eps_c, LMDA))
        recs = torch.zeros(world * rec.numel(), dtype=torch.float64)
        dist.all_gather_into_tensor(recs, rec)
        upd = merge_records(recs.numpy().reshape(world, -1), LMDA, H, 2)
        act = mine.act_sequence + upd              # do_rollouts already shifted (mppi.py:122-123)
        full.solve(x0, eps=eps.copy())
        out[rank] = dict(geos=[g.tolist() for g in geos], err=float(np.abs(act - full.act_sequence).max()),
                         act=act.tolist())
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_update_equals_unsharded():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        res = [out[r] for r in range(world)]
    # shards tile [0, K) exactly, in rank order
    geos = res[0]["geos"]
    assert geos == res[1]["geos"]
    assert geos[0][1] == 0 and geos[0][0] + geos[1][0] == K and geos[1][1] == geos[0][0]
    for r in res:
        assert r["err"] < 1e-12
    np.testing.assert_array_equal(res[0]["act"], res[1]["act"])      # numerically identical on all ranks


def test_shard_of_partitions():
    from autompc_b200.mppi import shard_of
    for K_, W_ in [(16384, 8), (101, 2), (7, 8), (1000, 3)]:
        spans = [shard_of(K_, W_, r) for r in range(W_)]
        assert sum(k for k, _ in spans) == K_
        off = 0
        for k, o in spans:
            assert o == off
            off += k
    with pytest.raises(ValueError):
        shard_of(10, 2, 2)
