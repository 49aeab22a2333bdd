"""Multi-rank host logic on CPU: env sharding covers the batch exactly once, and the episode-statistic reduction over a
world_size-2 gloo group equals the single-process statistics.  (The step path itself has no collective.)"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from jaxmarl_hft_b200 import dist as D


@pytest.mark.parametrize("n,world", [(16384, 1), (65536, 8), (10, 3), (7, 8), (131072, 8)])
def test_shards_partition_the_batch(n, world):
    shards = [D.shard_range(n, r, world) for r in range(world)]
    assert shards[0].start == 0 and sum(s.count for s in shards) == n
    for a, b in zip(shards, shards[1:]):
        assert a.start + a.count == b.start
    assert max(s.count for s in shards) - min(s.count for s in shards) <= 1
    with pytest.raises(ValueError):
        D.shard_range(n, world, world)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, w, _ = D.init_from_env(backend="gloo")
    rng = np.random.default_rng(0)
    rewards = torch.from_numpy(rng.normal(size=n))          # the same global array on every rank
    done = torch.from_numpy(rng.random(n) < 0.3)
    sh = D.shard_range(n, r, w)
    sl = slice(sh.start, sh.start + sh.count)
    local = torch.stack([D.local_episode_stats(rewards[sl]), D.local_episode_stats(rewards[sl], done[sl])])
    out = D.reduce_episode_stats(local)
    q.put((rank, {k: v.tolist() for k, v in out.items()}))
    dist.destroy_process_group()


def test_episode_stats_reduce_over_two_ranks():
    n, world = 1001, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(0)
    rewards = rng.normal(size=n)
    done = rng.random(n) < 0.3
    for rank in range(world):
        got = res[rank]
        for j, x in enumerate((rewards, rewards[done])):
            assert got["count"][j] == x.size
            np.testing.assert_allclose(got["mean"][j], x.mean(), rtol=1e-12)
            np.testing.assert_allclose(got["std"][j], x.std(), rtol=1e-9)
            assert got["min"][j] == x.min() and got["max"][j] == x.max()


def test_single_process_is_a_no_op():
    x = torch.arange(10, dtype=torch.float32)
    out = D.reduce_episode_stats(D.local_episode_stats(x))
    assert out["count"].item() == 10 and out["mean"].item() == 4.5
    empty = D.reduce_episode_stats(D.local_episode_stats(x, torch.zeros(10, dtype=torch.bool)))
    assert empty["count"].item() == 0
