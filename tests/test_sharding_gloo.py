"""CPU suite: the N>1 host path (one process per GPU, contiguous work-balanced shards, no data-path collective,
max-over-ranks timing) exercised with world_size-2 gloo process groups on 127.0.0.1."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from capycrypt_b200 import sharding


def test_contiguous_shards_cover_and_balance():
    rnd = np.random.default_rng(1)
    lens = np.exp(rnd.uniform(np.log(64), np.log(1 << 20), size=5000)).astype(np.int64)
    off = np.concatenate([[0], np.cumsum(lens)])
    cost = sharding.item_cost(off, 72)
    for world in (1, 2, 3, 4, 8):
        sh = sharding.contiguous_shards(cost, world)
        assert sh[0][0] == 0 and sh[-1][1] == len(cost)
        assert all(a[1] == b[0] for a, b in zip(sh, sh[1:]))
        loads = [int(cost[a:b].sum()) for a, b in sh]
        assert max(loads) <= sum(loads) / world + int(cost.max())  # within one item of perfect
    assert sharding.contiguous_shards(np.zeros(0, np.int64), 4) == [(0, 0)] * 4
    assert sharding.contiguous_shards(np.array([5]), 3) in ([(0, 0), (0, 1), (1, 1)], [(0, 1), (1, 1), (1, 1)], [(0, 0), (0, 0), (0, 1)])


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # every rank derives the same global batch description and takes its own shard: no exchange of data
    rnd = np.random.default_rng(7)
    lens = rnd.integers(0, 5000, size=1000)
    off = np.concatenate([[0], np.cumsum(lens)])
    i0, i1 = sharding.contiguous_shards(sharding.item_cost(off, 136), world)[rank]
    mine = torch.tensor([i1 - i0, int(lens[i0:i1].sum())], dtype=torch.int64)
    tot = mine.clone()
    dist.all_reduce(tot)  # test-only check that the shards cover the batch
    strided = sharding.strided_shard(len(lens), rank, world)
    # timing rule: max over ranks
    t = sharding.max_over_ranks(1.0 + rank, dist)
    dist.barrier()
    q.put((rank, int(tot[0]), int(tot[1]), len(strided), t))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_timing():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    rnd = np.random.default_rng(7)
    lens = rnd.integers(0, 5000, size=1000)
    for rank, n_tot, b_tot, n_strided, t in res:
        assert n_tot == 1000 and b_tot == int(lens.sum())
        assert t == 2.0  # max over ranks of (1 + rank)
    assert sum(r[3] for r in res) == 1000
