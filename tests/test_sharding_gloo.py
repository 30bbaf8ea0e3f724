"""CPU suite: the N>1 host logic, exercised with world_size-2 gloo process groups on 127.0.0.1.

Two things shard on this path (no collective touches the data):
  * one process per GPU (bench.py under torch.distributed.run): every rank derives its share of the cfg-5 batch with
    bench.lpt_shards -- longest-processing-time-first over ALL messages -- and the timing rule is the max over ranks;
  * one ctx over several devices (capy_gpu_init(devs, n)): capy_sha3_batch divides a ragged batch with lpt_shares of
    csrc/hostbatch.h (outliers dealt out longest first, the rest in contiguous ranges), exported for tests as
    capy_lpt_shares.  The same C++ function is also checked by tests/host/hostbatch_check.cpp."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

from capycrypt_b200 import _binding as B  # noqa: E402


def _owners(lens, parts, unit=72, per_item=2):
    lib = B.load()
    off = np.zeros(len(lens) + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens)
    owner = np.full(len(lens), 0xFFFFFFFF, dtype=np.uint32)
    assert lib.capy_lpt_shares(off.ctypes.data, len(lens), parts, unit, per_item, owner.ctypes.data) == 0
    return owner


def _mixed(n, seed):
    rnd = np.random.default_rng(seed)
    return np.exp(rnd.uniform(np.log(64), np.log(1 << 20), size=n)).astype(np.int64)


def test_lpt_shares_of_a_multi_device_ctx():
    for lens in (_mixed(20000, 1), np.sort(_mixed(20000, 2))[::-1].copy(), np.sort(_mixed(20000, 3))):
        cost = lens // 72 + 2
        for parts in (2, 4, 8):
            owner = _owners(lens, parts)
            assert owner.max() < parts  # every item has exactly one owner
            load = np.bincount(owner, weights=cost, minlength=parts)
            assert load.max() <= load.mean() * 1.03 + cost.max(), (parts, load)
            # the long chains are spread: no device holds more than its share (+1) of the 64 longest messages
            top = np.argsort(lens)[-64:]
            per_dev = np.bincount(owner[top], minlength=parts)
            assert per_dev.max() - per_dev.min() <= max(2, 64 // parts // 2), (parts, per_dev)
            # and the bulk still travels in few pieces: runs of consecutive items per device
            runs = 1 + int(np.count_nonzero(np.diff(owner.astype(np.int64)) != 0))
            assert runs <= 2 * 4096 + parts
    # a uniform batch is split into plain contiguous ranges
    owner = _owners(np.full(1 << 16, 64), 4, unit=136, per_item=1)
    assert np.all(np.diff(owner.astype(np.int64)) >= 0) and np.bincount(owner).tolist() == [1 << 14] * 4


def test_rank_level_lpt_of_the_bench():
    lens = _mixed(30000, 5)
    for world in (2, 4, 8):
        sh = bench.lpt_shards(lens, world)
        allidx = np.sort(np.concatenate(sh))
        assert np.array_equal(allidx, np.arange(len(lens)))
        loads = [int((lens[s] // 72 + 2).sum()) for s in sh]
        assert max(loads) - min(loads) <= int(lens.max() // 72 + 2)  # LPT: within one item of each other
        longest = np.argsort(lens)[-world * 4:]
        assert all(3 <= np.isin(longest, s).sum() <= 5 for s in sh)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # every rank derives the same global batch description and takes its own LPT share: no exchange of data
    lens = _mixed(5000, 7)
    mine = bench.lpt_shards(lens, world)[rank]
    t = torch.tensor([len(mine), int(lens[mine].sum()), int((lens[mine] // 72 + 2).sum())], dtype=torch.int64)
    tot = t.clone()
    dist.all_reduce(tot)  # test-only check that the shares cover the batch
    mx = t.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    # timing rule of bench.py: every multi-rank number is the MAX over ranks
    ms = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dist.barrier()
    q.put((rank, int(tot[0]), int(tot[1]), int(t[2]), int(mx[2]), float(ms.item())))
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
    lens = _mixed(5000, 7)
    for rank, n_tot, b_tot, my_cost, max_cost, t in res:
        assert n_tot == 5000 and b_tot == int(lens.sum())
        assert max_cost - my_cost <= int(lens.max() // 72 + 2)  # the two ranks carry the same work to within one message
        assert t == 2.0  # max over ranks of (1 + rank)
