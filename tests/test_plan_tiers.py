"""CPU: the host-side launch planner of chain-bound sponge batches (csrc/sha3_api.cu plan_tiers, exported for tests as
capy_plan_tiers).  No GPU needed: the planner is pure host logic over the length histogram.

Model behind the expectations: one block per SM; a block carries 4 c warp-tier (c = 1..3 chains per
scheduler), 64 pair-tier or 128 thread-tier items; per permutation a thread-tier chain takes 1 unit, a pair 0.77, a
warp 0.48 alone on its scheduler and more when it shares it."""
import ctypes as C

import numpy as np
import pytest

from capycrypt_b200 import _binding as B

NB = 1 << 14


@pytest.fixture(scope="module")
def lib():
    return B.load()


def plan(lib, lens_blocks, sm=148, with_c=False):
    lens_blocks = np.asarray(lens_blocks, dtype=np.int64)
    bins = np.minimum(lens_blocks, NB - 1)
    hist = np.bincount(bins, minlength=NB)
    longer = (len(bins) - np.cumsum(hist)).astype(np.uint32)  # items in bins > k
    w3, p = (C.c_uint64 * 3)(), C.c_uint64()
    rc = lib.capy_plan_tiers3(longer.ctypes.data, NB, len(bins), int(bins.max()) + 1, int((bins + 1).sum()), sm,
                              C.addressof(w3), C.addressof(p))
    assert rc == 0
    w2, p2 = C.c_uint64(), C.c_uint64()
    assert lib.capy_plan_tiers(longer.ctypes.data, NB, len(bins), int(bins.max()) + 1, int((bins + 1).sum()), sm,
                               C.addressof(w2), C.addressof(p2)) == 0 and (w2.value, p2.value) == (sum(w3), p.value)
    if with_c:
        return tuple(int(x) for x in w3), int(p.value)
    return int(sum(w3)), int(p.value)


def test_single_long_message_gets_a_warp(lib):
    assert plan(lib, [58254]) == (1, 0)          # 1 x 4 MiB of SHA3-512
    assert plan(lib, [14563] * 64) == (64, 0)    # 64 x 1 MiB: 16 blocks of 4 warps


def _blocks(w3, p):
    return sum((w + 4 * (c + 1) - 1) // (4 * (c + 1)) for c, w in enumerate(w3)) + (p + 63) // 64


def test_long_messages_share_schedulers_before_they_fall_back_to_the_pair_tier(lib):
    assert plan(lib, [14563] * 64, with_c=True) == ((64, 0, 0), 0)       # a scheduler each
    assert plan(lib, [14563] * 1024, with_c=True) == ((0, 1024, 0), 0)   # 256 blocks of 4 do not fit 148 SMs, 128 blocks of 8 do
    assert plan(lib, [14563] * 1600, with_c=True) == ((0, 0, 1600), 0)   # 134 blocks of 12
    w, p = plan(lib, [14563] * 2400)             # 200 blocks of 12 would not fit: two threads per message
    assert w == 0 and p == 2400


def test_the_longest_chains_keep_a_scheduler_to_themselves(lib):
    """Mixed classes in one plan: a few very long messages, many long ones, a crowd of medium ones."""
    lens = [14563] * 40 + [12000] * 400 + [9000] * 3000 + [100] * 200000
    (w1, w2, w3), p = plan(lib, lens, with_c=True)
    assert w1 + w2 + w3 + p > 0 and _blocks((w1, w2, w3), p) < 148
    # classes are nested by length: whoever shares a scheduler is not longer than whoever has one alone
    assert w1 >= 40 or w1 == 0


def test_work_bound_batch_has_no_fast_tier(lib):
    rng = np.random.default_rng(1)
    # 2^20 short messages: the longest chain is nowhere near the work-bound time (the C side would not even call the
    # planner here; called directly it must still return an empty plan)
    assert plan(lib, rng.integers(0, 20, size=1 << 20)) == (0, 0)


def _mixed(total_bytes, seed=5):
    rs = np.random.default_rng(seed)
    lens, acc = [], 0
    while acc < total_bytes:
        c = np.exp(rs.uniform(np.log(64), np.log(1 << 20), size=8192)).astype(np.int64)
        lens.append(c)
        acc += int(c.sum())
    lens = np.concatenate(lens)
    return lens[: int(np.searchsorted(np.cumsum(lens), total_bytes)) + 1] // 72


def test_mixed_batches_like_config_5(lib):
    # 16 GiB: almost work-bound -> a pair tier of about a thousand items, no warp tier
    w, p = plan(lib, _mixed(16 << 30))
    assert w == 0 and 500 <= p <= 3000
    # the 2 GiB shard of an 8-GPU run: chain-bound -> both fast tiers, all of their blocks resident at once
    w3, p = plan(lib, _mixed(2 << 30), with_c=True)
    w = sum(w3)
    assert w > 0 and p > 0 and _blocks(w3, p) < 148
    # every chain must fit the step the planner assumed: the first thread-tier item is shorter than the first pair item
    # times 0.77 / 1 and the first pair item shorter than the longest times 0.48 / 0.77
    lens = np.sort(_mixed(2 << 30))[::-1] + 1
    assert lens[w + p] <= lens[w] and lens[w] * 0.772 >= lens[0] * 0.478 * 0.99
    # the shards of a 2- and a 4-GPU run hold more long messages than there are schedulers: they get a warp tier too
    for total in (8 << 30, 4 << 30):
        w3, p = plan(lib, _mixed(total), with_c=True)
        assert sum(w3) > 0 and _blocks(w3, p) < 148, (total, w3, p)


def test_bad_arguments(lib):
    w, p = C.c_uint64(), C.c_uint64()
    assert lib.capy_plan_tiers(None, NB, 1, 1, 1, 148, C.addressof(w), C.addressof(p)) == B.ERR_BAD_ARG
