"""GPU: the pieces of round 2 that make the device-pointer entry points asynchronous and the seal a single launch.

  * capy_gpu_set_plan_cache: the launch plan of a ragged batch is kept per offsets array; a later call with the same
    device array does not touch the host.  The contract says a stale plan (the caller rewrote the offsets in place) costs
    speed, never correctness -- checked here on purpose.
  * sponge-AE seal: tag pass and keystream pass go out as ONE launch when the ciphertext has its own buffer and as two
    when the caller encrypts in place; both must give the reference's bytes.
  * capy_copy_probe: the link ceiling bench.py reports."""
import numpy as np
import pytest
import torch

from capycrypt_b200 import pack
from oracle import ref_sha3 as R

pytestmark = pytest.mark.gpu


def _ragged(rnd, n, hi):
    lens = rnd.integers(0, hi, size=n)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    return lens, off


def test_plan_cache_reuse_and_stale_plans(engine, oracle):
    rnd = np.random.default_rng(11)
    n = 3000
    # a chain-bound batch (tiers + longest-first order) so that the cached plan is not trivial
    lens = np.concatenate([rnd.integers(150_000, 400_000, size=24), rnd.integers(0, 3000, size=n - 24)])
    rnd.shuffle(lens)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    data = rnd.integers(0, 256, size=int(off[-1]) + 400_000 * 8, dtype=np.uint8)  # slack for the rewritten offsets below
    t_data = torch.from_numpy(data).cuda()
    t_off = torch.from_numpy(off).cuda()
    t_out = torch.zeros(n * 64, dtype=torch.uint8, device="cuda")
    want = oracle.sha3_batch(data, off.astype(np.uint64), 512, threads=0)
    engine.set_plan_cache(True)
    try:
        l0 = engine.launch_count
        engine.sha3_dev(t_data, t_off, 512, t_out)  # builds the plan (histogram, scan, scatter) and launches
        first = engine.launch_count - l0
        torch.cuda.synchronize()
        assert np.array_equal(t_out.cpu().numpy().reshape(n, 64), want)
        t_out.zero_()
        l0 = engine.launch_count
        engine.sha3_dev(t_data, t_off, 512, t_out)  # same offsets array: the sponge launch only
        assert engine.launch_count - l0 == 1 and first >= 4
        torch.cuda.synchronize()
        assert np.array_equal(t_out.cpu().numpy().reshape(n, 64), want)
        # the caller breaks the contract: other lengths under the same address.  The plan is stale (wrong tiers, wrong
        # order) but every digest must still be right.
        lens2 = lens.copy()
        rnd.shuffle(lens2)
        lens2[:8] += 17
        off2 = np.concatenate([[0], np.cumsum(lens2)]).astype(np.int64)
        t_off.copy_(torch.from_numpy(off2))
        t_out.zero_()
        l0 = engine.launch_count
        engine.sha3_dev(t_data, t_off, 512, t_out)
        assert engine.launch_count - l0 == 1
        torch.cuda.synchronize()
        assert np.array_equal(t_out.cpu().numpy().reshape(n, 64), oracle.sha3_batch(data, off2.astype(np.uint64), 512, threads=0))
        # KMAC with the same offsets array as message offsets: its own cache entry (other unit), same answers as uncached
        keys = torch.from_numpy(rnd.integers(0, 256, size=n * 32, dtype=np.uint8)).cuda()
        koff = (torch.arange(n + 1, dtype=torch.int64) * 32).cuda()
        o1 = torch.zeros(n * 64, dtype=torch.uint8, device="cuda")
        o2 = torch.zeros_like(o1)
        engine.kmac_xof_dev(keys, koff, t_data, t_off, 512, b"tag", 512, o1)
        engine.kmac_xof_dev(keys, koff, t_data, t_off, 512, b"tag", 512, o2)
        torch.cuda.synchronize()
        assert torch.equal(o1, o2)
    finally:
        engine.set_plan_cache(False)
    o3 = torch.zeros_like(o1)
    engine.kmac_xof_dev(keys, koff, t_data, t_off, 512, b"tag", 512, o3)
    torch.cuda.synchronize()
    assert torch.equal(o1, o3)
    idx = [0, 1, n // 2, n - 1]
    for i in idx:
        k = keys[32 * i:32 * i + 32].cpu().numpy().tobytes()
        m = data[int(off2[i]):int(off2[i + 1])].tobytes()
        assert o3[64 * i:64 * i + 64].cpu().numpy().tobytes() == R.kmac_xof(k, m, 512, b"tag", 512)


@pytest.mark.parametrize("d", [256, 512])
def test_seal_in_one_launch_and_in_place(engine, d):
    rnd = np.random.default_rng(12 + d)
    n = 700
    msgs = [bytes(rnd.integers(0, 256, size=int(k), dtype=np.uint8)) for k in rnd.integers(0, 900, size=n)]
    msgs[0], msgs[1], msgs[2] = b"", b"x" * 136, b"y" * 168
    pws = [bytes(rnd.integers(0, 256, size=int(k), dtype=np.uint8)) for k in rnd.integers(0, 40, size=n)]
    nonces = rnd.integers(0, 256, size=n * 512, dtype=np.uint8)
    md, mo = pack(msgs)
    pd, po = pack(pws)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    t_m, t_mo, t_p, t_po, t_n = t(md), t(mo.astype(np.int64)), t(pd if len(pd) else np.zeros(1, np.uint8)), t(po.astype(np.int64)), t(nonces)
    ct = torch.zeros_like(t_m)
    tag = torch.zeros(n * 64, dtype=torch.uint8, device="cuda")
    l0 = engine.launch_count
    engine.sponge_encrypt_dev(t_p, t_po, len(pd), t_n, 512, t_m, t_mo, d, ct, tag)
    separate = engine.launch_count - l0
    torch.cuda.synchronize()
    inplace = t_m.clone()
    tag2 = torch.zeros_like(tag)
    l0 = engine.launch_count
    engine.sponge_encrypt_dev(t_p, t_po, len(pd), t_n, 512, inplace, t_mo, d, inplace, tag2)
    aliased = engine.launch_count - l0
    torch.cuda.synchronize()
    # tag pass and keystream pass: ONE sponge launch when the ciphertext has its own buffer, two when the message is
    # overwritten (the second pass also plans its ragged batch again: histogram, scan, scatter)
    assert separate + 1 <= aliased <= separate + 4
    assert torch.equal(ct, inplace) and torch.equal(tag, tag2)
    ct_h, tag_h = ct.cpu().numpy(), tag.cpu().numpy().reshape(n, 64)
    for i in list(range(6)) + [n - 1]:
        c_ref, t_ref = R.sha3_encrypt(msgs[i], pws[i], d, nonces[512 * i:512 * i + 512].tobytes())
        assert ct_h[int(mo[i]):int(mo[i + 1])].tobytes() == c_ref and tag_h[i].tobytes() == t_ref, i
    out = torch.zeros_like(t_m)
    ok = torch.zeros(n, dtype=torch.uint8, device="cuda")
    engine.sponge_decrypt_dev(t_p, t_po, len(pd), t_n, 512, ct, t_mo, tag, d, out, ok)
    torch.cuda.synchronize()
    assert bool(ok.all().item()) and torch.equal(out, t_m)


def test_copy_probe_reports_a_link_time(engine):
    h_in, h_out = engine.pinned(8 << 20), engine.pinned(4 << 20)
    ms = engine.copy_probe(h_in, h_out, reps=3)
    assert 0.01 < ms < 50.0  # 8 MiB in + 4 MiB out: between 240 GB/s and 0.25 GB/s
    with pytest.raises(Exception):
        engine.copy_probe(h_in, h_out, reps=0)


def test_host_calls_plan_from_the_host_offsets(engine, oracle):
    """Host-buffer entry points read the lengths from the caller's offsets array (no histogram round trip): a uniform
    batch is ONE launch, a chain-bound ragged batch still gets its tiers and the longest-first order, same digests."""
    rnd = np.random.default_rng(21)
    n = 4096
    data = rnd.integers(0, 256, size=n * 200, dtype=np.uint8)
    off = (np.arange(n + 1, dtype=np.uint64) * 200)
    l0 = engine.launch_count
    got = engine.sha3(data, off, 512)
    assert engine.launch_count - l0 == 1  # uniform lengths through the ragged entry point: the sponge launch only
    assert np.array_equal(got, oracle.sha3_batch(data, off, 512, threads=0))
    keys = rnd.integers(0, 256, size=n * 32, dtype=np.uint8)
    koff = np.arange(n + 1, dtype=np.uint64) * 32
    l0 = engine.launch_count
    tags = engine.kmac_xof(keys, koff, data, off, 512, b"tag", 512, out=engine.pinned(n * 64).reshape(n, 64))
    assert engine.launch_count - l0 <= 2  # (+ the prefix state when "tag" is new to the cache)
    assert np.array_equal(tags[:64], oracle.kmac_xof_batch(keys[: 64 * 32], koff[:65], data[: 64 * 200], off[:65], 512, b"tag", 512))
    # chain-bound: a few long messages among short ones
    lens = np.concatenate([rnd.integers(200_000, 500_000, size=12), rnd.integers(0, 2000, size=3000)])
    rnd.shuffle(lens)
    off2 = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    data2 = rnd.integers(0, 256, size=int(off2[-1]), dtype=np.uint8)
    l0 = engine.launch_count
    got2 = engine.sha3(data2, off2, 512)
    assert engine.launch_count - l0 == 4  # histogram, scan, scatter, one tiered sponge launch -- and no D2H of the plan
    assert np.array_equal(got2, oracle.sha3_batch(data2, off2, 512, threads=0))
    # the same batch through the device-pointer entry point (plan fetched from the device) gives the same bytes
    t_out = torch.zeros(len(lens) * 64, dtype=torch.uint8, device="cuda")
    engine.sha3_dev(torch.from_numpy(data2).cuda(), torch.from_numpy(off2.astype(np.int64)).cuda(), 512, t_out)
    torch.cuda.synchronize()
    assert np.array_equal(t_out.cpu().numpy().reshape(-1, 64), got2)
