"""GPU parity at the FULL sizes BASELINE.json names for configs 2-5 (config 1: test_sha3_gpu.py::test_sha3_cfg1_full_size).

Each test runs the whole batch through the C ABI and compares a seeded sample of >= 4096 items with the C oracle
(oracle/ref_cpu.c, threads = all host cores: the oracle is the checker here, never the thing measured), plus a
size-independent property over the whole batch where the domain offers one."""
import numpy as np
import pytest

from capycrypt_b200 import pack

pytestmark = pytest.mark.gpu

SAMPLE = 4096


def _rows(buf, n, w, idx):
    return np.ascontiguousarray(buf.reshape(n, w)[idx]).reshape(-1)


def test_cfg2_kmacxof256_2p16_x_4kb_all_output_lengths(engine, oracle):
    """cfg 2: KMACXOF256 (D512) over 2^16 x 4 KB, 32-byte keys, S = "My Tagged Application", output 256 / 512 /
    4096 / 32768 bits (the last = 8 * len(m): the keystream shape of sha3/encryptable.rs:41); cSHAKE256 rides along."""
    n, mlen, klen = 1 << 16, 4096, 32
    rnd = np.random.default_rng(2)
    data = rnd.integers(0, 256, size=n * mlen, dtype=np.uint8)
    keys = rnd.integers(0, 256, size=n * klen, dtype=np.uint8)
    off = np.arange(n + 1, dtype=np.uint64) * mlen
    koff = np.arange(n + 1, dtype=np.uint64) * klen
    idx = np.sort(rnd.choice(n, size=SAMPLE, replace=False))
    s_data, s_keys = _rows(data, n, mlen, idx), _rows(keys, n, klen, idx)
    s_off = np.arange(SAMPLE + 1, dtype=np.uint64) * mlen
    s_koff = np.arange(SAMPLE + 1, dtype=np.uint64) * klen
    custom = b"My Tagged Application"
    for out_bits in (256, 512, 4096, 32768):
        got = engine.kmac_xof(keys, koff, data, off, out_bits, custom, 512)
        want = oracle.kmac_xof_batch(s_keys, s_koff, s_data, s_off, out_bits, custom, 512, threads=0)
        assert np.array_equal(got[idx], want), out_bits
        if out_bits == 512:  # prefix property of an XOF: shorter outputs are prefixes of longer ones
            first = got
        if out_bits == 32768:
            assert np.array_equal(got[:, :64], first)
    got = engine.cshake(data, off, 512, b"", b"Email Signature", 512)
    want = oracle.cshake_batch(s_data, s_off, 512, b"", b"Email Signature", 512, threads=0)
    assert np.array_equal(got[idx], want)


def test_cfg3_keygen_2p20(engine, oracle):
    """cfg 3: [s]G for 2^20 random 56-byte scalars, and KeyPair::new over 2^20 x 32-byte passwords."""
    n = 1 << 20
    rnd = np.random.default_rng(3)
    sc = rnd.integers(0, 256, size=n * 56, dtype=np.uint8)
    got = engine.ed448_fixed_base(sc)
    idx = np.sort(rnd.choice(n, size=SAMPLE, replace=False))
    assert np.array_equal(got[idx], oracle.fixed_base_batch(_rows(sc, n, 56, idx), threads=0))
    # every output is a point of the curve x^2 + y^2 = 1 + d x^2 y^2 (checked on a second sample with Python integers)
    from oracle import ref_ed448 as E

    for i in rnd.choice(n, size=64, replace=False):
        x = int.from_bytes(got[i, :56].tobytes(), "little")
        y = int.from_bytes(got[i, 56:].tobytes(), "little")
        assert (x * x + y * y - 1 - E.D * x * x * y * y) % E.P == 0
    pws = rnd.integers(0, 256, size=n * 32, dtype=np.uint8)
    poff = np.arange(n + 1, dtype=np.uint64) * 32
    got = engine.ed448_keygen(pws, poff, 512)
    want = oracle.keygen_batch(_rows(pws, n, 32, idx), np.arange(SAMPLE + 1, dtype=np.uint64) * 32, 512, threads=0)
    assert np.array_equal(got[idx], want)


def test_cfg4_sign_verify_2p18(engine, oracle):
    """cfg 4: Schnorr sign + verify for 2^18 x (32-byte password, 256-byte message), D512."""
    n, plen, mlen = 1 << 18, 32, 256
    rnd = np.random.default_rng(4)
    pws = rnd.integers(0, 256, size=n * plen, dtype=np.uint8)
    msgs = rnd.integers(0, 256, size=n * mlen, dtype=np.uint8)
    poff = np.arange(n + 1, dtype=np.uint64) * plen
    moff = np.arange(n + 1, dtype=np.uint64) * mlen
    pub = engine.ed448_keygen(pws, poff, 512)
    h, z = engine.ed448_sign(pws, poff, msgs, moff, 512)
    idx = np.sort(rnd.choice(n, size=SAMPLE, replace=False))
    s_pw, s_msg = _rows(pws, n, plen, idx), _rows(msgs, n, mlen, idx)
    s_poff = np.arange(SAMPLE + 1, dtype=np.uint64) * plen
    s_moff = np.arange(SAMPLE + 1, dtype=np.uint64) * mlen
    h_ref, z_ref = oracle.sign_batch(s_pw, s_poff, s_msg, s_moff, 512, threads=0)
    assert np.array_equal(h[idx], h_ref) and np.array_equal(z[idx], z_ref)
    assert np.array_equal(pub[idx], oracle.keygen_batch(s_pw, s_poff, 512, threads=0))
    # the oracle accepts the engine's signatures on the sample ...
    assert oracle.verify_batch(pub[idx].reshape(-1), s_msg, s_moff, h[idx].reshape(-1), z[idx].reshape(-1), 512, threads=0).all()
    # ... and the engine accepts all 2^18 of them, and rejects exactly the tampered ones
    rc, ok = engine.ed448_verify(pub, msgs, moff, h, z, 512)
    assert rc == 0 and ok.all()
    bad = np.sort(rnd.choice(n, size=257, replace=False))
    h2, z2, m2 = h.copy(), z.copy(), msgs.copy()
    h2[bad[:100], 7] ^= 1
    z2[bad[100:200], 55] ^= 0x80
    m2.reshape(n, mlen)[bad[200:], 0] ^= 1
    rc, ok = engine.ed448_verify(pub, m2, moff, h2, z2, 512)
    want_ok = np.ones(n, dtype=bool)
    want_ok[bad] = False
    assert rc == 0 and np.array_equal(ok.astype(bool), want_ok)
    got_ref = oracle.verify_batch(pub[bad[:64]].reshape(-1), _rows(m2, n, mlen, bad[:64]),
                                  np.arange(65, dtype=np.uint64) * mlen, h2[bad[:64]].reshape(-1), z2[bad[:64]].reshape(-1), 512,
                                  threads=0)
    assert not got_ref.any()


def test_cfg5_mixed_sha3_512_2gib_all_tiers(engine, oracle):
    """cfg 5 at the size of one rank's shard of the 8-GPU run: >= 2 GiB of SHA3-512 over lengths log-uniform in
    [64 B, 1 MiB] -- a chain-bound batch that goes through the warp, pair and thread tiers of the ragged launch.
    Sample: every message whose length hits a reference quirk (len % 72 == 71: no 0x80 is absorbed, Q1; len % 136 == 135:
    the 0x86 suffix chosen with the hard-coded rate 136, Q2) plus random ones, >= 4096 in total, against the C oracle."""
    rnd = np.random.default_rng(5)
    target = 2 << 30
    lens, acc = [], 0
    while acc < target:
        c = np.exp(rnd.uniform(np.log(64), np.log(1 << 20), size=4096)).astype(np.int64)
        lens.append(c)
        acc += int(c.sum())
    lens = np.concatenate(lens)
    lens = lens[: int(np.searchsorted(np.cumsum(lens), target)) + 1]
    # make sure the quirk lengths are present at every scale, and the extremes
    forced = [71, 135, 143, 64, 1 << 20, (1 << 20) - 1, 72 * 1000 - 1, 136 * 3000 - 1, 72 * 14000 + 71, 136 * 7000 + 135]
    lens[rnd.choice(len(lens), size=len(forced), replace=False)] = forced
    n = len(lens)
    off = np.zeros(n + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens)
    total = int(off[-1])
    block = rnd.integers(0, 256, size=64 << 20, dtype=np.uint8)
    data = np.tile(block, (total + len(block) - 1) // len(block))[:total]
    data[:: 4099] ^= np.arange(len(data[:: 4099]), dtype=np.uint64).astype(np.uint8)  # break the period of the tiling
    # what the planner does with this batch (host-side restatement of the histogram the device builds)
    got = engine.sha3(data, off, 512)
    quirk = np.nonzero((lens % 72 == 71) | (lens % 136 == 135))[0]
    rest = np.setdiff1d(np.arange(n), quirk)
    extra = rnd.choice(rest, size=max(0, SAMPLE - len(quirk)), replace=False)
    longest = np.argsort(lens)[-32:]  # the warp-tier items
    idx = np.unique(np.concatenate([quirk, extra, longest]))
    assert len(idx) >= SAMPLE and len(quirk) > 0
    s_data, s_off = pack([data[int(off[i]):int(off[i + 1])] for i in idx])
    want = oracle.sha3_batch(s_data, s_off, 512, threads=0)
    assert np.array_equal(got[idx], want)
    # the quirk lengths really are the ones where the reference leaves FIPS 202 (not a vacuous check)
    import hashlib

    q = int(quirk[0])
    m = data[int(off[q]):int(off[q + 1])].tobytes()
    assert got[q].tobytes() != hashlib.sha3_512(m).digest()
    nq = int(rest[0])
    m = data[int(off[nq]):int(off[nq + 1])].tobytes()
    assert got[nq].tobytes() == hashlib.sha3_512(m).digest()
    # device-pointer entry point, same batch: identical digests (the host path chunks, the device path does not)
    import torch

    t_data = torch.from_numpy(data).cuda()
    t_off = torch.from_numpy(off.astype(np.int64)).cuda()
    t_out = torch.zeros(n * 64, dtype=torch.uint8, device="cuda")
    engine.sha3_dev(t_data, t_off, 512, t_out)
    torch.cuda.synchronize()
    assert np.array_equal(t_out.cpu().numpy().reshape(n, 64), got)


def test_offsets_beyond_4gib(engine, oracle):
    """A device-resident batch of 5 GiB: byte offsets cross 2^32 (the reference indexes with usize; a 32-bit slip in an
    offset, a pointer difference or a block count would show on the far side of the boundary).  Mixed lengths so that the
    ragged launch orders and tiers the batch; the sample takes the messages around the 4 GiB mark, the last ones and random
    ones on both sides.  Matches compute_sha3_hash over many messages (sha3/hashable.rs:19-21)."""
    import torch

    rnd = np.random.default_rng(55)
    target = 5 << 30
    lens = np.exp(rnd.uniform(np.log(64), np.log(256 << 10), size=400_000)).astype(np.int64)
    lens = lens[: int(np.searchsorted(np.cumsum(lens), target)) + 1]
    n = len(lens)
    off = np.zeros(n + 1, dtype=np.int64)
    off[1:] = np.cumsum(lens)
    assert off[-1] > (1 << 32) + (1 << 29)
    data = torch.empty(int(off[-1]) + 16, dtype=torch.uint8, device="cuda")
    data.random_(0, 256)
    t_off = torch.from_numpy(off).cuda()
    out = torch.zeros(n * 64, dtype=torch.uint8, device="cuda")
    engine.sha3_dev(data, t_off, 512, out)
    torch.cuda.synchronize()
    k = int(np.searchsorted(off, 1 << 32))  # first message that starts at or beyond 4 GiB
    idx = np.unique(np.concatenate([np.arange(k - 8, k + 8), np.arange(n - 8, n), np.arange(8),
                                    rnd.choice(k, size=64, replace=False), k + rnd.choice(n - k, size=64, replace=False)]))
    msgs = [data[int(off[i]):int(off[i + 1])].cpu().numpy() for i in idx]
    s_data, s_off = pack(msgs)
    want = oracle.sha3_batch(s_data, s_off, 512, threads=0)
    got = out.view(n, 64)[torch.from_numpy(idx).cuda()].cpu().numpy()
    assert np.array_equal(got, want)
