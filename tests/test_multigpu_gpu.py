"""GPU (needs >= 2 devices, skipped otherwise): one ctx spanning several GPUs shards host batches across them by
contiguous ranges with no collective and returns the same bytes as a single-GPU ctx."""
import numpy as np
import pytest

from capycrypt_b200 import Engine, pack

pytestmark = pytest.mark.gpu


def _ndev():
    import torch

    return torch.cuda.device_count()


@pytest.mark.skipif("_ndev() < 2")
def test_multi_device_ctx_matches_single(engine, oracle):
    import torch

    eng2 = Engine(devices=list(range(min(_ndev(), 8))))
    assert eng2.device_count >= 2
    rnd = np.random.default_rng(3)
    lens = rnd.integers(0, 3000, size=20000)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    data = rnd.integers(0, 256, size=int(off[-1]), dtype=np.uint8)
    a = engine.sha3(data, off, 512)
    b = eng2.sha3(data, off, 512)
    assert np.array_equal(a, b)
    idx = rnd.choice(len(lens), 500, replace=False)
    msgs = [data[int(off[i]):int(off[i + 1])].tobytes() for i in idx]
    d2, o2 = pack(msgs)
    assert np.array_equal(a[idx], oracle.sha3_batch(d2, o2, 512, threads=0))
    fixed = rnd.integers(0, 256, size=100000 * 64, dtype=np.uint8)
    assert np.array_equal(engine.sha3_fixed(fixed, 64, 64, 100000, 256), eng2.sha3_fixed(fixed, 64, 64, 100000, 256))
    sc = rnd.integers(0, 256, size=3000 * 56, dtype=np.uint8)
    assert np.array_equal(engine.ed448_fixed_base(sc), eng2.ed448_fixed_base(sc))
    pws, po = pack([bytes(rnd.integers(0, 256, size=16, dtype=np.uint8)) for _ in range(600)])
    md, mo = pack([bytes(rnd.integers(0, 256, size=int(n), dtype=np.uint8)) for n in rnd.integers(0, 400, size=600)])
    h1, z1 = engine.ed448_sign(pws, po, md, mo, 512)
    h2, z2 = eng2.ed448_sign(pws, po, md, mo, 512)
    assert np.array_equal(h1, h2) and np.array_equal(z1, z2)
    pub = eng2.ed448_keygen(pws, po, 512)
    rc, ok = eng2.ed448_verify(pub, md, mo, h2, z2, 512)
    assert rc == 0 and ok.all()
    # authenticated encryption sharded the same way (ragged ciphertext ranges, per-device staging)
    nonces = rnd.integers(0, 256, size=600 * 512, dtype=np.uint8)
    c1, t1 = engine.sponge_encrypt(pws, po, nonces, 512, md, mo, 256)
    c2, t2 = eng2.sponge_encrypt(pws, po, nonces, 512, md, mo, 256)
    assert np.array_equal(c1, c2) and np.array_equal(t1, t2)
    out, ok = eng2.sponge_decrypt(pws, po, nonces, 512, c2, mo, t2, 256)
    assert ok.all() and np.array_equal(out, md)
    k = rnd.integers(0, 256, size=600 * 56, dtype=np.uint8)
    rc1, c1, t1, z1 = engine.ed448_key_encrypt(pub, k, md, mo, 512)
    rc2, c2, t2, z2 = eng2.ed448_key_encrypt(pub, k, md, mo, 512)
    assert rc1 == rc2 == 0 and np.array_equal(c1, c2) and np.array_equal(t1, t2) and np.array_equal(z1, z2)
    rc, out, ok = eng2.ed448_key_decrypt(pws, po, z2, c2, mo, t2, 512)
    assert rc == 0 and ok.all() and np.array_equal(out, md)
    eng2.close()


@pytest.mark.skipif("_ndev() < 2")
def test_multi_device_ctx_deals_out_long_messages(engine, oracle):
    """A ragged batch with outliers, in an order that is unkind to a contiguous split (all long messages first): the
    multi-device ctx deals the long chains out over its devices (lpt_shares) and still returns the digests in the
    caller's order, identical to a single-GPU ctx and to the oracle."""
    eng2 = Engine(devices=list(range(min(_ndev(), 8))))
    rnd = np.random.default_rng(9)
    lens = np.concatenate([rnd.integers(200_000, 600_000, size=37), rnd.integers(0, 800, size=30000),
                           rnd.integers(300_000, 500_000, size=5), [0, 71, 135, 72 * 3000 - 1]])
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    data = rnd.integers(0, 256, size=int(off[-1]), dtype=np.uint8)
    for d in (256, 512):
        a = engine.sha3(data, off, d)
        b = eng2.sha3(data, off, d)
        assert np.array_equal(a, b), d
    idx = np.concatenate([np.arange(40), rnd.choice(len(lens), 200, replace=False), np.arange(len(lens) - 9, len(lens))])
    d2, o2 = pack([data[int(off[i]):int(off[i + 1])].tobytes() for i in idx])
    assert np.array_equal(b[idx], oracle.sha3_batch(d2, o2, 512, threads=0))
    eng2.close()
